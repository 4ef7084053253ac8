"""Packed batch layout of include/phmm.h and the base / cigar-op encodings shared by the host code.

bases    uint8  A=0 C=1 G=2 T=3, anything else 4 (N)
cigar op uint32 (length << 2) | code with the SAM codes 0 = M, 1 = I, 2 = D that the reference writes
         straight back into the SAM record (reference nanopore/analyses/utils.py:173,602)
"""
import numpy as np

BASES = np.frombuffer(b"ACGTN", dtype=np.uint8)
_CODE = np.full(256, 4, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _CODE[_c] = _i
    _CODE[ord(chr(_c).lower())] = _i

OP_M, OP_I, OP_D = 0, 1, 2


def encode(seq):
    """str/bytes -> uint8 codes A=0 C=1 G=2 T=3 other=4."""
    if isinstance(seq, str):
        seq = seq.encode("ascii")
    return _CODE[np.frombuffer(seq, dtype=np.uint8)]


def decode(codes):
    return BASES[np.asarray(codes, dtype=np.uint8)].tobytes().decode("ascii")


def reverse_complement_codes(codes):
    c = np.asarray(codes, dtype=np.uint8)[::-1].copy()
    m = c < 4
    c[m] = 3 - c[m]
    return c


def pack_ops(ops):
    """[(code, length), ...] -> uint32 (length<<2)|code, dropping empty ops and merging neighbours of equal code."""
    a = np.asarray(ops, dtype=np.int64).reshape(-1, 2)
    a = a[a[:, 1] > 0]
    if len(a) == 0:
        return np.zeros(0, dtype=np.uint32)
    code, ln = a[:, 0], a[:, 1]
    first = np.concatenate(([True], code[1:] != code[:-1]))
    if not first.all():
        ln = np.add.reduceat(ln, np.flatnonzero(first))
        code = code[first]
    return ((ln << 2) | code).astype(np.uint32)


def unpack_ops(packed):
    p = np.asarray(packed, dtype=np.int64)
    return list(zip((p & 3).tolist(), (p >> 2).tolist()))



class Batch:
    """Packed batch in the layout of include/phmm.h."""

    def __init__(self, ref, reads, read_off, ref_start, ref_end, in_ops, in_off, names=None, reverse=None):
        self.ref = np.ascontiguousarray(ref, dtype=np.uint8)
        self.reads = np.ascontiguousarray(reads, dtype=np.uint8)
        self.read_off = np.ascontiguousarray(read_off, dtype=np.int64)
        self.ref_start = np.ascontiguousarray(ref_start, dtype=np.int64)
        self.ref_end = np.ascontiguousarray(ref_end, dtype=np.int64)
        self.in_ops = np.ascontiguousarray(in_ops, dtype=np.uint32)
        self.in_off = np.ascontiguousarray(in_off, dtype=np.int64)
        self.names = names
        self.reverse = reverse

    @property
    def n(self):
        return len(self.read_off) - 1

    def read(self, i):
        return self.reads[self.read_off[i]:self.read_off[i + 1]]

    def ops(self, i):
        return self.in_ops[self.in_off[i]:self.in_off[i + 1]]

    def subset(self, idx):
        """Reads idx (any order) as a new Batch sharing the reference; no per-read Python work."""
        idx = np.asarray(idx, dtype=np.int64)
        reads, read_off = _gather_ranges(self.reads, self.read_off, idx)
        ops, in_off = _gather_ranges(self.in_ops, self.in_off, idx)
        return Batch(self.ref, reads, read_off, self.ref_start[idx], self.ref_end[idx], ops, in_off,
                     [self.names[i] for i in idx] if self.names else None,
                     self.reverse[idx] if self.reverse is not None else None)


def _native():
    """libphmm_io.so when it is built (threaded C++); the numpy forms below compute the same."""
    try:
        from . import io_native
        return io_native if io_native.available() else None
    except Exception:
        return None


def _gather_ranges(data, off, idx, out=None):
    """data[off[i]:off[i+1]] for i in idx, concatenated (into `out` when given), with the new offsets."""
    nat = _native()
    if nat is not None and len(idx) > 64:
        return nat.gather_ranges(data, off, idx, out=out)
    lens = off[idx + 1] - off[idx] if len(idx) else np.zeros(0, dtype=np.int64)
    new_off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    total = int(new_off[-1])
    if total == 0:
        return (data[:0].copy() if out is None else out), new_off
    pos = np.arange(total, dtype=np.int64) + np.repeat(off[idx] - new_off[:-1], lens)
    if out is None:
        return data[pos], new_off
    np.take(data, pos, out=out)
    return out, new_off


def estimate_cells(batch, band=10, anchor_trim=14, split_side=3000):
    """Cheap host estimate of the DP cells of every read (the unit of work of the realigner): anchor runs cost
    2 n (band + 1) cells, the block between two anchor runs its band rectangle (dx + band + 1)(dy + band + 1), and a
    block larger than split_side^2 only its two corner rectangles.  Vectorised over all guide
    ops of the batch.  Used to balance shards and to bound the size of one library call; the exact count comes
    back from the library (phmm_batch_stats.cells)."""
    n = batch.n
    if n == 0:
        return np.zeros(0, dtype=np.int64)
    nat = _native()
    if nat is not None:
        return nat.estimate_cells(batch.in_ops, batch.in_off, batch.read_off, batch.ref_start, batch.ref_end, band, anchor_trim,
                                  split_side)
    return _estimate_cells_numpy(batch, band, anchor_trim, split_side)


def _estimate_cells_numpy(batch, band=10, anchor_trim=14, split_side=3000):
    n = batch.n
    code = (batch.in_ops & 3).astype(np.int64)
    ln = (batch.in_ops >> 2).astype(np.int64)
    nops = batch.in_off[1:] - batch.in_off[:-1]
    rid = np.repeat(np.arange(n, dtype=np.int64), nops)
    xadv = np.where((code == 0) | (code == 2), ln, 0)
    yadv = np.where((code == 0) | (code == 1), ln, 0)
    cx, cy = np.cumsum(xadv), np.cumsum(yadv)
    bx = np.concatenate(([0], cx))[batch.in_off[:-1]]          # totals before each read's first op
    by = np.concatenate(([0], cy))[batch.in_off[:-1]]
    x0 = cx - xadv - bx[rid]                                     # start of each op, read-local
    y0 = cy - yadv - by[rid]
    lX = (batch.ref_end - batch.ref_start).astype(np.int64)
    lY = (batch.read_off[1:] - batch.read_off[:-1]).astype(np.int64)
    e = int(band)
    a = (code == 0) & (ln > 2 * anchor_trim)
    ar, ax0, ay0 = rid[a], x0[a] + anchor_trim, y0[a] + anchor_trim
    an = ln[a] - 2 * anchor_trim
    cost = np.zeros(n, dtype=np.float64)
    np.add.at(cost, ar, 2.0 * an * (e + 1))
    # blocks: previous anchor end (or the origin) -> next anchor start (or the far corner)
    first = np.concatenate(([True], ar[1:] != ar[:-1])) if len(ar) else np.zeros(0, dtype=bool)
    last = np.concatenate((ar[1:] != ar[:-1], [True])) if len(ar) else np.zeros(0, dtype=bool)
    px = np.where(first, 0, np.concatenate(([0], (ax0 + an)[:-1])))
    py = np.where(first, 0, np.concatenate(([0], (ay0 + an)[:-1])))
    bdx = np.concatenate((ax0 - px, (lX[ar] - (ax0 + an))[last]))
    bdy = np.concatenate((ay0 - py, (lY[ar] - (ay0 + an))[last]))
    br = np.concatenate((ar, ar[last]))
    has = np.zeros(n, dtype=bool)
    has[ar] = True
    bdx = np.concatenate((bdx, lX[~has]))                         # reads without anchors: one block
    bdy = np.concatenate((bdy, lY[~has]))
    br = np.concatenate((br, np.flatnonzero(~has)))
    bdx, bdy = np.maximum(bdx, 0).astype(np.float64), np.maximum(bdy, 0).astype(np.float64)
    big = bdx * bdy > float(split_side) ** 2
    hx, hy = np.minimum(np.floor(bdx / 2), split_side), np.minimum(np.floor(bdy / 2), split_side)
    blk = np.where(big, 2.0 * (hx + e + 1) * (hy + e + 1), (bdx + e + 1) * (bdy + e + 1))
    np.add.at(cost, br, blk)
    return np.maximum(cost, 1.0).astype(np.int64)
