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
        idx = np.asarray(idx, dtype=np.int64)
        reads = [self.read(i) for i in idx]
        ops = [self.ops(i) for i in idx]
        return Batch(self.ref,
                     np.concatenate(reads) if reads else np.zeros(0, np.uint8),
                     np.concatenate(([0], np.cumsum([len(r) for r in reads]))).astype(np.int64),
                     self.ref_start[idx], self.ref_end[idx],
                     np.concatenate(ops) if ops else np.zeros(0, np.uint32),
                     np.concatenate(([0], np.cumsum([len(o) for o in ops]))).astype(np.int64),
                     [self.names[i] for i in idx] if self.names else None,
                     self.reverse[idx] if self.reverse is not None else None)
