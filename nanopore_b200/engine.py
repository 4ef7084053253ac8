"""Realigner: one GPU's view of the realignment path, on top of the C ABI (capi.PhmmContext).

Replaces the reference's per-read job fan-out (reference nanopore/analyses/utils.py:557-609) and the E-step
fan-out of cactus_expectationMaximisation (utils.py:528): a packed Batch goes in, CIGAR ops (or Baum-Welch
sufficient statistics) come out, in input order.  Large batches are cut into calls bounded by read bases so the
posterior-pair buffers stay a small part of HBM.

Statistics of the E-step travel as exact integers (`FixedStats`): per value a pair (hi, lo) with
value = hi + lo / 2^32 (log-likelihood: lo / 2^20).  Integer sums do not depend on evaluation order, so any
sharding of the reads over calls, ranks or GPUs yields bit-identical HMMs.
"""
import numpy as np

from . import capi
from .batch import estimate_cells, Batch

N_STATS = 106
LO_BITS = np.array([32] * 105 + [20], dtype=np.int64)


class FixedStats:
    """Exact sums of Baum-Welch statistics: [0..24] transitions from*5+to, [25..104] emissions
    state*16+x*4+y, [105] log-likelihood."""

    def __init__(self, hi=None, lo=None):
        self.hi = np.zeros(N_STATS, dtype=np.int64) if hi is None else np.asarray(hi, dtype=np.int64).copy()
        self.lo = np.zeros(N_STATS, dtype=np.int64) if lo is None else np.asarray(lo, dtype=np.int64).copy()

    def __iadd__(self, o):
        self.hi += o.hi
        self.lo += o.lo
        return self.normalised()

    def normalised(self):
        """Carries lo into hi (lo stays in [0, 2^bits)); keeps lo far from int64 overflow when many are summed."""
        carry = self.lo >> LO_BITS
        self.hi += carry
        self.lo -= carry << LO_BITS
        return self

    def values(self):
        """float64[106]; a single rounding per value, applied after the exact sum."""
        self.normalised()
        return self.hi.astype(np.float64) + self.lo.astype(np.float64) / np.exp2(LO_BITS.astype(np.float64))

    def as_tensor_array(self):
        return np.concatenate([self.hi, self.lo])

    @staticmethod
    def from_tensor_array(a):
        a = np.asarray(a, dtype=np.int64)
        return FixedStats(a[:N_STATS], a[N_STATS:]).normalised()

    def __eq__(self, o):
        a, b = FixedStats(self.hi, self.lo).normalised(), FixedStats(o.hi, o.lo).normalised()
        return bool(np.array_equal(a.hi, b.hi) and np.array_equal(a.lo, b.lo))


def chunk_bounds(batch, max_bases, max_cells=None, params=None):
    """Cuts [0, n) into contiguous runs of reads (>= 1 read each) for one library call.

    The bound that matters is the work and scratch of a call -- diagonal records, pair buffers and forward windows all
    scale with the DP cells -- so a run holds at most max_cells ESTIMATED cells (batch.estimate_cells).  Raw window
    length says little: chained-global records span the whole contig (reference utils.py:344-346), yet after
    splitting only the aligned part is swept.  max_bases (read + window bases) is kept as a second, optional cap."""
    size = (batch.read_off[1:] - batch.read_off[:-1]) + (batch.ref_end - batch.ref_start)
    cells = None
    if max_cells is not None:
        cells = estimate_cells(batch) if params is None else estimate_cells(batch, params.band, params.anchor_trim, params.split_side)
    bounds, start, acc, acc_c = [], 0, 0, 0
    for i in range(batch.n):
        over = (max_bases is not None and acc + size[i] > max_bases) or (cells is not None and acc_c + cells[i] > max_cells)
        if i > start and over:
            bounds.append((start, i))
            start, acc, acc_c = i, 0, 0
        acc += int(size[i])
        if cells is not None:
            acc_c += int(cells[i])
    if batch.n > start:
        bounds.append((start, batch.n))
    return bounds


class Realigner:
    """ctx: anything with the PhmmContext methods (tests inject a CPU checker here; the product never does)."""

    def __init__(self, device=0, hmm=None, ctx=None, max_bases_per_call=None, max_cells_per_call=150_000_000_000):
        if ctx is None:
            if hmm is None:
                ctx = capi.PhmmContext(device)
            else:
                t, e = hmm.arrays()
                ctx = capi.PhmmContext(device, t, e, hmm.type)
        self.ctx = ctx
        self.max_bases = int(max_bases_per_call) if max_bases_per_call is not None else None
        self.max_cells = int(max_cells_per_call) if max_cells_per_call is not None else None
        self.cells = 0
        self._em_batch = None          # batch whose E-step plan is resident in the context (EM runs hundreds of iterations on it)
        self._em_params = None

    def close(self):
        self.ctx.close()

    def set_hmm(self, hmm):
        if hmm is None:
            self.ctx.set_model(None, None, 1)
        else:
            t, e = hmm.arrays()
            self.ctx.set_model(t, e, hmm.type)

    def set_reference(self, codes):
        self._em_batch = None
        self.ctx.set_reference(codes)

    def realign(self, batch, params, want_posteriors=False):
        """-> (ops uint32, off int64[n+1], posteriors dict or None), reads in input order."""
        ops_l, off_l, posts = [], [np.zeros(1, dtype=np.int64)], []
        self._em_batch = None
        base = 0
        pbase = 0
        self.cells = 0
        for a, b in chunk_bounds(batch, self.max_bases, self.max_cells, params):
            sub = batch if (a, b) == (0, batch.n) else batch.subset(np.arange(a, b))
            ops, off, post = self.ctx.realign_batch(sub.reads, sub.read_off, sub.ref_start, sub.ref_end, sub.in_ops,
                                                    sub.in_off, params, want_posteriors=want_posteriors)
            self.cells += int(self.ctx.stats()["cells"])
            ops_l.append(ops)
            off_l.append(off[1:] + base)
            base += int(off[-1])
            if want_posteriors:
                posts.append((post, pbase))
                pbase += int(post["off"][-1])
        ops = np.concatenate(ops_l) if ops_l else np.zeros(0, dtype=np.uint32)
        off = np.concatenate(off_l)
        post = None
        if want_posteriors:
            post = {"off": np.concatenate([np.zeros(1, dtype=np.int64)] + [p["off"][1:] + pb for p, pb in posts]),
                    "ref_pos": np.concatenate([p["ref_pos"] for p, _ in posts] or [np.zeros(0, np.int32)]),
                    "read_pos": np.concatenate([p["read_pos"] for p, _ in posts] or [np.zeros(0, np.int32)]),
                    "prob_1e7": np.concatenate([p["prob_1e7"] for p, _ in posts] or [np.zeros(0, np.int32)])}
        return ops, off, post

    def base_expectations(self, batch, params, masks=None):
        """Per-reference-position sums of posterior match probability by read base (the table
        marginAlignSnpCaller.py:149-155 builds from every read's --outputAllPosteriorProbs file), accumulated on the
        device: -> int64[len(masks) or 1, reference length, 5] in units of 1e-7 (read base A C G T other).  masks: boolean arrays over the reads
        of `batch`, one per sample of reads (coverage replicates); all of them share ONE forward/backward pass."""
        self._em_batch = None
        ref_len = len(batch.ref)
        ms = [None] if masks is None else [np.ascontiguousarray(m, dtype=np.uint8) for m in masks]
        self.ctx.base_expectations_reset(len(ms))
        self.cells = 0
        for a, b in chunk_bounds(batch, self.max_bases, self.max_cells, params):
            sub = batch if (a, b) == (0, batch.n) else batch.subset(np.arange(a, b))
            self.ctx.prepare(sub.reads, sub.read_off, sub.ref_start, sub.ref_end, sub.in_ops, sub.in_off, params)
            self.ctx.run()
            self.cells += int(self.ctx.stats()["cells"])
            for k, m in enumerate(ms):
                if m is None or m[a:b].any():
                    self.ctx.add_base_expectations(None if m is None else m[a:b], k)
        return np.stack([self.ctx.base_expectations_fetch(ref_len, k) for k in range(len(ms))])

    def expectations(self, batch, params):
        """E-step over the batch -> FixedStats (exact, order independent)."""
        tot = FixedStats()
        bounds = chunk_bounds(batch, self.max_bases, self.max_cells, params)
        if len(bounds) == 1 and hasattr(self.ctx, "expectations_prepare"):
            # resident E-step: plan + upload once per batch, one kernel launch per EM iteration (utils.py:509-523 runs
            # 3 trials x 100 iterations over the same alignments; set_hmm between iterations keeps the plan)
            key = (params.band, params.anchor_trim, params.split_side, params.min_diags, params.tb_diags, params.threshold)
            if self._em_batch is not batch or self._em_params != key:
                self.ctx.expectations_prepare(batch.reads, batch.read_off, batch.ref_start, batch.ref_end, batch.in_ops,
                                              batch.in_off, params)
                self._em_batch, self._em_params = batch, key
            hi, lo = self.ctx.expectations_run_fixed()
            self.cells = int(self.ctx.stats()["cells"])
            return FixedStats(hi, lo)
        self._em_batch = None
        self.cells = 0
        for a, b in bounds:
            sub = batch if (a, b) == (0, batch.n) else batch.subset(np.arange(a, b))
            hi, lo = self.ctx.expectations_batch_fixed(sub.reads, sub.read_off, sub.ref_start, sub.ref_end, sub.in_ops,
                                                       sub.in_off, params)
            self.cells += int(self.ctx.stats()["cells"])
            tot += FixedStats(hi, lo)
        return tot
