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
from .batch import Batch

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


def chunk_bounds(batch, max_bases):
    """Cuts [0, n) into contiguous runs of reads with at most max_bases read + window bases each (>= 1 read)."""
    size = (batch.read_off[1:] - batch.read_off[:-1]) + (batch.ref_end - batch.ref_start)
    bounds, start, acc = [], 0, 0
    for i in range(batch.n):
        if i > start and acc + size[i] > max_bases:
            bounds.append((start, i))
            start, acc = i, 0
        acc += int(size[i])
    if batch.n > start:
        bounds.append((start, batch.n))
    return bounds


class Realigner:
    """ctx: anything with the PhmmContext methods (tests inject a CPU checker here; the product never does)."""

    def __init__(self, device=0, hmm=None, ctx=None, max_bases_per_call=400_000_000):
        if ctx is None:
            if hmm is None:
                ctx = capi.PhmmContext(device)
            else:
                t, e = hmm.arrays()
                ctx = capi.PhmmContext(device, t, e, hmm.type)
        self.ctx = ctx
        self.max_bases = int(max_bases_per_call)
        self.cells = 0

    def close(self):
        self.ctx.close()

    def set_hmm(self, hmm):
        if hmm is None:
            self.ctx.set_model(None, None, 1)
        else:
            t, e = hmm.arrays()
            self.ctx.set_model(t, e, hmm.type)

    def set_reference(self, codes):
        self.ctx.set_reference(codes)

    def realign(self, batch, params, want_posteriors=False):
        """-> (ops uint32, off int64[n+1], posteriors dict or None), reads in input order."""
        ops_l, off_l, posts = [], [np.zeros(1, dtype=np.int64)], []
        base = 0
        pbase = 0
        self.cells = 0
        for a, b in chunk_bounds(batch, self.max_bases):
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

    def expectations(self, batch, params):
        """E-step over the batch -> FixedStats (exact, order independent)."""
        tot = FixedStats()
        self.cells = 0
        for a, b in chunk_bounds(batch, self.max_bases):
            sub = batch if (a, b) == (0, batch.n) else batch.subset(np.arange(a, b))
            hi, lo = self.ctx.expectations_batch_fixed(sub.reads, sub.read_off, sub.ref_start, sub.ref_end, sub.in_ops,
                                                       sub.in_off, params)
            self.cells += int(self.ctx.stats()["cells"])
            tot += FixedStats(hi, lo)
        return tot
