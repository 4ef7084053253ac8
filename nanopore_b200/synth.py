"""Seeded synthetic workloads for the realignment path.

No genomes or nanopore reads are available offline, so the benchmark inputs are
generated: an i.i.d. ACGT reference, reads drawn from a window of it through a
substitution/insertion/deletion channel, and the guide alignment a mapper would
hand to the realigner in the chained-global form `mergeChainedAlignedReads`
emits (reference nanopore/analyses/utils.py:344-346,366-382: pos = 0, leading
`D` up to the window start, trailing `D` to the reference end).

The error channel follows the spirit of the reference's own
`mutateSequence` (nanopore/analyses/mutate_reference.py:11-12), extended with
indels.
"""
import numpy as np

from .batch import (BASES, OP_M, OP_I, OP_D, Batch, decode, encode, pack_ops,  # noqa: F401
                    reverse_complement_codes, unpack_ops)


def random_reference(length, rng):
    return rng.integers(0, 4, size=length, dtype=np.uint8)


def _rle(ops):
    """uint8 op stream -> packed uint32 runs."""
    if len(ops) == 0:
        return np.zeros(0, dtype=np.uint32)
    cut = np.flatnonzero(np.diff(ops)) + 1
    starts = np.concatenate(([0], cut))
    ends = np.concatenate((cut, [len(ops)]))
    return ((ends - starts).astype(np.uint32) << 2) | ops[starts].astype(np.uint32)


def _fold_short_matches(runs, min_match_run):
    """Matched runs shorter than min_match_run become D(k) I(k); every gap block between the surviving
    M runs is then canonicalised as one D followed by one I (vectorised)."""
    code = (runs & 3).astype(np.int64)
    ln = (runs >> 2).astype(np.int64)
    keep_m = (code == OP_M) & (ln >= min_match_run)
    dl = np.where((code == OP_D) | ((code == OP_M) & ~keep_m), ln, 0)
    il = np.where((code == OP_I) | ((code == OP_M) & ~keep_m), ln, 0)
    # gap block g = number of kept M runs before it
    g = np.cumsum(keep_m) - keep_m
    nm = int(keep_m.sum())
    dsum = np.bincount(g[~keep_m], weights=dl[~keep_m], minlength=nm + 1).astype(np.int64)
    isum = np.bincount(g[~keep_m], weights=il[~keep_m], minlength=nm + 1).astype(np.int64)
    mlen = ln[keep_m]
    # interleave: D_0 I_0 M_0 D_1 I_1 M_1 ... D_nm I_nm
    out_code = np.empty(3 * nm + 2, dtype=np.int64)
    out_len = np.empty(3 * nm + 2, dtype=np.int64)
    out_code[0::3] = OP_D
    out_len[0::3] = dsum
    out_code[1::3] = OP_I
    out_len[1::3] = isum
    out_code[2::3] = OP_M
    out_len[2::3] = mlen
    ok = out_len > 0
    out_code, out_len = out_code[ok], out_len[ok]
    # merge neighbouring M runs that became adjacent (empty gap block between them)
    if len(out_code) > 1:
        first = np.concatenate(([True], out_code[1:] != out_code[:-1]))
        grp = np.cumsum(first) - 1
        out_len = np.bincount(grp, weights=out_len).astype(np.int64)
        out_code = out_code[first]
    return ((out_len.astype(np.uint32) << 2) | out_code.astype(np.uint32)).astype(np.uint32)


def simulate_read(ref, start, length, rng, sub=0.05, ins=0.04, dele=0.06, geo_p=0.6, min_match_run=8):
    """One read from ref[start:start+length].

    Returns (read codes, local guide ops packed) where the ops describe the true
    edit script degraded the way a seed-and-extend mapper would report it:
    matched runs shorter than `min_match_run` are folded into the neighbouring
    indels.  Ops span exactly `length` reference bases and
    len(read) read bases.
    """
    w = ref[start:start + length]
    n = len(w)
    # deletions: runs of reference bases dropped
    dstart = rng.random(n) < dele
    dlen = rng.geometric(geo_p, size=n)
    mark = np.zeros(n + 1, dtype=np.int64)
    idx = np.flatnonzero(dstart)
    np.add.at(mark, idx, 1)
    np.add.at(mark, np.minimum(idx + dlen[idx], n), -1)
    deleted = np.cumsum(mark[:n]) > 0
    # never delete the first/last base of the window: keeps the local script anchored
    deleted[0] = deleted[-1] = False
    # insertions before kept positions
    ilen = np.where(rng.random(n) < ins, rng.geometric(geo_p, size=n), 0)
    ilen[0] = 0
    ilen[deleted] = 0
    # column stream: for each reference position: ilen[i] I-columns then one M or D column
    cols_per = ilen + 1
    total = int(cols_per.sum())
    ops = np.full(total, OP_I, dtype=np.uint8)
    last = np.cumsum(cols_per) - 1
    ops[last] = np.where(deleted, OP_D, OP_M)
    # read bases
    kept = ~deleted
    bases = w.copy()
    smask = (rng.random(n) < sub) & kept
    bases[smask] = (bases[smask] + rng.integers(1, 4, size=int(smask.sum()), dtype=np.uint8)) % 4
    read = np.empty(int(ilen.sum() + kept.sum()), dtype=np.uint8)
    is_read_col = ops != OP_D
    col_is_m = ops == OP_M
    read_cols = np.flatnonzero(is_read_col)
    m_in_read = col_is_m[read_cols]
    read[m_in_read] = bases[kept]
    read[~m_in_read] = rng.integers(0, 4, size=int((~m_in_read).sum()), dtype=np.uint8)
    runs = _rle(ops)
    if min_match_run > 1 and len(runs):
        runs = _fold_short_matches(runs, min_match_run)
    return read, runs


def globalise(local_ops, start, length, ref_len):
    """Chained-global form (utils.py:344-346,366-382): leading/trailing D to span the reference."""
    ops = [(OP_D, start)] + unpack_ops(local_ops) + [(OP_D, ref_len - start - length)]
    return pack_ops(ops)


def make_batch(n_reads, read_len, ref_len, seed, sub=0.05, ins=0.04, dele=0.06, global_form=True,
               min_match_run=8, lengths=None, ref=None):
    """n_reads reads of ~read_len reference bases each from one random reference.

    lengths: optional per-read reference-window lengths (mixed-length config).
    With global_form the guide alignment spans the whole reference the way the
    chaining step hands it to the realigner; otherwise ref_start/ref_end are
    the read's own window.
    """
    rng = np.random.default_rng(seed)
    if ref is None:
        ref = random_reference(ref_len, rng)
    else:
        ref = np.ascontiguousarray(ref, dtype=np.uint8)
        ref_len = len(ref)
    reads, ops, rs, re_, names, rev = [], [], [], [], [], []
    for i in range(n_reads):
        ln = int(lengths[i]) if lengths is not None else read_len
        ln = min(ln, ref_len)
        start = int(rng.integers(0, ref_len - ln + 1))
        r, o = simulate_read(ref, start, ln, rng, sub, ins, dele, min_match_run=min_match_run)
        if global_form:
            o = globalise(o, start, ln, ref_len)
            rs.append(0)
            re_.append(ref_len)
        else:
            rs.append(start)
            re_.append(start + ln)
        reads.append(r)
        ops.append(o)
        names.append("read_%d" % i)
        rev.append(bool(rng.random() < 0.5))
    read_off = np.concatenate(([0], np.cumsum([len(r) for r in reads]))).astype(np.int64)
    in_off = np.concatenate(([0], np.cumsum([len(o) for o in ops]))).astype(np.int64)
    return Batch(ref, np.concatenate(reads), read_off, rs, re_, np.concatenate(ops), in_off, names,
                 np.array(rev, dtype=bool))


def pareto_lengths(n, seed, lo=500, hi=50000, alpha=1.2):
    """Mixed-length config: clip(lo * Pareto(alpha), lo, hi)."""
    rng = np.random.default_rng(seed)
    return np.clip(lo * (1.0 + rng.pareto(alpha, size=n)), lo, hi).astype(np.int64)
