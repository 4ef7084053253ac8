"""Parity of the sm_100a kernels (through the C ABI) against the CPU oracle.  `-m gpu` only.

Bar (north_star / SURVEY.md 8c): CIGAR ops bit-exact; posterior pairs identical as integers in units
of 1e-7 (tolerance stated by north_star is 1e-5; the kernels reproduce the oracle's arithmetic
operation for operation, so the test asks for equality); Baum-Welch statistics identical (fixed-point
accumulation) and the summed log-likelihood within 1e-9 relative.
"""
import os

import numpy as np
import pytest

import oracle
from nanopore_b200 import capi, synth
from nanopore_b200.hmm import Hmm

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def trained(golden_dir):
    return Hmm.loadHmm(os.path.join(golden_dir, "blasr_hmm_0.txt")).arrays()


def gpu_vs_oracle(ctx, model, b, band, **kw):
    params = capi.default_params(band=band, **kw)
    okw = {"split_side": kw.get("split_side", 3000), "trim": kw.get("anchor_trim", 14),
           "gap_gamma": kw.get("gap_gamma", 0.5), "match_gamma": kw.get("match_gamma", 0.0),
           "threshold": kw.get("threshold", 0.01), "min_diags": kw.get("min_diags", 1000),
           "tb_diags": kw.get("tb_diags", 40)}
    op = oracle.make_params(expansion=band, **okw)
    ops, off, post = ctx.realign_batch(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off, params,
                                       want_posteriors=True)
    assert off[0] == 0 and len(off) == b.n + 1
    cells = 0
    for i in range(b.n):
        X = b.ref[b.ref_start[i]:b.ref_end[i]]
        r = oracle.realign(model, X, b.read(i), b.ops(i), op)
        cells += r["cells"]
        g = ops[off[i]:off[i + 1]]
        assert len(g) == len(r["ops"]) and (g == r["ops"]).all(), "CIGAR of read %d differs" % i
        o = np.lexsort((r["py"], r["px"]))
        s = slice(post["off"][i], post["off"][i + 1])
        assert post["off"][i + 1] - post["off"][i] == len(o), "pair count of read %d differs" % i
        assert (post["ref_pos"][s] == r["px"][o]).all() and (post["read_pos"][s] == r["py"][o]).all()
        assert (post["prob_1e7"][s] == r["pw"][o]).all(), "posterior weights of read %d differ" % i
    st = ctx.stats()
    assert st["cells"] == cells                      # the unit of the roofline is counted identically
    assert st["launches"] >= 3
    return ops, off, post


@pytest.mark.parametrize("n,L,R,band,seed,global_form", [
    (4, 300, 900, 10, 1, True),
    (8, 1000, 3000, 20, 2, True),
    (6, 2000, 2000, 50, 3, False),
    (3, 3000, 12000, 50, 4, True),
    (5, 700, 2500, 100, 5, True),
    (4, 2500, 2500, 0, 6, False),
])
def test_stock_model(n, L, R, band, seed, global_form):
    ctx = capi.PhmmContext(0)
    b = synth.make_batch(n, L, R, seed=seed, global_form=global_form)
    ctx.set_reference(b.ref)
    gpu_vs_oracle(ctx, oracle.Model(), b, band)
    ctx.close()


@pytest.mark.parametrize("band,split", [(10, 3000), (10, 100), (50, 300)])
def test_trained_model_and_splitting(trained, band, split):
    t, e = trained
    ctx = capi.PhmmContext(0, t, e, 1)
    b = synth.make_batch(4, 600, 15000, seed=9 + band + split)
    ctx.set_reference(b.ref)
    gpu_vs_oracle(ctx, oracle.Model(t, e), b, band, split_side=split)
    assert ctx.stats()["n_regions"] >= b.n
    ctx.close()


def test_high_error_reads_and_n_bases(trained):
    t, e = trained
    ctx = capi.PhmmContext(0, t, e, 1)
    b = synth.make_batch(5, 1500, 5000, seed=21, sub=0.10, ins=0.06, dele=0.09)
    # sprinkle Ns into reads and reference (symbol 4)
    rng = np.random.default_rng(3)
    b.reads[rng.random(len(b.reads)) < 0.01] = 4
    b.ref[rng.random(len(b.ref)) < 0.01] = 4
    ctx.set_reference(b.ref)
    gpu_vs_oracle(ctx, oracle.Model(t, e), b, 50)
    ctx.close()


def test_knobs(trained):
    t, e = trained
    ctx = capi.PhmmContext(0, t, e, 1)
    b = synth.make_batch(3, 2500, 2500, seed=33, global_form=False)
    ctx.set_reference(b.ref)
    m = oracle.Model(t, e)
    gpu_vs_oracle(ctx, m, b, 20, gap_gamma=0.0, match_gamma=0.3)
    gpu_vs_oracle(ctx, m, b, 20, gap_gamma=0.9, threshold=0.05)
    gpu_vs_oracle(ctx, m, b, 20, min_diags=200, tb_diags=20, anchor_trim=4)
    ctx.close()


def test_mixed_lengths_and_degenerate_reads():
    ctx = capi.PhmmContext(0)
    lengths = [40, 3000, 1, 900, 17, 1500]
    b = synth.make_batch(len(lengths), 0, 4000, seed=44, lengths=lengths, global_form=False)
    ctx.set_reference(b.ref)
    gpu_vs_oracle(ctx, oracle.Model(), b, 10)
    # a read with no bases (pure deletion), a window with no bases (pure insertion), and an empty pair
    ref = b.ref
    reads = np.array([0, 1, 2, 3, 0], dtype=np.uint8)
    read_off = np.array([0, 0, 5, 5], dtype=np.int64)
    ref_start = np.array([10, 20, 30], dtype=np.int64)
    ref_end = np.array([25, 20, 30], dtype=np.int64)
    in_ops = synth.pack_ops([(2, 15)]).tolist() + synth.pack_ops([(1, 5)]).tolist()
    in_off = np.array([0, 1, 2, 2], dtype=np.int64)
    deg = synth.Batch(ref, reads, read_off, ref_start, ref_end, np.array(in_ops, dtype=np.uint32), in_off)
    ops, off, post = gpu_vs_oracle(ctx, oracle.Model(), deg, 10)
    assert synth.unpack_ops(ops[off[0]:off[1]]) == [(2, 15)]
    assert synth.unpack_ops(ops[off[1]:off[2]]) == [(1, 5)]
    assert off[3] == off[2]
    # empty batch
    z64 = np.zeros(1, dtype=np.int64)
    ops, off, _ = ctx.realign_batch(np.zeros(0, np.uint8), z64, np.zeros(0, np.int64), np.zeros(0, np.int64),
                                    np.zeros(0, np.uint32), z64, capi.default_params())
    assert len(ops) == 0 and off.tolist() == [0]
    ctx.close()


@pytest.mark.parametrize("legacy", [0, 1])
def test_expectations_match_oracle(trained, legacy):
    """E-step on the windowed kernel (k_fb2<EXPECT>) and on the first-generation kernel (k_fwdbwd<EXPECT>)."""
    t, e = trained
    ctx = capi.PhmmContext(0, t, e, 1)
    ctx.set_option("legacy_kernel", legacy)
    model = oracle.Model(t, e)
    b = synth.make_batch(5, 800, 2400, seed=11)
    ctx.set_reference(b.ref)
    for band, split in [(10, 300), (50, 3000)]:
        out = ctx.expectations_batch(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off,
                                     capi.default_params(band=band, split_side=split))
        T, E, ll = np.zeros(25), np.zeros(80), 0.0
        op = oracle.make_params(expansion=band, split_side=split)
        for i in range(b.n):
            T, E, ll, _ = oracle.expectations(model, b.ref[b.ref_start[i]:b.ref_end[i]], b.read(i), b.ops(i), op, T, E, ll)
        assert np.array_equal(out[:25], T) and np.array_equal(out[25:105], E)
        assert abs(out[105] - ll) <= 1e-9 * abs(ll)
    # stock (symmetric, with sX<->sY switch) model
    ctx.set_model(None, None)
    out = ctx.expectations_batch(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off,
                                 capi.default_params(band=10, split_side=300))
    T, E, ll = np.zeros(25), np.zeros(80), 0.0
    for i in range(b.n):
        T, E, ll, _ = oracle.expectations(oracle.Model(), b.ref[b.ref_start[i]:b.ref_end[i]], b.read(i), b.ops(i),
                                          oracle.make_params(expansion=10, split_side=300), T, E, ll)
    assert np.array_equal(out[:25], T) and np.array_equal(out[25:105], E)
    ctx.close()


def test_errors_are_reported_not_swallowed():
    ctx = capi.PhmmContext(0)
    b = synth.make_batch(2, 200, 600, seed=5)
    with pytest.raises(capi.PhmmError):            # no reference yet
        ctx.realign_batch(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off, capi.default_params())
    ctx.set_reference(b.ref)
    with pytest.raises(capi.PhmmError) as ei:       # odd band (upstream asserts expansion % 2 == 0)
        ctx.realign_batch(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off, capi.default_params(band=11))
    assert "even" in str(ei.value)
    bad = b.in_ops.copy()
    bad[0] += 4                                      # guide no longer spans the sequences (utils.py:381-382)
    with pytest.raises(capi.PhmmError) as ei:
        ctx.realign_batch(b.reads, b.read_off, b.ref_start, b.ref_end, bad, b.in_off, capi.default_params())
    assert "spans" in str(ei.value)
    with pytest.raises(capi.PhmmError):             # window outside the reference
        ctx.realign_batch(b.reads, b.read_off, b.ref_start, b.ref_end + 10, b.in_ops, b.in_off, capi.default_params())
    # the context stays usable after errors
    ops, off, _ = ctx.realign_batch(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off, capi.default_params())
    assert off[-1] == len(ops) > 0
    ctx.close()


def test_size_independent_properties_at_bench_shape():
    """10 kb reads vs a 50 kb reference, band 50 (BASELINE.json config 2 shape, fewer reads): results do not
    depend on batch composition or order, CIGARs span both sequences, realigning is repeatable."""
    ctx = capi.PhmmContext(0)
    b = synth.make_batch(48, 10000, 50000, seed=2)
    ctx.set_reference(b.ref)
    p = capi.default_params(band=50)
    ops, off, post = ctx.realign_batch(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off, p, True)
    st = ctx.stats()
    for i in range(b.n):
        o = synth.unpack_ops(ops[off[i]:off[i + 1]])
        assert sum(l for c, l in o if c in (0, 2)) == b.ref_end[i] - b.ref_start[i]
        assert sum(l for c, l in o if c in (0, 1)) == len(b.read(i))
        assert all(o[k][0] != o[k + 1][0] for k in range(len(o) - 1))
        matched = sum(l for c, l in o if c == 0)
        assert matched > 0.8 * len(b.read(i))
    w = post["prob_1e7"]
    assert w.min() >= 100000 and w.max() <= 10000000
    # permutation invariance + repeatability
    perm = np.random.default_rng(0).permutation(b.n)
    bp = b.subset(perm)
    ops2, off2, _ = ctx.realign_batch(bp.reads, bp.read_off, bp.ref_start, bp.ref_end, bp.in_ops, bp.in_off, p)
    for k, i in enumerate(perm):
        assert np.array_equal(ops2[off2[k]:off2[k + 1]], ops[off[i]:off[i + 1]])
    assert ctx.stats()["cells"] == st["cells"]
    # spot check three reads against the oracle at full size
    model = oracle.Model()
    for i in (0, 17, 47):
        r = oracle.realign(model, b.ref[b.ref_start[i]:b.ref_end[i]], b.read(i), b.ops(i), oracle.make_params(expansion=50))
        assert np.array_equal(r["ops"], ops[off[i]:off[i + 1]])
    # feeding the realigned CIGAR back as the guide gives a valid alignment again
    ops3, off3, _ = ctx.realign_batch(b.reads, b.read_off, b.ref_start, b.ref_end, ops, off, p)
    for i in range(b.n):
        o1 = synth.unpack_ops(ops[off[i]:off[i + 1]])
        o3 = synth.unpack_ops(ops3[off3[i]:off3[i + 1]])
        assert sum(l for c, l in o3 if c in (0, 2)) == b.ref_end[i] - b.ref_start[i]
        assert sum(l for c, l in o3 if c in (0, 1)) == len(b.read(i))
        assert sum(l for c, l in o3 if c == 0) > 0.8 * len(b.read(i)) and len(o1) > 0
    ctx.close()


def test_prepare_run_fetch_split_form():
    ctx = capi.PhmmContext(0)
    b = synth.make_batch(6, 1500, 6000, seed=8)
    ctx.set_reference(b.ref)
    p = capi.default_params(band=50)
    ops, off, _ = ctx.realign_batch(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off, p)
    ctx.prepare(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off, p)
    ctx.run()
    ctx.run()                                        # repeatable with everything resident in HBM
    ops2, off2, _ = ctx.fetch()
    assert np.array_equal(ops, ops2) and np.array_equal(off, off2)
    st = ctx.stats()
    assert st["ms_fwdbwd"] > 0 and st["n_slots"] >= 1
    # tiny memory budget: fewer resident regions, same answer
    ctx.set_memory_budget(64 << 20)
    ops3, off3, _ = ctx.realign_batch(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off, p)
    assert np.array_equal(ops, ops3)
    ctx.close()


@pytest.mark.parametrize("opts", [
    {"legacy_kernel": 1},
    {"decode_full_sweep": 1},
    {"decode_block": 1},                        # k_decode on the band for every region instead of k_decode_w on the envelope
    {"warps": 2}, {"warps": 4}, {"warps": 8},
    {"smem_columns": 64},                       # most diagonals take the wide (global buffer) path
    {"smem_columns": 64, "warps": 2},
    {"smem_columns": 1024, "warps": 8},
])
def test_kernel_variants_agree_with_the_oracle(opts):
    """Every kernel configuration (first-generation kernel, warps per region, shared-memory width with the
    global-buffer fallback for wider diagonals) gives the oracle's bits."""
    ctx = capi.PhmmContext(0)
    for k, v in opts.items():
        ctx.set_option(k, v)
    b = synth.make_batch(5, 1800, 7000, seed=77)
    ctx.set_reference(b.ref)
    gpu_vs_oracle(ctx, oracle.Model(), b, 50)
    gpu_vs_oracle(ctx, oracle.Model(), b, 10, min_diags=150, tb_diags=30)
    b2 = synth.make_batch(4, 2500, 2500, seed=78, global_form=False)
    ctx.set_reference(b2.ref)
    gpu_vs_oracle(ctx, oracle.Model(), b2, 20)
    ctx.close()


def test_pair_buffers_grow_when_a_low_threshold_overflows_them():
    """threshold 1e-6 keeps most band cells as pairs: far more than the first capacity guess (8 per row), so the
    library re-runs with larger buffers (x4 per attempt) and must still return the oracle's pairs and CIGARs."""
    ctx = capi.PhmmContext(0)
    b = synth.make_batch(3, 500, 1500, seed=91, global_form=False)
    ctx.set_reference(b.ref)
    ops, off, post = gpu_vs_oracle(ctx, oracle.Model(), b, 50, threshold=1e-6)
    per_row = (post["off"][1:] - post["off"][:-1]) / np.array([len(b.read(i)) for i in range(b.n)])
    assert per_row.max() > 10                                   # the first guess (8 per row + 1024) was exceeded
    ctx.close()


def test_option_errors():
    ctx = capi.PhmmContext(0)
    with pytest.raises(capi.PhmmError):
        ctx.set_option("warps", 3)
    with pytest.raises(capi.PhmmError):
        ctx.set_option("no_such_option", 1)
    ctx.close()
