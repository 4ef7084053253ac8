"""Full-size parity (`-m gpu`): posterior pairs and Baum-Welch integers, not only CIGARs, at the shapes
BASELINE.json names -- where the long traceback windows, the posterior-candidate shortcut of k_fb2 see inputs the small cases never produce.

    config 2   10 kb reads vs a 50 kb reference, band 50            (utils.py:587 flags, band per BASELINE.json)
    config 3   5 kb template/complement reads vs a 1 Mb contig, band 100, chained-global guides whose
               leading / trailing deletions are split away (split 3000)
    config 4   E-step, 8 kb reads, band 50, split 300               (utils.py:511)
    config 5   one 50 kb read (tail of the Pareto length distribution)

The oracle runs one read per host thread (ctypes drops the GIL); sizes are chosen so that the whole
file stays within a few minutes of CPU time on the GPU box.
"""
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import oracle
from nanopore_b200 import capi, synth
from nanopore_b200.hmm import Hmm

pytestmark = pytest.mark.gpu

THREADS = max(1, min(32, os.cpu_count() or 1))


def oracle_realign_all(model, b, op):
    oracle.lib()
    with ThreadPoolExecutor(THREADS) as ex:
        return list(ex.map(lambda i: oracle.realign(model, b.ref[b.ref_start[i]:b.ref_end[i]], b.read(i), b.ops(i), op),
                           range(b.n)))


def assert_same_as_oracle(b, ops, off, post, want):
    for i, r in enumerate(want):
        g = ops[off[i]:off[i + 1]]
        assert len(g) == len(r["ops"]) and (g == r["ops"]).all(), "CIGAR of read %d differs" % i
        o = np.lexsort((r["py"], r["px"]))
        s = slice(post["off"][i], post["off"][i + 1])
        assert post["off"][i + 1] - post["off"][i] == len(o), "pair count of read %d differs" % i
        assert (post["ref_pos"][s] == r["px"][o]).all() and (post["read_pos"][s] == r["py"][o]).all()
        assert (post["prob_1e7"][s] == r["pw"][o]).all(), "posterior weights of read %d differ" % i


@pytest.fixture(scope="module")
def trained(golden_dir):
    return Hmm.loadHmm(os.path.join(golden_dir, "blasr_hmm_0.txt")).arrays()


@pytest.fixture(scope="module")
def config2():
    """16 reads of the bench workload's shape with the oracle's answers (computed once for the module)."""
    b = synth.make_batch(16, 10000, 50000, seed=2002)
    want = oracle_realign_all(oracle.Model(), b, oracle.make_params(expansion=50))
    return b, want


@pytest.mark.parametrize("opts", [
    {},                                            # the shipped configuration
    {"candidate_eps_ppm": 0},                      # no total is within 0 of the first one: every window re-reads all cells
    {"candidate_cap": 64},                         # the candidate list overflows in every window
    {"warps": 2},
    {"warps": 8},
    {"decode_block": 1},                           # block-per-region decode on the band instead of the envelope kernel
])
def test_config2_shape_posteriors_and_cigars(config2, opts):
    b, want = config2
    ctx = capi.PhmmContext(0)
    for k, v in opts.items():
        ctx.set_option(k, v)
    ctx.set_reference(b.ref)
    ops, off, post = ctx.realign_batch(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off,
                                       capi.default_params(band=50), want_posteriors=True)
    assert ctx.stats()["cells"] == sum(r["cells"] for r in want)
    assert_same_as_oracle(b, ops, off, post, want)
    ctx.close()


def test_config3_shape_megabase_contig(trained):
    t, e = trained
    b = synth.make_batch(6, 5000, 1_000_000, seed=3003, sub=0.10, ins=0.06, dele=0.09)
    ctx = capi.PhmmContext(0, t, e, 1)
    ctx.set_reference(b.ref)
    ops, off, post = ctx.realign_batch(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off,
                                       capi.default_params(band=100), want_posteriors=True)
    st = ctx.stats()
    assert st["n_regions"] > b.n                     # the megabase leading / trailing deletions were split away
    want = oracle_realign_all(oracle.Model(t, e), b, oracle.make_params(expansion=100))
    assert st["cells"] == sum(r["cells"] for r in want)
    assert_same_as_oracle(b, ops, off, post, want)
    ctx.close()


def test_config4_shape_estep_integers(trained):
    t, e = trained
    b = synth.make_batch(8, 8000, 8000, seed=4004, global_form=False)      # EM inputs are global: window = contig (utils.py:492-496)
    ctx = capi.PhmmContext(0, t, e, 1)
    ctx.set_reference(b.ref)
    hi, lo = ctx.expectations_batch_fixed(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off,
                                          capi.default_params(band=50, split_side=300))
    model = oracle.Model(t, e)
    op = oracle.make_params(expansion=50, split_side=300)
    oracle.lib()

    def one(i):
        return oracle.expectations_fixed(model, b.ref[b.ref_start[i]:b.ref_end[i]], b.read(i), b.ops(i), op)

    with ThreadPoolExecutor(THREADS) as ex:
        parts = list(ex.map(one, range(b.n)))
    # exact integers: hi * 2^bits + lo, 32 fractional bits for the 105 expectations, 20 for the log-likelihood
    bits = [32] * 105 + [20]
    want = [sum((int(h[k]) << bits[k]) + int(l[k]) for h, l, _ in parts) for k in range(106)]
    got = [(int(hi[k]) << bits[k]) + int(lo[k]) for k in range(106)]
    assert got == want
    assert sum(got[:25]) > 0 and sum(got[25:105]) > 0
    ctx.close()


def test_config5_tail_one_50kb_read():
    b = synth.make_batch(1, 50000, 60000, seed=5005)
    ctx = capi.PhmmContext(0)
    ctx.set_reference(b.ref)
    ops, off, post = ctx.realign_batch(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off,
                                       capi.default_params(band=50), want_posteriors=True)
    want = oracle_realign_all(oracle.Model(), b, oracle.make_params(expansion=50))
    assert ctx.stats()["cells"] == want[0]["cells"]
    assert_same_as_oracle(b, ops, off, post, want)
    ctx.close()
