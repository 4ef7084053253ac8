"""The plugin path on the real library (`-m gpu`): SAM in -> chain -> batched realignment on the B200 -> SAM out, EM
through the E-step kernel, posterior epilogues; every result compared with the same host code driven by the CPU
checker (bit-exact CIGARs, identical trained HMM files)."""
import filecmp
import os
import shutil

import numpy as np
import pytest

import oracle
from nanopore_b200 import capi, em, posteriors, realign, synth
from nanopore_b200.engine import FixedStats, Realigner
from nanopore_b200.mappers.abstractMapper import AbstractMapper
from nanopore_b200.target import Stack, Target

from helpers_sam import make_experiment
from oracle_ctx import oracle_realigner_factory

pytestmark = pytest.mark.gpu


class MapperRealign(AbstractMapper):
    kw = {}

    def run(self):
        self.realignSamFile(**self.kw)


def run_both(tmp_path, cls, seed, **exp):
    """Runs the mapper plugin twice on copies of one experiment: GPU library, then CPU checker."""
    outs = []
    for tag, factory in (("gpu", None), ("cpu", oracle_realigner_factory())):
        d = str(tmp_path / tag)
        ref_fa, fq, sam_path, _ = make_experiment(d, seed=seed, **exp)
        prev = realign.setRealignerFactory(factory) if factory else None
        try:
            m = cls(fq, "2D", ref_fa, sam_path, emptyHmmFile=os.path.join(d, "hmm.txt"))
            assert Stack(m).startJobTree(None) == 0
        finally:
            if factory:
                realign.setRealignerFactory(prev)
        outs.append(d)
    return outs


def test_realign_sam_file_on_gpu_equals_checker(tmp_path):
    g, c = run_both(tmp_path, MapperRealign, seed=3, n_reads=12, read_len=700, contig_lens=(4000, 2500))
    assert filecmp.cmp(os.path.join(g, "mapping.sam"), os.path.join(c, "mapping.sam"), shallow=False)

    class Trained(MapperRealign):
        kw = dict(useTrainedModel=True, trainedModelFile="blasr_hmm_40.txt", gapGamma=0.3)
    g, c = run_both(tmp_path / "t", Trained, seed=4, n_reads=8, read_len=500)
    assert filecmp.cmp(os.path.join(g, "mapping.sam"), os.path.join(c, "mapping.sam"), shallow=False)


def test_em_on_gpu_trains_the_same_model_as_the_checker(tmp_path, monkeypatch):
    fast = em.Options()
    fast.modelType, fast.randomStart, fast.trials, fast.iterations, fast.trainEmissions = "fiveStateAsymmetric", True, 2, 3, True
    fast.outputTrialHmms = True
    orig = em.learnModelFromSamFileTargetFn
    monkeypatch.setattr(em, "learnModelFromSamFileTargetFn", lambda t, *a: orig(t, *a, options=fast))

    class Em(MapperRealign):
        kw = dict(doEm=True)
    g, c = run_both(tmp_path, Em, seed=5, n_reads=10, read_len=400)
    for f in ("hmm.txt_unnormalised", "hmm.txt", "hmm.txt.xml", "hmm.txt_unnormalised_0", "hmm.txt_unnormalised_1", "mapping.sam"):
        assert filecmp.cmp(os.path.join(g, f), os.path.join(c, f), shallow=False), f


def test_expectations_fixed_point_and_chunking_on_gpu():
    b = synth.make_batch(9, 600, 2000, seed=13)
    p = capi.default_params(band=10, split_side=300)
    r = Realigner(0)
    r.set_reference(b.ref)
    st = r.expectations(b, p)
    hi, lo = np.zeros(106, np.int64), np.zeros(106, np.int64)
    model, op = oracle.Model(), oracle.make_params(expansion=10, split_side=300)
    for i in range(b.n):
        hi, lo, _ = oracle.expectations_fixed(model, b.ref[b.ref_start[i]:b.ref_end[i]], b.read(i), b.ops(i), op, hi, lo)
    assert st == FixedStats(hi, lo)                                      # exact integers, log-likelihood included
    small = Realigner(0, max_bases_per_call=6000)
    small.set_reference(b.ref)
    assert small.expectations(b, p) == st                                # any chunking gives the same integers
    o1, f1, p1 = r.realign(b, p, want_posteriors=True)
    o2, f2, p2 = small.realign(b, p, want_posteriors=True)
    assert np.array_equal(o1, o2) and np.array_equal(f1, f2) and all(np.array_equal(p1[k], p2[k]) for k in p1)
    r.close(); small.close()


@pytest.mark.parametrize("band,lengths", [
    (100, [5000] * 3),                                                   # config 3 shape: 5 kb reads, band 100
    (50, synth.pareto_lengths(10, seed=4, lo=500, hi=20000).tolist()),   # config 5 shape: Pareto mixed lengths
])
def test_other_baseline_configs_at_small_scale(band, lengths):
    b = synth.make_batch(len(lengths), 0, 30000, seed=band, lengths=lengths, sub=0.10, ins=0.06, dele=0.09)
    ctx = capi.PhmmContext(0)
    ctx.set_reference(b.ref)
    ops, off, _ = ctx.realign_batch(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off, capi.default_params(band=band))
    model, op = oracle.Model(), oracle.make_params(expansion=band)
    for i in range(b.n):
        r = oracle.realign(model, b.ref[b.ref_start[i]:b.ref_end[i]], b.read(i), b.ops(i), op)
        assert np.array_equal(r["ops"], ops[off[i]:off[i + 1]]), i
    ctx.close()


def test_alignment_uncertainty_on_gpu_equals_checker(tmp_path):
    outs = []
    for tag, factory in (("gpu", None), ("cpu", oracle_realigner_factory())):
        ref_fa, fq, sam_path, _ = make_experiment(str(tmp_path / tag), n_reads=6, seed=8, hits=(1,))
        outdir = str(tmp_path / tag / "analysis")
        os.makedirs(outdir)
        prev = realign.setRealignerFactory(factory) if factory else None
        try:
            assert Stack(posteriors.AlignmentUncertainty(fq, "2D", ref_fa, sam_path, outdir)).startJobTree(None) == 0
        finally:
            if factory:
                realign.setRealignerFactory(prev)
        outs.append(os.path.join(outdir, "alignmentUncertainty.xml"))
    assert filecmp.cmp(outs[0], outs[1], shallow=False)


def test_base_expectation_tables_on_gpu_equal_the_host_sum_over_returned_pairs():
    """k_base_expect (phmm_batch_add_base_expectations): the device scatter-add of posterior mass by reference position
    and read base equals, as integers, the host sum over the pairs phmm_posteriors returns -- with read masks, with N
    read bases, and when the batch is cut into several library calls."""
    b = synth.make_batch(40, 600, 3000, seed=77)
    b.reads[::53] = 4                                                   # N read bases go to the fifth column
    p = posteriors.posteriorParams()
    r = Realigner()
    r.set_reference(b.ref)
    rng = np.random.default_rng(1)
    masks = [np.ones(b.n, np.uint8)] + [(rng.random(b.n) < f).astype(np.uint8) for f in (0.5, 0.1)]
    t = r.base_expectations(b, p, masks=masks)
    _, _, post = r.realign(b, p, want_posteriors=True)
    for k, m in enumerate(masks):
        want = np.zeros((len(b.ref), 5), dtype=np.int64)
        for i in np.nonzero(m)[0]:
            s = slice(post["off"][i], post["off"][i + 1])
            np.add.at(want, (b.ref_start[i] + post["ref_pos"][s].astype(np.int64), np.minimum(b.read(i)[post["read_pos"][s]], 4)),
                      post["prob_1e7"][s].astype(np.int64))
        assert np.array_equal(t[k], want), k
    assert t[0][:, 4].sum() > 0 and t[0][:, :4].sum() > t[1][:, :4].sum() > t[2][:, :4].sum() > 0
    r.max_cells = int(r.cells // 5)                                     # ~5 library calls
    assert np.array_equal(r.base_expectations(b, p, masks=masks), t)
    r.close()


def test_margin_align_snp_caller_on_gpu_equals_checker(tmp_path):
    from nanopore_b200.analyses import marginAlignSnpCaller as masc
    from test_host_snpcaller import make_snp_experiment

    class Small(masc.MarginAlignSnpCaller):
        hmmTypes = ("cactus", "trained_40")
        coverages = (1000000, 2, 1)
        seed = 3
    outs = []
    for tag, factory in (("gpu", None), ("cpu", oracle_realigner_factory())):
        ref_fa, fq, chained, _, _, _ = make_snp_experiment(str(tmp_path / tag), seed=19)
        outdir = str(tmp_path / tag / "analysis")
        os.makedirs(outdir)
        prev = realign.setRealignerFactory(factory) if factory else None
        try:
            assert Stack(Small(fq, "2D", ref_fa, chained, outdir)).startJobTree(None) == 0
        finally:
            if factory:
                realign.setRealignerFactory(prev)
        outs.append(os.path.join(outdir, "marginaliseConsensus.xml"))
    assert filecmp.cmp(outs[0], outs[1], shallow=False)


def test_base_expectation_api_error_paths():
    """Call-order and argument errors of phmm_base_expectations_* come back as error codes with text, not as crashes."""
    b = synth.make_batch(3, 200, 800, seed=5)
    p = posteriors.posteriorParams()
    ctx = capi.PhmmContext(0)
    with pytest.raises(capi.PhmmError):
        ctx.base_expectations_reset()                                   # no reference yet
    ctx.set_reference(b.ref)
    with pytest.raises(capi.PhmmError):
        ctx.base_expectations_fetch(len(b.ref))                         # nothing accumulated
    ctx.base_expectations_reset(2)
    with pytest.raises(capi.PhmmError):
        ctx.add_base_expectations()                                     # no batch prepared
    ctx.prepare(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off, p)
    with pytest.raises(capi.PhmmError):
        ctx.add_base_expectations(table=2)                              # only tables 0 and 1 exist
    with pytest.raises(ValueError):
        ctx.add_base_expectations(read_mask=np.ones(2, np.uint8))       # one byte per read
    ctx.add_base_expectations(table=1)                                  # runs the batch itself when it has not run yet
    t0, t1 = ctx.base_expectations_fetch(len(b.ref), 0), ctx.base_expectations_fetch(len(b.ref), 1)
    assert t0.sum() == 0 and t1.sum() > 0
    with pytest.raises(capi.PhmmError):
        ctx.base_expectations_fetch(len(b.ref) - 1, 1)                  # wrong size
    ctx.set_reference(b.ref)                                            # a new reference drops the tables
    with pytest.raises(capi.PhmmError):
        ctx.base_expectations_fetch(len(b.ref), 1)
    ctx.close()
