"""The realignment plugin path end to end on the HOST side (chain -> pack -> batched call -> SAM out), with the CPU
checker standing in for the GPU library so the plumbing is testable without a device.  Mirrors the only test the
reference has for this path -- the `make test` smoke run -- and adds the assertions it lacks."""
import os

import numpy as np
import pytest

import oracle
from nanopore_b200 import realign, synth
from nanopore_b200.analyses.abstractAnalysis import AbstractAnalysis
from nanopore_b200.analyses.gpuRealign import GpuRealign, GpuRealignTrainedModel
from nanopore_b200.bioio import reverseComplement
from nanopore_b200.hmm import Hmm
from nanopore_b200.mappers.abstractMapper import AbstractMapper, trainedModelPath
from nanopore_b200.sam import Samfile
from nanopore_b200.target import Stack

from helpers_sam import make_experiment
from oracle_ctx import oracle_realigner_factory


@pytest.fixture()
def oracle_engine():
    prev = realign.setRealignerFactory(oracle_realigner_factory())
    yield
    realign.setRealignerFactory(prev)


def spans(cigar):
    return sum(l for op, l in cigar if op in (0, 2)), sum(l for op, l in cigar if op in (0, 1))


def test_chain_makes_one_global_alignment_per_read(tmp_path):
    ref_fa, fq, sam_path, truth = make_experiment(str(tmp_path), n_reads=8, seed=5)
    out = str(tmp_path / "chained.sam")
    realign.chainSamFile(sam_path, out, fq, ref_fa)
    refs, reads = realign.getFastaDictionary(ref_fa), realign.getFastqDictionary(fq)
    sam = Samfile(out, "r")
    recs = list(sam)
    assert sorted(r.qname for r in recs) == sorted(truth)               # one record per mapped read, unmapped dropped
    for r in recs:
        cname, start, reverse, L = truth[r.qname]
        assert sam.getrname(r.rname) == cname and r.pos == 0 and r.is_reverse == reverse
        assert r.seq == (reverseComplement(reads[r.qname]) if reverse else reads[r.qname])     # utils.py:330-333
        assert spans(r.cigar) == (len(refs[cname]), len(reads[r.qname]))                          # utils.py:381-382
        assert all(op in (0, 1, 2) for op, _ in r.cigar)
        assert r.cigar[0][0] == 2 and r.cigar[0][1] >= start                                      # leading D up to the first hit
    sam.close()


def test_chain_picks_the_best_same_strand_chain():
    from nanopore_b200.sam import AlignedRead, parse_cigar

    def hit(pos, cig, seq, reverse=False):
        a = AlignedRead()
        a.qname, a.rname, a.pos, a.cigar, a.seq, a.is_reverse = "r", 0, pos, parse_cigar(cig), seq, reverse
        return a
    read = "A" * 100
    ref = "A" * 1000
    h1 = hit(10, "30M70S", read)                 # read 0..29   -> ref 10..39
    h2 = hit(60, "35S40M25S", read)              # read 35..74  -> ref 60..99   (gap 20+5 <= 200: chains with h1)
    h3 = hit(700, "80S20M", read)                # read 80..99  -> ref 700..719 (gap 600 > 200: cannot chain)
    h4 = hit(300, "50M50S", read, reverse=True)  # other strand
    chain = realign.chainFn([h3, h2, h4, h1], ref, read)
    assert chain == [h1, h2]
    g = realign.mergeChainedAlignedReads(chain, ref, read)
    assert g.cigar == ((2, 10), (0, 30), (2, 20), (1, 5), (0, 40), (2, 900), (1, 25))
    assert realign.chainFn([h4, h3], ref, read) == [h4]                 # 50 aligned positions beat 20
    g = realign.mergeChainedAlignedReads([h4], ref, read)
    assert g.is_reverse and spans(g.cigar) == (1000, 100) and g.cigar[0] == (2, 300)


class FakeMapperRealign(AbstractMapper):
    """What e.g. LastRealign does (reference nanopore/mappers/last.py): produce a SAM, then realignSamFile()."""
    kw = {}

    def run(self):
        self.realignSamFile(**self.kw)


def run_mapper(cls, tmp_path, seed=7, **exp):
    ref_fa, fq, sam_path, truth = make_experiment(str(tmp_path), seed=seed, **exp)
    out = str(tmp_path / "out.sam")
    import shutil
    shutil.copyfile(sam_path, out)
    m = cls(fq, "2D", ref_fa, out, emptyHmmFile=str(tmp_path / "hmm.txt"))
    assert Stack(m).startJobTree(None) == 0
    return ref_fa, fq, out, truth


def check_against_oracle(ref_fa, fq, out_sam, chained_sam, model, gap_gamma=0.5, match_gamma=0.0):
    refs = realign.getFastaDictionary(ref_fa)
    a, b = Samfile(out_sam, "r"), Samfile(chained_sam, "r")
    assert a.header_lines == b.header_lines                                 # header copied (utils.py:596)
    ra, rb = list(a), list(b)
    assert [r.qname for r in ra] == [r.qname for r in rb]                   # input order kept (utils.py:597)
    op = oracle.make_params(expansion=10, split_side=3000, gap_gamma=gap_gamma, match_gamma=match_gamma)
    changed = 0
    for x, y in zip(ra, rb):
        assert (x.flag, x.rname, x.pos, x.seq) == (y.flag, y.rname, y.pos, y.seq)
        X = synth.encode(refs[a.getrname(x.rname)])
        r = oracle.realign(model, X, synth.encode(y.query), synth.pack_ops(list(y.cigar)), op)
        assert tuple(synth.unpack_ops(r["ops"])) == x.cigar
        assert spans(x.cigar) == spans(y.cigar)
        changed += x.cigar != y.cigar
    assert changed > 0
    a.close(); b.close()


def test_realign_sam_file_stock_model(tmp_path, oracle_engine):
    ref_fa, fq, out, truth = run_mapper(FakeMapperRealign, tmp_path)
    chained = str(tmp_path / "chained.sam")
    realign.chainSamFile(str(tmp_path / "mapping.sam"), chained, fq, ref_fa)
    check_against_oracle(ref_fa, fq, out, chained, oracle.Model())


def test_realign_sam_file_trained_model_and_gammas(tmp_path, oracle_engine, golden_dir):
    class M(FakeMapperRealign):
        kw = dict(gapGamma=0.2, matchGamma=0.3, useTrainedModel=True, trainedModelFile="blasr_hmm_20.txt")
    ref_fa, fq, out, truth = run_mapper(M, tmp_path, seed=8)
    chained = str(tmp_path / "chained.sam")
    realign.chainSamFile(str(tmp_path / "mapping.sam"), chained, fq, ref_fa)
    t, e = Hmm.loadHmm(trainedModelPath("blasr_hmm_20.txt", str(tmp_path))).arrays()
    check_against_oracle(ref_fa, fq, out, chained, oracle.Model(t, e), 0.2, 0.3)
    # the derived blasr_hmm_20 equals the file the reference ships, to the digits it prints
    g = Hmm.loadHmm(os.path.join(golden_dir, "blasr_hmm_20.txt"))
    assert np.allclose(e, g.emissions, rtol=0, atol=5e-12) and np.allclose(t, g.transitions, rtol=0, atol=5e-12)


def test_em_and_trained_together_is_an_error(tmp_path, oracle_engine):
    class M(FakeMapperRealign):
        kw = dict(doEm=True, useTrainedModel=True)
    ref_fa, fq, sam_path, _ = make_experiment(str(tmp_path), seed=1)
    m = M(fq, "2D", ref_fa, sam_path)
    assert Stack(m).startJobTree(None) == 1                               # abstractMapper.py:29-30 raises


def test_large_batches_are_chunked_with_identical_results(tmp_path):
    b = synth.make_batch(7, 300, 900, seed=12)
    from nanopore_b200 import capi
    p = capi.default_params(band=10)
    f_all, f_small = oracle_realigner_factory(), oracle_realigner_factory(max_bases_per_call=2500)
    r1, r2 = f_all(), f_small()
    r1.set_reference(b.ref); r2.set_reference(b.ref)
    o1, f1, p1 = r1.realign(b, p, want_posteriors=True)
    o2, f2, p2 = r2.realign(b, p, want_posteriors=True)
    assert len(r2.ctx.calls) > len(r1.ctx.calls) == 1
    assert np.array_equal(o1, o2) and np.array_equal(f1, f2) and r1.cells == r2.cells
    for k in p1:
        assert np.array_equal(p1[k], p2[k])
    assert r1.expectations(b, p) == r2.expectations(b, p)                 # exact integers: order independent


def test_gpu_realign_analysis_plugin(tmp_path, oracle_engine):
    ref_fa, fq, sam_path, truth = make_experiment(str(tmp_path / "exp"), seed=9)
    outdir = str(tmp_path / "analysis_GpuRealign")
    os.makedirs(outdir)
    assert not AbstractAnalysis.isFinished(outdir)
    a = GpuRealign(fq, "2D", ref_fa, sam_path, outdir)                    # the 5 positional args of pipeline.py:140
    assert Stack(a).startJobTree(None) == 0
    assert AbstractAnalysis.isFinished(outdir) and os.path.exists(os.path.join(outdir, "realigned.sam"))
    chained = str(tmp_path / "chained.sam")
    realign.chainSamFile(sam_path, chained, fq, ref_fa)
    check_against_oracle(ref_fa, fq, os.path.join(outdir, "realigned.sam"), chained, oracle.Model())
    AbstractAnalysis.reset(outdir)
    assert not AbstractAnalysis.isFinished(outdir)
    assert AbstractAnalysis.formatRatio(1, 0) != AbstractAnalysis.formatRatio(1, 0) and AbstractAnalysis.formatRatio(1, 2) == 0.5
    outdir2 = str(tmp_path / "analysis_Trained")
    os.makedirs(outdir2)
    assert Stack(GpuRealignTrainedModel(fq, "2D", ref_fa, sam_path, outdir2)).startJobTree(None) == 0
    assert AbstractAnalysis.isFinished(outdir2)


def test_realign_cigar_target_fn_single_read(tmp_path, oracle_engine):
    from nanopore_b200.bioio import cigarRead
    from nanopore_b200.target import Target
    b = synth.make_batch(1, 300, 800, seed=4)
    ops = " ".join("%s %d" % ("MID"[c], l) for c, l in synth.unpack_ops(b.ops(0)))
    line = "cigar: read_0 0 %d + ref 0 %d + 1 %s" % (len(b.read(0)), len(b.ref), ops)
    out = str(tmp_path / "o.cig")
    realign.realignCigarTargetFn(Target(), line, "ref", synth.decode(b.ref), "read_0", synth.decode(b.read(0)), out, None, 0.5, 0.0)
    pAs = list(cigarRead(out))
    assert len(pAs) == 1                                                   # utils.py:588-589
    r = oracle.realign(oracle.Model(), b.ref, b.read(0), b.ops(0), oracle.make_params(expansion=10))
    assert [(o.type, o.length) for o in pAs[0].operationList] == synth.unpack_ops(r["ops"])


def test_command_line_entry_point(tmp_path, oracle_engine):
    """scripts/realign_sam.py: SAM in, realigned SAM out (single process; under torchrun rank 0 does the same)."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("realign_sam", os.path.join(root, "scripts", "realign_sam.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    ref_fa, fq, sam_path, truth = make_experiment(str(tmp_path), seed=13)
    out = str(tmp_path / "realigned.sam")
    assert mod.main([sam_path, fq, ref_fa, out, "--gapGamma", "0.4"]) == 0
    chained = str(tmp_path / "chained.sam")
    realign.chainSamFile(sam_path, chained, fq, ref_fa)
    check_against_oracle(ref_fa, fq, out, chained, oracle.Model(), 0.4, 0.0)


def test_realign_variant_classes_of_a_base_mapper(tmp_path, oracle_engine, monkeypatch):
    """The six `*Chain` / `*Realign*` variants the reference writes out for every base mapper (last_params.py:10-38),
    derived from a stand-in base mapper: each one runs the base, then chains / realigns with its variant's arguments."""
    from nanopore_b200 import em
    from nanopore_b200.mappers import realignVariants as rv
    from nanopore_b200.sam import Samfile
    fast = em.Options()                                       # the reference's 3 x 100 schedule is for GPUs, not for the CPU checker
    fast.modelType, fast.randomStart, fast.trials, fast.iterations, fast.trainEmissions = "fiveStateAsymmetric", True, 1, 2, True
    orig = em.learnModelFromSamFileTargetFn
    monkeypatch.setattr(em, "learnModelFromSamFileTargetFn", lambda t, *a: orig(t, *a, options=fast))
    assert sorted(rv.realignVariants(rv.SamFileMapper)) == sorted("SamFileMapper" + s for s in rv.VARIANTS)
    ref_fa, fq, sam_path, _ = make_experiment(str(tmp_path / "exp"), n_reads=5, read_len=300, seed=41)
    outs = {}
    for name in ("SamFileMapperChain", "SamFileMapperRealign", "SamFileMapperRealignTrainedModel40", "SamFileMapperRealignEm"):
        out = str(tmp_path / (name + ".sam"))
        m = getattr(rv, name)(fq, "2D", ref_fa, out, emptyHmmFile=str(tmp_path / (name + "_hmm.txt")), mappedSamFile=sam_path)
        assert Stack(m).startJobTree(None) == 0
        outs[name] = [(aR.qname, aR.pos, tuple(aR.cigar)) for aR in Samfile(out, "r") if aR.rname != -1]
    chain = outs["SamFileMapperChain"]
    assert len(chain) == 5 and all(pos == 0 for _, pos, _ in chain)                   # one global alignment per read
    for name in ("SamFileMapperRealign", "SamFileMapperRealignTrainedModel40", "SamFileMapperRealignEm"):
        assert [q for q, _, _ in outs[name]] == [q for q, _, _ in chain]                   # same records, same order
        assert any(c != c0 for (_, _, c), (_, _, c0) in zip(outs[name], chain))            # new cigars
    assert outs["SamFileMapperRealignTrainedModel40"] != outs["SamFileMapperRealign"]      # the trained model was used
    assert os.path.exists(str(tmp_path / "SamFileMapperRealignEm_hmm.txt"))                # doEm trained into emptyHmmFile
