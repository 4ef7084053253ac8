"""Baum-Welch driver on the host side (rows a8, a9, a12): option parsing, M-step, trials, XML schema, resume, and
the size-independent EM property that the data likelihood does not decrease.  The CPU checker stands in for the
GPU library (host logic only)."""
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from nanopore_b200 import em, realign, synth
from nanopore_b200.engine import FixedStats
from nanopore_b200.hmm import Hmm
from nanopore_b200.mappers.abstractMapper import AbstractMapper
from nanopore_b200.target import Stack, Target

from helpers_sam import make_experiment
from oracle_ctx import oracle_realigner_factory


@pytest.fixture()
def oracle_engine():
    prev = realign.setRealignerFactory(oracle_realigner_factory())
    yield
    realign.setRealignerFactory(prev)


def test_parse_realign_options():
    p = em.parseRealignOptions("--diagonalExpansion=10 --splitMatrixBiggerThanThis=300")      # utils.py:511
    assert (p.band, p.split_side, p.anchor_trim) == (10, 300, 14)
    with pytest.raises(RuntimeError):
        em.parseRealignOptions("--noSuchFlag=1")


def test_fixed_stats_are_order_independent():
    rng = np.random.default_rng(0)
    parts = [FixedStats(rng.integers(-2**40, 2**40, 106), rng.integers(0, 2**32, 106) % (1 << 20)) for _ in range(9)]
    a, b = FixedStats(), FixedStats()
    for p in parts:
        a += p
    for p in reversed(parts):
        b += p
    assert a == b and np.array_equal(a.values(), b.values())
    assert FixedStats.from_tensor_array(a.as_tensor_array()) == a


def test_m_step_normalises_rows_and_states():
    h = Hmm("fiveStateAsymmetric")
    v = np.arange(1, 107, dtype=np.float64)
    em.mStep(h, v, trainEmissions=True)
    assert np.allclose(np.array(h.transitions).reshape(5, 5).sum(1), 1.0)
    assert np.allclose(np.array(h.emissions).reshape(5, 16).sum(1), 1.0)
    assert h.likelihood == 106.0
    e0 = list(h.emissions)
    em.mStep(h, v[::-1].copy(), trainEmissions=False)
    assert h.emissions == e0


def test_em_trials_end_to_end(tmp_path, oracle_engine):
    ref_fa, fq, sam_path, truth = make_experiment(str(tmp_path), n_reads=6, read_len=300, seed=2)
    chained = str(tmp_path / "chained.sam")
    realign.chainSamFile(sam_path, chained, fq, ref_fa)
    out = str(tmp_path / "hmm.txt")
    opts = em.Options()
    opts.modelType = "fiveStateAsymmetric"
    opts.randomStart = True
    opts.trials, opts.iterations, opts.trainEmissions, opts.outputTrialHmms = 2, 4, True, True

    class T(Target):
        def run(self):
            self.addChildTargetFn(em.learnModelFromSamFileTargetFn, args=(chained, fq, ref_fa, out, opts))

    assert Stack(T()).startJobTree(None) == 0
    for f in ("hmm.txt", "hmm.txt_unnormalised", "hmm.txt.xml", "hmm.txt_unnormalised_0", "hmm.txt_unnormalised_1"):
        assert os.path.exists(str(tmp_path / f)), f
    raw, final = Hmm.loadHmm(out + "_unnormalised"), Hmm.loadHmm(out)
    assert raw.type == 1 and np.allclose(np.array(raw.transitions).reshape(5, 5).sum(1), 1.0)
    assert raw.transitions[1 * 5 + 2] == 0.0 and raw.transitions[2 * 5 + 1] == 0.0              # asymmetric: no switch
    e = np.array(final.emissions).reshape(5, 4, 4)
    assert np.allclose(e[1:], 1.0 / 16)                                                         # utils.py:626-629
    assert np.allclose(e[0].sum(1), 0.25)                                                       # utils.py:614-619, GC 0.5
    # XML schema the in-tree consumers parse (hmm.py:18-84)
    root = ET.parse(out + ".xml").getroot()
    tr = root.findall("transition")
    assert len(tr) == 25 and {"from", "to", "avg", "std"} <= set(tr[0].attrib)
    emn = root.findall("emission")
    assert len(emn) == 80 and emn[0].attrib["x"] in "ACGT" and emn[0].attrib["state"] == "0"
    trials = root.findall("hmm")
    assert len(trials) == 2
    for t in trials:
        rl = [float(v) for v in t.attrib["runningLikelihoods"].split()]
        assert len(rl) == 4
        assert all(b >= a - 1e-6 * abs(a) for a, b in zip(rl[1:], rl[2:])), rl                  # EM: non-decreasing
    best = max(float(t.attrib["likelihood"]) for t in trials)
    assert raw.likelihood == best
    # resume: an existing _unnormalised file skips training (utils.py:527)
    mt = os.path.getmtime(out + "_unnormalised")
    os.remove(out)
    assert Stack(T()).startJobTree(None) == 0
    assert os.path.getmtime(out + "_unnormalised") == mt and os.path.exists(out)


def test_do_em_through_the_mapper(tmp_path, oracle_engine, monkeypatch):
    """AbstractMapper.realignSamFile(doEm=True): train into emptyHmmFile, then realign with it (abstractMapper.py:32-33)."""
    fast = em.Options()
    fast.modelType, fast.randomStart, fast.trials, fast.iterations, fast.trainEmissions = "fiveStateAsymmetric", False, 1, 2, True
    orig = em.learnModelFromSamFileTargetFn
    monkeypatch.setattr(em, "learnModelFromSamFileTargetFn", lambda t, *a: orig(t, *a, options=fast))
    ref_fa, fq, sam_path, truth = make_experiment(str(tmp_path), n_reads=4, read_len=300, seed=6)

    class M(AbstractMapper):
        def run(self):
            self.realignSamFile(doEm=True)

    hmm_file = str(tmp_path / "hmm.txt")
    assert Stack(M(fq, "2D", ref_fa, sam_path, emptyHmmFile=hmm_file)).startJobTree(None) == 0
    assert os.path.exists(hmm_file) and os.path.exists(hmm_file + ".xml")
    from nanopore_b200.sam import Samfile
    recs = list(Samfile(sam_path, "r"))
    assert len(recs) == 4 and all(r.pos == 0 for r in recs)


def test_symmetric_model_type_stays_symmetric_and_unsupported_options_raise(tmp_path):
    """Options.modelType defaults to fiveState (the symmetric model): its M-step ties the X and Y gap states;
    updateTheBand is refused instead of being ignored."""
    rng = np.random.default_rng(3)
    h = Hmm("fiveState")
    assert h.type == 0
    values = np.concatenate((rng.random(25) + 0.1, rng.random(80) + 0.1, [-123.0]))
    em.mStep(h, values, trainEmissions=True)
    t = np.array(h.transitions).reshape(5, 5)
    e = np.array(h.emissions).reshape(5, 4, 4)
    perm = [0, 2, 1, 4, 3]
    assert np.allclose(t, t[np.ix_(perm, perm)]) and np.allclose(t.sum(axis=1), 1.0)
    assert np.allclose(e, e[perm].transpose(0, 2, 1)) and np.allclose(e.reshape(5, -1).sum(axis=1), 1.0)
    a = Hmm("fiveStateAsymmetric")
    em.mStep(a, values, trainEmissions=True)
    ta = np.array(a.transitions).reshape(5, 5)
    assert not np.allclose(ta, ta[np.ix_(perm, perm)])                 # the asymmetric model is left alone
    o = em.Options()
    o.updateTheBand = True
    with pytest.raises(NotImplementedError):
        em.expectationMaximisationTrials(None, "", str(tmp_path / "none.cig"), str(tmp_path / "m.txt"), o)
