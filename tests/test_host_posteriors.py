"""Posterior outputs on the host side (row a13): rescoring of the original alignment, per-position base
expectations, the `refPos readPos prob` file format, the AlignmentUncertainty analysis.  CPU checker as the engine."""
import io
import os
import xml.etree.ElementTree as ET

import numpy as np
import pytest

import oracle
from nanopore_b200 import posteriors, realign, synth
from nanopore_b200.hmm import Hmm
from nanopore_b200.target import Stack

from helpers_sam import make_experiment
from oracle_ctx import oracle_realigner_factory


@pytest.fixture()
def oracle_engine():
    prev = realign.setRealignerFactory(oracle_realigner_factory())
    yield
    realign.setRealignerFactory(prev)


def brute_rescore(b, i, r):
    d = {(int(x), int(y)): int(w) for x, y, w in zip(r["px"], r["py"], r["pw"])}
    x = y = 0
    tot, n = 0, 0
    for c, l in synth.unpack_ops(b.ops(i)):
        if c == 0:
            for k in range(l):
                tot += d.get((x + k, y + k), 0)
            n += l; x += l; y += l
        elif c == 1:
            y += l
        else:
            x += l
    return tot / 1e7 / n, n


def test_rescore_and_expectations_match_a_direct_computation():
    b = synth.make_batch(5, 300, 900, seed=17, global_form=False)
    r = oracle_realigner_factory()(None)
    r.set_reference(b.ref)
    p = posteriors.posteriorParams()
    assert (p.band, p.split_side) == (10, 100)                                   # alignmentUncertainty.py:41
    ops, off, post = r.realign(b, p, want_posteriors=True)
    avg, pairs = posteriors.rescoreOriginalAlignments(b, post)
    model, op = oracle.Model(), oracle.make_params(expansion=10, split_side=100)
    exp = np.zeros((len(b.ref), 4))
    for i in range(b.n):
        o = oracle.realign(model, b.ref[b.ref_start[i]:b.ref_end[i]], b.read(i), b.ops(i), op)
        a, n = brute_rescore(b, i, o)
        assert pairs[i] == n and abs(avg[i] - a) < 1e-12
        assert 0.0 < avg[i] <= 1.0
        for x, y, w in zip(o["px"], o["py"], o["pw"]):
            base = b.read(i)[y]
            if base < 4:
                exp[b.ref_start[i] + x, base] += w / 1e7
    pr = realign.PackedReference({"ref": synth.decode(b.ref)})
    got = posteriors.baseExpectations(b, post, pr)
    assert np.allclose(got, exp, rtol=0, atol=1e-9) and got.sum() > 0
    # file format parsed by marginAlignSnpCaller.py:149
    buf = io.StringIO()
    posteriors.writeAllPosteriorProbs(buf, post, 2)
    rows = [list(map(float, ln.split())) for ln in buf.getvalue().splitlines()]
    s = slice(post["off"][2], post["off"][3])
    assert len(rows) == s.stop - s.start
    assert [int(v[0]) for v in rows] == post["ref_pos"][s].tolist() and all(0.01 <= v[2] <= 1.0 for v in rows)


def test_alignment_uncertainty_analysis(tmp_path, oracle_engine):
    ref_fa, fq, sam_path, truth = make_experiment(str(tmp_path / "exp"), n_reads=5, seed=21, hits=(1,))
    outdir = str(tmp_path / "analysis_AlignmentUncertainty")
    os.makedirs(outdir)
    assert Stack(posteriors.AlignmentUncertainty(fq, "2D", ref_fa, sam_path, outdir)).startJobTree(None) == 0
    assert posteriors.AlignmentUncertainty.isFinished(outdir)
    node = ET.parse(os.path.join(outdir, "alignmentUncertainty.xml")).getroot()
    per_read = [float(v) for v in node.attrib["averagePosteriorMatchProbabilitesPerRead"].split(",")]
    pairs = [int(v) for v in node.attrib["alignedPairsInCigar"].split(",")]
    assert len(per_read) == len(pairs) == 5 and all(0.0 < v <= 1.0 for v in per_read)
    assert abs(float(node.attrib["averagePosteriorMatchProbabilityPerRead"]) - sum(per_read) / 5) < 1e-12
    w = sum(a * n for a, n in zip(per_read, pairs)) / sum(pairs)
    assert abs(float(node.attrib["averagePosteriorMatchProbability"]) - w) < 1e-12
