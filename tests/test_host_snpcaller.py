"""MarginAlignSnpCaller on the host side: the vectorised calling arithmetic against the
reference's scalar definition, the sampling rule, and the whole analysis (CPU checker as the engine) against a
dictionary-based restatement of reference nanopore/analyses/marginAlignSnpCaller.py:82-246 written in the reference's
own loop structure."""
import os
import random
import xml.etree.ElementTree as ET

import numpy as np
import pytest

from nanopore_b200 import posteriors, realign
from nanopore_b200.analyses import marginAlignSnpCaller as masc
from nanopore_b200.bioio import fastaRead, fastaWrite
from nanopore_b200.mappers.abstractMapper import trainedModelPath
from nanopore_b200.sam import Samfile
from nanopore_b200.target import Stack

from helpers_sam import make_experiment
from oracle_ctx import oracle_realigner_factory


@pytest.fixture()
def oracle_engine():
    prev = realign.setRealignerFactory(oracle_realigner_factory())
    yield
    realign.setRealignerFactory(prev)


def test_vectorised_posteriors_equal_the_scalar_definition(tmp_path):
    rng = np.random.default_rng(5)
    err = masc.loadHmmErrorSubstitutionMatrix(trainedModelPath("blasr_hmm_20.txt", str(tmp_path)))
    for k in range(4):                                              # rows of a substitution matrix are distributions
        assert abs(sum(err[(masc.bases[k], b)] for b in masc.bases) - 1.0) < 1e-12
    flat, null = masc.getJukesCantorTypeSubstitutionMatrix(), masc.getNullSubstitutionMatrix()
    obs = rng.random((50, 4))
    obs[::7, 1] = 0.0
    obs /= obs.sum(axis=1, keepdims=True)
    ref = rng.integers(0, 4, 50)
    for e in (err, flat):
        got = masc.calcBasePosteriorProbsAtPositions(obs, ref, null, e)
        for i in range(50):
            want = masc.calcBasePosteriorProbs(dict(zip(masc.bases, obs[i])), masc.bases[ref[i]], null, e)
            assert np.allclose(got[i], [want[b] for b in masc.bases], rtol=1e-13, atol=0)
            assert abs(got[i].sum() - 1.0) < 1e-12


def test_bucket_and_sampling_rules():
    # int(round(p * 100)) with Python 2's round (half away from zero), cumulative from the top
    b = masc.SnpCalls.bucket([0.005, 0.015, 0.994, 0.995, 1.0, 0.5])
    assert b[100] == 2 and b[99] == 3 and b[50] == 4 and b[2] == 5 and b[1] == 6 and b[0] == 6
    assert masc.SnpCalls.bucket([]) == [0.0] * 101
    # reads are taken until bases // reference length reaches the coverage; the check comes before each read
    taken = masc.sampleReads([400] * 10, 1000, 2, random.Random(1))
    assert len(taken) == 5 and len(set(taken)) == 5                 # 4 reads = 1600 -> 1; the 5th makes 2000 -> 2: stop
    assert sorted(masc.sampleReads([400] * 10, 1000, 1000000, random.Random(1))) == list(range(10))


def make_snp_experiment(d, seed):
    """Chained (global-form) SAM over a reference that carries held-out substitutions (<reference>_Index.txt)."""
    ref_fa, fq, sam_path, _ = make_experiment(d, n_reads=9, read_len=300, contig_lens=(700,), seed=seed, hits=(1, 2), unmapped=1)
    chained = os.path.join(d, "chained.sam")
    realign.chainSamFile(sam_path, chained, fq, ref_fa)
    name, mutated = next(iter(fastaRead(ref_fa)))
    name = name.split()[0]
    rng = np.random.default_rng(seed)
    true = list(mutated)
    for i in rng.choice(len(true), size=25, replace=False):
        true[i] = "ACGT"[("ACGT".index(true[i]) + 1 + int(rng.integers(0, 3))) % 4]
    with open(ref_fa + "_Index.txt", "w") as fh:
        fastaWrite(fh, name, "".join(true))
        fastaWrite(fh, name + "_mutated", mutated)
    return ref_fa, fq, chained, name, mutated, "".join(true)


def reference_style_calls(samFile, refName, refSeq, trueSeq, readSequences, hmmFile, errMatrices):
    """marginAlignSnpCaller.py:82-246 for the all-reads sample, with dictionaries and per-read posterior lists."""
    from nanopore_b200.realign import PackedReference, loadHmmOrNone, makeRealigner, packAlignedReads, samIterator
    sam = Samfile(samFile, "r")
    reads = list(samIterator(sam))
    packedRef = PackedReference({refName: refSeq})
    batch = packAlignedReads(reads, sam, packedRef)
    sam.close()
    r = makeRealigner(hmm=loadHmmOrNone(hmmFile))
    r.set_reference(packedRef.codes)
    _, _, post = r.realign(batch, posteriors.posteriorParams(), want_posteriors=True)
    r.close()
    expectations, frequencies, totalAlignedPairs = {}, {}, 0
    for i, aR in enumerate(reads):
        readSeq = aR.query
        for q, x in aR.aligned_pairs:
            if q is None or x is None:
                continue
            totalAlignedPairs += 1
            frequencies.setdefault(x, dict.fromkeys("ACGT", 0.0))
            if aR.seq[q].upper() in "ACGT":
                frequencies[x][aR.seq[q].upper()] += 1
        s = slice(post["off"][i], post["off"][i + 1])
        for x, y, w in zip(post["ref_pos"][s], post["read_pos"][s], post["prob_1e7"][s]):
            expectations.setdefault(int(x), dict.fromkeys("ACGT", 0.0))
            if readSeq[int(y)].upper() in "ACGT":
                expectations[int(x)][readSeq[int(y)].upper()] += w / 1e7
    out = {}
    for tag, (err, table) in {"marginAlignMaxExpectedSnpCalls": (errMatrices[0], expectations), "marginAlignMaxLikelihoodSnpCalls": (errMatrices[1], expectations),
                              "maxFrequencySnpCalls": (errMatrices[0], frequencies), "maximumLikelihoodSnpCalls": (errMatrices[1], frequencies)}.items():
        tp, fp, notCalled = [], [], 0
        for x in range(len(refSeq)):
            if x not in table:
                notCalled += 1
                continue
            tot = sum(table[x].values())
            if tot > 0.0:
                pp = masc.calcBasePosteriorProbs({b: table[x][b] / tot for b in "ACGT"}, refSeq[x].upper(), masc.getNullSubstitutionMatrix(), err)
                for chosen in "ACGT":
                    if chosen != refSeq[x].upper():
                        (tp if trueSeq[x] != refSeq[x] and trueSeq[x] == chosen else fp).append(pp[chosen])
        out[tag] = (sorted(tp), sorted(fp), notCalled, totalAlignedPairs)
    return out


def test_analysis_equals_a_reference_style_restatement(tmp_path, oracle_engine):
    d = str(tmp_path / "exp")
    ref_fa, fq, chained, name, mutated, true = make_snp_experiment(d, seed=12)
    outdir = str(tmp_path / "analysis_MarginAlignSnpCaller")
    os.makedirs(outdir)

    class Small(masc.MarginAlignSnpCaller):
        hmmTypes = ("cactus", "trained_20")
        coverages = (1000000, 1)
        seed = 7
    assert Stack(Small(fq, "2D", ref_fa, chained, outdir)).startJobTree(None) == 0
    assert Small.isFinished(outdir)
    root = ET.parse(os.path.join(outdir, "marginaliseConsensus.xml")).getroot()
    assert root.tag == "marginAlignComparison" and len(root) == 2 * 4 * 4          # HMMs x samples x call sets
    readSequences = realign.getFastqDictionary(fq)
    flat = masc.getJukesCantorTypeSubstitutionMatrix()
    err = masc.loadHmmErrorSubstitutionMatrix(trainedModelPath("blasr_hmm_20.txt", str(tmp_path)))
    for hmmType, hmmFile in (("cactus", None), ("trained_20", trainedModelPath("blasr_hmm_20.txt", str(tmp_path)))):
        want = reference_style_calls(chained, name, mutated, true, readSequences, hmmFile, (flat, err))
        for tag, (tp, fp, notCalled, pairs) in want.items():
            el = [e for e in root if e.tag == tag + "_" + hmmType and e.attrib["coverage"] == "1000000"]
            assert len(el) == 1
            a = el[0].attrib
            assert int(a["totalNoCalls"]) == notCalled and int(a["totalAlignedPairs"]) == pairs
            assert int(a["totalHeldOut"]) == 25 and int(a["totalReferenceLength"]) == len(mutated)
            calls = masc.SnpCalls(25)
            calls.truePositives, calls.falsePositives = [np.array(tp)], [np.array(fp)]
            assert a["recallByProbability"] == " ".join(map(str, calls.getRecallByProbability()))
            assert a["precisionByProbability"] == " ".join(map(str, calls.getPrecisionByProbability()))
            assert len(tp) + len(fp) > 0
    # the low-coverage replicates sample fewer reads than the all-reads sample
    low = [e for e in root if e.attrib["coverage"] == "1" and e.tag.startswith("maxFrequencySnpCalls_cactus")]
    assert len(low) == 3 and all(int(e.attrib["totalSampledReads"]) < int(e.attrib["totalReads"]) for e in low)


def test_sharded_tables_equal_single_rank_tables(oracle_engine):
    """engine.base_expectations: masks select reads; chunking does not change the integer tables."""
    from nanopore_b200 import capi, synth
    b = synth.make_batch(6, 200, 700, seed=4)
    p = capi.default_params(band=10, split_side=100)
    r = realign.makeRealigner()
    r.set_reference(b.ref)
    masks = [np.array([1, 1, 1, 1, 1, 1], np.uint8), np.array([1, 0, 0, 1, 0, 0], np.uint8), np.zeros(6, np.uint8)]
    t = r.base_expectations(b, p, masks=masks)
    assert t.shape == (3, 700, 5) and t[2].sum() == 0 and 0 < t[1].sum() < t[0].sum()
    _, _, post = r.realign(b, p, want_posteriors=True)
    pr = realign.PackedReference({"ref": synth.decode(b.ref)})
    assert np.array_equal(t[0][:, :4], np.rint(posteriors.baseExpectations(b, post, pr) * 1e7).astype(np.int64))
    r.max_cells = 1                                                  # one read per library call
    assert np.array_equal(r.base_expectations(b, p, masks=masks), t)
    assert np.array_equal(r.base_expectations(b, p)[0], t[0])
