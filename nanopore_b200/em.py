"""Baum-Welch re-estimation of the 5-state pair-HMM on the GPU (rows a8, a9, a12 of SURVEY.md 8(a)).

Mirrors the EM entry points the reference drives (reference nanopore/analyses/utils.py:471-538):
`learnModelFromSamFileTargetFn` / `learnModelFromSamFileTargetFn2` are the reference's own functions; `Options`
and `expectationMaximisationTrials` stand in for `cactus.bar.cactus_expectationMaximisation` (absent upstream
module, called at utils.py:509-528) with the option names the reference sets.

Each iteration is ONE batched E-step on the GPU (phmm_expectations_batch_fixed) over all alignments -- the
upstream module fans `cactus_realign --outputExpectations` out over alignment chunks and sums expectation files.
Statistics come back as exact integers, so the trained model does not depend on how reads were sharded.

XML written to options.outputXMLModelFile follows the schema the in-tree consumers parse
(nanopore/analyses/hmm.py:18-84, nanopore/metaAnalyses/hmmMetaAnalysis.py:27-47): a root with
<transition from to avg std>, <emission state x y avg std> and one <hmm runningLikelihoods="..."> per trial.
"""
import os
import random
import xml.etree.ElementTree as ET

import numpy as np

from . import capi
from .batch import Batch, encode, pack_ops
from .bioio import cigarRead, fastaRead, fastaWrite, logger, reverseComplement
from .hmm import SYMBOL_NUMBER, Hmm, normaliseHmmByReferenceGCContent, setHmmIndelEmissionsToBeFlat
from .sam import Samfile

BASES = "ACGT"


class Options:
    """Defaults as upstream's Options(); the reference overrides the ones listed at utils.py:509-523."""

    def __init__(self):
        self.modelType = "fiveState"
        self.optionsToRealign = "--diagonalExpansion=10 --splitMatrixBiggerThanThis=300"
        self.randomStart = False
        self.trials = 3
        self.outputTrialHmms = False
        self.iterations = 10
        self.maxAlignmentLengthPerJob = 1000000      # job granularity of the upstream fan-out; unused here
        self.maxAlignmentLengthToSample = 50000000
        self.outputXMLModelFile = None
        self.trainEmissions = False
        self.tieEmissions = False
        self.useDefaultModelAsStart = False
        self.setJukesCantorStartingEmissions = None
        self.updateTheBand = False
        self.inputModel = None
        self.numberOfAlignmentsPerJob = 100
        self.seed = 0                                # new: makes randomStart reproducible


def parseRealignOptions(optionsToRealign):
    """`--diagonalExpansion=10 --splitMatrixBiggerThanThis=300` -> phmm_params (utils.py:511)."""
    kw = {}
    for tok in optionsToRealign.split():
        if not tok.startswith("--"):
            raise RuntimeError("Unrecognised realign option: %s" % tok)
        name, _, value = tok[2:].partition("=")
        if name == "diagonalExpansion":
            kw["band"] = int(value)
        elif name == "splitMatrixBiggerThanThis":
            kw["split_side"] = int(value)
        elif name == "constraintDiagonalTrim":
            kw["anchor_trim"] = int(value)
        elif name == "gapGamma":
            kw["gap_gamma"] = float(value)
        elif name == "matchGamma":
            kw["match_gamma"] = float(value)
        else:
            raise RuntimeError("Unrecognised realign option: %s" % tok)
    return capi.default_params(**kw)


def _stock_start(modelType):
    """Probabilities of the stock cPecan model (the --loadHmm-less model of utils.py:587)."""
    hmm = Hmm(modelType)
    t = np.zeros((5, 5))
    t[0, 0] = np.exp(-0.030064059121770816)
    t[1, 0] = t[2, 0] = np.exp(-1.272871422049609)
    t[3, 0] = t[4, 0] = np.exp(-5.673280173170473)
    t[0, 1] = t[0, 2] = np.exp(-4.34381910900448)
    t[1, 1] = t[2, 2] = np.exp(-0.3388262689231553)
    t[1, 2] = t[2, 1] = np.exp(-4.910694825551255)
    t[0, 3] = t[0, 4] = np.exp(-6.30810595366929)
    t[3, 3] = t[4, 4] = np.exp(-0.003442492794189331)
    hmm.transitions = t.reshape(-1).tolist()
    m = np.empty((4, 4))
    for x in range(4):
        for y in range(4):
            m[x, y] = np.exp(-2.1149196655034745 if x == y else (-3.9833860032220842 if (x ^ y) == 2 else -4.5691014376830479))
    e = [m.reshape(-1) / m.sum()] + [np.full(16, 1.0 / 16)] * 4
    hmm.emissions = np.concatenate(e).tolist()
    hmm.normalise()
    return hmm


def _random_start(modelType, rng):
    """Random transitions on the allowed arcs of the model type and random emissions, normalised."""
    hmm = Hmm(modelType)
    allowed = np.array(_stock_start(modelType).transitions).reshape(5, 5) > 0.0
    if hmm.type == 1:                                    # fiveStateAsymmetric: no shortGapX <-> shortGapY switch
        allowed[1, 2] = allowed[2, 1] = False
    t = np.where(allowed, np.array([[rng.random() for _ in range(5)] for _ in range(5)]), 0.0)
    hmm.transitions = t.reshape(-1).tolist()
    hmm.emissions = [rng.random() for _ in range(5 * SYMBOL_NUMBER ** 2)]
    hmm.normalise()
    return hmm


def _tie_emissions(hmm):
    """One probability for matches, one for mismatches (state 0); gap states flat."""
    e = np.array(hmm.emissions).reshape(5, 4, 4)
    d = np.trace(e[0]) / 4.0
    o = (e[0].sum() - np.trace(e[0])) / 12.0
    e[0] = np.where(np.eye(4, dtype=bool), d, o)
    e[1:] = 1.0 / 16
    hmm.emissions = e.reshape(-1).tolist()


def loadAlignments(sequenceFiles, alignmentsFile, maxAlignmentLengthToSample, seed=0):
    """FASTA files + exonerate cigar file -> (Batch over a packed reference, number of cigars used).
    contig1 of each cigar is the reference (X), contig2 the read (Y) (utils.py:168-180).  If the alignments are
    longer in total than maxAlignmentLengthToSample a seeded random subset is used (upstream samples too)."""
    seqs = {}
    for f in sequenceFiles:
        for name, seq in fastaRead(f):
            seqs[name.split()[0]] = seq
    cigars = list(cigarRead(alignmentsFile))
    total = sum((pA.end1 - pA.start1) + (pA.end2 - pA.start2) for pA in cigars)
    if total > maxAlignmentLengthToSample:
        random.Random(seed).shuffle(cigars)
        acc, keep = 0, []
        for pA in cigars:
            if acc >= maxAlignmentLengthToSample:
                break
            keep.append(pA)
            acc += (pA.end1 - pA.start1) + (pA.end2 - pA.start2)
        cigars = keep
    ref_names = []
    for pA in cigars:
        if pA.contig1 not in seqs or pA.contig2 not in seqs:
            raise RuntimeError("Cigar refers to a sequence that is not in the fasta files: %s %s" % (pA.contig1, pA.contig2))
        if pA.contig1 not in ref_names:
            ref_names.append(pA.contig1)
    off, parts, o = {}, [], 0
    for n in ref_names:
        off[n] = o
        parts.append(encode(seqs[n]))
        o += len(parts[-1])
    ref = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    reads, ops, rs, re_ = [], [], [], []
    for pA in cigars:
        if not (pA.strand1 and pA.strand2):
            raise RuntimeError("Only forward-strand cigars are supported (reverse reads are written reverse complemented, utils.py:478-484)")
        reads.append(encode(seqs[pA.contig2][pA.start2:pA.end2]))
        ops.append(pack_ops([(op.type, op.length) for op in pA.operationList]))
        rs.append(off[pA.contig1] + pA.start1)
        re_.append(off[pA.contig1] + pA.end1)
    read_off = np.concatenate(([0], np.cumsum([len(r) for r in reads]))).astype(np.int64)
    in_off = np.concatenate(([0], np.cumsum([len(x) for x in ops]))).astype(np.int64)
    batch = Batch(ref, np.concatenate(reads) if reads else np.zeros(0, np.uint8), read_off, rs, re_,
                  np.concatenate(ops) if ops else np.zeros(0, np.uint32), in_off, [pA.contig2 for pA in cigars])
    return batch, len(cigars)


def _tie_symmetric(hmm):
    """fiveState (type 0) is the symmetric model: the X and Y gap states share their parameters.  After normalising,
    every parameter is averaged with its mirror image under X <-> Y (states sX <-> sY, lX <-> lY; emissions transposed);
    rows stay normalised because a row and its mirror row both sum to one."""
    perm = [0, 2, 1, 4, 3]
    t = np.array(hmm.transitions, dtype=np.float64).reshape(5, 5)
    hmm.transitions = (0.5 * (t + t[np.ix_(perm, perm)])).reshape(-1).tolist()
    e = np.array(hmm.emissions, dtype=np.float64).reshape(5, 4, 4)
    hmm.emissions = (0.5 * (e + e[perm].transpose(0, 2, 1))).reshape(-1).tolist()


def mStep(hmm, values, trainEmissions=True, tieEmissions=False):
    """Expectations (float64[106]) -> next model, in place.  Model type 0 (fiveState) keeps its X / Y symmetry."""
    nxt = Hmm(hmm.type)
    nxt.transitions = [float(v) for v in values[:25]]
    nxt.emissions = [float(v) for v in values[25:105]] if trainEmissions else list(hmm.emissions)
    nxt.likelihood = float(values[105])
    keep = list(nxt.emissions)
    nxt.normalise()
    if not trainEmissions:
        nxt.emissions = keep
    elif tieEmissions:
        _tie_emissions(nxt)
    if nxt.type == 0:
        keep = list(nxt.emissions)
        _tie_symmetric(nxt)
        if not trainEmissions:
            nxt.emissions = keep
    hmm.transitions, hmm.emissions, hmm.likelihood = nxt.transitions, nxt.emissions, nxt.likelihood


def expectationMaximisation(realigner, batch, hmm, params, iterations, trainEmissions=True, tieEmissions=False,
                            onIteration=None):
    """Runs `iterations` E+M steps from `hmm` (modified in place). Returns the running likelihoods."""
    running = []
    for it in range(iterations):
        realigner.set_hmm(hmm)
        stats = realigner.expectations(batch, params)          # FixedStats, identical on every rank
        values = stats.values()
        mStep(hmm, values, trainEmissions, tieEmissions)
        running.append(hmm.likelihood)
        if onIteration is not None:
            onIteration(it, hmm)
    return running


def writeXml(xmlFile, trialHmms, runningLikelihoods):
    """avg / std over trials per parameter + the running likelihoods of each trial (schema: hmm.py:18-84)."""
    root = ET.Element("hmms")
    T = np.array([h.transitions for h in trialHmms])
    E = np.array([h.emissions for h in trialHmms])
    for f in range(5):
        for t in range(5):
            v = T[:, f * 5 + t]
            ET.SubElement(root, "transition", {"from": str(f), "to": str(t), "avg": repr(float(v.mean())), "std": repr(float(v.std()))})
    for s in range(5):
        for x in range(4):
            for y in range(4):
                v = E[:, s * 16 + x * 4 + y]
                ET.SubElement(root, "emission", {"state": str(s), "x": BASES[x], "y": BASES[y],
                                                 "avg": repr(float(v.mean())), "std": repr(float(v.std()))})
    for h, rl in zip(trialHmms, runningLikelihoods):
        ET.SubElement(root, "hmm", {"type": str(h.type), "likelihood": repr(float(h.likelihood)),
                                    "runningLikelihoods": " ".join(repr(float(v)) for v in rl)})
    ET.ElementTree(root).write(xmlFile)


def expectationMaximisationTrials(target, sequences, alignments, outputModel, options):
    """options.trials independent EM runs; the model with the highest final likelihood is written to
    outputModel (utils.py:528).  sequences: space separated FASTA paths; alignments: exonerate cigar file."""
    from .realign import makeRealigner
    if options.updateTheBand:
        raise NotImplementedError("updateTheBand (re-deriving the guide alignments after every EM iteration) is not supported")
    batch, n = loadAlignments(sequences.split(), alignments, options.maxAlignmentLengthToSample, options.seed)
    if n == 0:
        raise RuntimeError("No alignments to train on in %s" % alignments)
    params = parseRealignOptions(options.optionsToRealign)
    rng = random.Random(options.seed)
    realigner = makeRealigner(hmm=None)
    trialHmms, running = [], []
    try:
        realigner.set_reference(batch.ref)
        for trial in range(options.trials):
            if options.useDefaultModelAsStart:
                hmm = _stock_start(options.modelType)
            elif options.inputModel is not None:
                hmm = Hmm.loadHmm(options.inputModel)
            elif options.randomStart:
                hmm = _random_start(options.modelType, rng)
            else:
                hmm = _stock_start(options.modelType)
            if options.setJukesCantorStartingEmissions is not None:
                r = float(options.setJukesCantorStartingEmissions)
                hmm.emissions[:16] = [((1.0 - r) if i % 4 == i // 4 else r / 3.0) / 4.0 for i in range(16)]
            rl = expectationMaximisation(realigner, batch, hmm, params, options.iterations, options.trainEmissions,
                                         options.tieEmissions)
            if target is not None:
                target.logToMaster("EM trial %d of %d: %d alignments, final likelihood %s" % (trial + 1, options.trials, n, hmm.likelihood))
            if options.outputTrialHmms:
                hmm.write(outputModel + "_%i" % trial)
            trialHmms.append(hmm)
            running.append(rl)
    finally:
        realigner.close()
    best = max(range(len(trialHmms)), key=lambda i: trialHmms[i].likelihood)
    trialHmms[best].write(outputModel)
    if options.outputXMLModelFile is not None:
        writeXml(options.outputXMLModelFile, trialHmms, running)
    return trialHmms[best]


# ---------------------------------------------------------------------------------------------------------
# the reference's own EM driver functions
# ---------------------------------------------------------------------------------------------------------
def learnModelFromSamFileTargetFn(target, samFile, readFastqFile, referenceFastaFile, outputModel, options=None):
    """EM on the chained SAM file (utils.py:471-531).  `options` (new, optional) overrides the reference's
    hard-coded schedule of 3 random-start trials x 100 iterations."""
    from .realign import getExonerateCigarFormatString, getFastaDictionary, getFastqDictionary
    refSequences = getFastaDictionary(referenceFastaFile)
    readSequences = getFastqDictionary(readFastqFile)
    reads = os.path.join(target.getGlobalTempDir(), "temp.fa")
    with open(reads, "w") as fH:
        for name, seq in readSequences.items():
            fastaWrite(fH, name, seq)
            fastaWrite(fH, name + "_reverse", reverseComplement(seq))
    cigars = os.path.join(target.getGlobalTempDir(), "temp.cigar")
    with open(cigars, "w") as fH:
        sam = Samfile(samFile, "r")
        for aR in sam:
            # global alignments, reverse complements in reversed coordinates (utils.py:492-496)
            assert aR.pos == 0
            assert aR.qstart == 0
            assert aR.qend == len(readSequences[aR.qname])
            assert aR.aend == len(refSequences[sam.getrname(aR.rname)])
            assert len(aR.query) == len(readSequences[aR.qname])
            if aR.is_reverse:
                assert aR.query.upper() == reverseComplement(readSequences[aR.qname]).upper()
                aR.qname += "_reverse"
            else:
                assert aR.query.upper() == readSequences[aR.qname].upper()
            fH.write(getExonerateCigarFormatString(aR, sam) + "\n")
        sam.close()
    if options is None:
        options = Options()
        options.modelType = "fiveStateAsymmetric"
        options.optionsToRealign = "--diagonalExpansion=10 --splitMatrixBiggerThanThis=300"
        options.randomStart = True
        options.trials = 3
        options.outputTrialHmms = True
        options.iterations = 100
        options.maxAlignmentLengthPerJob = 700000
        options.maxAlignmentLengthToSample = 50000000
        options.trainEmissions = True
    options.outputXMLModelFile = outputModel + ".xml"
    unnormalisedOutputModel = outputModel + "_unnormalised"
    if not os.path.exists(unnormalisedOutputModel):          # resume: training is skipped if its output exists
        target.addChildTargetFn(expectationMaximisationTrials, args=(" ".join([reads, referenceFastaFile]), cigars,
                                                                     unnormalisedOutputModel, options))
    target.setFollowOnTargetFn(learnModelFromSamFileTargetFn2, args=(unnormalisedOutputModel, outputModel))


def learnModelFromSamFileTargetFn2(target, unnormalisedOutputModel, outputModel):
    """utils.py:533-538"""
    hmm = Hmm.loadHmm(unnormalisedOutputModel)
    setHmmIndelEmissionsToBeFlat(hmm)
    normaliseHmmByReferenceGCContent(hmm, 0.5)
    hmm.write(outputModel)
