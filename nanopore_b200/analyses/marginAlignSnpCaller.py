"""MarginAlignSnpCaller: the reference's SNP-calling analysis (reference nanopore/analyses/marginAlignSnpCaller.py:38-309)
on the batched kernel -- the heaviest consumer of the realignment path.

For each of 4 HMMs ("cactus" = stock model, "trained_0/20/40" = blasr_hmm_<R>.txt) and 13 samples of the reads
(coverage 1000000, and 3 replicates each of 120, 60, 30, 10; marginAlignSnpCaller.py:46-48) the reference forks one
`cactus_realign --diagonalExpansion=10 --splitMatrixBiggerThanThis=100 --outputAllPosteriorProbs=F [--loadHmm=H]` per
sampled read, parses F and sums the posterior match probabilities by reference position and read base
(:136-155); then calls bases against four (error model, observations) combinations and writes
<outputDir>/marginaliseConsensus.xml (:205-303).

Here the posteriors of a read do not depend on the sample it is drawn into, so each HMM costs ONE forward/backward
pass over all mapped reads; the 13 samples are 13 read masks over the posterior pairs resident in HBM, scatter-added
into 13 per-position tables by k_base_expect (phmm_batch_add_base_expectations) -- the pairs never leave the GPU, and
under torchrun the int64 tables of the ranks are summed with one all-reduce.  Function and attribute names, the
sampling rule, the calling arithmetic (same operation order) and the XML schema are the reference's.

Deviations, all on inputs the reference cannot process: reference bases outside ACGT are not called (the reference
raises KeyError in getProb); `random` is seeded when `seed` is set (the reference shuffles unseeded).
"""
import math
import os
import random
import xml.etree.ElementTree as ET
from itertools import product

import numpy as np

from ..bioio import prettyXml
from ..hmm import Hmm
from ..mappers.abstractMapper import trainedModelPath
from ..posteriors import alignedPairsOfOps, posteriorParams
from ..realign import PackedReference, getFastaDictionary, getFastqDictionary, loadHmmOrNone, makeRealigner, packAlignedReads, samIterator
from ..sam import Samfile
from .abstractAnalysis import AbstractAnalysis

bases = "ACGT"


def getProb(subMatrix, start, end):
    return subMatrix[(start, end)]


def calcBasePosteriorProbs(baseObservations, refBase, evolutionarySubstitionMatrix, errorSubstutionMatrix):
    """Posterior over the true base at one position (marginAlignSnpCaller.py:18-23): prior row of the evolutionary
    matrix times the multinomial likelihood of the observed read-base fractions under the error matrix."""
    logBaseProbs = []
    for missingBase in bases:
        acc = None
        for observedBase in bases:
            term = math.log(getProb(errorSubstutionMatrix, missingBase, observedBase)) * baseObservations[observedBase]
            acc = term if acc is None else acc + term
        logBaseProbs.append(math.log(getProb(evolutionarySubstitionMatrix, refBase, missingBase)) + acc)
    totalLogProb = logBaseProbs[0]
    for y in logBaseProbs[1:]:
        totalLogProb = totalLogProb + math.log(1 + math.exp(y - totalLogProb))
    return dict(zip(bases, [math.exp(lp - totalLogProb) for lp in logBaseProbs]))


def loadHmmErrorSubstitutionMatrix(hmmFile):
    """Match emissions of the HMM, each row (reference base) normalised (marginAlignSnpCaller.py:25-29)."""
    hmm = Hmm.loadHmm(hmmFile)
    m = hmm.emissions[:len(bases) ** 2]
    m = [m[i] / sum(m[4 * (i // 4):4 * (1 + i // 4)]) for i in range(len(m))]
    return dict(zip(product(bases, bases), m))


def getNullSubstitutionMatrix():
    return dict(zip(product(bases, bases), [1.0] * len(bases) ** 2))


def getJukesCantorTypeSubstitutionMatrix():
    return dict(zip(product(bases, bases), [0.8 if x[0] == x[1] else (0.2 / 3) for x in product(bases, bases)]))


def _matrix(subMatrix):
    return np.array([[subMatrix[(a, b)] for b in bases] for a in bases], dtype=np.float64)


def calcBasePosteriorProbsAtPositions(observations, refCodes, evolutionarySubstitionMatrix, errorSubstutionMatrix):
    """calcBasePosteriorProbs for many positions at once, same operation order per position.
    observations float64[n, 4] (fractions), refCodes int[n] in 0..3 -> float64[n, 4]."""
    logErr = np.log(_matrix(errorSubstutionMatrix))                   # [missing, observed]
    logEvo = np.log(_matrix(evolutionarySubstitionMatrix))            # [ref, missing]
    lbp = np.empty((len(refCodes), 4), dtype=np.float64)
    for mb in range(4):
        acc = logErr[mb, 0] * observations[:, 0]
        for ob in range(1, 4):
            acc = acc + logErr[mb, ob] * observations[:, ob]
        lbp[:, mb] = logEvo[refCodes, mb] + acc
    total = lbp[:, 0].copy()
    for mb in range(1, 4):
        total = total + np.log(1 + np.exp(lbp[:, mb] - total))
    return np.exp(lbp - total[:, None])


class SnpCalls:
    """marginAlignSnpCaller.py:162-198 with the call lists as arrays."""

    def __init__(self, totalHeldOut=0):
        self.truePositives = []              # arrays of (prob,) per batch of positions; locations kept alongside
        self.falsePositives = []
        self.truePositiveLocations = []
        self.falsePositiveLocations = []
        self.falseNegatives = []             # never filled by the reference either (the code that did is commented out)
        self.notCalled = 0
        self.totalHeldOut = totalHeldOut

    @staticmethod
    def bucket(calls):
        """Counts of calls by probability rounded to 1/100 (half away from zero, Python 2's round), made cumulative
        from the top."""
        buckets = np.zeros(101, dtype=np.float64)
        if len(calls):
            idx = np.floor(np.asarray(calls, dtype=np.float64) * 100 + 0.5).astype(np.int64)
            buckets += np.bincount(idx, minlength=101)[:101]
        return np.cumsum(buckets[::-1])[::-1].tolist()

    def _probs(self, which):
        return np.concatenate(which) if which else np.zeros(0, dtype=np.float64)

    def getPrecisionByProbability(self):
        tPs, fPs = self.bucket(self._probs(self.truePositives)), self.bucket(self._probs(self.falsePositives))
        return [float(tPs[i]) / (tPs[i] + fPs[i]) if tPs[i] + fPs[i] != 0 else 0 for i in range(len(tPs))]

    def getRecallByProbability(self):
        return [i / self.totalHeldOut if self.totalHeldOut != 0 else 0 for i in self.bucket(self._probs(self.truePositives))]

    def getTruePositiveLocations(self):
        return np.concatenate(self.truePositiveLocations).tolist() if self.truePositiveLocations else []

    def getFalsePositiveLocations(self):
        return np.concatenate(self.falsePositiveLocations).tolist() if self.falsePositiveLocations else []

    def getFalseNegativeLocations(self):
        return [x[0] for x in self.falseNegatives]


def sampleReads(readLengths, totalReferenceLength, coverage, rng):
    """The reference's sampling rule (marginAlignSnpCaller.py:89-97): shuffle, then take reads until the bases taken so
    far, integer-divided by the reference length, reach the coverage.  -> indices in sampling order."""
    order = list(range(len(readLengths)))
    rng.shuffle(order)
    taken, totalReadLength = [], 0
    for i in order:
        if totalReadLength // totalReferenceLength >= coverage:
            break
        totalReadLength += readLengths[i]
        taken.append(i)
    return taken


def loadHeldOutSnps(referenceFastaFile, refSequences):
    """(name, position) -> true base, from <reference>_Index.txt when it exists (marginAlignSnpCaller.py:62-80)."""
    snpSet = {}
    referenceAlignmentFile = referenceFastaFile + "_Index.txt"
    if os.path.exists(referenceAlignmentFile):
        seqsAndMutatedSeqs = getFastaDictionary(referenceAlignmentFile)
        count = 0
        for name in seqsAndMutatedSeqs:
            if name in refSequences:
                count += 1
                trueSeq = seqsAndMutatedSeqs[name]
                mutatedSeq = seqsAndMutatedSeqs[name + "_mutated"]
                assert mutatedSeq == refSequences[name]
                for i in range(len(trueSeq)):
                    if trueSeq[i] != mutatedSeq[i]:
                        snpSet[(name, i)] = trueSeq[i]
            else:
                assert name.split("_")[-1] == "mutated"
        assert count == len(refSequences.keys())
    return snpSet


class MarginAlignSnpCaller(AbstractAnalysis):
    """Calculates stats on snp calling."""
    hmmTypes = ("cactus", "trained_0", "trained_20", "trained_40")
    coverages = (1000000, 120, 60, 30, 10)
    seed = None                                   # set for reproducible samples (tests)

    def hmmFileOfType(self, hmmType):
        if hmmType == "cactus":
            return None
        return trainedModelPath("blasr_hmm_%s.txt" % hmmType.split("_")[1], self.getLocalTempDir())

    def run(self):
        AbstractAnalysis.run(self)
        refSequences = getFastaDictionary(self.referenceFastaFile)
        readSequences = getFastqDictionary(self.readFastqFile)
        rng = random.Random(self.seed) if self.seed is not None else random
        snpSet = loadHeldOutSnps(self.referenceFastaFile, refSequences)
        totalReferenceLength = sum(map(len, refSequences.values()))
        totalHeldOut = len(snpSet)
        totalNotHeldOut = totalReferenceLength - totalHeldOut

        sam = Samfile(self.samFile, "r")
        reads = list(samIterator(sam))
        packedRef = PackedReference(refSequences)
        for aR in reads:                                          # the chained-global form (marginAlignSnpCaller.py:129-132)
            assert aR.pos == 0
            assert aR.aend == len(refSequences[sam.getrname(aR.rname)])
        batch = packAlignedReads(reads, sam, packedRef)
        sam.close()
        readLengths = [len(readSequences[aR.qname]) for aR in reads]

        # the SAM's own aligned pairs, per read: absolute reference index and read base (marginAlignSnpCaller.py:117-124)
        pairRef, pairBase = [], []
        for i in range(batch.n):
            xs, ys = alignedPairsOfOps(batch.ops(i))
            pairRef.append(int(batch.ref_start[i]) + xs)
            pairBase.append(batch.read(i)[ys])

        # true / mutated base codes along the packed reference
        refCodes = packedRef.codes.astype(np.int64)
        trueCodes = refCodes.copy()
        for (name, i), b in snpSet.items():
            trueCodes[packedRef.offset[name] + i] = bases.find(b.upper()) if b.upper() in bases else 4

        nullSubstitionMatrix = getNullSubstitutionMatrix()
        flatSubstitutionMatrix = getJukesCantorTypeSubstitutionMatrix()
        hmmErrorSubstitutionMatrix = loadHmmErrorSubstitutionMatrix(trainedModelPath("blasr_hmm_20.txt", self.getLocalTempDir()))

        node = ET.Element("marginAlignComparison")
        params = posteriorParams()
        for hmmType in self.hmmTypes:
            samples = [(coverage, replicate, sampleReads(readLengths, totalReferenceLength, coverage, rng))
                       for coverage in self.coverages for replicate in range(3 if coverage < 1000000 else 1)]
            masks = np.zeros((len(samples), batch.n), dtype=np.uint8)
            for k, (_, _, taken) in enumerate(samples):
                masks[k, taken] = 1
            realigner = makeRealigner(hmm=loadHmmOrNone(self.hmmFileOfType(hmmType)))
            try:
                realigner.set_reference(packedRef.codes)
                tables = realigner.base_expectations(batch, params, masks=list(masks))     # ONE pass of the kernel per HMM
            finally:
                realigner.close()
            for k, (coverage, replicate, taken) in enumerate(samples):
                totalSampledReads = len(taken)
                totalReadLength = sum(readLengths[i] for i in taken)
                frequencies = np.zeros((len(refCodes), 5), dtype=np.float64)
                totalAlignedPairs = 0
                for i in taken:
                    totalAlignedPairs += len(pairRef[i])
                    np.add.at(frequencies, (pairRef[i], np.minimum(pairBase[i], 4)), 1.0)
                expectations = tables[k].astype(np.float64) / 1e7
                callSets = self.callSnps(expectations, frequencies, refCodes, trueCodes, packedRef, flatSubstitutionMatrix,
                                         hmmErrorSubstitutionMatrix, nullSubstitionMatrix, totalHeldOut)
                for snpCalls, tagName in callSets:
                    self.writeCalls(node, snpCalls, tagName, hmmType, coverage, replicate, totalAlignedPairs, totalReferenceLength,
                                    len(reads), totalReadLength, totalSampledReads, totalHeldOut, totalNotHeldOut)
        with open(os.path.join(self.outputDir, "marginaliseConsensus.xml"), "w") as f:
            f.write(prettyXml(node))
        self.finish()

    @staticmethod
    def callSnps(expectations, frequencies, refCodes, trueCodes, packedRef, flatSubstitutionMatrix, hmmErrorSubstitutionMatrix,
                 nullSubstitionMatrix, totalHeldOut):
        """The four call sets of marginAlignSnpCaller.py:199-246 from the two [reference length, 5] observation tables
        (columns A C G T other; a row with no mass at all is a position the reference has no dictionary entry for)."""
        out = []
        for errorSubstitutionMatrix, observations, tagName in (
                (flatSubstitutionMatrix, expectations, "marginAlignMaxExpectedSnpCalls"),
                (hmmErrorSubstitutionMatrix, expectations, "marginAlignMaxLikelihoodSnpCalls"),
                (flatSubstitutionMatrix, frequencies, "maxFrequencySnpCalls"),
                (hmmErrorSubstitutionMatrix, frequencies, "maximumLikelihoodSnpCalls")):
            snpCalls = SnpCalls(totalHeldOut)
            for name in packedRef.names:                          # per contig: call locations are contig coordinates
                o, n = packedRef.offset[name], packedRef.length[name]
                obs, ref, true = observations[o:o + n], refCodes[o:o + n], trueCodes[o:o + n]
                present = obs.sum(axis=1) > 0
                callable_ = ref < 4
                snpCalls.notCalled += int((~present & callable_).sum())
                total = obs[:, :4].sum(axis=1)
                sel = np.nonzero(present & callable_ & (total > 0.0))[0]
                if len(sel) == 0:
                    continue
                probs = calcBasePosteriorProbsAtPositions(obs[sel, :4] / total[sel, None], ref[sel], nullSubstitionMatrix,
                                                          errorSubstitutionMatrix)
                for chosen in range(4):
                    other = ref[sel] != chosen
                    tp = other & (true[sel] != ref[sel]) & (true[sel] == chosen)
                    fp = other & ~tp
                    snpCalls.truePositives.append(probs[tp, chosen])
                    snpCalls.truePositiveLocations.append(sel[tp])
                    snpCalls.falsePositives.append(probs[fp, chosen])
                    snpCalls.falsePositiveLocations.append(sel[fp])
            out.append((snpCalls, tagName))
        return out

    @staticmethod
    def writeCalls(node, snpCalls, tagName, hmmType, coverage, replicate, totalAlignedPairs, totalReferenceLength, totalReads,
                   totalReadLength, totalSampledReads, totalHeldOut, totalNotHeldOut):
        """One element per call set with the reference's attribute names (marginAlignSnpCaller.py:251-283)."""
        recall = snpCalls.getRecallByProbability()
        precision = snpCalls.getPrecisionByProbability()
        assert len(recall) == len(precision)
        fScore, pIndex = max((2 * recall[i] * precision[i] / (recall[i] + precision[i]) if recall[i] + precision[i] > 0 else 0.0, i)
                             for i in range(len(recall)))
        optimumProbThreshold = float(pIndex) / 100.0
        ET.SubElement(node, tagName + "_" + hmmType, {
            "coverage": str(coverage),
            "actualCoverage": str(float(totalAlignedPairs) / totalReferenceLength),
            "totalAlignedPairs": str(totalAlignedPairs),
            "totalReferenceLength": str(totalReferenceLength),
            "replicate": str(replicate),
            "totalReads": str(totalReads),
            "avgSampledReadLength": str(float(totalReadLength) / totalSampledReads) if totalSampledReads else "nan",
            "totalSampledReads": str(totalSampledReads),
            "totalHeldOut": str(totalHeldOut),
            "totalNonHeldOut": str(totalNotHeldOut),
            "recall": str(recall[pIndex]),
            "precision": str(precision[pIndex]),
            "fScore": str(fScore),
            "optimumProbThreshold": str(optimumProbThreshold),
            "totalNoCalls": str(snpCalls.notCalled),
            "recallByProbability": " ".join(map(str, recall)),
            "precisionByProbability": " ".join(map(str, precision))})
