"""Posterior match probability outputs of the realignment kernel (row a13 of SURVEY.md 8(a); "next" row f1).

The reference reaches these through extra `cactus_realign` flags, one process per read:
  --outputAllPosteriorProbs=F            lines `refPos readPos prob` for every pair with prob >= 0.01, parsed and
                                         summed into per-reference-position base expectations
                                         (reference nanopore/analyses/marginAlignSnpCaller.py:136-155)
  --rescoreOriginalAlignment --rescoreByPosteriorProbIgnoringGaps
                                         the returned cigar keeps the ORIGINAL ops and its score becomes the mean
                                         posterior match probability over that cigar's aligned pairs
                                         (nanopore/analyses/alignmentUncertainty.py:41-48)
both with `--diagonalExpansion=10 --splitMatrixBiggerThanThis=100`.  Here one batched library call returns the
posterior pairs of all reads (phmm_posteriors, units of 1e-7 as upstream quantises them); the epilogues below are
integer / gather work over those arrays.
"""
import os
import xml.etree.ElementTree as ET

import numpy as np

from . import capi
from .analyses.abstractAnalysis import AbstractAnalysis
from .batch import unpack_ops
from .realign import (PackedReference, getFastaDictionary, loadHmmOrNone, makeRealigner, packAlignedReads, samIterator)
from .sam import Samfile

POSTERIOR_DIAGONAL_EXPANSION = 10          # alignmentUncertainty.py:41, marginAlignSnpCaller.py:136
POSTERIOR_SPLIT_MATRIX_BIGGER_THAN = 100
PROB_1 = 10000000.0


def posteriorParams(band=POSTERIOR_DIAGONAL_EXPANSION, split=POSTERIOR_SPLIT_MATRIX_BIGGER_THAN):
    return capi.default_params(band=band, split_side=split)


def alignedPairsOfOps(ops):
    """(ref positions, read positions) of the M columns of a packed cigar, relative to the window / read start."""
    xs, ys, x, y = [], [], 0, 0
    for code, ln in unpack_ops(ops):
        if code == 0:
            xs.append(np.arange(x, x + ln))
            ys.append(np.arange(y, y + ln))
            x += ln
            y += ln
        elif code == 1:
            y += ln
        else:
            x += ln
    if not xs:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    return np.concatenate(xs), np.concatenate(ys)


def rescoreOriginalAlignments(batch, post):
    """Per read: (mean posterior match probability over the pairs of the ORIGINAL guide cigar, number of those pairs).
    Pairs whose posterior is below the extraction threshold contribute 0, as upstream's sparse pair list does
    (--rescoreOriginalAlignment --rescoreByPosteriorProbIgnoringGaps, alignmentUncertainty.py:41-52)."""
    scores = np.zeros(batch.n, dtype=np.float64)
    counts = np.zeros(batch.n, dtype=np.int64)
    for i in range(batch.n):
        xs, ys = alignedPairsOfOps(batch.ops(i))
        counts[i] = len(xs)
        if len(xs) == 0:
            scores[i] = float("nan")
            continue
        s = slice(post["off"][i], post["off"][i + 1])
        ly = int(batch.read_off[i + 1] - batch.read_off[i]) + 1
        key = post["ref_pos"][s].astype(np.int64) * ly + post["read_pos"][s]            # sorted: pairs come (ref, read) ordered
        want = xs * ly + ys
        j = np.searchsorted(key, want)
        j[j >= len(key)] = 0
        hit = (key[j] == want) if len(key) else np.zeros(len(want), dtype=bool)
        tot = int(post["prob_1e7"][s][j[hit]].astype(np.int64).sum()) if len(key) else 0
        scores[i] = tot / PROB_1 / len(xs)
    return scores, counts


def baseExpectations(batch, post, packedRef):
    """expectations[contig offset + refPos][base] += prob over all reads (marginAlignSnpCaller.py:149-155); read bases
    outside ACGT are skipped.  Returns float64[len(packed reference), 4]; contig c occupies rows
    packedRef.offset[c] : packedRef.offset[c] + packedRef.length[c]."""
    out = np.zeros((len(packedRef.codes), 4), dtype=np.int64)
    for i in range(batch.n):
        s = slice(post["off"][i], post["off"][i + 1])
        bases = batch.read(i)[post["read_pos"][s]]
        ok = bases < 4
        np.add.at(out, (batch.ref_start[i] + post["ref_pos"][s][ok].astype(np.int64), bases[ok]), post["prob_1e7"][s][ok].astype(np.int64))
    return out / PROB_1


def writeAllPosteriorProbs(fileHandle, post, i):
    """One read's pairs in the `--outputAllPosteriorProbs` format the reference parses with map(float, line.split())
    (marginAlignSnpCaller.py:149)."""
    s = slice(post["off"][i], post["off"][i + 1])
    for x, y, w in zip(post["ref_pos"][s], post["read_pos"][s], post["prob_1e7"][s]):
        fileHandle.write("%i %i %.7f\n" % (x, y, w / PROB_1))


def posteriorsOfSamFile(samFile, referenceFastaFile, hmmFile=None, params=None):
    """Runs the kernel over every mapped record of samFile. Returns (records, batch, packedRef, realigned ops, off, post)."""
    refSequences = getFastaDictionary(referenceFastaFile)
    sam = Samfile(samFile, "r")
    records = list(samIterator(sam))
    packedRef = PackedReference(refSequences)
    batch = packAlignedReads(records, sam, packedRef)
    sam.close()
    realigner = makeRealigner(hmm=loadHmmOrNone(hmmFile))
    try:
        realigner.set_reference(packedRef.codes)
        ops, off, post = realigner.realign(batch, params or posteriorParams(), want_posteriors=True)
    finally:
        realigner.close()
    return records, batch, packedRef, ops, off, post


class AlignmentUncertainty(AbstractAnalysis):
    """The reference's AlignmentUncertainty analysis (alignmentUncertainty.py:9-70) on the batched kernel: average
    posterior match probability of each read's existing alignment under the trained model blasr_hmm_0.txt.  Writes
    <outputDir>/alignmentUncertainty.xml with the reference's attribute names (the R histogram is not produced)."""

    def run(self):
        AbstractAnalysis.run(self)
        from .mappers.abstractMapper import trainedModelPath
        hmmFile = trainedModelPath("blasr_hmm_0.txt", self.getLocalTempDir())
        records, batch, _, _, _, post = posteriorsOfSamFile(self.samFile, self.referenceFastaFile, hmmFile)
        avg, pairs = rescoreOriginalAlignments(batch, post)
        for aR, n in zip(records, pairs):                          # alignmentUncertainty.py:52
            assert n == sum(1 for q, r in aR.aligned_pairs if q is not None and r is not None)
        avgs = [float(v) for v in avg]
        node = ET.Element("alignmentUncertainty", {
            "averagePosteriorMatchProbabilityPerRead": str(self.formatRatio(sum(avgs), len(avgs))),
            "averagePosteriorMatchProbability": str(self.formatRatio(float(sum(a * n for a, n in zip(avgs, pairs))), int(pairs.sum()))),
            "averagePosteriorMatchProbabilitesPerRead": ",".join(str(v) for v in avgs),
            "alignedPairsInCigar": ",".join(str(int(v)) for v in pairs)})
        ET.ElementTree(node).write(os.path.join(self.outputDir, "alignmentUncertainty.xml"))
        self.finish()
