"""Host side of the realignment path: the functions of reference nanopore/analyses/utils.py that sit on it, with
the same names, argument meaning and error behaviour, re-based on one batched GPU call.

reference                                              here
-----------------------------------------------------  ---------------------------------------------------
utils.py:168-180  getExonerateCigarFormatString        same text line (kept for the cigar wire format)
utils.py:233-245  getFastaDictionary/getFastqDictionary same
utils.py:287-293  samIterator                          same
utils.py:295-386  mergeChainedAlignedReads             same global record (pos 0, leading/trailing D and I)
utils.py:388-426  chainFn                              same chain (same tie-breaking), O(n^2) over hits
utils.py:441-469  chainSamFile                         same
utils.py:540-555  realignSamFileTargetFn               chain -> [EM child] -> follow-on realignSamFile2TargetFn
utils.py:557-574  realignSamFile2TargetFn              packs ALL mapped records into one Batch and calls the
                                                       library once (the reference adds one job per read)
utils.py:576-589  realignCigarTargetFn                 one cigar through the same library call (batch of one)
utils.py:591-609  realignSamFile3TargetFn              fan-in: aR.cigar = new ops, input order, header copied

The knobs the reference hard-codes on the cactus_realign command line (utils.py:587:
--diagonalExpansion=10 --splitMatrixBiggerThanThis=3000) are module constants here.  There is no CPU
fallback: without the CUDA library and a B200 these functions raise.
"""
import os

import numpy as np

from . import capi, io_native
from .batch import Batch, encode, pack_ops, unpack_ops
from .bioio import (PairwiseAlignment, cigarReadFromString, fastaRead, fastqRead, logger, reverseComplement)
from .engine import Realigner
from .hmm import Hmm
from .sam import AlignedRead, Samfile

# File-level steps (chain, pack, fan-in) go through libphmm_io.so (C++, threaded) when it is built; the Python
# functions below restate the same reference code, serve custom chain functions, and are what the tests compare the
# native path with, byte for byte.  NANOPORE_B200_PYTHON_IO=1 forces them.
def useNativeIo():
    return io_native.available() and os.environ.get("NANOPORE_B200_PYTHON_IO", "0") != "1"


REALIGN_DIAGONAL_EXPANSION = 10          # utils.py:587
REALIGN_SPLIT_MATRIX_BIGGER_THAN = 3000  # utils.py:587


# ---------------------------------------------------------------------------------------------------------
# sequence dictionaries, SAM iteration, cigar text
# ---------------------------------------------------------------------------------------------------------
def getFastaDictionary(fastaFile):
    """First word of each fasta header -> sequence (utils.py:233-238)."""
    pairs = [(name.split()[0], seq) for name, seq in fastaRead(fastaFile)]
    assert len(pairs) == len(set(n for n, _ in pairs))      # names are unique
    return dict(pairs)


def getFastqDictionary(fastqFile):
    """First word of each fastq header -> sequence (utils.py:240-245)."""
    pairs = [(name.split()[0], seq) for name, seq, _ in fastqRead(fastqFile)]
    assert len(pairs) == len(set(n for n, _ in pairs))
    return dict(pairs)


def samIterator(sam):
    """Aligned reads of a SAM file, skipping reads with no reference alignment (utils.py:287-293)."""
    for aR in sam:
        if aR.rname != -1:
            yield aR


def getExonerateCigarFormatString(alignedRead, sam):
    """Complete exonerate-like cigar line for the SAM record (utils.py:168-180); clipping ops are dropped."""
    for op, length in alignedRead.cigar:
        assert op in (0, 1, 2, 4, 5)
    translation = {0: "M", 1: "I", 2: "D"}
    cigarString = " ".join("%s %i" % (translation[op], length) for op, length in alignedRead.cigar if op in translation)
    completeCigarString = "cigar: %s %i %i + %s %i %i + 1 %s" % (
        alignedRead.qname, 0, alignedRead.qend - alignedRead.qstart,
        sam.getrname(alignedRead.rname), alignedRead.pos, alignedRead.aend, cigarString)
    pA = cigarReadFromString(completeCigarString)            # checks it is an okay cigar
    assert sum(op.length for op in pA.operationList if op.type == PairwiseAlignment.PAIRWISE_MATCH) == \
        sum(length for op, length in alignedRead.cigar if op == 0)
    return completeCigarString


def getAbsoluteReadOffset(alignedRead, refSeq, readSeq):
    """Signed coordinate in the original read of the first non-clipped base (utils.py:155-166); for a reverse
    strand record coordinates run from -(len-1) up to 0."""
    readOffset = alignedRead.cigar[0][1] if alignedRead.cigar[0][0] == 5 else 0
    if alignedRead.is_reverse:
        readOffset = -(len(readSeq) - 1 - readOffset)
    return readOffset + alignedRead.qstart


# ---------------------------------------------------------------------------------------------------------
# chaining (the step immediately before the path; defines its input shape)
# ---------------------------------------------------------------------------------------------------------
def _readOffsetFromArrays(aR, codes, lens, readSeq):
    """getAbsoluteReadOffset on the array form of the cigar."""
    readOffset = int(lens[0]) if len(codes) and codes[0] == 5 else 0
    if aR.is_reverse:
        readOffset = -(len(readSeq) - 1 - readOffset)
    nc = codes != 5
    c2, l2 = codes[nc], lens[nc]
    return readOffset + (int(l2[0]) if len(c2) and c2[0] == 4 else 0)


def _hit_summary(aR, refSeq, readSeq):
    """(aligned pair count, refPos first, signed readPos first, refPos last, signed readPos last): what the
    reference derives by materialising AlignedPair.iterator (utils.py:388-396).  Works on the array form of the
    cigar (prefix sums instead of a loop over ops)."""
    codes, lens = aR.cigar_arrays()
    off = _readOffsetFromArrays(aR, codes, lens, readSeq)
    m = codes == 0
    if not m.any():
        raise RuntimeError("alignment of %s has no aligned positions" % aR.qname)
    qadv = np.where((codes == 0) | (codes == 1), lens, 0)
    radv = np.where((codes == 0) | (codes == 2), lens, 0)
    q0 = np.cumsum(qadv) - qadv                      # query / reference position at the start of each op
    r0 = aR.pos + np.cumsum(radv) - radv
    i0, i1 = np.flatnonzero(m)[[0, -1]]
    return (int(lens[m].sum()), int(r0[i0]), off + int(q0[i0]), int(r0[i1] + lens[i1] - 1), off + int(q0[i1] + lens[i1] - 1))


def chainFn(alignedReads, refSeq, readSeq, scoreFn=None, maxGap=200):
    """Highest scoring chain of same-strand local alignments; score = aligned positions (utils.py:388-426)."""
    info = {id(aR): _hit_summary(aR, refSeq, readSeq) for aR in alignedReads}
    score = {id(aR): (scoreFn(aR, refSeq, readSeq) if scoreFn else info[id(aR)][0]) for aR in alignedReads}
    pointers = {}
    alignedReads = sorted(alignedReads, key=lambda aR: info[id(aR)][1])          # by reference coordinate (stable)
    for i, aR in enumerate(alignedReads):
        _, rStart, qStart, rEnd, qEnd = info[id(aR)]
        own = score[id(aR)]
        for j in range(i):
            aR2 = alignedReads[j]
            _, rStart2, qStart2, rEnd2, qEnd2 = info[id(aR2)]
            assert rStart2 <= rStart
            if rStart > rEnd2 and qStart > qEnd2 and aR.is_reverse == aR2.is_reverse and \
                    rStart - rEnd2 + qStart - qEnd2 <= maxGap and own + score[id(aR2)] > score[id(aR)]:
                score[id(aR)] = own + score[id(aR2)]
                pointers[id(aR)] = aR2
    aR = sorted(alignedReads, key=lambda a: score[id(a)])[-1]
    chain = [aR]
    while id(aR) in pointers:
        aR = pointers[id(aR)]
        chain.append(aR)
    chain.reverse()
    return chain


def mergeChainedAlignedReads(chainedAlignedReads, refSequence, readSequence):
    """One global alignment for the chain (utils.py:295-386): pos = 0, seq = the read (reverse complemented for a
    reverse-strand chain), leading/trailing D and I so that the cigar spans the whole reference and read.  The ops
    are assembled as arrays; `cAR.cigar` materialises the tuple form on demand."""
    cAR = AlignedRead()
    first = chainedAlignedReads[0]
    cAR.qname = first.qname
    cAR.rnext = -1
    cAR.pos = 0
    cAR.is_reverse = first.is_reverse
    cAR.seq = reverseComplement(readSequence) if cAR.is_reverse else readSequence
    cAR.rname = first.rname
    code_parts, len_parts = [], []

    def gap(code, length):
        code_parts.append(np.array([code], dtype=np.uint8))
        len_parts.append(np.array([length], dtype=np.int64))

    pPos = 0                                                 # reference positions covered so far
    pQPos = -(len(readSequence) - 1) if cAR.is_reverse else 0   # signed read position reached so far
    for aR in chainedAlignedReads:
        assert cAR.is_reverse == aR.is_reverse
        assert aR.pos >= pPos
        if aR.pos > pPos:                                    # unaligned reference positions before this hit
            gap(2, aR.pos - pPos)
            pPos = aR.pos
        codes, lens = aR.cigar_arrays()
        assert np.isin(codes, (0, 1, 2, 4, 5)).all()
        qPos = _readOffsetFromArrays(aR, codes, lens, readSequence)
        assert qPos >= pQPos
        if qPos > pQPos:                                     # unaligned read positions before this hit
            gap(1, qPos - pQPos)
            pQPos = qPos
        keep = codes <= 2                                    # the hit's own ops, clipping filtered
        code_parts.append(codes[keep])
        len_parts.append(lens[keep])
        pPos += int(lens[(codes == 0) | (codes == 2)].sum())
        pQPos += int(lens[(codes == 0) | (codes == 1)].sum())
    assert pPos <= len(refSequence)
    if pPos < len(refSequence):
        gap(2, len(refSequence) - pPos)
    if cAR.is_reverse:
        assert pQPos <= 1
        if pQPos < 1:
            gap(1, -pQPos + 1)
    else:
        assert pQPos <= len(readSequence)
        if pQPos < len(readSequence):
            gap(1, len(readSequence) - pQPos)
    codes, lens = np.concatenate(code_parts), np.concatenate(len_parts)
    assert int(lens[(codes == 0) | (codes == 2)].sum()) == len(refSequence)
    assert int(lens[(codes == 0) | (codes == 1)].sum()) == len(readSequence)
    cAR.set_cigar_arrays(codes, lens)
    return cAR


def chainSamFile(samFile, outputSamFile, readFastqFile, referenceFastaFile, chainFn=chainFn):
    """Each (read, reference) pair is covered by a single maximal global alignment (utils.py:441-469)."""
    if chainFn is globals()["chainFn"] and useNativeIo():
        io = io_native.NativeIo()
        try:
            io.load_reference(referenceFastaFile)
            io.load_reads(readFastqFile)
            io.chain_sam(samFile)
            io.write_sam(outputSamFile)
        finally:
            io.close()
        return
    sam = Samfile(samFile, "r")
    refSequences = getFastaDictionary(referenceFastaFile)
    readSequences = getFastqDictionary(readFastqFile)
    readsToAlignedReads = {}
    for aR in samIterator(sam):
        if aR.qname not in readSequences:
            raise RuntimeError("Aligned read name: %s not in read sequences names: %s" % (aR.qname, list(readSequences.keys())[:10]))
        readsToAlignedReads.setdefault((aR.qname, aR.rname), []).append(aR)
    outputSam = Samfile(outputSamFile, "wh", template=sam)
    chained = []
    for (readName, refID), alignedReads in readsToAlignedReads.items():
        refSeq = refSequences[sam.getrname(refID)]
        readSeq = readSequences[readName]
        chained.append(mergeChainedAlignedReads(chainFn(alignedReads, refSeq, readSeq), refSeq, readSeq))
    chained.sort()           # by reference id, then name: deterministic (pysam's order here is a struct compare)
    for cAR in chained:
        outputSam.write(cAR)
    sam.close()
    outputSam.close()


# ---------------------------------------------------------------------------------------------------------
# packing SAM records for the library
# ---------------------------------------------------------------------------------------------------------
class PackedReference:
    """All reference contigs concatenated into one code array (phmm_set_reference takes one array;
    ref_start/ref_end index into it)."""

    def __init__(self, refSequences):
        self.names = list(refSequences.keys())
        self.offset, parts, o = {}, [], 0
        for n in self.names:
            self.offset[n] = o
            c = encode(refSequences[n])
            parts.append(c)
            o += len(c)
        self.length = {n: len(refSequences[n]) for n in self.names}
        self.codes = np.concatenate(parts) if parts else np.zeros(0, dtype=np.uint8)


def packAlignedReads(alignedReads, sam, packedRef):
    """SAM records -> Batch.  X = reference[pos, aend) of the record's contig, Y = aR.query, guide = the M/I/D ops
    (exactly what getExonerateCigarFormatString puts on the wire, utils.py:168-180,570)."""
    reads, ops, rs, re_, names = [], [], [], [], []
    for aR in alignedReads:
        rn = sam.getrname(aR.rname)
        if rn not in packedRef.offset:
            raise RuntimeError("Reference sequence %s of read %s not in the reference fasta" % (rn, aR.qname))
        codes, lens = aR.cigar_arrays()                      # no per-op Python work
        assert np.isin(codes, (0, 1, 2, 4, 5)).all()
        keep = codes <= 2
        o = pack_ops(np.stack([codes[keep].astype(np.int64), lens[keep]], axis=1))
        # aR.query / aR.aend from the arrays: soft clips bound the query, M and D consume the reference
        seq = aR.seq or ""
        nclip = codes != 5
        c2, l2 = codes[nclip], lens[nclip]
        qstart = int(l2[0]) if len(c2) and c2[0] == 4 else 0
        qend = len(seq) - (int(l2[-1]) if len(c2) > 1 and c2[-1] == 4 else 0)
        aend = aR.pos + int(lens[(codes == 0) | (codes == 2)].sum())
        if aend > packedRef.length[rn]:
            raise RuntimeError("Alignment of %s runs past the end of %s" % (aR.qname, rn))
        reads.append(encode(seq[qstart:qend]))
        ops.append(o)
        rs.append(packedRef.offset[rn] + aR.pos)
        re_.append(packedRef.offset[rn] + aend)
        names.append(aR.qname)
    read_off = np.concatenate(([0], np.cumsum([len(r) for r in reads]))).astype(np.int64)
    in_off = np.concatenate(([0], np.cumsum([len(o) for o in ops]))).astype(np.int64)
    return Batch(packedRef.codes,
                 np.concatenate(reads) if reads else np.zeros(0, np.uint8), read_off, rs, re_,
                 np.concatenate(ops) if ops else np.zeros(0, np.uint32), in_off, names)


def loadHmmOrNone(hmmFile):
    return Hmm.loadHmm(hmmFile) if hmmFile is not None else None


def realignParams(gapGamma, matchGamma, band=None, split=None):
    """The knobs of the cactus_realign command line (utils.py:587); the module constants are read at call time so that
    a driver (bench.py: band 50 per BASELINE.json) can set them the way the reference edits its literal string."""
    band = REALIGN_DIAGONAL_EXPANSION if band is None else band
    split = REALIGN_SPLIT_MATRIX_BIGGER_THAN if split is None else split
    return capi.default_params(band=band, split_side=split, gap_gamma=float(gapGamma), match_gamma=float(matchGamma))


# the Realigner the target functions use; replaced by parallel.ShardedRealigner under torchrun, by a checker-backed
# object in the CPU tests of the host logic
_realigner_factory = Realigner


def setRealignerFactory(factory):
    """factory(hmm=None) -> object with set_reference / realign / expectations / set_hmm / close."""
    global _realigner_factory
    prev = _realigner_factory
    _realigner_factory = factory
    return prev


def makeRealigner(hmm=None):
    return _realigner_factory(hmm=hmm)


# ---------------------------------------------------------------------------------------------------------
# the realignment targets
# ---------------------------------------------------------------------------------------------------------
def realignSamFileTargetFn(target, samFile, outputSamFile, readFastqFile, referenceFastaFile, gapGamma, matchGamma,
                           hmmFile=None, trainHmmFile=False, chainFn=chainFn):
    """Chains, optionally trains the HMM by EM, then realigns the global alignments (utils.py:540-555)."""
    tempSamFile = os.path.join(target.getGlobalTempDir(), "temp.sam")
    chainSamFile(samFile, tempSamFile, readFastqFile, referenceFastaFile, chainFn)
    if hmmFile is not None and trainHmmFile:
        from .em import learnModelFromSamFileTargetFn
        target.addChildTargetFn(learnModelFromSamFileTargetFn, args=(tempSamFile, readFastqFile, referenceFastaFile, hmmFile))
    else:
        assert not trainHmmFile
    target.setFollowOnTargetFn(realignSamFile2TargetFn, args=(tempSamFile, outputSamFile, readFastqFile, referenceFastaFile,
                                                              hmmFile, gapGamma, matchGamma))


def realignSamFile2TargetFn(target, samFile, outputSamFile, readFastqFile, referenceFastaFile, hmmFile, gapGamma, matchGamma):
    """Realigns every mapped record of samFile in one batched call and hands the ops to the fan-in
    (utils.py:557-574; the per-read child jobs and their temp cigar files are gone)."""
    if useNativeIo():
        io = io_native.NativeIo()
        try:
            io.load_reference(referenceFastaFile)
            io.load_sam(samFile)
            batch = io.batch()
            realigner = makeRealigner(hmm=loadHmmOrNone(hmmFile))
            try:
                realigner.set_reference(batch.ref)
                ops, off, _ = realigner.realign(batch, realignParams(gapGamma, matchGamma))
            finally:
                realigner.close()
            assert len(off) == batch.n + 1                 # exactly one cigar per read (utils.py:588-589)
            target.logToMaster("Realigned %d reads (%d DP cells) from %s" % (batch.n, getattr(realigner, "cells", 0), samFile))
            io.write_realigned_sam(outputSamFile, ops, off)   # the fan-in of realignSamFile3TargetFn (utils.py:591-609)
        finally:
            io.close()
        return
    refSequences = getFastaDictionary(referenceFastaFile)
    sam = Samfile(samFile, "r")
    records = list(samIterator(sam))
    packedRef = PackedReference(refSequences)
    batch = packAlignedReads(records, sam, packedRef)
    sam.close()
    realigner = makeRealigner(hmm=loadHmmOrNone(hmmFile))
    try:
        realigner.set_reference(packedRef.codes)
        ops, off, _ = realigner.realign(batch, realignParams(gapGamma, matchGamma))
    finally:
        realigner.close()
    assert len(off) == len(records) + 1                    # exactly one cigar per read (utils.py:588-589)
    cigars = [ops[off[i]:off[i + 1]] for i in range(len(records))]          # packed uint32, see realignSamFile3TargetFn
    target.logToMaster("Realigned %d reads (%d DP cells) from %s" % (len(records), getattr(realigner, "cells", 0), samFile))
    realignSamFile3TargetFn(target, samFile, outputSamFile, cigars)


def realignCigarTargetFn(target, exonerateCigarString, referenceSequenceName, referenceSequence, querySequenceName,
                         querySequence, outputCigarFile, hmmFile, gapGamma, matchGamma):
    """One cigar in, one cigar file out (utils.py:576-589), for callers that still drive the path read by read."""
    from .bioio import AlignmentOperation, cigarWrite
    pA = cigarReadFromString(exonerateCigarString)
    X = encode(referenceSequence)
    batch = Batch(X, encode(querySequence[pA.start2:pA.end2]), [0, pA.end2 - pA.start2], [pA.start1], [pA.end1],
                  pack_ops([(op.type, op.length) for op in pA.operationList]),
                  [0, len(pack_ops([(op.type, op.length) for op in pA.operationList]))])
    realigner = makeRealigner(hmm=loadHmmOrNone(hmmFile))
    try:
        realigner.set_reference(X)
        ops, off, _ = realigner.realign(batch, realignParams(gapGamma, matchGamma))
    finally:
        realigner.close()
    pA.operationList = [AlignmentOperation(c, ln) for c, ln in unpack_ops(ops)]
    with open(outputCigarFile, "w") as fh:
        cigarWrite(fh, pA)


def realignSamFile3TargetFn(target, samFile, outputSamFile, cigars):
    """Fan-in (utils.py:591-609): replaces each mapped record's cigar, input order, header copied.  `cigars` holds,
    per mapped record, either a sequence of (op, length) tuples or a packed uint32 array (length << 2 | op) as the
    library returns it (the reference reads the ops back from one temp cigar file per read)."""
    sam = Samfile(samFile, "r")
    outputSam = Samfile(outputSamFile, "wh", template=sam)
    n = 0
    for aR, cigar in zip(samIterator(sam), cigars):
        if isinstance(cigar, np.ndarray) and cigar.dtype == np.uint32:
            aR.set_cigar_arrays(cigar & 3, cigar >> 2)
        else:
            aR.cigar = tuple((int(op), int(length)) for op, length in cigar)
        outputSam.write(aR)
        n += 1
    assert n == len(cigars)
    sam.close()
    outputSam.close()
