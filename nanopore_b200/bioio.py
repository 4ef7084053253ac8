"""The slice of `sonLib.bioio` the realignment path uses (sonLib is an empty submodule in the reference).

Call sites mirrored (reference nanopore/analyses/utils.py:2): `fastaRead`, `fastqRead`, `fastaWrite`,
`fastqWrite`, `cigarRead`, `cigarReadFromString`, `cigarWrite`, `PairwiseAlignment`, `reverseComplement`,
`system`, `nameValue`, `logger`.

Cigar wire format (exonerate style, written at utils.py:168-180 and read back at utils.py:588-599):
    cigar: <contig2> <start2> <end2> <strand2> <contig1> <start1> <end1> <strand1> <score> (<op> <len>)*
contig1 is the target (reference, X), contig2 the query (read, Y).  `op.type` values are the SAM op codes
(0 = M, 1 = I, 2 = D): the reference writes them straight into `aR.cigar` (utils.py:602) and encodes the same
mapping at utils.py:173, so that identity is part of the in-tree contract.
"""
import logging
import re
import subprocess

import numpy as np

logger = logging.getLogger("nanopore_b200")

_COMPLEMENT = str.maketrans("ACGTNacgtn", "TGCANtgcan")


def reverseComplement(seq):
    return seq.translate(_COMPLEMENT)[::-1]


def system(command):
    """Runs a shell command, raising on a non-zero exit (the error convention of the reference boundary)."""
    logger.debug("Running the command: %s", command)
    sts = subprocess.call(command, shell=True)
    if sts != 0:
        raise RuntimeError("Command: %s exited with non-zero status %i" % (command, sts))
    return sts


def prettyXml(elem):
    """Indented text of an ElementTree element (sonLib bioio.prettyXml, used by marginAlignSnpCaller.py:304)."""
    import xml.etree.ElementTree as ET
    from xml.dom import minidom
    return minidom.parseString(ET.tostring(elem, "utf-8")).toprettyxml(indent="  ")


def nameValue(name, value, valueType=str):
    """`--name=value`, or the empty string when value is None (utils.py:586)."""
    if valueType == bool:
        return "--%s" % name if value else ""
    if value is None:
        return ""
    return "--%s=%s" % (name, str(value))


def _open(fileHandleOrFile, mode="r"):
    if isinstance(fileHandleOrFile, str):
        return open(fileHandleOrFile, mode), True
    return fileHandleOrFile, False


def fastaRead(fileHandleOrFile):
    """Yields (header without '>', sequence)."""
    fh, own = _open(fileHandleOrFile)
    try:
        name, chunks = None, []
        for line in fh:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if name is not None:
                    yield name, "".join(chunks)
                name, chunks = line[1:], []
            elif name is not None:
                chunks.append("".join(line.split()))
        if name is not None:
            yield name, "".join(chunks)
    finally:
        if own:
            fh.close()


def fastaWrite(fileHandleOrFile, name, seq, mode="w"):
    fh, own = _open(fileHandleOrFile, mode)
    try:
        fh.write(">%s\n" % name)
        for i in range(0, len(seq), 100):
            fh.write(seq[i:i + 100] + "\n")
    finally:
        if own:
            fh.close()


def fastqRead(fileHandleOrFile):
    """Yields (name, sequence, quality values as a list of ints or None)."""
    fh, own = _open(fileHandleOrFile)
    try:
        line = fh.readline()
        while line != "":
            if line[0] == "@":
                name = line[1:].rstrip("\r\n")
                seq = fh.readline().rstrip("\r\n")
                plus = fh.readline()
                if not plus.startswith("+"):
                    raise RuntimeError("Got unexpected line: %s" % plus)
                quals = fh.readline().rstrip("\r\n")
                if len(seq) != len(quals):
                    logger.critical("Got a mismatch between the number of sequence characters (%s) and number of "
                                    "qual values (%s) for sequence: %s, ignoring returning None", len(seq), len(quals), name)
                    qv = None
                else:
                    raw = quals.encode("latin-1")
                    q8 = np.frombuffer(raw, dtype=np.uint8)
                    if len(q8) and (q8.min() < 33 or q8.max() > 126):
                        raise RuntimeError("Got a qual value out of range for sequence %s" % name)
                    qv = list(raw)
                yield name, seq, qv
            line = fh.readline()
    finally:
        if own:
            fh.close()


def fastqWrite(fileHandleOrFile, name, seq, qualValues, mode="w"):
    fh, own = _open(fileHandleOrFile, mode)
    try:
        if qualValues is not None and len(seq) != len(qualValues):
            raise RuntimeError("Got a mismatch between the number of sequence characters (%s) and number of qual "
                               "values (%s) for sequence: %s" % (len(seq), len(qualValues), name))
        q = "".join(chr(v) for v in qualValues) if qualValues is not None else "*"
        fh.write("@%s\n%s\n+\n%s\n" % (name, seq, q))
    finally:
        if own:
            fh.close()


class AlignmentOperation:
    def __init__(self, opType, length, score=0.0):
        self.type = opType
        self.length = length
        self.score = score

    def __eq__(self, o):
        return self.type == o.type and self.length == o.length

    def __repr__(self):
        return "AlignmentOperation(%r, %r)" % (self.type, self.length)


class PairwiseAlignment:
    """contig1 / X = target (reference), contig2 / Y = query (read)."""
    PAIRWISE_MATCH = 0        # 'M'  == SAM op 0
    PAIRWISE_INDEL_Y = 1      # 'I'  == SAM op 1: read-only columns
    PAIRWISE_INDEL_X = 2      # 'D'  == SAM op 2: reference-only columns
    PAIRWISE_PLUS = "+"
    PAIRWISE_MINUS = "-"

    def __init__(self, contig1, start1, end1, strand1, contig2, start2, end2, strand2, score, operationList):
        self.contig1, self.start1, self.end1, self.strand1 = contig1, start1, end1, strand1
        self.contig2, self.start2, self.end2, self.strand2 = contig2, start2, end2, strand2
        self.score = score
        self.operationList = operationList


_OP_LETTER = {PairwiseAlignment.PAIRWISE_MATCH: "M", PairwiseAlignment.PAIRWISE_INDEL_Y: "I",
              PairwiseAlignment.PAIRWISE_INDEL_X: "D"}
_LETTER_OP = {v: k for k, v in _OP_LETTER.items()}
_CIGAR_RE = re.compile(r"cigar:\s+(\S+)\s+([0-9]+)\s+([0-9]+)\s+([\+\-\.])\s+(\S+)\s+([0-9]+)\s+([0-9]+)\s+([\+\-\.])\s+(\S+)(.*)")


def cigarReadFromString(line):
    m = _CIGAR_RE.match(line.strip())
    if m is None:
        raise RuntimeError("Not a cigar line: %s" % line[:80])
    toks = m.group(10).split()
    if len(toks) % 2:
        raise RuntimeError("Odd number of cigar operation tokens: %s" % line[:80])
    ops = []
    for i in range(0, len(toks), 2):
        if toks[i] not in _LETTER_OP:
            raise RuntimeError("Unknown cigar operation %s" % toks[i])
        ops.append(AlignmentOperation(_LETTER_OP[toks[i]], int(toks[i + 1])))
    strand = lambda s: s != "-"
    return PairwiseAlignment(m.group(5), int(m.group(6)), int(m.group(7)), strand(m.group(8)),
                             m.group(1), int(m.group(2)), int(m.group(3)), strand(m.group(4)),
                             float(m.group(9)), ops)


def cigarRead(fileHandleOrFile):
    fh, own = _open(fileHandleOrFile)
    try:
        for line in fh:
            if line.startswith("cigar:"):
                yield cigarReadFromString(line)
    finally:
        if own:
            fh.close()


def cigarWrite(fileHandle, pairwiseAlignment, withProbs=False):
    pA = pairwiseAlignment
    s = lambda b: "+" if b else "-"
    fileHandle.write("cigar: %s %i %i %s %s %i %i %s %f" % (pA.contig2, pA.start2, pA.end2, s(pA.strand2),
                                                         pA.contig1, pA.start1, pA.end1, s(pA.strand1), pA.score))
    for op in pA.operationList:
        fileHandle.write(" %s %i" % (_OP_LETTER[op.type], op.length))
    fileHandle.write("\n")
