"""Host-side formats of the path: sonLib-style cigar / FASTA / FASTQ I/O, the SAM reader/writer, the exonerate
cigar line (reference nanopore/analyses/utils.py:168-180), and the jobTree Target shim's ordering."""
import io
import os

import numpy as np
import pytest

from nanopore_b200 import bioio, sam as samlib
from nanopore_b200.realign import getExonerateCigarFormatString, getFastaDictionary, getFastqDictionary, samIterator
from nanopore_b200.target import Stack, Target

from helpers_sam import make_experiment


def test_fasta_fastq_roundtrip(tmp_path):
    fa = str(tmp_path / "x.fa")
    with open(fa, "w") as fh:
        bioio.fastaWrite(fh, "a first", "ACGT" * 60)
        bioio.fastaWrite(fh, "b", "")
        bioio.fastaWrite(fh, "c", "NNNNACGTacgt")
    assert list(bioio.fastaRead(fa)) == [("a first", "ACGT" * 60), ("b", ""), ("c", "NNNNACGTacgt")]
    assert getFastaDictionary(fa) == {"a": "ACGT" * 60, "b": "", "c": "NNNNACGTacgt"}
    fq = str(tmp_path / "x.fq")
    with open(fq, "w") as fh:
        bioio.fastqWrite(fh, "r1 x", "ACGTN", [33, 40, 50, 60, 126])
        bioio.fastqWrite(fh, "r2", "GG", [50, 50])
    assert list(bioio.fastqRead(fq)) == [("r1 x", "ACGTN", [33, 40, 50, 60, 126]), ("r2", "GG", [50, 50])]
    assert getFastqDictionary(fq) == {"r1": "ACGTN", "r2": "GG"}
    with pytest.raises(RuntimeError):
        bioio.fastqWrite(io.StringIO(), "bad", "ACG", [50])
    assert bioio.reverseComplement("AACGTNacg") == "cgtNACGTT"
    assert bioio.nameValue("loadHmm", None) == "" and bioio.nameValue("loadHmm", "f") == "--loadHmm=f"
    with pytest.raises(RuntimeError):
        bioio.system("exit 3")


def test_duplicate_names_are_rejected(tmp_path):
    fa = str(tmp_path / "d.fa")
    with open(fa, "w") as fh:
        bioio.fastaWrite(fh, "a 1", "AC")
        bioio.fastaWrite(fh, "a 2", "GT")
    with pytest.raises(AssertionError):          # utils.py:237
        getFastaDictionary(fa)


def test_cigar_wire_format_roundtrip():
    line = "cigar: read_1 0 12 + ref0 5 18 + 1 M 5 D 3 M 4 I 2 M 1"
    pA = bioio.cigarReadFromString(line)
    assert (pA.contig2, pA.start2, pA.end2, pA.strand2) == ("read_1", 0, 12, True)
    assert (pA.contig1, pA.start1, pA.end1, pA.strand1) == ("ref0", 5, 18, True)
    assert [(o.type, o.length) for o in pA.operationList] == [(0, 5), (2, 3), (0, 4), (1, 2), (0, 1)]   # SAM op codes
    assert pA.score == 1.0
    buf = io.StringIO()
    bioio.cigarWrite(buf, pA)
    pB = next(bioio.cigarRead(io.StringIO(buf.getvalue())))
    assert pB.operationList == pA.operationList and pB.contig1 == "ref0" and pB.end2 == 12
    assert bioio.cigarReadFromString("cigar: q 0 0 + t 0 0 + 0").operationList == []
    with pytest.raises(RuntimeError):
        bioio.cigarReadFromString("cigar: q 0 3 + t 0 3 + 0 Z 3")
    with pytest.raises(RuntimeError):
        bioio.cigarReadFromString("not a cigar")


def test_sam_reader_writer_and_aligned_read_coordinates(tmp_path):
    ref_fa, fq, sam_path, truth = make_experiment(str(tmp_path), n_reads=5, seed=3)
    sam = samlib.Samfile(sam_path, "r")
    assert sam.references == ["ref0", "ref1"] and sam.lengths == [1500, 1100]
    recs = list(sam)
    assert any(r.rname == -1 for r in recs) and len(list(samIterator(iter(recs)))) == len(recs) - 1
    reads = getFastqDictionary(fq)
    refs = getFastaDictionary(ref_fa)
    for r in samIterator(iter(recs)):
        assert r.qend - r.qstart == len(r.query) == sum(l for op, l in r.cigar if op in (0, 1))
        assert r.aend - r.pos == sum(l for op, l in r.cigar if op in (0, 2)) == r.alen
        pairs = r.aligned_pairs
        assert sum(1 for q, t in pairs if q is not None and t is not None) == sum(l for op, l in r.cigar if op == 0)
        assert [q for q, t in pairs if q is not None] == list(range(len(r.query)))
        assert [t for q, t in pairs if t is not None] == list(range(r.pos, r.aend))
        assert r.is_reverse == truth[r.qname][2]
        line = getExonerateCigarFormatString(r, sam)
        pA = bioio.cigarReadFromString(line)
        assert (pA.contig2, pA.start2, pA.end2) == (r.qname, 0, len(r.query))
        assert (pA.contig1, pA.start1, pA.end1) == (sam.getrname(r.rname), r.pos, r.aend)
        assert [(o.type, o.length) for o in pA.operationList] == [(op, l) for op, l in r.cigar if op in (0, 1, 2)]
    # write -> read is the identity on every field
    out = str(tmp_path / "copy.sam")
    o = samlib.Samfile(out, "wh", template=sam)
    for r in recs:
        o.write(r)
    o.close()
    sam.close()
    assert open(out).read() == open(sam_path).read()
    assert samlib.parse_cigar("*") is None and samlib.format_cigar(None) == "*"
    with pytest.raises(ValueError):
        samlib.parse_cigar("5M3")
    a = samlib.AlignedRead()
    a.is_reverse = True
    assert a.flag == 16
    a.is_reverse = False
    assert a.flag == 0 and a.aend is None


def test_target_children_run_before_follow_on():
    log = []

    class T(Target):
        def run(self):
            log.append("root")
            self.addChildTargetFn(lambda t, x: (log.append("child%d" % x), t.setFollowOnFn(lambda: log.append("child%d.follow" % x))), args=(1,))
            self.addChildTargetFn(lambda t, x: log.append("child%d" % x), args=(2,))
            self.setFollowOnTargetFn(lambda t: log.append("follow"))

    assert Stack(T()).startJobTree(None) == 0
    assert log == ["root", "child1", "child1.follow", "child2", "follow"]

    class Bad(Target):
        def run(self):
            self.addChildTargetFn(lambda t: 1 / 0)
            self.setFollowOnFn(lambda: log.append("never"))

    assert Stack(Bad()).startJobTree(None) == 1 and "never" not in log      # pipeline.py:207-210 raises on failed jobs

    seen = {}

    class Tmp(Target):
        def run(self):
            seen["g"] = self.getGlobalTempDir()
            open(os.path.join(seen["g"], "f"), "w").close()
            self.setFollowOnTargetFn(lambda t, g: seen.setdefault("alive", os.path.exists(os.path.join(g, "f"))), args=(seen["g"],))

    assert Stack(Tmp()).startJobTree(None) == 0 and seen["alive"] and not os.path.exists(seen["g"])
