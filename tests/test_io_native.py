"""libphmm_io.so (native ingest / chain / pack / emit) against the Python restatement of the same reference
functions (nanopore_b200/realign.py, sam.py; reference nanopore/analyses/utils.py:441-469,557-574,591-609):
byte-identical files, identical packed batches.  CPU only."""
import os

import numpy as np
import pytest

from helpers_sam import make_experiment
from nanopore_b200 import io_native, realign
from nanopore_b200.sam import Samfile


@pytest.fixture(scope="module", params=[0, 1, 2])
def experiment(request, tmp_path_factory):
    d = str(tmp_path_factory.mktemp("exp%d" % request.param))
    seed = request.param
    kw = [dict(n_reads=12, read_len=400, contig_lens=(1500, 1100), hits=(1, 2, 3), unmapped=2),
          dict(n_reads=40, read_len=900, contig_lens=(5000,), hits=(2, 3, 4), unmapped=0),
          dict(n_reads=7, read_len=150, contig_lens=(700, 900, 800), hits=(1,), unmapped=3)][request.param]
    ref_fa, fq, sam, truth = make_experiment(d, seed=seed, **kw)
    return d, ref_fa, fq, sam


def test_library_exports_every_declared_symbol():
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    decl = set(re.findall(r"\b(phmm_io_[a-z_]+)\s*\(", open(os.path.join(root, "include", "phmm_io.h")).read()))
    assert decl == set(io_native.SYMBOLS)
    io_native.load_library()                         # getattr on every symbol


@pytest.mark.parametrize("threads", [1, 4])
def test_chain_pack_and_emit_match_the_python_path(experiment, threads):
    d, ref_fa, fq, sam = experiment
    # Python path
    py_chain = os.path.join(d, "py_chained.sam")
    realign.chainSamFile(sam, py_chain, fq, ref_fa)
    # native path
    io = io_native.NativeIo(threads)
    io.load_reference(ref_fa)
    io.load_reads(fq)
    io.chain_sam(sam)
    nat_chain = os.path.join(d, "nat_chained_%d.sam" % threads)
    io.write_sam(nat_chain)
    assert open(nat_chain, "rb").read() == open(py_chain, "rb").read()
    # packed batch == packAlignedReads of the chained file
    s = Samfile(py_chain, "r")
    records = list(realign.samIterator(s))
    pref = realign.PackedReference(realign.getFastaDictionary(ref_fa))
    want = realign.packAlignedReads(records, s, pref)
    s.close()
    got = io.batch()
    assert io.counts() == (len(records), len(records))
    for name in ("ref", "reads", "read_off", "ref_start", "ref_end", "in_ops", "in_off"):
        assert np.array_equal(getattr(got, name), getattr(want, name)), name
    # fan-in: new cigars into the chained records (any ops will do: reversed guide ops per read)
    ops = np.concatenate([got.ops(i)[::-1] for i in range(got.n)]) if got.n else np.zeros(0, np.uint32)
    off = got.in_off
    py_out = os.path.join(d, "py_realigned.sam")
    realign.realignSamFile3TargetFn(None, py_chain, py_out, [ops[off[i]:off[i + 1]] for i in range(got.n)])
    nat_out = os.path.join(d, "nat_realigned_%d.sam" % threads)
    io.write_realigned_sam(nat_out, ops, off)
    assert open(nat_out, "rb").read() == open(py_out, "rb").read()
    io.close()


def test_load_sam_keeps_unmapped_records_out_of_the_batch_and_the_output(experiment):
    d, ref_fa, fq, sam = experiment
    io = io_native.NativeIo(2)
    io.load_reference(ref_fa)
    io.load_sam(sam)
    s = Samfile(sam, "r")
    allrecs = list(s)
    s.close()
    mapped = [a for a in allrecs if a.rname != -1]
    assert io.counts() == (len(allrecs), len(mapped))
    # the untouched file round-trips byte for byte (every field is re-serialised the way sam.Samfile does)
    rt = os.path.join(d, "roundtrip.sam")
    io.write_sam(rt)
    py_rt = os.path.join(d, "py_roundtrip.sam")
    s = Samfile(sam, "r")
    o = Samfile(py_rt, "wh", template=s)
    for a in s:
        o.write(a)
    o.close()
    s.close()
    assert open(rt, "rb").read() == open(py_rt, "rb").read()
    b = io.batch()
    assert b.n == len(mapped)
    s = Samfile(sam, "r")
    want = realign.packAlignedReads(list(realign.samIterator(s)), s, realign.PackedReference(realign.getFastaDictionary(ref_fa)))
    s.close()
    for name in ("reads", "read_off", "ref_start", "ref_end", "in_ops", "in_off"):
        assert np.array_equal(getattr(b, name), getattr(want, name)), name
    io.close()


def test_errors_carry_messages(tmp_path, experiment):
    d, ref_fa, fq, sam = experiment
    io = io_native.NativeIo(1)
    with pytest.raises(io_native.PhmmIoError) as ei:
        io.load_reference(str(tmp_path / "missing.fa"))
    assert "cannot open" in str(ei.value)
    with pytest.raises(io_native.PhmmIoError):           # chaining needs the sequences
        io.chain_sam(sam)
    io.load_reference(ref_fa)
    io.load_reads(fq)
    bad = tmp_path / "bad.sam"
    bad.write_text("@HD\tVN:1.0\nr1\t0\tref0\t1\t30\t5M3Q\t*\t0\t0\tACGTACGT\t*\n")
    with pytest.raises(io_native.PhmmIoError) as ei:
        io.chain_sam(str(bad))
    assert "not in read sequences" in str(ei.value) or "malformed CIGAR" in str(ei.value)
    short = tmp_path / "short.sam"
    short.write_text("r1\t0\tref0\t1\n")
    with pytest.raises(io_native.PhmmIoError) as ei:
        io.load_sam(str(short))
    assert "fields" in str(ei.value)
    dup = tmp_path / "dup.fa"
    dup.write_text(">a x\nACGT\n>a y\nAC\n")
    with pytest.raises(io_native.PhmmIoError) as ei:
        io.load_reference(str(dup))
    assert "duplicate" in str(ei.value)
    io.close()
