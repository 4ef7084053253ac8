"""tests/tools/differential.py: the plumbing of the differential run -- FASTA files, the exonerate cigar
on stdin, flags, the printed cigar -- exercised against a stand-in `cactus_realign` executable that answers with the CPU
checker (the real binary does not exist in this environment; with it the same script pins or refutes upstream parity)."""
import importlib.util
import os
import stat
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

SHIM = r'''#!%s
import sys
sys.path.insert(0, %r)
import numpy as np
import oracle
from nanopore_b200.batch import encode, pack_ops, unpack_ops
from nanopore_b200.bioio import cigarReadFromString, fastaRead
args = [a for a in sys.argv[1:] if not a.startswith("--")]
opts = dict(a[2:].split("=") for a in sys.argv[1:] if a.startswith("--"))
X = encode(next(iter(fastaRead(args[0])))[1]); Y = encode(next(iter(fastaRead(args[1])))[1])
pA = cigarReadFromString(sys.stdin.readline())
ops = pack_ops(np.array([[op.type, op.length] for op in pA.operationList], dtype=np.int64).reshape(-1, 2))
p = oracle.make_params(expansion=int(opts["diagonalExpansion"]), split_side=int(opts["splitMatrixBiggerThanThis"]),
                       gap_gamma=float(opts["gapGamma"]), match_gamma=float(opts["matchGamma"]))
r = oracle.realign(oracle.Model(), X, Y, ops, p)
print("cigar: read 0 %%d + ref 0 %%d + 1 %%s" %% (len(Y), len(X), " ".join("%%s %%d" %% ("MID"[c], l) for c, l in unpack_ops(r["ops"]))))
'''


def test_differential_script_against_a_stand_in_binary(tmp_path, capsys):
    shim = tmp_path / "cactus_realign"
    shim.write_text(SHIM % (sys.executable, ROOT))
    shim.chmod(shim.stat().st_mode | stat.S_IEXEC)
    spec = importlib.util.spec_from_file_location("differential", os.path.join(ROOT, "tests", "tools", "differential.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.main(["--binary", str(tmp_path / "missing")]) == 2
    assert mod.main(["--binary", str(shim), "--reads", "3", "--read-len", "300"]) == 0
    out = capsys.readouterr().out
    assert "shipped arithmetic       3 / 3" in out and "all upstream switches" in out
