"""Differential run of a real `cactus_realign` binary against the CPU oracle.

    python tests/tools/differential.py --binary /path/to/cactus_realign [--reads 20] [--read-len 2000] [--band 10]

For every synthetic read the script does what the reference does (reference nanopore/analyses/utils.py:576-589): writes
ref.fa / read.fa, pipes the exonerate cigar line into
    cactus_realign ref.fa read.fa --diagonalExpansion=B --splitMatrixBiggerThanThis=S --gapGamma=G --matchGamma=M
parses the cigar it prints, and compares the M/I/D ops with the oracle -- first with the arithmetic the CUDA library
implements, then with each documented deviation of the oracle undone (oracle.set_upstream_arithmetic: unfused Horner,
libm exp, greedy ordering) and with all three, so that a mismatch is attributed to a deviation.  The binary cannot be
built in this environment (its sources are empty submodules); without --binary, or when the file is missing, the script
says so and exits 2: parity with upstream stays UNPINNED until someone runs this where cactus exists.
Test infrastructure: imports oracle/, never used by the product path."""
import argparse
import os
import subprocess
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np                                                     # noqa: E402

import oracle                                                          # noqa: E402
from nanopore_b200 import synth                                        # noqa: E402
from nanopore_b200.bioio import cigarReadFromString, fastaWrite       # noqa: E402
from nanopore_b200.batch import pack_ops, unpack_ops                   # noqa: E402


def cigar_line(name, contig, lx, ly, ops):
    body = " ".join("%s %i" % ("MID"[c], l) for c, l in unpack_ops(ops))
    return "cigar: %s 0 %i + %s 0 %i + 1 %s" % (name, ly, contig, lx, body)


def run_binary(binary, d, X, Y, ops, a):
    fastaWrite(os.path.join(d, "ref.fa"), "ref", synth.decode(X))
    fastaWrite(os.path.join(d, "read.fa"), "read", synth.decode(Y))
    cmd = [binary, os.path.join(d, "ref.fa"), os.path.join(d, "read.fa"), "--diagonalExpansion=%d" % a.band,
           "--splitMatrixBiggerThanThis=%d" % a.split, "--gapGamma=%s" % a.gap_gamma, "--matchGamma=%s" % a.match_gamma]
    out = subprocess.run(cmd, input=cigar_line("read", "ref", len(X), len(Y), ops) + "\n", capture_output=True, text=True, check=True).stdout
    pA = cigarReadFromString(out.strip().splitlines()[-1])
    code = {pA.PAIRWISE_MATCH: 0, pA.PAIRWISE_INDEL_Y: 1, pA.PAIRWISE_INDEL_X: 2}
    return pack_ops(np.array([[code[op.type], op.length] for op in pA.operationList], dtype=np.int64).reshape(-1, 2))


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--binary", default=os.environ.get("CACTUS_REALIGN", ""))
    ap.add_argument("--reads", type=int, default=20)
    ap.add_argument("--read-len", type=int, default=2000)
    ap.add_argument("--band", type=int, default=10)
    ap.add_argument("--split", type=int, default=3000)
    ap.add_argument("--gap-gamma", type=float, default=0.5)
    ap.add_argument("--match-gamma", type=float, default=0.0)
    a = ap.parse_args(argv)
    if not a.binary or not os.path.exists(a.binary):
        print("no cactus_realign binary (%r): nothing compared; parity with upstream remains unpinned" % a.binary)
        return 2
    b = synth.make_batch(a.reads, a.read_len, 4 * a.read_len, seed=2024, global_form=False)
    p = oracle.make_params(expansion=a.band, split_side=a.split, gap_gamma=a.gap_gamma, match_gamma=a.match_gamma)
    m = oracle.Model()
    modes = [("shipped arithmetic", 0), ("unfused Horner", 1), ("libm exp", 2), ("greedy ordering", 4), ("all upstream switches", 7)]
    agree = {name: 0 for name, _ in modes}
    with tempfile.TemporaryDirectory() as d:
        for i in range(b.n):
            X, Y = b.ref[b.ref_start[i]:b.ref_end[i]], b.read(i)
            theirs = run_binary(a.binary, d, X, Y, b.ops(i), a)
            for name, flags in modes:
                oracle.set_upstream_arithmetic(flags)
                try:
                    ours = oracle.realign(m, X, Y, b.ops(i), p)["ops"]
                finally:
                    oracle.set_upstream_arithmetic(0)
                agree[name] += int(np.array_equal(ours, theirs))
    for name, _ in modes:
        print("%-24s %d / %d reads with identical cigar ops" % (name, agree[name], b.n))
    return 0 if agree["shipped arithmetic"] == b.n else 1


if __name__ == "__main__":
    sys.exit(main())
