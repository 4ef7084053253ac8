"""Generates tests/golden/oracle_vectors.json: seeded inputs -> CIGARs, posterior checksums, MEA scores and E-step
integers of the CPU oracle.  Committed so that later refactors of the oracle (and, through the GPU parity tests, of
the kernels) are pinned to today's arithmetic.  Parity with upstream cactus_realign stays unpinned.
usage: python tests/tools/make_golden.py"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle                                   # noqa: E402
from nanopore_b200 import synth                 # noqa: E402
from nanopore_b200.hmm import Hmm               # noqa: E402

CASES = [
    {"name": "stock_band10", "n": 3, "L": 300, "R": 900, "seed": 101, "band": 10, "split": 3000, "model": "stock", "global_form": True},
    {"name": "stock_band50_local", "n": 2, "L": 800, "R": 800, "seed": 102, "band": 50, "split": 3000, "model": "stock", "global_form": False},
    {"name": "trained_band10_split100", "n": 3, "L": 400, "R": 6000, "seed": 103, "band": 10, "split": 100, "model": "blasr_hmm_0.txt", "global_form": True},
    {"name": "trained_band20_em300", "n": 2, "L": 600, "R": 2000, "seed": 104, "band": 20, "split": 300, "model": "blasr_hmm_0.txt", "global_form": True},
]


def cigar_string(ops):
    return "".join("%d%s" % (l, "MID"[c]) for c, l in synth.unpack_ops(ops))


def run_case(c):
    b = synth.make_batch(c["n"], c["L"], c["R"], seed=c["seed"], global_form=c["global_form"])
    if c["model"] == "stock":
        m = oracle.Model()
    else:
        t, e = Hmm.loadHmm(os.path.join(ROOT, "tests", "golden", c["model"])).arrays()
        m = oracle.Model(t, e)
    p = oracle.make_params(expansion=c["band"], split_side=c["split"])
    reads = []
    hi, lo = np.zeros(106, np.int64), np.zeros(106, np.int64)
    for i in range(b.n):
        X = b.ref[b.ref_start[i]:b.ref_end[i]]
        r = oracle.realign(m, X, b.read(i), b.ops(i), p)
        o = np.lexsort((r["py"], r["px"]))
        h = hashlib.sha256(np.stack([r["px"][o], r["py"][o], r["pw"][o]]).astype(np.int64).tobytes()).hexdigest()
        reads.append({"guide": cigar_string(b.ops(i)), "cigar": cigar_string(r["ops"]), "cells": int(r["cells"]),
                      "pairs": int(len(o)), "pairs_sha256": h, "weight_sum": int(r["pw"].sum()), "mea_score": int(r["mea_score"])})
        hi, lo, _ = oracle.expectations_fixed(m, X, b.read(i), b.ops(i), p, hi, lo)
    return {"case": c, "input_sha256": hashlib.sha256(b.ref.tobytes() + b.reads.tobytes() + b.in_ops.tobytes()).hexdigest(),
            "reads": reads, "expect_hi": hi.tolist(), "expect_lo": lo.tolist()}


if __name__ == "__main__":
    out = {"generator": "tests/tools/make_golden.py", "note": "oracle outputs (parity with upstream unpinned)",
           "logadd": [[x, y, oracle.logadd(x, y).hex()] for x, y in [(0.0, 0.0), (-1.5, -0.25), (-3.0, 2.0), (1.0, -6.4), (-10.0, 0.0), (-700.0, -701.0)]],
           "exp": [[x, oracle.exp(x).hex()] for x in (0.0, -0.01, -4.60517, -20.0, -699.0)],
           "cases": [run_case(c) for c in CASES]}
    path = os.path.join(ROOT, "tests", "golden", "oracle_vectors.json")
    json.dump(out, open(path, "w"), indent=1)
    print(path, os.path.getsize(path), "bytes")
