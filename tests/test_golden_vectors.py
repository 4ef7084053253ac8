"""Committed golden vectors (tests/golden/oracle_vectors.json, made by tests/tools/make_golden.py) pin the oracle's
arithmetic: logAdd / exp bit patterns, CIGARs, posterior pair sets, MEA scores and E-step integers.  The GPU test
checks the library against the same file, so kernel and oracle are pinned to one committed answer."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "tools"))
import make_golden                                  # noqa: E402

import oracle                                       # noqa: E402
from nanopore_b200 import synth                     # noqa: E402

GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_vectors.json")))


def test_arithmetic_bit_patterns():
    for x, y, h in GOLD["logadd"]:
        assert oracle.logadd(x, y).hex() == h
    for x, h in GOLD["exp"]:
        assert oracle.exp(x).hex() == h


@pytest.mark.parametrize("g", GOLD["cases"], ids=[g["case"]["name"] for g in GOLD["cases"]])
def test_oracle_reproduces_golden(g):
    assert make_golden.run_case(g["case"]) == g


@pytest.mark.gpu
@pytest.mark.parametrize("g", GOLD["cases"], ids=[g["case"]["name"] for g in GOLD["cases"]])
def test_gpu_reproduces_golden(g):
    import hashlib
    from nanopore_b200 import capi
    from nanopore_b200.hmm import Hmm
    c = g["case"]
    b = synth.make_batch(c["n"], c["L"], c["R"], seed=c["seed"], global_form=c["global_form"])
    if c["model"] == "stock":
        ctx = capi.PhmmContext(0)
    else:
        t, e = Hmm.loadHmm(os.path.join(ROOT, "tests", "golden", c["model"])).arrays()
        ctx = capi.PhmmContext(0, t, e, 1)
    ctx.set_reference(b.ref)
    p = capi.default_params(band=c["band"], split_side=c["split"])
    ops, off, post = ctx.realign_batch(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off, p, want_posteriors=True)
    for i, r in enumerate(g["reads"]):
        assert make_golden.cigar_string(ops[off[i]:off[i + 1]]) == r["cigar"]
        s = slice(post["off"][i], post["off"][i + 1])
        h = hashlib.sha256(np.stack([post["ref_pos"][s], post["read_pos"][s], post["prob_1e7"][s]]).astype(np.int64).tobytes()).hexdigest()
        assert h == r["pairs_sha256"] and int(post["prob_1e7"][s].astype(np.int64).sum()) == r["weight_sum"]
    hi, lo = ctx.expectations_batch_fixed(b.reads, b.read_off, b.ref_start, b.ref_end, b.in_ops, b.in_off, p)
    assert hi.tolist() == g["expect_hi"] and lo.tolist() == g["expect_lo"]
    ctx.close()
