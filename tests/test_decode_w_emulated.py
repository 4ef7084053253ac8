"""k_decode_w's own source (nanopore_b200/csrc/phmm_decode_w.cuh), compiled for the host and run under a 32-lane
fiber emulation of the warp intrinsics (tests/tools/warp_emu/), against the checker: the checker's posterior pairs in,
the checker's maximum-expected-accuracy chain and score out.  Needs g++ only; the same comparison runs on the real
kernel in tests/test_gpu_parity.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from nanopore_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "tools", "warp_emu")


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("warp_emu") / "libdecode_w_emu.so")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I", EMU_DIR, "-o", out,
                           os.path.join(EMU_DIR, "decode_w_emu.cpp")])
    lib = C.CDLL(out)
    i32p = C.POINTER(C.c_int32)
    lib.emu_decode_w.restype = C.c_int
    lib.emu_decode_w.argtypes = [C.c_int, C.c_int, C.c_int, i32p, i32p, i32p, C.c_double, C.c_double, C.c_int, C.c_uint,
                                 i32p, i32p, i32p, C.c_int, C.POINTER(C.c_int64), i32p]
    return lib


def run_emu(lib, lx, ly, px, py, pw, gap_gamma=0.5, match_gamma=0.0, regular=1, seed=0):
    px, py, pw = (np.ascontiguousarray(v, dtype=np.int32) for v in (px, py, pw))
    cap = min(lx, ly) + 1
    mrx, mry, mrn = (np.zeros(cap, dtype=np.int32) for _ in range(3))
    score, env = C.c_int64(0), C.c_int32(0)
    p = lambda v: v.ctypes.data_as(C.POINTER(C.c_int32))
    n = lib.emu_decode_w(lx, ly, len(px), p(px), p(py), p(pw), gap_gamma, match_gamma, regular, seed, p(mrx), p(mry), p(mrn), cap,
                         C.byref(score), C.byref(env))
    if n < 0:
        return None
    cx, cy = [], []
    for r in range(n - 1, -1, -1):                       # runs come out last first
        cx.extend(range(mrx[r], mrx[r] + mrn[r]))
        cy.extend(range(mry[r], mry[r] + mrn[r]))
    return np.array(cx, dtype=np.int64), np.array(cy, dtype=np.int64), score.value, env.value


def one_region_reads(n, read_len, ref_len, seed, band, **kw):
    b = synth.make_batch(n, read_len, ref_len, seed=seed, **kw)
    model, params = oracle.Model(), oracle.make_params(expansion=band)
    for i in range(b.n):
        X = b.ref[b.ref_start[i]:b.ref_end[i]]
        if len(oracle.regions(b.ops(i), len(X), len(b.read(i)), 14, 3000)) != 1:
            continue
        yield X, b.read(i), oracle.realign(model, X, b.read(i), b.ops(i), params)


@pytest.mark.parametrize("seed", [0, 1, 12345])
def test_emulated_kernel_returns_the_checkers_chain(emu, seed):
    n = 0
    for X, Y, want in one_region_reads(4, 1200, 5000, seed=31, band=50):
        got = run_emu(emu, len(X), len(Y), want["px"], want["py"], want["pw"], seed=seed)
        assert got is not None
        cx, cy, score, env = got
        assert score == want["mea_score"]
        assert np.array_equal(cx, want["cx"]) and np.array_equal(cy, want["cy"])
        assert env < want["cells"] // 8                   # the envelope is a small part of the band
        n += 1
    for X, Y, want in one_region_reads(3, 900, 900, seed=32, band=20, global_form=False):
        cx, cy, score, _ = run_emu(emu, len(X), len(Y), want["px"], want["py"], want["pw"], seed=seed)
        assert score == want["mea_score"] and np.array_equal(cx, want["cx"]) and np.array_equal(cy, want["cy"])
        n += 1
    assert n >= 5


def test_emulated_kernel_edge_cases(emu):
    # no pairs at all: empty chain, score 0
    cx, cy, score, _ = run_emu(emu, 40, 30, [], [], [])
    assert len(cx) == 0 and score == 0
    # one pair whose reweighted mass is not positive: unusable, empty chain, and the envelope is the path through it
    cx, cy, score, env = run_emu(emu, 40, 30, [7], [9], [1000])
    assert len(cx) == 0 and score == 0 and env < 3 * 71
    # a single pair, pairs in the corners, a one-base read
    cx, cy, score, _ = run_emu(emu, 5, 5, [0, 4], [0, 4], [9000000, 9000000])
    assert list(cx) == [0, 4] and list(cy) == [0, 4]
    cx, cy, score, _ = run_emu(emu, 300, 1, [123], [0], [9999999])
    assert list(cx) == [123] and list(cy) == [0]
    # irregular band flag, more than 32 pairs on one diagonal, an envelope wider than the ring: left to k_decode
    assert run_emu(emu, 50, 50, [3], [3], [5000000], regular=0) is None
    xs = np.arange(40)
    assert run_emu(emu, 60, 60, xs, 39 - xs, np.full(40, 200000)) is None
    assert run_emu(emu, 400, 400, [0, 399], [399, 0], [5000000, 5000000]) is None


def test_emulated_kernel_tie_rules_on_random_bands(emu):
    """Weights in {0, 1, 2} (gap_gamma 0, so the kernel's reweighting leaves them as they are): equal scores everywhere.
    The kernel never sees the band -- it must still pick the chain the full sweep picks on a random regular band that
    holds the pairs (tests/test_decode_skip_theory.py::full_sweep)."""
    import random
    from test_decode_narrow_theory import random_case
    from test_decode_skip_theory import full_sweep
    rng = random.Random(23)
    n = 0
    for trial in range(500):
        case = random_case(rng)
        if case is None:
            continue
        lx, ly, lo, hi, pairs = case
        want_score, want_ids = full_sweep(lo, hi, lx, ly, pairs)
        by_id = {k: c for c, (_, k) in pairs.items()}
        cells = list(pairs)
        got = run_emu(emu, lx, ly, [c[0] - 1 for c in cells], [c[1] - 1 for c in cells], [pairs[c][0] for c in cells],
                      gap_gamma=0.0, seed=trial)
        assert got is not None
        cx, cy, score, _ = got
        assert score == want_score, (lx, ly, pairs)
        assert [(int(x) + 1, int(y) + 1) for x, y in zip(cx, cy)] == [by_id[k] for k in want_ids], (lx, ly, lo, hi, pairs)
        n += 1
    assert n > 200
