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


# ---- the block-per-region kernel k_decode<2> (takes the regions k_decode_w leaves) under the same emulation ----
@pytest.fixture(scope="module")
def emu_block(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("warp_emu") / "libdecode_emu.so")
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-I", EMU_DIR, "-o", out,
                           os.path.join(EMU_DIR, "decode_emu.cpp")])
    lib = C.CDLL(out)
    vp = C.c_void_p
    lib.emu_decode_block.restype = C.c_int
    lib.emu_decode_block.argtypes = [C.c_int, C.c_int, C.c_int, vp, C.c_int, C.c_int, vp, vp, vp, C.c_double, C.c_double, C.c_int,
                                     C.c_uint, vp, vp, vp, C.c_int, vp, vp]
    return lib


def run_block_emu(lib, lx, ly, runs, band, px, py, pw, gap_gamma=0.5, full_sweep=0, seed=0):
    px, py, pw = (np.ascontiguousarray(v, dtype=np.int32) for v in (px, py, pw))
    r = np.ascontiguousarray(np.array(runs, dtype=np.int32).reshape(-1, 3))
    cap = min(lx, ly) + 1
    mrx, mry, mrn = (np.zeros(cap, dtype=np.int32) for _ in range(3))
    score, regular = C.c_int64(0), C.c_int32(-1)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    n = lib.emu_decode_block(lx, ly, len(r), vp(r), band, len(px), vp(px), vp(py), vp(pw), gap_gamma, 0.0, full_sweep, seed,
                             vp(mrx), vp(mry), vp(mrn), cap, C.byref(score), C.byref(regular))
    cx, cy = [], []
    for k in range(n - 1, -1, -1):
        cx.extend(range(mrx[k], mrx[k] + mrn[k]))
        cy.extend(range(mry[k], mry[k] + mrn[k]))
    return np.array(cx, dtype=np.int64), np.array(cy, dtype=np.int64), score.value, regular.value


def anchor_runs(ops, trim=14):
    runs, x, y = [], 0, 0
    for o in ops:
        ln, code = int(o) >> 2, int(o) & 3
        if code == 0:
            if ln > 2 * trim:
                runs.append((x + trim, y + trim, ln - 2 * trim))
            x += ln; y += ln
        elif code == 1:
            y += ln
        else:
            x += ln
    return runs


@pytest.mark.parametrize("full_sweep", [0, 1])
def test_emulated_block_kernel_returns_the_checkers_chain(emu_block, full_sweep):
    b = synth.make_batch(3, 1200, 5000, seed=31)
    model, params = oracle.Model(), oracle.make_params(expansion=50)
    n = 0
    for i in range(b.n):
        X = b.ref[b.ref_start[i]:b.ref_end[i]]
        if len(oracle.regions(b.ops(i), len(X), len(b.read(i)), 14, 3000)) != 1:
            continue
        want = oracle.realign(model, X, b.read(i), b.ops(i), params)
        cx, cy, score, regular = run_block_emu(emu_block, len(X), len(b.read(i)), anchor_runs(b.ops(i)), 50, want["px"], want["py"],
                                               want["pw"], full_sweep=full_sweep, seed=i + 1)
        assert regular == 1
        assert score == want["mea_score"] and np.array_equal(cx, want["cx"]) and np.array_equal(cy, want["cy"])
        n += 1
    assert n >= 2


def test_both_emulated_kernels_agree_under_ties(emu, emu_block):
    """Random pairs with weights in {0, 1, 2} inside the band of a few anchor runs: k_decode<2> with and without its
    skipping, k_decode_w (which never sees the band) and the plain-Python full sweep pick the same chain."""
    import random
    from test_decode_skip_theory import full_sweep
    rng = random.Random(29)
    n = 0
    for trial in range(120):
        lx, ly = rng.randint(20, 90), rng.randint(20, 90)
        # two or three anchor runs along a staircase, band 4..10
        runs, x, y = [], rng.randint(0, 5), rng.randint(0, 5)
        for _ in range(rng.randint(1, 3)):
            ln = rng.randint(2, 8)
            if x + ln >= lx or y + ln >= ly:
                break
            runs.append((x, y, ln))
            x += ln + rng.randint(1, 25); y += ln + rng.randint(1, 25)
        band = 2 * rng.randint(2, 5)
        ax = np.array([r[0] + k for r in runs for k in range(r[2])], dtype=np.int64)
        ay = np.array([r[1] + k for r in runs for k in range(r[2])], dtype=np.int64)
        L, R = oracle.band(ax, ay, lx, ly, band)
        d = np.arange(lx + ly + 1)
        lo, hi = ((d + L) // 2).tolist(), ((d + R) // 2).tolist()
        cells = [(xx, dd - xx) for dd in range(2, lx + ly + 1) for xx in range(lo[dd], hi[dd] + 1)
                 if xx >= 1 and dd - xx >= 1 and lo[dd - 2] <= xx - 1 <= hi[dd - 2]]
        if len(cells) < 4:
            continue
        centre = rng.choice(cells)
        near = [c for c in cells if abs(c[0] + c[1] - centre[0] - centre[1]) <= rng.randint(1, 10)] + rng.sample(cells, 2)
        chosen = list(dict.fromkeys(rng.sample(near, min(len(near), rng.randint(1, 14)))))
        pairs = {c: (rng.choice([0, 1, 1, 2]), k) for k, c in enumerate(chosen)}
        want_score, want_ids = full_sweep(lo, hi, lx, ly, pairs)
        by_id = {k: c for c, (_, k) in pairs.items()}
        want_chain = [by_id[k] for k in want_ids]
        args = ([c[0] - 1 for c in chosen], [c[1] - 1 for c in chosen], [pairs[c][0] for c in chosen])
        for fs in (0, 1):
            cx, cy, score, regular = run_block_emu(emu_block, lx, ly, runs, band, *args, gap_gamma=0.0, full_sweep=fs, seed=trial + 1)
            assert regular == 1
            assert score == want_score and [(int(a) + 1, int(b_) + 1) for a, b_ in zip(cx, cy)] == want_chain, (lx, ly, runs, band, pairs, fs)
        cx, cy, score, _ = run_emu(emu, lx, ly, *args, gap_gamma=0.0, seed=trial + 1)
        assert score == want_score and [(int(a) + 1, int(b_) + 1) for a, b_ in zip(cx, cy)] == want_chain
        n += 1
    assert n > 60
