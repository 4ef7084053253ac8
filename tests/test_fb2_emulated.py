"""The sources of k_geometry, k_records and k_fb2 (nanopore_b200/csrc/), compiled for the host and run under the fiber
emulation of tests/tools/warp_emu/ (one fiber per thread of the block, switching at every __syncthreads in shuffled
order), against the checker: posterior pairs of every DP region bit for bit, cell counts equal.  No GPU needed; the
same comparison runs on the device in tests/test_gpu_parity.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from nanopore_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "tools", "warp_emu")
TRIM, SPLIT = 14, 3000


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("warp_emu") / "libfb2_emu.so")
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-I", EMU_DIR, "-o", out,
                           os.path.join(EMU_DIR, "fb2_emu.cpp")])
    lib = C.CDLL(out)
    vp = C.c_void_p
    lib.emu_fb2_region.restype = C.c_int
    lib.emu_fb2_region.argtypes = [vp, C.c_int64, vp, C.c_int64, vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.c_double,
                                   C.c_int, C.c_int, C.c_uint, vp, vp, vp, C.c_int, vp]
    lib.emu_fwdbwd_region.restype = C.c_int
    lib.emu_fwdbwd_region.argtypes = lib.emu_fb2_region.argtypes
    lib.emu_fb2_region_expect.restype = C.c_int
    lib.emu_fb2_region_expect.argtypes = [vp, C.c_int64, vp, C.c_int64, vp, C.c_int, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_uint, vp, vp, vp, vp]
    return lib


def anchors_and_runs(ops):
    ax, ay, runs = [], [], []
    x = y = 0
    for o in ops:
        ln, code = int(o) >> 2, int(o) & 3
        if code == 0:
            if ln > 2 * TRIM:
                runs.append((x + TRIM, y + TRIM, ln - 2 * TRIM))
                ax.extend(range(x + TRIM, x + ln - TRIM))
                ay.extend(range(y + TRIM, y + ln - TRIM))
            x += ln; y += ln
        elif code == 1:
            y += ln
        else:
            x += ln
    return np.array(ax, dtype=np.int64), np.array(ay, dtype=np.int64), runs


def check_read(lib, X, Y, ops, band, warps=4, wcap=0, seed=0, min_diags=1000, tb_diags=40, threshold=0.01, kernel="emu_fb2_region"):
    model = oracle.Model()
    params = oracle.make_params(expansion=band, min_diags=min_diags, tb_diags=tb_diags, threshold=threshold)
    m60 = np.ascontiguousarray(model.dump(), dtype=np.float64)
    ax, ay, runs = anchors_and_runs(ops)
    X, Y = np.ascontiguousarray(X, dtype=np.uint8), np.ascontiguousarray(Y, dtype=np.uint8)
    j, n_regions = 0, 0
    for (x1, y1, x2, y2, a0, a1, rl, rr) in oracle.regions(ops, len(X), len(Y), TRIM, SPLIT):
        mine = []
        while j < len(runs) and runs[j][0] + runs[j][1] < x2 + y2:          # plan_read() of phmm_api.cu
            mine.append((runs[j][0] - x1, runs[j][1] - y1, runs[j][2]))
            j += 1
        want = oracle.posteriors(model, X[x1:x2], Y[y1:y2], ax[a0:a1] - x1, ay[a0:a1] - y1, params, bool(rl), bool(rr))
        cap = 8 * int(min(x2 - x1, y2 - y1)) + 1024
        px, py, pw = (np.zeros(cap, dtype=np.int32) for _ in range(3))
        region = np.array([x1, y1, x2, y2, rl, rr], dtype=np.int64)
        r = np.array(mine, dtype=np.int32).reshape(-1, 3)
        cells = C.c_int64(0)
        vp = lambda a: a.ctypes.data_as(C.c_void_p)
        n = getattr(lib, kernel)(vp(X), len(X), vp(Y), len(Y), vp(region), len(mine), vp(r), vp(m60), band, min_diags, tb_diags,
                               threshold, warps, wcap, seed, vp(px), vp(py), vp(pw), cap, C.byref(cells))
        assert n >= 0
        assert cells.value == want["cells"]
        o = np.lexsort((py[:n], px[:n]))
        ow = np.lexsort((want["py"], want["px"]))
        assert n == len(want["px"])
        assert np.array_equal(px[:n][o], want["px"][ow]) and np.array_equal(py[:n][o], want["py"][ow])
        assert np.array_equal(pw[:n][o], want["pw"][ow]), "posterior integers differ"
        n_regions += 1
    return n_regions


def test_emulated_fb2_equals_the_checker_bit_for_bit(emu):
    b = synth.make_batch(2, 1200, 5000, seed=41)                      # chained-global: leading / trailing deletions, several windows
    for i in range(b.n):
        assert check_read(emu, b.ref[b.ref_start[i]:b.ref_end[i]], b.read(i), b.ops(i), band=50, seed=i + 1) >= 1
    b = synth.make_batch(2, 900, 900, seed=42, global_form=False)
    for i in range(b.n):
        check_read(emu, b.ref[b.ref_start[i]:b.ref_end[i]], b.read(i), b.ops(i), band=20, seed=7)


@pytest.mark.parametrize("warps,wcap", [(2, 0), (8, 0), (4, 64)])
def test_emulated_fb2_variants(emu, warps, wcap):
    """2 and 8 warps per region; 64 shared-memory columns, so that most diagonals take the guarded / global-buffer paths;
    short traceback windows."""
    b = synth.make_batch(1, 800, 3000, seed=43)
    check_read(emu, b.ref[b.ref_start[0]:b.ref_end[0]], b.read(0), b.ops(0), band=50, warps=warps, wcap=wcap, seed=3,
               min_diags=150, tb_diags=30)


@pytest.mark.parametrize("warps", [1, 2, 4])
def test_emulated_first_generation_kernel(emu, warps):
    """k_fwdbwd (option legacy_kernel; the fall-back for E-step windows too long for the windowed kernel's ring)."""
    b = synth.make_batch(1, 700, 2500, seed=44)
    check_read(emu, b.ref[b.ref_start[0]:b.ref_end[0]], b.read(0), b.ops(0), band=20, warps=warps, seed=warps, min_diags=200,
               tb_diags=40, kernel="emu_fwdbwd_region")


def regions_with_runs(ops, lX, lY, split):
    """(region row, region-local anchor runs) as plan_read() of phmm_api.cu assigns them."""
    _, _, runs = anchors_and_runs(ops)
    j, out = 0, []
    for row in oracle.regions(ops, lX, lY, TRIM, split):
        x1, y1, x2, y2 = (int(v) for v in row[:4])
        mine = []
        while j < len(runs) and runs[j][0] + runs[j][1] < x2 + y2:
            mine.append((runs[j][0] - x1, runs[j][1] - y1, runs[j][2]))
            j += 1
        out.append((row, mine))
    return out


@pytest.mark.parametrize("warps,band,split", [(4, 10, 300), (2, 10, 300), (4, 50, 3000)])
def test_emulated_estep_integers_equal_the_checkers(emu, warps, band, split):
    """k_fb2<.., EXPECT>: the 105 expectations in 2^-32 fixed point and the log-likelihood in 2^-20, summed over the
    regions of a read the way expectations_reduce() of phmm_api.cu does, against po_expectations_fixed."""
    b = synth.make_batch(2, 700, 700, seed=51, global_form=False)          # EM inputs are global: window = contig (utils.py:492-496)
    model = oracle.Model()
    m60 = np.ascontiguousarray(model.dump(), dtype=np.float64)
    params = oracle.make_params(expansion=band, split_side=split, min_diags=200, tb_diags=40)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    for i in range(b.n):
        X = np.ascontiguousarray(b.ref[b.ref_start[i]:b.ref_end[i]], dtype=np.uint8)
        Y = np.ascontiguousarray(b.read(i), dtype=np.uint8)
        acc = [0] * 106
        cells = 0
        for row, mine in regions_with_runs(b.ops(i), len(X), len(Y), split):
            region = np.array([row[0], row[1], row[2], row[3], row[6], row[7]], dtype=np.int64)
            r = np.array(mine, dtype=np.int32).reshape(-1, 3)
            T, E = np.zeros(25, dtype=np.uint64), np.zeros(80, dtype=np.uint64)
            LL, c = C.c_double(0.0), C.c_int64(0)
            rc = emu.emu_fb2_region_expect(vp(X), len(X), vp(Y), len(Y), vp(region), len(mine), vp(r), vp(m60), band, 200, 40, warps, 0,
                                           i + 5, vp(T), vp(E), C.byref(LL), C.byref(c))
            assert rc >= 0
            cells += c.value
            for k in range(25):
                acc[k] += int(T[k])
            for k in range(80):
                acc[25 + k] += int(E[k])
            acc[105] += int(np.rint(LL.value * 1048576.0))
        hi, lo, want_cells = oracle.expectations_fixed(model, X, Y, b.ops(i), params)
        assert cells == want_cells
        for k in range(106):
            bits = 32 if k < 105 else 20
            h = acc[k] >> bits
            assert (h, acc[k] - (h << bits)) == (int(hi[k]), int(lo[k])), "statistic %d differs" % k
