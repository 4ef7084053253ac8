"""ctypes wrapper over oracle/libphmm_oracle.so (the CPU oracle).

TEST INFRASTRUCTURE ONLY -- see the header of oracle/phmm_oracle.c.  Only
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this package; nanopore_b200/ never does.  Parity unpinned (the upstream
cactus/sonLib sources are absent from the reference tree).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libphmm_oracle.so")


def build(force=False):
    src = os.path.join(_HERE, "phmm_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libphmm_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


class _Params(C.Structure):
    _fields_ = [("expansion", C.c_int64), ("trim", C.c_int64), ("split_side", C.c_int64),
                ("min_diags", C.c_int64), ("tb_diags", C.c_int64), ("threshold", C.c_double),
                ("gap_gamma", C.c_double), ("match_gamma", C.c_double)]


class _Stats(C.Structure):
    _fields_ = [("cells", C.c_int64), ("diagonals", C.c_int64), ("tracebacks", C.c_int64),
                ("max_live_cells", C.c_int64)]


class _Result(C.Structure):
    _fields_ = [("n_ops", C.c_int64), ("ops", C.POINTER(C.c_uint32)),
                ("n_pairs", C.c_int64), ("px", C.POINTER(C.c_int64)), ("py", C.POINTER(C.c_int64)),
                ("pw", C.POINTER(C.c_int64)),
                ("n_chain", C.c_int64), ("cx", C.POINTER(C.c_int64)), ("cy", C.POINTER(C.c_int64)),
                ("mea_score", C.c_int64), ("stats", _Stats)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.po_logadd.restype = C.c_double
        L.po_logadd.argtypes = [C.c_double, C.c_double]
        L.po_exp.restype = C.c_double
        L.po_exp.argtypes = [C.c_double]
        L.po_model_create.restype = C.c_void_p
        L.po_model_create.argtypes = [C.c_void_p, C.c_void_p]
        L.po_model_destroy.argtypes = [C.c_void_p]
        L.po_model_dump.argtypes = [C.c_void_p, C.c_void_p]
        L.po_params_default.argtypes = [C.POINTER(_Params)]
        L.po_realign.restype = C.POINTER(_Result)
        L.po_realign.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                 C.POINTER(_Params)]
        L.po_posteriors.restype = C.POINTER(_Result)
        L.po_posteriors.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                    C.c_int64, C.POINTER(_Params), C.c_int, C.c_int]
        L.po_result_free.argtypes = [C.POINTER(_Result)]
        L.po_expectations.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                      C.POINTER(_Params), C.c_void_p, C.c_void_p, C.POINTER(C.c_double),
                                      C.POINTER(_Stats)]
        L.po_expectations_fixed.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64,
                                            C.POINTER(_Params), C.c_void_p, C.c_void_p, C.POINTER(_Stats)]
        L.po_band.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
        L.po_regions.restype = C.c_int64
        L.po_regions.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_int64, C.c_void_p, C.c_int64]
        L.po_realign_batch.restype = C.c_int64
        L.po_realign_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.POINTER(_Params), C.c_void_p, C.c_int64, C.c_void_p,
                                       C.POINTER(C.c_int64)]
        L.po_set_exact_logadd.argtypes = [C.c_int]
        L.po_set_upstream_arithmetic.argtypes = [C.c_int]
        L.po_get_upstream_arithmetic.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def logadd(x, y):
    return lib().po_logadd(float(x), float(y))


def exp(x):
    return lib().po_exp(float(x))


def set_exact_logadd(on):
    lib().po_set_exact_logadd(int(bool(on)))


UP_UNFUSED_HORNER, UP_LIBM_EXP, UP_GREEDY_ORDER = 1, 2, 4


def set_upstream_arithmetic(flags):
    """Differential runs only (tests/tools/differential.py): bit mask of UP_* that replaces the documented deviations of the
    oracle (header of phmm_oracle.c, items 1-3) by the recalled upstream behaviour.  0 = the arithmetic the CUDA library
    implements."""
    lib().po_set_upstream_arithmetic(int(flags))


def get_upstream_arithmetic():
    return int(lib().po_get_upstream_arithmetic())


def make_params(expansion=10, trim=14, split_side=3000, gap_gamma=0.5, match_gamma=0.0,
                min_diags=1000, tb_diags=40, threshold=0.01):
    p = _Params()
    lib().po_params_default(C.byref(p))
    p.expansion, p.trim, p.split_side = expansion, trim, split_side
    p.gap_gamma, p.match_gamma = gap_gamma, match_gamma
    p.min_diags, p.tb_diags, p.threshold = min_diags, tb_diags, threshold
    return p


class Model:
    """stateMachine5 built from HMM probabilities (trans[25], emis[80]) or the stock model."""

    def __init__(self, trans=None, emis=None):
        if trans is None:
            self.h = lib().po_model_create(None, None)
        else:
            t = np.ascontiguousarray(trans, dtype=np.float64)
            e = np.ascontiguousarray(emis, dtype=np.float64)
            assert t.size == 25 and e.size == 80
            with np.errstate(divide="ignore"):
                self.h = lib().po_model_create(_p(t), _p(e))

    def dump(self):
        out = np.zeros(60, dtype=np.float64)
        lib().po_model_dump(self.h, _p(out))
        return out

    def __del__(self):
        try:
            lib().po_model_destroy(self.h)
        except Exception:
            pass


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _take(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True)


def _unpack(r, want_pairs=True):
    rr = r.contents
    out = {
        "ops": _take(rr.ops, rr.n_ops, np.uint32) if rr.n_ops else np.zeros(0, np.uint32),
        "mea_score": rr.mea_score,
        "cells": rr.stats.cells, "diagonals": rr.stats.diagonals, "tracebacks": rr.stats.tracebacks,
        "max_live_cells": rr.stats.max_live_cells,
    }
    if want_pairs:
        out["px"] = _take(rr.px, rr.n_pairs, np.int64)
        out["py"] = _take(rr.py, rr.n_pairs, np.int64)
        out["pw"] = _take(rr.pw, rr.n_pairs, np.int64)
        out["cx"] = _take(rr.cx, rr.n_chain, np.int64)
        out["cy"] = _take(rr.cy, rr.n_chain, np.int64)
    lib().po_result_free(r)
    return out


def realign(model, X, Y, in_ops, params):
    """One read. X = reference window codes, Y = read codes, in_ops uint32 (len<<2)|code."""
    X, Y = _u8(X), _u8(Y)
    ops = np.ascontiguousarray(in_ops, dtype=np.uint32)
    r = lib().po_realign(model.h, _p(X), len(X), _p(Y), len(Y), _p(ops), len(ops), C.byref(params))
    return _unpack(r)


def posteriors(model, X, Y, ax, ay, params, ragged_left=False, ragged_right=False):
    X, Y = _u8(X), _u8(Y)
    ax = np.ascontiguousarray(ax, dtype=np.int64)
    ay = np.ascontiguousarray(ay, dtype=np.int64)
    r = lib().po_posteriors(model.h, _p(X), len(X), _p(Y), len(Y), _p(ax), _p(ay), len(ax), C.byref(params),
                            int(ragged_left), int(ragged_right))
    return _unpack(r)


def expectations(model, X, Y, in_ops, params, T=None, E=None, loglik=0.0):
    """E-step of one read, ADDED into T[25], E[80], loglik (fresh zeros if None)."""
    X, Y = _u8(X), _u8(Y)
    ops = np.ascontiguousarray(in_ops, dtype=np.uint32)
    T = np.zeros(25) if T is None else T
    E = np.zeros(80) if E is None else E
    ll = C.c_double(loglik)
    st = _Stats()
    lib().po_expectations(model.h, _p(X), len(X), _p(Y), len(Y), _p(ops), len(ops), C.byref(params), _p(T), _p(E),
                          C.byref(ll), C.byref(st))
    return T, E, ll.value, st.cells


def expectations_fixed(model, X, Y, in_ops, params, hi=None, lo=None):
    """E-step of one read as exact integers, ADDED into hi[106], lo[106] (see po_expectations_fixed)."""
    X, Y = _u8(X), _u8(Y)
    ops = np.ascontiguousarray(in_ops, dtype=np.uint32)
    hi = np.zeros(106, dtype=np.int64) if hi is None else hi
    lo = np.zeros(106, dtype=np.int64) if lo is None else lo
    st = _Stats()
    lib().po_expectations_fixed(model.h, _p(X), len(X), _p(Y), len(Y), _p(ops), len(ops), C.byref(params), _p(hi), _p(lo),
                                C.byref(st))
    return hi, lo, st.cells


def band(ax, ay, lX, lY, expansion):
    ax = np.ascontiguousarray(ax, dtype=np.int64)
    ay = np.ascontiguousarray(ay, dtype=np.int64)
    L = np.zeros(lX + lY + 1, dtype=np.int64)
    R = np.zeros(lX + lY + 1, dtype=np.int64)
    lib().po_band(_p(ax), _p(ay), len(ax), lX, lY, expansion, _p(L), _p(R))
    return L, R


def regions(in_ops, lX, lY, trim, split_side):
    ops = np.ascontiguousarray(in_ops, dtype=np.uint32)
    cap = 4096
    out = np.zeros(cap * 8, dtype=np.int64)
    n = lib().po_regions(_p(ops), len(ops), lX, lY, trim, split_side, _p(out), cap)
    return out[: n * 8].reshape(n, 8)


def realign_batch(model, ref, reads, read_off, ref_start, ref_end, in_ops, in_off, params):
    """Batch driver with the same packed layout as include/phmm.h. Returns (ops, off, cells)."""
    ref, reads = _u8(ref), _u8(reads)
    read_off = np.ascontiguousarray(read_off, dtype=np.int64)
    ref_start = np.ascontiguousarray(ref_start, dtype=np.int64)
    ref_end = np.ascontiguousarray(ref_end, dtype=np.int64)
    in_ops = np.ascontiguousarray(in_ops, dtype=np.uint32)
    in_off = np.ascontiguousarray(in_off, dtype=np.int64)
    n = len(read_off) - 1
    cap = int(2 * (len(reads) + int((ref_end - ref_start).sum())) + 16 * n + 16)
    out = np.zeros(cap, dtype=np.uint32)
    off = np.zeros(n + 1, dtype=np.int64)
    cells = C.c_int64(0)
    tot = lib().po_realign_batch(model.h, _p(ref), n, _p(reads), _p(read_off), _p(ref_start), _p(ref_end), _p(in_ops),
                                 _p(in_off), C.byref(params), _p(out), cap, _p(off), C.byref(cells))
    if tot < 0:
        raise RuntimeError("oracle output overflow")
    return out[:tot].copy(), off, cells.value
