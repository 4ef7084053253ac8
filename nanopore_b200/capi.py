"""ctypes binding of libphmm_sm100.so (C ABI declared in include/phmm.h).

This is the only door from the Python host code into the CUDA kernels.  It
raises if the shared library is missing or no CUDA device answers -- there is
no CPU fallback (the CPU restatement in oracle/ is test infrastructure and is
never imported from here).

Replaces, for the whole batch at once, the per-read `cactus_realign` process
the reference spawns (reference nanopore/analyses/utils.py:576-589).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PHMM_LIB") or os.path.join(_HERE, "libphmm_sm100.so")      # PHMM_LIB: tuning builds of the same library

PHMM_OK = 0
EXPORTS = [
    "phmm_version", "phmm_default_params", "phmm_create", "phmm_create_error", "phmm_destroy",
    "phmm_last_error", "phmm_set_stream", "phmm_set_model", "phmm_set_reference", "phmm_realign_batch",
    "phmm_expectations_batch", "phmm_expectations_batch_fixed", "phmm_batch_prepare", "phmm_batch_run", "phmm_batch_fetch",
    "phmm_batch_get_stats", "phmm_set_memory_budget", "phmm_set_option", "phmm_free", "phmm_free_posteriors",
    "phmm_expectations_prepare", "phmm_expectations_run_fixed",
    "phmm_base_expectations_reset", "phmm_batch_add_base_expectations", "phmm_base_expectations_fetch",
]


class PhmmError(RuntimeError):
    pass


class Params(C.Structure):
    """struct phmm_params"""
    _fields_ = [("band", C.c_int32), ("anchor_trim", C.c_int32), ("split_side", C.c_int64),
                ("min_diags", C.c_int32), ("tb_diags", C.c_int32), ("threshold", C.c_double),
                ("gap_gamma", C.c_double), ("match_gamma", C.c_double)]


class Posteriors(C.Structure):
    """struct phmm_posteriors"""
    _fields_ = [("n", C.c_int64), ("off", C.POINTER(C.c_int64)), ("ref_pos", C.POINTER(C.c_int32)),
                ("read_pos", C.POINTER(C.c_int32)), ("prob_1e7", C.POINTER(C.c_int32))]


class BatchStats(C.Structure):
    """struct phmm_batch_stats"""
    _fields_ = [("n_reads", C.c_int64), ("n_regions", C.c_int64), ("cells", C.c_int64),
                ("diagonals", C.c_int64), ("pairs", C.c_int64), ("launches", C.c_int64),
                ("ms_geometry", C.c_double), ("ms_fwdbwd", C.c_double), ("ms_decode", C.c_double),
                ("ms_total", C.c_double), ("slot_bytes", C.c_int64), ("n_slots", C.c_int64),
                ("run_launches", C.c_int64), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def load_library():
    """Loads libphmm_sm100.so; raises PhmmError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PhmmError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    L.phmm_version.restype = i32
    L.phmm_default_params.argtypes = [C.POINTER(Params)]
    L.phmm_create.restype = vp
    L.phmm_create.argtypes = [i32, vp, vp, i32]
    L.phmm_create_error.restype = C.c_char_p
    L.phmm_destroy.argtypes = [vp]
    L.phmm_last_error.restype = C.c_char_p
    L.phmm_last_error.argtypes = [vp]
    L.phmm_set_stream.argtypes = [vp, vp]
    L.phmm_set_model.argtypes = [vp, vp, vp, i32]
    L.phmm_set_reference.argtypes = [vp, vp, i64]
    batch_in = [vp, i64, vp, vp, vp, vp, vp, vp, C.POINTER(Params)]
    L.phmm_realign_batch.argtypes = batch_in + [C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.POINTER(C.c_int64)),
                                                C.POINTER(Posteriors)]
    L.phmm_expectations_batch.argtypes = batch_in + [vp]
    L.phmm_expectations_batch_fixed.argtypes = batch_in + [vp, vp]
    L.phmm_expectations_prepare.argtypes = batch_in
    L.phmm_expectations_run_fixed.argtypes = [vp, vp, vp]
    L.phmm_batch_prepare.argtypes = batch_in
    L.phmm_batch_run.argtypes = [vp]
    L.phmm_batch_fetch.argtypes = [vp, C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.POINTER(C.c_int64)),
                                   C.POINTER(Posteriors)]
    L.phmm_batch_get_stats.argtypes = [vp, C.POINTER(BatchStats)]
    L.phmm_base_expectations_reset.argtypes = [vp, i32]
    L.phmm_batch_add_base_expectations.argtypes = [vp, vp, i32]
    L.phmm_base_expectations_fetch.argtypes = [vp, i32, vp, i64]
    L.phmm_set_memory_budget.argtypes = [vp, i64]
    L.phmm_set_option.argtypes = [vp, C.c_char_p, i64]
    L.phmm_free.argtypes = [vp]
    L.phmm_free_posteriors.argtypes = [C.POINTER(Posteriors)]
    _lib = L
    return L


def default_params(**kw):
    """phmm_params initialised with what the reference passes (utils.py:587; abstractMapper.py:25)."""
    p = Params()
    load_library().phmm_default_params(C.byref(p))
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError("unknown parameter %s" % k)
        setattr(p, k, v)
    return p


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class PhmmContext:
    """One library context = one GPU.  trans/emis None selects the stock model."""

    def __init__(self, device=0, trans=None, emis=None, model_type=1):
        self._lib = load_library()
        self._h = None
        t, e = self._model_arrays(trans, emis)
        h = self._lib.phmm_create(int(device), _ptr(t) if t is not None else None,
                                  _ptr(e) if e is not None else None, int(model_type))
        if not h:
            raise PhmmError("phmm_create failed: %s" % self._lib.phmm_create_error().decode())
        self._h = h
        self.device = device

    @staticmethod
    def _model_arrays(trans, emis):
        if trans is None and emis is None:
            return None, None
        t = np.ascontiguousarray(trans, dtype=np.float64)
        e = np.ascontiguousarray(emis, dtype=np.float64)
        if t.size != 25 or e.size != 80:
            raise ValueError("trans must have 25 and emis 80 entries")
        return t, e

    def close(self):
        if self._h:
            self._lib.phmm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != PHMM_OK:
            raise PhmmError("libphmm error %d: %s" % (rc, self._lib.phmm_last_error(self._h).decode()))

    def set_model(self, trans=None, emis=None, model_type=1):
        t, e = self._model_arrays(trans, emis)
        self._check(self._lib.phmm_set_model(self._h, _ptr(t) if t is not None else None,
                                             _ptr(e) if e is not None else None, int(model_type)))

    def set_stream(self, cuda_stream):
        """cuda_stream: integer handle of a cudaStream_t (e.g. torch.cuda.Stream().cuda_stream); 0/None = own stream."""
        self._check(self._lib.phmm_set_stream(self._h, C.c_void_p(int(cuda_stream or 0))))

    def set_reference(self, codes):
        ref = np.ascontiguousarray(codes, dtype=np.uint8)
        self._check(self._lib.phmm_set_reference(self._h, _ptr(ref), ref.size))

    def set_option(self, name, value):
        """Library tuning / test switches, see phmm_set_option in include/phmm.h."""
        self._check(self._lib.phmm_set_option(self._h, name.encode(), int(value)))

    def set_memory_budget(self, nbytes):
        self._check(self._lib.phmm_set_memory_budget(self._h, int(nbytes)))

    @staticmethod
    def _batch_arrays(reads, read_off, ref_start, ref_end, in_ops, in_off):
        arrs = (np.ascontiguousarray(reads, dtype=np.uint8), np.ascontiguousarray(read_off, dtype=np.int64),
                np.ascontiguousarray(ref_start, dtype=np.int64), np.ascontiguousarray(ref_end, dtype=np.int64),
                np.ascontiguousarray(in_ops, dtype=np.uint32), np.ascontiguousarray(in_off, dtype=np.int64))
        n = arrs[1].size - 1
        if n < 0 or arrs[2].size != n or arrs[3].size != n or arrs[5].size != n + 1:
            raise ValueError("inconsistent batch arrays")
        return n, arrs

    def _take_outputs(self, n, ops_p, off_p, post):
        off = np.ctypeslib.as_array(off_p, shape=(n + 1,)).copy()
        tot = int(off[n])
        ops = np.ctypeslib.as_array(ops_p, shape=(max(tot, 1),))[:tot].copy()
        self._lib.phmm_free(ops_p)
        self._lib.phmm_free(off_p)
        pdict = None
        if post is not None:
            m = int(post.n)
            poff = np.ctypeslib.as_array(post.off, shape=(n + 1,)).copy()

            def grab(p):
                return np.ctypeslib.as_array(p, shape=(max(m, 1),))[:m].copy()
            pdict = {"off": poff, "ref_pos": grab(post.ref_pos), "read_pos": grab(post.read_pos),
                     "prob_1e7": grab(post.prob_1e7)}
            self._lib.phmm_free_posteriors(C.byref(post))
        return ops, off, pdict

    def realign_batch(self, reads, read_off, ref_start, ref_end, in_ops, in_off, params, want_posteriors=False):
        """Returns (ops uint32, off int64[n+1], posteriors dict or None)."""
        n, a = self._batch_arrays(reads, read_off, ref_start, ref_end, in_ops, in_off)
        ops_p = C.POINTER(C.c_uint32)()
        off_p = C.POINTER(C.c_int64)()
        post = Posteriors() if want_posteriors else None
        self._check(self._lib.phmm_realign_batch(self._h, n, *[_ptr(x) for x in a], C.byref(params), C.byref(ops_p),
                                                 C.byref(off_p), C.byref(post) if post is not None else None))
        return self._take_outputs(n, ops_p, off_p, post)

    def prepare(self, reads, read_off, ref_start, ref_end, in_ops, in_off, params):
        n, a = self._batch_arrays(reads, read_off, ref_start, ref_end, in_ops, in_off)
        self._n_prepared = n
        self._check(self._lib.phmm_batch_prepare(self._h, n, *[_ptr(x) for x in a], C.byref(params)))

    def run(self):
        self._check(self._lib.phmm_batch_run(self._h))

    def fetch(self, want_posteriors=False):
        ops_p = C.POINTER(C.c_uint32)()
        off_p = C.POINTER(C.c_int64)()
        post = Posteriors() if want_posteriors else None
        self._check(self._lib.phmm_batch_fetch(self._h, C.byref(ops_p), C.byref(off_p),
                                               C.byref(post) if post is not None else None))
        return self._take_outputs(self._n_prepared, ops_p, off_p, post)

    def stats(self):
        s = BatchStats()
        self._check(self._lib.phmm_batch_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def base_expectations_reset(self, n_tables=1):
        """Sizes and zeroes n_tables device tables of per-reference-position base expectations."""
        self._check(self._lib.phmm_base_expectations_reset(self._h, int(n_tables)))

    def add_base_expectations(self, read_mask=None, table=0):
        """Adds the posterior pairs of the prepared batch to a device table; read_mask: one byte per read."""
        m = None
        if read_mask is not None:
            m = np.ascontiguousarray(read_mask, dtype=np.uint8)
            if m.size != self._n_prepared:
                raise ValueError("read_mask must have one entry per read of the prepared batch")
        self._check(self._lib.phmm_batch_add_base_expectations(self._h, _ptr(m) if m is not None else None, int(table)))

    def base_expectations_fetch(self, ref_len, table=0):
        """-> int64[ref_len, 5] in units of 1e-7 (read base A, C, G, T, other)."""
        out = np.zeros((int(ref_len), 5), dtype=np.int64)
        self._check(self._lib.phmm_base_expectations_fetch(self._h, int(table), _ptr(out), out.size))
        return out

    def expectations_batch(self, reads, read_off, ref_start, ref_end, in_ops, in_off, params):
        """Returns float64[106]: 25 transition + 80 emission expectations + log-likelihood."""
        n, a = self._batch_arrays(reads, read_off, ref_start, ref_end, in_ops, in_off)
        out = np.zeros(106, dtype=np.float64)
        self._check(self._lib.phmm_expectations_batch(self._h, n, *[_ptr(x) for x in a], C.byref(params), _ptr(out)))
        return out

    def expectations_batch_fixed(self, reads, read_off, ref_start, ref_end, in_ops, in_off, params):
        """Returns (hi int64[106], lo int64[106]): exact sums, value = hi + lo / 2^32 (log-likelihood: / 2^20)."""
        n, a = self._batch_arrays(reads, read_off, ref_start, ref_end, in_ops, in_off)
        hi = np.zeros(106, dtype=np.int64)
        lo = np.zeros(106, dtype=np.int64)
        self._check(self._lib.phmm_expectations_batch_fixed(self._h, n, *[_ptr(x) for x in a], C.byref(params),
                                                            _ptr(hi), _ptr(lo)))
        return hi, lo

    def expectations_prepare(self, reads, read_off, ref_start, ref_end, in_ops, in_off, params):
        """Resident E-step: plans and uploads the batch once; expectations_run_fixed() then runs per EM iteration
        (set_model in between keeps the plan)."""
        n, a = self._batch_arrays(reads, read_off, ref_start, ref_end, in_ops, in_off)
        self._check(self._lib.phmm_expectations_prepare(self._h, n, *[_ptr(x) for x in a], C.byref(params)))

    def expectations_run_fixed(self):
        hi = np.zeros(106, dtype=np.int64)
        lo = np.zeros(106, dtype=np.int64)
        self._check(self._lib.phmm_expectations_run_fixed(self._h, _ptr(hi), _ptr(lo)))
        return hi, lo
