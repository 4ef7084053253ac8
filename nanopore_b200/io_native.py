"""ctypes binding of libphmm_io.so (include/phmm_io.h): native ingest / chain / pack / emit.

Host C++ with threads; no GPU needed.  `NativeIo` mirrors the file-level steps of the reference's realignment
driver (nanopore/analyses/utils.py:441-469 chainSamFile, :557-574 the per-read job inputs, :591-609 the fan-in)
on whole files and yields the packed `Batch` the realigner takes.  The pure-Python functions of
nanopore_b200/realign.py remain the readable restatement (and serve custom chain functions); tests compare the two
byte for byte.
"""
import ctypes as C
import os

import numpy as np

from .batch import Batch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "libphmm_io.so")

SYMBOLS = ["phmm_io_version", "phmm_io_create", "phmm_io_destroy", "phmm_io_last_error", "phmm_io_load_reference",
           "phmm_io_load_reads", "phmm_io_chain_sam", "phmm_io_load_sam", "phmm_io_counts", "phmm_io_batch_view",
           "phmm_io_write_sam", "phmm_io_write_realigned_sam", "phmm_io_estimate_cells", "phmm_io_gather_ranges"]


class IoBatch(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("ref", C.POINTER(C.c_uint8)), ("ref_len", C.c_int64),
                ("reads", C.POINTER(C.c_uint8)), ("read_off", C.POINTER(C.c_int64)),
                ("ref_start", C.POINTER(C.c_int64)), ("ref_end", C.POINTER(C.c_int64)),
                ("ops", C.POINTER(C.c_uint32)), ("ops_off", C.POINTER(C.c_int64))]


class PhmmIoError(RuntimeError):
    pass


_lib = None


def available():
    return os.path.exists(LIB)


def load_library():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            raise PhmmIoError("%s is missing: run `python -m nanopore_b200.build`" % LIB)
        L = C.CDLL(LIB)
        for s in SYMBOLS:
            getattr(L, s)
        L.phmm_io_create.restype = C.c_void_p
        L.phmm_io_create.argtypes = [C.c_int]
        L.phmm_io_destroy.argtypes = [C.c_void_p]
        L.phmm_io_last_error.restype = C.c_char_p
        L.phmm_io_last_error.argtypes = [C.c_void_p]
        for s in ("phmm_io_load_reference", "phmm_io_load_reads", "phmm_io_chain_sam", "phmm_io_load_sam", "phmm_io_write_sam"):
            getattr(L, s).argtypes = [C.c_void_p, C.c_char_p]
        L.phmm_io_counts.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
        L.phmm_io_batch_view.argtypes = [C.c_void_p, C.POINTER(IoBatch)]
        L.phmm_io_write_realigned_sam.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int64]
        L.phmm_io_estimate_cells.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                             C.c_int64, C.c_int, C.c_void_p]
        L.phmm_io_gather_ranges.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def estimate_cells(in_ops, in_off, read_off, ref_start, ref_end, band, anchor_trim, split_side, threads=0):
    """int64[n] estimated DP cells per read (phmm_io_estimate_cells)."""
    n = len(in_off) - 1
    out = np.zeros(n, dtype=np.int64)
    a = [np.ascontiguousarray(x, dtype=dt) for x, dt in ((in_ops, np.uint32), (in_off, np.int64), (read_off, np.int64),
                                                         (ref_start, np.int64), (ref_end, np.int64))]
    rc = load_library().phmm_io_estimate_cells(n, *[_vp(x) for x in a], int(band), int(anchor_trim), int(split_side), int(threads), _vp(out))
    if rc != 0:
        raise PhmmIoError("phmm_io_estimate_cells failed (%d)" % rc)
    return out


def gather_ranges(data, off, idx, threads=0, out=None):
    """(rows idx of the ragged array (data, off) concatenated, new offsets) via phmm_io_gather_ranges.  out: optional
    destination (contiguous, same dtype, exactly the gathered size) -- e.g. a view into a staging buffer."""
    data = np.ascontiguousarray(data)
    off = np.ascontiguousarray(off, dtype=np.int64)
    idx = np.ascontiguousarray(idx, dtype=np.int64)
    lens = off[idx + 1] - off[idx] if len(idx) else np.zeros(0, dtype=np.int64)
    new_off = np.concatenate(([0], np.cumsum(lens))).astype(np.int64)
    if out is None:
        out = np.empty(int(new_off[-1]), dtype=data.dtype)
    elif out.dtype != data.dtype or out.size != int(new_off[-1]) or not out.flags.c_contiguous:
        raise ValueError("out must be a contiguous %s array of %d elements" % (data.dtype, int(new_off[-1])))
    rc = load_library().phmm_io_gather_ranges(_vp(data), _vp(off), _vp(idx), len(idx), data.dtype.itemsize, int(threads), _vp(out), _vp(new_off))
    if rc != 0:
        raise PhmmIoError("phmm_io_gather_ranges failed (%d)" % rc)
    return out, new_off


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,))


class NativeIo:
    """One handle = one experiment's reference, reads and SAM records."""

    def __init__(self, threads=0):
        self._lib = load_library()
        self._h = self._lib.phmm_io_create(int(threads))
        if not self._h:
            raise PhmmIoError("phmm_io_create failed")

    def close(self):
        if self._h:
            self._lib.phmm_io_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise PhmmIoError(self._lib.phmm_io_last_error(self._h).decode("utf-8", "replace"))

    def load_reference(self, fasta):
        self._check(self._lib.phmm_io_load_reference(self._h, os.fsencode(fasta)))

    def load_reads(self, fastq):
        self._check(self._lib.phmm_io_load_reads(self._h, os.fsencode(fastq)))

    def chain_sam(self, sam):
        self._check(self._lib.phmm_io_chain_sam(self._h, os.fsencode(sam)))

    def load_sam(self, sam):
        self._check(self._lib.phmm_io_load_sam(self._h, os.fsencode(sam)))

    def counts(self):
        a, b = C.c_int64(0), C.c_int64(0)
        self._check(self._lib.phmm_io_counts(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def batch(self):
        """Packed Batch of the mapped records.  The arrays are views into the handle (valid until the next load)."""
        v = IoBatch()
        self._check(self._lib.phmm_io_batch_view(self._h, C.byref(v)))
        n = v.n_reads
        read_off = _arr(v.read_off, n + 1, np.int64)
        ops_off = _arr(v.ops_off, n + 1, np.int64)
        b = Batch(_arr(v.ref, v.ref_len, np.uint8), _arr(v.reads, int(read_off[-1]), np.uint8), read_off,
                  _arr(v.ref_start, n, np.int64), _arr(v.ref_end, n, np.int64),
                  _arr(v.ops, int(ops_off[-1]), np.uint32), ops_off)
        b._owner = self
        return b

    def write_sam(self, path):
        self._check(self._lib.phmm_io_write_sam(self._h, os.fsencode(path)))

    def write_realigned_sam(self, path, ops, off):
        ops = np.ascontiguousarray(ops, dtype=np.uint32)
        off = np.ascontiguousarray(off, dtype=np.int64)
        self._check(self._lib.phmm_io_write_realigned_sam(self._h, os.fsencode(path), ops.ctypes.data_as(C.c_void_p),
                                                          off.ctypes.data_as(C.c_void_p), len(off) - 1))
