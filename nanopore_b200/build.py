"""In-tree build of libphmm_sm100.so (sm_100a only; nvcc cross-compiles without a GPU)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libphmm_sm100.so")
SOURCES = ["phmm_api.cu"]
HEADERS = ["phmm_device.cuh", "phmm_kernels.cuh", "phmm_fb2.cuh", "phmm_decode_w.cuh", os.path.join("..", "..", "include", "phmm.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "--fmad=false", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


IO_LIB = os.path.join(_HERE, "libphmm_io.so")
IO_SOURCES = ["phmm_io.cpp"]
IO_HEADERS = [os.path.join("..", "..", "include", "phmm_io.h")]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-shared", "-pthread", "-Wall", "-Wextra"]


def build_io(force=False):
    """Compiles the host-only ingest / emit library (g++; no CUDA)."""
    if not force and os.path.exists(IO_LIB):
        t = os.path.getmtime(IO_LIB)
        if not any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in IO_SOURCES + IO_HEADERS):
            return IO_LIB
    cxx = os.environ.get("CXX", "g++")
    subprocess.check_call([cxx] + CXX_FLAGS + ["-o", IO_LIB] + IO_SOURCES, cwd=CSRC)
    return IO_LIB


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False):
    """Compiles the CUDA library (and the host-only IO library) in-tree. Returns the path of the CUDA .so."""
    build_io(force)
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + SOURCES
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB


def build_tune(force=False):
    """Tuning build (-DPHMM_TUNE: the timing_experiment switches exist) under build/; scripts/tune.py loads it
    explicitly, nanopore_b200 never does."""
    out_dir = os.path.join(os.path.dirname(_HERE), "build")
    os.makedirs(out_dir, exist_ok=True)
    out = os.path.join(out_dir, "libphmm_tune.so")
    if not force and os.path.exists(out):
        t = os.path.getmtime(out)
        if not any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS):
            return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    subprocess.check_call([nvcc] + NVCC_FLAGS + ["-DPHMM_TUNE", "-o", out] + SOURCES, cwd=CSRC)
    return out


if __name__ == "__main__":
    print(build(force=True, verbose=True))
