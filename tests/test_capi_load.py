"""The C-ABI library loads without a GPU and exports exactly what include/phmm.h declares."""
import ctypes
import os
import re

import pytest

from nanopore_b200 import build as B
from nanopore_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "phmm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(phmm_[a-z_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    B.build()
    lib = ctypes.CDLL(capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), "libphmm_sm100.so does not export %s" % n
    assert sorted(capi.EXPORTS) == names


def test_version_and_defaults_need_no_gpu():
    L = capi.load_library()
    assert L.phmm_version() == 1
    p = capi.default_params()
    # what utils.py:587 and abstractMapper.py:25 pass
    assert (p.band, p.split_side, p.gap_gamma, p.match_gamma) == (10, 3000, 0.5, 0.0)
    assert (p.anchor_trim, p.min_diags, p.tb_diags, p.threshold) == (14, 1000, 40, 0.01)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.PhmmError) as e:
        capi.PhmmContext(0)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_product_never_imports_the_oracle():
    """The package, the C ABI headers and the command-line scripts never touch oracle/: only tests/ (incl. tests/tools),
    __graft_entry__.smoke() and the CPU legs of bench.py do."""
    for top in ("nanopore_b200", "scripts", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".sh")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                    assert "phmm_oracle" not in txt and "libphmm_oracle" not in txt, f
