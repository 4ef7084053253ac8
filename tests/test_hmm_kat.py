"""The one numeric path the reference's own files pin (rows a9-a11):
blasr_hmm_{20,40}.txt == scripts/modifyHmm.py(blasr_hmm_0.txt, gc=0.5, rate=0.2/0.4).

tests/golden/blasr_hmm_*.txt are byte copies of reference nanopore/mappers/blasr_hmm_*.txt (data
fixtures, 2 lines each)."""
import os

import numpy as np
import pytest

from nanopore_b200 import hmm as H


def _load(golden_dir, name):
    return H.Hmm.loadHmm(os.path.join(golden_dir, name))


@pytest.mark.parametrize("name,rate", [("blasr_hmm_20.txt", 0.2), ("blasr_hmm_40.txt", 0.4)])
def test_modify_hmm_reproduces_reference_files(golden_dir, name, rate):
    h = _load(golden_dir, "blasr_hmm_0.txt")
    want = _load(golden_dir, name)
    # scripts/modifyHmm.py:17,22
    H.normaliseHmmByReferenceGCContent(h, 0.5)
    H.modifyHmmEmissionsByExpectedVariationRate(h, rate)
    assert h.type == want.type == 1 and h.stateNumber == 5
    assert np.array_equal(np.array(h.transitions), np.array(want.transitions))
    assert np.abs(np.array(h.emissions) - np.array(want.emissions)).max() < 1e-12
    assert h.likelihood == want.likelihood


def test_file_invariants(golden_dir):
    for name in ("blasr_hmm_0.txt", "blasr_hmm_20.txt", "blasr_hmm_40.txt"):
        h = _load(golden_dir, name)
        t = np.array(h.transitions).reshape(5, 5)
        e = np.array(h.emissions).reshape(5, 4, 4)
        assert np.allclose(t.sum(axis=1), 1.0, atol=1e-9)
        assert np.allclose(e.sum(axis=(1, 2)), 1.0, atol=1e-9)
        assert np.allclose(e[0].sum(axis=1), 0.25, atol=1e-9)       # GC 0.5 normalisation (utils.py:537,619)
        assert np.allclose(e[1:], 0.0625)                           # setHmmIndelEmissionsToBeFlat (utils.py:626-629)
        # asymmetric model: no shortGapX <-> shortGapY switch
        assert t[1, 2] == 0.0 and t[2, 1] == 0.0


def test_write_load_round_trip(tmp_path, golden_dir):
    h = _load(golden_dir, "blasr_hmm_0.txt")
    p = tmp_path / "out.hmm"
    h.write(str(p))
    g = H.Hmm.loadHmm(str(p))
    assert g.type == h.type and g.transitions == h.transitions and g.emissions == h.emissions
    assert g.likelihood == h.likelihood
    assert len(open(str(p)).read().splitlines()) == 2


def test_flat_indels_and_normalise():
    h = H.Hmm("fiveStateAsymmetric")
    rng = np.random.default_rng(0)
    h.transitions = rng.random(25).tolist()
    h.emissions = rng.random(80).tolist()
    h.normalise()
    assert np.allclose(np.array(h.transitions).reshape(5, 5).sum(axis=1), 1.0)
    assert np.allclose(np.array(h.emissions).reshape(5, 16).sum(axis=1), 1.0)
    H.setHmmIndelEmissionsToBeFlat(h)
    assert h.emissions[16:] == [1.0 / 16] * 64
    before = list(h.emissions)
    H.normaliseHmmByReferenceGCContent(h, 0.4)
    e = np.array(h.emissions).reshape(5, 4, 4)
    assert np.allclose(e[0].sum(axis=1), [0.3, 0.2, 0.2, 0.3])
    assert h.emissions[32:48] == before[32:48] and h.emissions[64:] == before[64:]   # insert states untouched


def test_loader_rejects_wrong_sizes(tmp_path):
    p = tmp_path / "bad.hmm"
    p.write_text("1 0.5 0.5 0.0\n0.1 0.2\n")
    with pytest.raises(RuntimeError):
        H.Hmm.loadHmm(str(p))


@pytest.mark.parametrize("name,rate", [("blasr_hmm_20.txt", "0.2"), ("blasr_hmm_40.txt", "0.4")])
def test_modify_hmm_command_line(golden_dir, tmp_path, name, rate):
    """scripts/modifyHmm.py IN GC RATE OUT (reference scripts/modifyHmm.py:7-30) regenerates the reference's files."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = tmp_path / "out.hmm"
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "modifyHmm.py"),
                        os.path.join(golden_dir, "blasr_hmm_0.txt"), "0.5", rate, str(out)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "Got GC content 0.5" in r.stdout and "For state, ref frequencies" in r.stdout
    got, want = H.Hmm.loadHmm(str(out)), _load(golden_dir, name)
    assert got.transitions == want.transitions and got.likelihood == want.likelihood
    assert np.abs(np.array(got.emissions) - np.array(want.emissions)).max() < 1e-12
    # wrong argument count: usage, non-zero exit
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "modifyHmm.py")], capture_output=True, text=True)
    assert r.returncode == 2
