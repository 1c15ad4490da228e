import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def golden_names(small_only=False):
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    if small_only:
        names = [n for n in names if n.startswith("g_")]
    return names


_problem_cache = {}


def load_golden(name):
    """golden dict + the regenerated instance (A, b, mu) checked against the stored probes"""
    from oracle.lasso_oracle import make_problem
    if name in _problem_cache:
        return _problem_cache[name]
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    N, K = int(g["N"]), int(g["K"])
    A, x_true, b, mu = make_problem(N, K, float(g["den"]), int(g["seed"]))
    assert np.array_equal(A[:4, :8], g["A_probe"]), "instance regeneration drifted"
    assert np.array_equal(b, g["b"]) and mu == float(g["mu"])
    out = (g, A, b, mu)
    _problem_cache[name] = out
    return out


@pytest.fixture(scope="session")
def lib():
    from convex_optimization_b200 import _lib
    return _lib.load()


def assert_support(x, x_ref, TYPE):
    """The support bar of BASELINE.json's north star ("the nonzero pattern is checked exactly"):
    fp64 -- the nonzero patterns are identical; fp32 -- identical on every entry the fp32 tolerance
    can resolve: an index where the patterns differ must be negligible in BOTH solutions,
    |x_ref| < 1e-5 max|x_ref| and |x| < 1e-5 max|x_ref| (the reference keeps entries like -1.95e-10
    as "nonzero", which an fp32 matrix may legitimately round to an exact 0).  Enforced by every
    fp32 parity test (golden, ragged, transposed, bench geometries, multi-GPU)."""
    x = np.asarray(x).reshape(-1)
    x_ref = np.asarray(x_ref).reshape(-1)
    mism = (x != 0) != (x_ref != 0)
    if TYPE == "double":
        assert not mism.any(), "support differs at %d entries" % int(mism.sum())
        return 0
    scale = 1e-5 * np.abs(x_ref).max()
    bad = mism & ((np.abs(x_ref) >= scale) | (np.abs(x) >= scale))
    assert not bad.any(), "support differs at %d resolvable entries (max |x_ref| there %.3e)" % (
        int(bad.sum()), float(np.abs(x_ref[bad]).max()))
    return int(mism.sum())
