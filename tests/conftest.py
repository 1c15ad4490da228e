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
