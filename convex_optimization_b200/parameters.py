# -*- coding: utf-8 -*-
"""Problem generator, mirror of the reference's parameters.py:13-69.

``parameters(N, K, den, SAVE_FLAG, READ_FLAG, SILENCE=False)`` returns
``(A, x_true, b, mu)``: dense Gaussian A with unit-l2 rows, a sparse x_true (csc),
b = A x_true + e with e ~ N(0, 1e-4), and mu = 0.1 ||A^T b||_inf.  The legacy global
NumPy stream is used in the reference's call order so a pinned seed gives the
reference's instance.  ``seed=None`` keeps the reference behaviour (seed = int(time())).
"""
import os
from time import time

import numpy as np
import scipy.sparse as sparse

from . import settings

_FILES = ("A_matrix.txt", "x_true.txt", "b_vector.txt", "parameters.txt")


def _store_dir():
    if settings.HOME is None:
        settings.init()
    return os.path.join(settings.HOME, "Documents", "python")


def parameters(N, K, den, SAVE_FLAG, READ_FLAG, SILENCE=False, seed=None):
    if READ_FLAG:
        base = _store_dir()
        A = np.loadtxt(os.path.join(base, _FILES[0]), delimiter=",")
        x_true = np.loadtxt(os.path.join(base, _FILES[1]))[:, np.newaxis]
        b = np.loadtxt(os.path.join(base, _FILES[2]))[:, np.newaxis]
        N, K, den, mu = np.loadtxt(os.path.join(base, _FILES[3]))
        if not SILENCE:
            print("Parameters @@loaded with N: %d, K: %d, DENSITY: %f, mu: %f." % (N, K, den, mu))
    else:
        np.random.seed(int(time()) if seed is None else int(seed))
        A = np.random.randn(N, K)
        A /= np.linalg.norm(A, ord=2, axis=1, keepdims=True)
        x_true = sparse.random(K, 1, density=den, format="csc", data_rvs=np.random.randn)
        noise = np.random.normal(0, np.sqrt(1e-4), (N, 1))
        b = A @ x_true + noise
        mu = 0.1 * np.max(np.abs(A.T @ b))
        if not SILENCE:
            print("Parameters @@created with N: %d, K: %d, DENSITY: %f, mu: %f." % (N, K, den, mu))

    if SAVE_FLAG:
        base = _store_dir()
        os.makedirs(base, exist_ok=True)
        np.savetxt(os.path.join(base, _FILES[0]), A, delimiter=",")
        xt = x_true.todense() if sparse.issparse(x_true) else x_true
        np.savetxt(os.path.join(base, _FILES[1]), xt)
        np.savetxt(os.path.join(base, _FILES[2]), b)
        np.savetxt(os.path.join(base, _FILES[3]), [N, K, den, mu])
        if not SILENCE:
            print("Paramenters @@saved!")
    return (A, x_true, b, mu)
