# -*- coding: utf-8 -*-
"""CPU oracle for the lasso hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

This is a NumPy float64 restatement of the reference's block proximal iteration
(kingold5/convex_optimization).  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product path (``convex_optimization_b200``) never does and fails loudly when
its CUDA library is missing.

Parity status: PINNED.  The reference ships no tests or golden vectors
(SURVEY.md section 4), so the oracle is pinned against outputs of the unmodified
reference ``ClassLassoCPU`` executed in the build container
(``oracle/run_reference.py`` -> ``tests/golden/*.npz``); ``tests/test_oracle.py``
re-checks the oracle against those fixtures on every run.

Each function cites the reference file:line it follows.
"""
import time

import numpy as np


# --------------------------------------------------------------------------
# helpers (reference: cpu_calculation.py)
# --------------------------------------------------------------------------
def soft_thresholding(u, thr):
    """sign(u) * max(|u| - thr, 0)            [cpu_calculation.py:5-6]"""
    return np.sign(u) * np.maximum(np.abs(u) - thr, 0)


def element_proj(v, lo, hi):
    """clip v onto [lo, hi]                    [cpu_calculation.py:10-11]"""
    return np.maximum(np.minimum(v, hi), lo)


def error_crit(g, x, mu):
    """|| g - clip(g - x, -mu, mu) ||_inf     [cpu_calculation.py:15-20]"""
    return np.max(np.abs(g - element_proj(g - x, -mu, mu)))


def diag_ata(A, BLOCK):
    """column squared norms as (BLOCK, w, 1)   [cpu_calculation.py:35-42]"""
    K = A.shape[1]
    return np.sum(np.square(A), axis=0).reshape(BLOCK, K // BLOCK, 1)


def objective(A, b, x, mu):
    """0.5 ||Ax-b||^2 + mu ||x||_1            [lasso.py:46-47]"""
    r = A @ x - b
    return 0.5 * float(np.sum(r * r)) + mu * float(np.sum(np.abs(x)))


# --------------------------------------------------------------------------
# problem recipe (reference: parameters.py:17-33)
# --------------------------------------------------------------------------
def make_problem(N, K, den, seed):
    """Gaussian A with unit-l2 rows, sparse x_true, b = A x_true + e,
    mu = 0.1 ||A^T b||_inf.  Same legacy global-RNG call order as
    parameters.py:17-33 so a pinned seed reproduces the reference's instance."""
    import scipy.sparse as sparse
    np.random.seed(int(seed))
    A = np.random.randn(N, K)
    A = A / np.linalg.norm(A, ord=2, axis=1, keepdims=True)
    x_true = sparse.random(K, 1, density=den, format="csc",
                           data_rvs=np.random.randn)
    e = np.random.normal(0, np.sqrt(1e-4), (N, 1))
    b = A @ x_true + e
    mu = 0.1 * np.max(np.abs(np.dot(np.transpose(A), b)))
    return A, x_true, np.asarray(b), float(mu)


# --------------------------------------------------------------------------
# the iteration (reference: lasso.py:102-157)
# --------------------------------------------------------------------------
def lasso_oracle(A, b, mu, BLOCK, ITER_MAX, ERR_BOUND=None, P=1, order=None,
                 faithful=True, x0=None, dtype=np.float64, time_limit=None):
    """Run the reference iteration.  Returns a dict with the final ``x`` (K,1),
    per-iteration ``err`` and ``gamma`` traces, ``iters`` (= t+1 of the last
    executed loop body, the reference's ``t+1`` in rlt_display), ``stopped``.

    faithful=True keeps the reference's per-block ``Ax`` array and rebuilds the
    residual with np.sum(Ax) - b every iteration (lasso.py:91,105,155); False
    keeps a running residual r += gamma*q (same iterates up to summation
    order) which is what the CUDA path does.
    P reproduces the worker split of block m (cpu_calculation.py:23-27): the
    block gradient is a concatenation of P slices, A_m d is a sum of P partial
    products (lasso.py:107-126).
    dtype=float32 gives the all-fp32 NumPy variant used for tolerance studies.
    """
    A = np.asarray(A, dtype=dtype)
    b = np.asarray(b, dtype=dtype).reshape(-1, 1)
    N, K = A.shape
    w = K // BLOCK
    assert w * BLOCK == K and w % P == 0
    mu = dtype(mu)
    d = np.sum(np.square(A), axis=0, dtype=dtype).reshape(BLOCK, w, 1)
    d_rec = [np.divide(dtype(1), d[i]) for i in range(BLOCK)]      # lasso.py:29-30
    bounded = isinstance(ERR_BOUND, float)                          # lasso.py:74-77

    x_block = np.zeros((BLOCK, w, 1), dtype=dtype)                  # lasso.py:89-90
    if x0 is not None:
        x_block[:] = np.asarray(x0, dtype=dtype).reshape(BLOCK, w, 1)
    if faithful:
        Ax = np.zeros((BLOCK, N, 1), dtype=dtype)                   # lasso.py:91
        if x0 is not None:
            for m in range(BLOCK):
                Ax[m] = A[:, m * w:(m + 1) * w] @ x_block[m]
    else:
        r = -b.copy()
        if x0 is not None:
            r = A @ x_block.reshape(K, 1) - b

    errs = np.zeros(ITER_MAX)
    gammas = np.zeros(ITER_MAX)
    block_cnt = 0
    gamma = dtype(0)
    stopped = False
    t = -1
    start = time.time()
    for t in range(ITER_MAX):
        m = (t % BLOCK) if order is None else int(order[t])         # lasso.py:104
        A_m = A[:, m * w:(m + 1) * w]
        s11 = (np.sum(Ax, axis=0) - b) if faithful else r           # lasso.py:105
        wp = w // P
        s13 = np.vstack([A_m[:, p * wp:(p + 1) * wp].T @ s11
                         for p in range(P)])                        # lasso.py:107-111
        rx = np.multiply(d[m], x_block[m]) - s13                    # lasso.py:114
        soft_t = soft_thresholding(rx, mu)                          # lasso.py:115
        Bx = np.multiply(d_rec[m], soft_t)                          # lasso.py:117
        dD = Bx - x_block[m]                                        # lasso.py:119
        s23 = np.sum([A_m[:, p * wp:(p + 1) * wp] @ dD[p * wp:(p + 1) * wp]
                      for p in range(P)], axis=0)                   # lasso.py:121-126
        r_1 = (s11.T @ s23).item() + mu * (np.linalg.norm(Bx, ord=1) -
                                           np.linalg.norm(x_block[m], ord=1))
        r_2 = (s23.T @ s23).item()                                  # lasso.py:129-132
        if r_2 != 0.0:                                              # lasso.py:133-136
            gamma = dtype(element_proj(-r_1 / r_2, 0, 1))
        err = float(error_crit(s13, x_block[m], mu))                # lasso.py:138-143
        errs[t] = err
        gammas[t] = gamma
        if bounded:                                                 # lasso.py:141-150
            if err < ERR_BOUND:
                block_cnt += 1
            if BLOCK - 1 == m:
                if block_cnt == BLOCK:
                    stopped = True
                    break
                block_cnt = 0
        x_block[m] += gamma * dD                                    # lasso.py:153
        if faithful:
            Ax[m] += gamma * s23                                    # lasso.py:155
        else:
            r = r + gamma * s23
        if time_limit is not None and time.time() - start > time_limit:
            break
    elapsed = time.time() - start
    x = x_block.reshape(K, 1).astype(np.float64)
    return dict(x=x, err=errs[:t + 1].copy(), gamma=gammas[:t + 1].copy(),
                iters=t + 1, stopped=stopped, elapsed=elapsed,
                objective=objective(np.asarray(A, np.float64),
                                    np.asarray(b, np.float64), x, float(mu)))


def sweep_bytes(N, K, BLOCK, itemsize):
    """Algorithmic HBM bytes of one sweep, SURVEY.md section 8(d):
    W = 2*N*K*s + 5*BLOCK*N*s + 4*K*s."""
    return 2 * N * K * itemsize + 5 * BLOCK * N * itemsize + 4 * K * itemsize
