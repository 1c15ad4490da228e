# -*- coding: utf-8 -*-
"""Regularisation path with warm starts, and the trace files ``compare.py`` plots.

Neither exists in the reference: it always starts from ``x = 0`` with one fixed ``mu``
(lasso.py:34,89) and nothing in-tree writes the four text files ``compare.py`` loads
(compare.py:7-10).  SURVEY.md section 8(f) ranks both as the first steps after the hot
path, BASELINE.json config 5 (20 decreasing lambdas with warm starts) needs the former.
"""
import ctypes
import os

import numpy as np

from . import _lib
from . import distributed
from . import settings


def lambda_grid(mu_max, n=20, hi=0.9, lo=0.009):
    """``n`` log-spaced values from ``hi*mu_max`` down to ``lo*mu_max``."""
    if n == 1:
        return np.array([hi * mu_max])
    return mu_max * hi * (lo / hi) ** (np.arange(n) / (n - 1.0))


def lasso_path(gpu_cal, b, mus, BLOCK, ITER_MAX, ERR_BOUND=1e-4, collect_x=True):
    """Solve the lasso for every ``mu`` in ``mus`` (decreasing), each solve warm-started from
    the previous solution: the device keeps ``x`` and the running residual ``r = Ax - b``
    between solves, only the threshold changes.  Returns a list of dicts
    ``{mu, iters, stopped, objective, kernel_ms, x}`` (``x`` is ``(K,1)`` float64 or None).

    On several GPUs (``distributed.connect`` done) call it with the same arguments on every
    rank; ``x`` is then the local slice, ``objective`` the objective of the whole instance (the
    l1 term is all-reduced over the ranks)."""
    lib, ctx = gpu_cal._lib, gpu_cal.ctx
    K = gpu_cal.MAT_WIDTH_ALL
    if BLOCK != gpu_cal.Block:
        raise ValueError("BLOCK=%d does not match the GPU_Calculation (%d blocks)" % (BLOCK, gpu_cal.Block))
    bb = np.ascontiguousarray(b, dtype=np.float64).reshape(-1)
    _lib.check(lib.b200l_set_problem(ctx, _lib.dptr(bb)))          # x = 0, r = -b
    out = []
    steps, stopped, kms, obj = ctypes.c_int64(), ctypes.c_int32(), ctypes.c_double(), ctypes.c_double()
    for mu in mus:
        # warm start: x and r stay; the per-sweep stop counter and the cyclic order restart
        _lib.check(lib.b200l_restart_counters(ctx))
        _lib.check(lib.b200l_run(ctx, None, int(ITER_MAX), float(mu),
                                 float(ERR_BOUND) if isinstance(ERR_BOUND, float) else -1.0, None, None,
                                 ctypes.byref(steps), ctypes.byref(stopped), ctypes.byref(kms)))
        obj.value = distributed.objective(gpu_cal, mu)          # l1 term summed over the column shards
        x = None
        if collect_x:
            x = np.empty((K, 1), np.float64)
            _lib.check(lib.b200l_get_x(ctx, _lib.dptr(x)))
        out.append(dict(mu=float(mu), iters=int(steps.value), stopped=bool(stopped.value),
                        objective=float(obj.value), kernel_ms=float(kms.value), x=x))
    return out


def save_traces(prefix, time_iter, err_iter, iters, directory=None):
    """Write ``<prefix>_time.txt`` / ``<prefix>_errors.txt`` (prefix 'GPU' or 'CPU') into
    ``settings.Dir_PERFORMANCE`` -- the files compare.py:7-10 loads -- from the traces a
    ``run(err_iter=..., time_iter=...)`` filled.  Returns the two paths."""
    if prefix not in ("GPU", "CPU"):
        raise ValueError("prefix must be 'GPU' or 'CPU'")
    if directory is None:
        if settings.Dir_PERFORMANCE is None:
            settings.init()
        directory = settings.Dir_PERFORMANCE
    os.makedirs(directory, exist_ok=True)
    n = int(iters)
    t = np.asarray(time_iter, dtype=np.float64)[:n]          # time_iter[t]: start of iteration t (lasso.py:60-62)
    e = np.asarray(err_iter, dtype=np.float64)[:n]
    m = min(len(t), len(e))
    keep = e[:m] > 0                                          # compare.py takes log10
    pt = os.path.join(directory, prefix + "_time.txt")
    pe = os.path.join(directory, prefix + "_errors.txt")
    np.savetxt(pt, t[:m][keep])
    np.savetxt(pe, e[:m][keep])
    return pt, pe
