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


def parameters_device(N, K, BLOCK, den, seed, gpu_cal_cls=None, group=None, SILENCE=True):
    """The recipe of ``parameters()`` (parameters.py:20-33) for instances that never exist on the
    host: ``A`` is generated on the device by ``b200l_gen_gaussian``
    (counter-based Philox keyed by (seed, row, global column), so every column sharding yields
    the same matrix), rows are scaled to unit l2 norm, ``b = A x_true + e`` and
    ``mu = 0.1 |A^T b|_inf`` come from the library's mat-vecs.  ``K`` is the GLOBAL column count;
    with an initialised ``torch.distributed`` group every rank builds its column slice of every
    block (``distributed.local_columns``) and the row norms, ``b`` and ``mu`` are all-reduced.
    ``x_true`` and the noise are drawn on the host from ``RandomState(seed + 1)`` (identical on
    all ranks; not the reference's ``sparse.random`` stream).
    Returns ``(gpu_cal, x_true (K,1), b (N,1), mu)``."""
    import torch
    from . import _lib
    from .gpu_calculation import GPU_Calculation
    cls = gpu_cal_cls or GPU_Calculation
    dist = None
    rank, world = 0, 1
    try:
        import torch.distributed as tdist
        if tdist.is_available() and tdist.is_initialized():
            dist = tdist
            rank, world = dist.get_rank(group), dist.get_world_size(group)
    except Exception:
        dist = None
    if K % (BLOCK * world):
        raise ValueError("K=%d is not divisible by BLOCK*world=%d" % (K, BLOCK * world))
    Kl = K // world
    wl = Kl // BLOCK
    device = torch.device("cuda", cls.DEVICE)
    tdt = torch.float64 if cls.TYPE == 'double' else torch.float32
    ld = cls.padded_ld(N, Kl, BLOCK)
    rows = N if cls.LAYOUT == 'row' else wl
    store = torch.zeros((BLOCK, rows, ld), dtype=tdt, device=device)
    cal = cls.from_device_blocks(store, N, Kl, BLOCK)
    lib, ctx = cal._lib, cal.ctx

    def allreduce(a, op):
        if dist is None:
            return a
        t = torch.from_numpy(a)
        if dist.get_backend(group) == "nccl":
            t = t.to(device)
        dist.all_reduce(t, op=op, group=group)
        return t.cpu().numpy()

    _lib.check(lib.b200l_gen_gaussian(ctx, int(seed) & 0xFFFFFFFFFFFFFFFF, rank, world))
    ss = np.empty(N)
    _lib.check(lib.b200l_row_sumsq(ctx, _lib.dptr(ss)))
    ss = allreduce(ss, None if dist is None else dist.ReduceOp.SUM)
    scale = np.ascontiguousarray(1.0 / np.sqrt(ss))
    _lib.check(lib.b200l_scale_rows(ctx, _lib.dptr(scale)))

    rs = np.random.RandomState(int(seed) + 1)
    mask = rs.rand(K) < den
    x_true = (rs.randn(K) * mask)[:, np.newaxis]
    noise = rs.normal(0, np.sqrt(1e-4), N)
    w = K // BLOCK                                       # global block width
    b = np.zeros(N)
    q = np.empty(N)
    for m in range(BLOCK):
        xm = np.ascontiguousarray(x_true[m * w + rank * wl:m * w + (rank + 1) * wl, 0])
        _lib.check(lib.b200l_gemv_n(ctx, m, _lib.dptr(xm), _lib.dptr(q)))
        b += q
    b = allreduce(b, None if dist is None else dist.ReduceOp.SUM) + noise
    b = np.ascontiguousarray(b)
    g = np.empty(wl)
    gmax = np.zeros(1)
    for m in range(BLOCK):
        _lib.check(lib.b200l_gemv_t(ctx, m, _lib.dptr(b), _lib.dptr(g)))
        gmax[0] = max(gmax[0], float(np.max(np.abs(g))))
    gmax = allreduce(gmax, None if dist is None else dist.ReduceOp.MAX)
    mu = 0.1 * float(gmax[0])
    if not SILENCE and rank == 0:
        print("Parameters @@created on the device with N: %d, K: %d, DENSITY: %f, mu: %f." % (N, K, den, mu))
    return cal, x_true, b[:, np.newaxis], mu
