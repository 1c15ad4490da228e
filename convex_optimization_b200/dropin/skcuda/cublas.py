# -*- coding: utf-8 -*-
"""``skcuda.cublas`` names used by the reference, on the B200 library.

* ``cublasCreate`` / ``cublasDestroy`` (cpu_vs_gpu.py:93,202): the handle is a plain object;
  the library needs none.
* ``cublasDgemv`` (lasso.py:350-353,405-408,416-419): when the matrix argument is a column
  block of a ``GPU_Calculation`` (``gpu_cal.A_b_gpu[m].gpudata``) the call is the library's
  own mat-vec on device vectors (``b200l_gemv_t_dev`` / ``b200l_gemv_n_dev``: the fp32 / fp64,
  row-major / pre-transposed block as it lies in HBM).  cuBLAS is column-major, so for the
  row-major (N, w) block op 'N' is ``A_m^T v`` and op 'T' is ``A_m v`` (lasso.py:336,342).
  Any other fp64 matrix goes through a torch mat-vec (a plain library GEMV, off the hot path).
* the level-1 routines of the reference's "pure cuBLAS" class (lasso.py:424-454,562-576):
  ``cublasDcopy/Daxpy/Dscal/Ddot/Dnrm2/Dasum`` on torch views of the device vectors.
"""
import ctypes

import numpy as np

_CUBLAS_OP = {0: 0, 'n': 0, 'N': 0, 1: 1, 't': 1, 'T': 1, 2: 2, 'c': 2, 'C': 2}


class _Handle:
    def __init__(self):
        self.alive = True


def cublasCreate():
    return _Handle()


def cublasDestroy(handle):
    if isinstance(handle, _Handle):
        handle.alive = False


def cublasSetStream(handle, stream):
    pass


def cublasGetVersion(handle=None):
    return 0


class _Span:
    """__cuda_array_interface__ view of n strided doubles at a raw device address"""

    def __init__(self, addr, n, inc=1, keep=None):
        self.__cuda_array_interface__ = {
            "shape": (int(n),), "typestr": "<f8", "data": (int(addr), False), "version": 2,
            "strides": None if inc == 1 else (8 * int(inc),)}
        self._keep = keep


def _vec(addr, n, inc=1):
    import torch
    return torch.as_tensor(_Span(addr, n, inc, getattr(addr, "_owner", None)), device="cuda")


def cublasDcopy(handle, n, x, incx, y, incy):
    _vec(y, n, incy).copy_(_vec(x, n, incx))


def cublasDaxpy(handle, n, alpha, x, incx, y, incy):
    _vec(y, n, incy).add_(_vec(x, n, incx), alpha=float(alpha))


def cublasDscal(handle, n, alpha, x, incx):
    _vec(x, n, incx).mul_(float(alpha))


def cublasDdot(handle, n, x, incx, y, incy):
    return float(_vec(x, n, incx).dot(_vec(y, n, incy)).item())


def cublasDnrm2(handle, n, x, incx):
    return float(_vec(x, n, incx).norm().item())


def cublasDasum(handle, n, x, incx):
    return float(_vec(x, n, incx).abs().sum().item())


def cublasDgemv(handle, trans, m, n, alpha, A, lda, x, incx, beta, y, incy):
    """y = alpha * op(A) x + beta * y, A column-major (m, n) with leading dimension lda."""
    import torch
    op = _CUBLAS_OP[trans] if not isinstance(trans, int) else trans
    ny, nx = (m, n) if op == 0 else (n, m)
    block = getattr(A, "block", None)
    if block is not None and incx == 1 and incy == 1:
        gpu_cal, idx = block
        from convex_optimization_b200 import _lib
        if (m, n) != (gpu_cal.MAT_WIDTH, gpu_cal.MAT_HEIGHT):
            raise ValueError("cublasDgemv shim: block is (%d, %d) column-major, got m=%d n=%d"
                             % (gpu_cal.MAT_WIDTH, gpu_cal.MAT_HEIGHT, m, n))
        plain = float(alpha) == 1.0 and float(beta) == 0.0
        out = y if plain else torch.empty(ny, dtype=torch.float64, device="cuda").data_ptr()
        keep = None
        if not plain:
            keep = torch.empty(ny, dtype=torch.float64, device="cuda")
            out = keep.data_ptr()
        fn = gpu_cal._lib.b200l_gemv_t_dev if op == 0 else gpu_cal._lib.b200l_gemv_n_dev
        _lib.check(fn(gpu_cal.ctx, int(idx), ctypes.c_void_p(int(x)), ctypes.c_void_p(int(out))))
        if not plain:
            yv = _vec(y, ny)
            yv.mul_(float(beta)).add_(keep, alpha=float(alpha))
        return
    # a plain fp64 column-major matrix that is not a block of the solver's A
    lda = int(getattr(A, "ld", None) or lda)
    Acm = torch.as_tensor(_MatSpan(A, m, n, lda), device="cuda")          # (n, lda) row-major = A^T padded
    At = Acm[:, :m]                                                       # (n, m) = A^T
    xv, yv = _vec(x, nx, incx), _vec(y, ny, incy)
    res = (At.t() @ xv) if op == 0 else (At @ xv)
    if float(beta) == 0.0:
        yv.copy_(res * float(alpha))
    else:
        yv.mul_(float(beta)).add_(res, alpha=float(alpha))


class _MatSpan:
    def __init__(self, addr, m, n, lda):
        self.__cuda_array_interface__ = {"shape": (int(n), int(lda)), "typestr": "<f8",
                                         "data": (int(addr), False), "version": 2, "strides": None}
        self._keep = getattr(addr, "_owner", None)
