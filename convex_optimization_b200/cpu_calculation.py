# -*- coding: utf-8 -*-
"""Host-side (NumPy) helpers of the lasso iteration.

Same eight entry points as the reference's cpu_calculation.py (:5-50), NumPy in /
NumPy out.  They define the prox / error semantics the CUDA kernels are checked against
and are used by the step-wise solver loop in lasso.py and by ``ClassLassoCPU``.
"""
import numpy as np


def soft_thresholding(tensor, threshold):
    """S_thr(u) = sign(u) * max(|u| - thr, 0), elementwise (ref cpu_calculation.py:5-6)."""
    mag = np.abs(tensor) - threshold
    return np.sign(tensor) * np.where(mag > 0, mag, 0.0)


def element_proj(vec, lower_bound, upper_bound):
    """Project onto the box [lower_bound, upper_bound] (ref cpu_calculation.py:10-11)."""
    return np.clip(vec, lower_bound, upper_bound)


def error_crit(grad_fx, x, mu):
    """Optimality measure ||g - P_[-mu,mu](g - x)||_inf (ref cpu_calculation.py:15-20)."""
    resid = grad_fx - element_proj(grad_fx - x, -mu, mu)
    return np.abs(resid).max()


def A_bp_get(A, BLOCK, P):
    """(BLOCK, P, N, w/P) view of A: block m, worker p -> its column slice
    (ref cpu_calculation.py:23-27).  K must be divisible by BLOCK*P."""
    N, K = A.shape
    if K % (BLOCK * P) != 0:
        raise ValueError("K=%d is not divisible by BLOCK*P=%d" % (K, BLOCK * P))
    return A.T.reshape(BLOCK, P, K // (BLOCK * P), N).swapaxes(2, 3)


def fun_s12(A_bp, s11):
    """Slice of the block gradient: A_bp^T s11 (ref cpu_calculation.py:30-31)."""
    return A_bp.T @ s11


def fun_diag_ATA(A_bp):
    """Column squared norms per block, shape (BLOCK, w, 1) (ref cpu_calculation.py:35-42)."""
    BLOCK, P, N, wp = A_bp.shape
    sq = np.einsum('bpnk,bpnk->bpk', A_bp, A_bp)
    return sq.reshape(BLOCK, P * wp, 1)


def fun_s22(A_bp, s21):
    """Partial product A_bp s21 (ref cpu_calculation.py:45-46)."""
    return A_bp @ s21


def fun_dd_p(P, descent_d):
    """Split the direction into the P worker slices (ref cpu_calculation.py:49-50)."""
    return descent_d.reshape(P, -1, 1)
