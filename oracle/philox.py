# -*- coding: utf-8 -*-
"""TEST INFRASTRUCTURE ONLY (see oracle/lasso_oracle.py): NumPy restatement of the device
instance generator ``b200l_gen_gaussian`` (include/b200lasso.h): Philox4x32-10 (Salmon et al.,
SC'11; the published round function and constants), counter = (column/4, row, column>>34,
0x4c41534f), key = seed, Box-Muller on 24-bit uniforms.  The reference itself draws A with
``np.random.randn`` (parameters.py:21); there is nothing of this in the reference."""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    c0, c1, c2, c3 = (np.asarray(v, dtype=np.uint64) & MASK for v in (c0, c1, c2, c3))
    k0, k1 = int(k0) & 0xFFFFFFFF, int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        n0 = (p1 >> np.uint64(32)) ^ c1 ^ np.uint64(k0)
        n1 = p1 & MASK
        n2 = (p0 >> np.uint64(32)) ^ c3 ^ np.uint64(k1)
        n3 = p0 & MASK
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def gauss_matrix(seed, N, gcols):
    """(N, len(gcols)) float32 entries of the generator for rows 0..N-1 and the given GLOBAL columns"""
    gcols = np.asarray(gcols, dtype=np.uint64)
    rows = np.arange(N, dtype=np.uint64)[:, None]
    grp = (gcols >> np.uint64(2))[None, :] + np.zeros_like(rows)
    hi = (gcols >> np.uint64(34))[None, :] + np.zeros_like(rows)
    x = philox4x32_10(grp & MASK, rows + np.zeros_like(grp), hi, np.uint64(0x4c41534f) + np.zeros_like(grp),
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    e = (gcols & np.uint64(3))[None, :] + np.zeros_like(rows)
    second = (e & np.uint64(2)) != 0
    a = np.where(second, x[2], x[0])
    b = np.where(second, x[3], x[1])
    u1 = ((a >> np.uint64(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    u2 = ((b >> np.uint64(8)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 16777216.0)
    rad = np.sqrt(np.float32(-2.0) * np.log(u1)).astype(np.float32)
    ang = (np.float32(2.0) * u2).astype(np.float64) * np.pi
    return np.where((e & np.uint64(1)) != 0, rad * np.sin(ang), rad * np.cos(ang)).astype(np.float32)
