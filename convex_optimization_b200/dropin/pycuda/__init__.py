# -*- coding: utf-8 -*-
"""Shim namespace ``pycuda``: the handful of PyCUDA names the
reference's drivers and solver classes touch (lasso.py:13-16, cpu_vs_gpu.py:9,201,
gpu_calculation.py:3-6), on torch device memory.  Not a PyCUDA re-implementation."""
VERSION = (0, 0, 0)
VERSION_TEXT = "b200lasso-shim"
