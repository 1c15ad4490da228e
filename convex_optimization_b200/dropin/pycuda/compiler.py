# -*- coding: utf-8 -*-
"""``pycuda.compiler.SourceModule`` (gpu_calculation.py:5,139): the reference compiles its
hand kernels from a Jinja2 string at import.  Those kernels are what libb200lasso.so
replaces, so there is nothing to compile; constructing a SourceModule is refused."""


class SourceModule:
    def __init__(self, *args, **kwargs):
        raise RuntimeError("pycuda shim: SourceModule is not available; the reference's hand kernels "
                           "(gpu_calculation.py:9-138) are replaced by convex_optimization_b200.gpu_calculation")
