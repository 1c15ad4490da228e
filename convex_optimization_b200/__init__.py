# -*- coding: utf-8 -*-
"""B200-native lasso hot path with the entry points of kingold5/convex_optimization.

Modules mirror the reference's flat files: ``lasso`` (solver classes),
``gpu_calculation`` (``GPU_Calculation``), ``cpu_calculation`` (NumPy helpers),
``parameters`` / ``settings`` (problem recipe and paths).  The compute path is
``libb200lasso.so`` (hand-written sm_100a CUDA, C ABI in include/b200lasso.h) bound with
ctypes in ``_lib``; there is no CPU fallback for the device classes.
"""
from . import _lib                                   # noqa: F401
from . import cpu_calculation, parameters, settings  # noqa: F401

__all__ = ["_lib", "cpu_calculation", "parameters", "settings", "gpu_calculation", "lasso", "path",
           "distributed"]
__version__ = "0.1.0"


def __getattr__(name):
    # gpu_calculation / lasso import torch lazily through GPU_Calculation only
    if name in ("gpu_calculation", "lasso", "path", "distributed"):
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
