# -*- coding: utf-8 -*-
"""Top-level name ``gpu_calculation`` of the reference (gpu_calculation.py) -> convex_optimization_b200.gpu_calculation."""
from convex_optimization_b200.gpu_calculation import *          # noqa: F401,F403
from convex_optimization_b200 import gpu_calculation as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
