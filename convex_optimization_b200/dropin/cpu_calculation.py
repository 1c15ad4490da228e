# -*- coding: utf-8 -*-
"""Top-level name ``cpu_calculation`` of the reference (cpu_calculation.py) -> convex_optimization_b200.cpu_calculation."""
from convex_optimization_b200.cpu_calculation import *          # noqa: F401,F403
from convex_optimization_b200 import cpu_calculation as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
