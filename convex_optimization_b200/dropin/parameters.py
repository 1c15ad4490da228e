# -*- coding: utf-8 -*-
"""Top-level name ``parameters`` of the reference (parameters.py) -> convex_optimization_b200.parameters."""
from convex_optimization_b200.parameters import *          # noqa: F401,F403
from convex_optimization_b200 import parameters as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
