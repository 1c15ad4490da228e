# -*- coding: utf-8 -*-
"""Top-level name ``average`` of the reference (average.py) -> convex_optimization_b200.average."""
from convex_optimization_b200.average import *          # noqa: F401,F403
from convex_optimization_b200 import average as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
