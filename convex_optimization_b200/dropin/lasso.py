# -*- coding: utf-8 -*-
"""Top-level name ``lasso`` of the reference (lasso.py) -> convex_optimization_b200.lasso."""
from convex_optimization_b200.lasso import *          # noqa: F401,F403
from convex_optimization_b200 import lasso as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
