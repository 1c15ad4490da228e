# -*- coding: utf-8 -*-
"""Top-level name ``settings`` of the reference (settings.py:11-19).  ``init()`` sets the
globals of the package module and mirrors them here, so both ``settings.HOME`` (this
module, what the reference's drivers read) and the package's own copy agree."""
from convex_optimization_b200 import settings as _impl

HOME = None
Dir_PERFORMANCE = None


def init():
    global HOME, Dir_PERFORMANCE
    _impl.init()
    HOME = _impl.HOME
    Dir_PERFORMANCE = _impl.Dir_PERFORMANCE
