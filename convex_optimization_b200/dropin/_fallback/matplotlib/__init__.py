# -*- coding: utf-8 -*-
"""Headless stand-in for matplotlib, importable only when matplotlib itself is not installed
(``convex_optimization_b200.dropin.install()`` appends this directory to the END of
``sys.path``).  It implements what compare.py:13-20 and cpu_vs_gpu.py:28-37 call and writes
the curves as text instead of opening a window."""
__version__ = "0+b200lasso.headless"


def use(backend, **_):
    pass
