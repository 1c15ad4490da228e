# -*- coding: utf-8 -*-
"""Flat-module drop-in directory: runs the reference's own drivers unchanged.

The reference (kingold5/convex_optimization) is flat Python: ``cpu_vs_gpu.py`` and
``compare.py`` import ``lasso``, ``gpu_calculation``, ``cpu_calculation``, ``parameters``,
``settings``, ``average`` by their top-level names, plus ``pycuda.driver``,
``skcuda.cublas`` and ``matplotlib.pyplot`` (cpu_vs_gpu.py:5-16, compare.py:1-3).  This
directory provides exactly those names on top of the B200 library:

* ``lasso.py`` ... ``average.py``: re-exports of the package modules of the same name;
* ``pycuda/`` and ``skcuda/``: thin shim namespaces -- the driver's
  ``cublas.cublasCreate()`` / ``cublasDestroy(h)`` / ``cuda.stop_profiler()`` calls
  (cpu_vs_gpu.py:93,201-202), device vectors (``gpuarray``) and the cuBLAS calls the
  reference's solver classes make (``cublasDgemv`` and the level-1 routines,
  lasso.py:336-353,441-599) on the library's own kernels;
* ``_fallback/matplotlib``: a headless stand-in, used only when matplotlib is not installed.

Run a reference driver against the package, without editing it::

    python -m convex_optimization_b200.dropin /path/to/reference/cpu_vs_gpu.py
    python -m convex_optimization_b200.dropin /path/to/reference/compare.py

or, inside Python, ``convex_optimization_b200.dropin.install()`` and then
``runpy.run_path(script, run_name="__main__")``.
"""
import os
import sys

DIR = os.path.dirname(os.path.abspath(__file__))
FALLBACK_DIR = os.path.join(DIR, "_fallback")
FLAT_MODULES = ("lasso", "gpu_calculation", "cpu_calculation", "parameters", "settings", "average")


def install():
    """Put the drop-in modules in front of ``sys.path`` (the fallbacks at its end, so an
    installed matplotlib wins) and make sure this repository's package is importable."""
    root = os.path.dirname(os.path.dirname(DIR))
    for name in FLAT_MODULES + ("pycuda", "skcuda"):
        mod = sys.modules.get(name)
        if mod is not None and not os.path.abspath(getattr(mod, "__file__", "") or "").startswith(DIR):
            raise RuntimeError("module %r is already imported from %s; install() must run first"
                               % (name, getattr(mod, "__file__", "?")))
    if root not in sys.path:
        sys.path.insert(0, root)
    if DIR in sys.path:
        sys.path.remove(DIR)
    sys.path.insert(0, DIR)
    if FALLBACK_DIR not in sys.path:
        sys.path.append(FALLBACK_DIR)
    return DIR


def run(script, argv=()):
    """Execute ``script`` as ``__main__`` with the drop-in modules installed."""
    import runpy
    install()
    old = sys.argv
    sys.argv = [script] + list(argv)
    try:
        return runpy.run_path(script, run_name="__main__")
    finally:
        sys.argv = old
