#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""TEST / BASELINE INFRASTRUCTURE -- not part of the product path.

Recipe for ``oracle/_ref/``: an UNMODIFIED copy of the reference's pure-Python files, taken
from where they lie under /root/reference (or $B200L_REFERENCE).  ``oracle/_ref/`` is
git-ignored (the reference's sources never enter this repository's history) but it is not
gpurun-ignored, so the copy travels to the GPU box like the built .so files do.  There it is
used by

* ``bench.py --impl reference``: times the reference's own ``ClassLassoCPU.run``
  (lasso.py:70-169, multiprocessing.Pool at :101) on the box's host cores;
* ``tests/test_dropin_drivers.py``: runs the unmodified drivers ``cpu_vs_gpu.py`` and
  ``compare.py`` against convex_optimization_b200 through the drop-in modules.

``__graft_entry__.build()`` calls ``make()`` whenever the reference tree is present.  A
manifest with the sha256 of every copied file is written next to the copies, and ``check()``
verifies that the copies are byte-identical to their manifest (nobody edited them).
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = ("lasso.py", "cpu_calculation.py", "parameters.py", "settings.py", "average.py",
         "gpu_calculation.py", "cpu_vs_gpu.py", "compare.py")
MANIFEST = "MANIFEST.json"


def reference_dir():
    return os.environ.get("B200L_REFERENCE", "/root/reference")


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def make(src=None, verbose=False):
    """copy the reference files to oracle/_ref/; returns the destination or None when the
    reference tree is not present (the GPU box: the copy made here has travelled)"""
    src = src or reference_dir()
    if not os.path.isdir(src):
        return None
    os.makedirs(DEST, exist_ok=True)
    manifest = {"source": src, "files": {}}
    for name in FILES:
        s = os.path.join(src, name)
        if not os.path.exists(s):
            raise RuntimeError("reference file %s is missing" % s)
        d = os.path.join(DEST, name)
        shutil.copyfile(s, d)
        manifest["files"][name] = _sha(d)
        if verbose:
            print("copied %s -> %s" % (s, d))
    with open(os.path.join(DEST, MANIFEST), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    with open(os.path.join(DEST, "README.txt"), "w") as f:
        f.write("Byte-identical copies of the reference's own files (see MANIFEST.json), made by oracle/make_ref.py\n"
                "from %s.  NOT this repository's code and not tracked by git (.gitignore: oracle/_ref/):\n"
                "the directory exists so that bench.py --impl reference can time the UNMODIFIED ClassLassoCPU on\n"
                "the GPU box (VERDICT round 1, \"Make the CPU arm the reference\") and so that the drop-in tests\n"
                "can run the unmodified drivers.  Nothing under convex_optimization_b200/ imports it.\n" % src)
    return DEST


def available():
    return os.path.exists(os.path.join(DEST, MANIFEST))


def check():
    """True when every copy is byte-identical to what make() recorded"""
    if not available():
        return False
    with open(os.path.join(DEST, MANIFEST)) as f:
        manifest = json.load(f)
    return all(os.path.exists(os.path.join(DEST, n)) and _sha(os.path.join(DEST, n)) == h
               for n, h in manifest["files"].items())


def import_reference(stub_gpu=True):
    """import the UNMODIFIED reference modules from oracle/_ref (returns lasso, parameters,
    cpu_calculation).  The reference's lasso.py imports pycuda / skcuda at module top
    (lasso.py:13-16); ``ClassLassoCPU`` uses neither, so empty stand-ins are enough."""
    import types
    if not available():
        raise RuntimeError("oracle/_ref is not populated (run oracle/make_ref.py where the reference tree exists)")
    if stub_gpu:
        for name in ("pycuda", "pycuda.gpuarray", "pycuda.autoinit", "pycuda.elementwise", "pycuda.driver",
                     "pycuda.compiler", "skcuda", "skcuda.cublas"):
            if name not in sys.modules:
                sys.modules[name] = types.ModuleType(name)
        sys.modules["pycuda"].gpuarray = sys.modules["pycuda.gpuarray"]
        if not hasattr(sys.modules["pycuda.elementwise"], "ElementwiseKernel"):
            sys.modules["pycuda.elementwise"].ElementwiseKernel = object
        sys.modules["skcuda"].cublas = sys.modules["skcuda.cublas"]
    for name in ("lasso", "parameters", "cpu_calculation", "settings"):
        mod = sys.modules.get(name)
        if mod is not None and os.path.dirname(os.path.abspath(getattr(mod, "__file__", ""))) != DEST:
            raise RuntimeError("a different module named %r is already imported (%s)" % (name, mod.__file__))
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    import lasso as ref_lasso
    import parameters as ref_parameters
    import cpu_calculation as ref_cpu
    assert os.path.dirname(os.path.abspath(ref_lasso.__file__)) == DEST
    return ref_lasso, ref_parameters, ref_cpu


if __name__ == "__main__":
    out = make(verbose=True)
    print("oracle/_ref: %s" % (out or "reference tree not present, nothing copied"))
    sys.exit(0 if out else 1)
