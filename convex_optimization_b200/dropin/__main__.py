# -*- coding: utf-8 -*-
"""``python -m convex_optimization_b200.dropin <script.py> [args]``: run a driver of the
reference (cpu_vs_gpu.py, compare.py) unchanged against the B200 library."""
import sys

from . import run

if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.stderr.write("usage: python -m convex_optimization_b200.dropin <script.py> [args]\n")
        sys.exit(2)
    run(sys.argv[1], sys.argv[2:])
