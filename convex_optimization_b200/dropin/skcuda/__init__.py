# -*- coding: utf-8 -*-
"""Shim namespace ``skcuda``: ``skcuda.cublas`` as used by the
reference (cpu_vs_gpu.py:8,93,202; lasso.py:15,336-353,404-454,562-576)."""
