# -*- coding: utf-8 -*-
"""Position-wise mean of ragged lists (mirror of the reference's average.py:6-24)."""
import numpy as np


def list_aver(lists):
    longest = max(len(item) for item in lists)
    total = np.zeros(longest)
    count = np.zeros(longest)
    for item in lists:
        n = len(item)
        total[:n] += np.asarray(item, dtype=float)
        count[:n] += 1
    return list(total / count)
