# -*- coding: utf-8 -*-
"""``matplotlib.pyplot`` subset of the reference's plots: ``subplots`` -> (fig, ax) with
``ax.plot`` / ``ax.legend``, ``xlabel`` / ``ylabel``, ``show`` / ``savefig``.  ``show()``
prints a one-line summary per curve and, when ``B200L_PLOT_DIR`` is set, writes
``figure_<n>.txt`` (label, x, y columns) there."""
import os

import numpy as np

_figures = []
_labels = {"x": "", "y": ""}


class _Line:
    def __init__(self, x, y, label, color):
        self.x = np.asarray(x, dtype=float).reshape(-1)
        self.y = np.asarray(y, dtype=float).reshape(-1)
        self.label = label or ""
        self.color = color


class Axes:
    def __init__(self):
        self.lines = []
        self.legend_opts = None
        self.xlabel = ""
        self.ylabel = ""

    def plot(self, *args, label=None, color=None, **_):
        if len(args) == 1:
            y = np.asarray(args[0])
            x = np.arange(y.size)
        else:
            x, y = args[0], args[1]
        line = _Line(x, y, label, color)
        self.lines.append(line)
        return [line]

    def legend(self, *_, **opts):
        self.legend_opts = opts
        return self

    def set_xlabel(self, text, **_):
        self.xlabel = text

    def set_ylabel(self, text, **_):
        self.ylabel = text


class Figure:
    def __init__(self):
        self.axes = []

    def savefig(self, path, **_):
        _write(self, path)


def subplots(*_, **__):
    fig = Figure()
    ax = Axes()
    fig.axes.append(ax)
    _figures.append(fig)
    return fig, ax


def figure(*_, **__):
    return subplots()[0]


def gca():
    if not _figures:
        subplots()
    return _figures[-1].axes[-1]


def plot(*args, **kw):
    return gca().plot(*args, **kw)


def legend(*args, **kw):
    return gca().legend(*args, **kw)


def xlabel(text, **_):
    gca().xlabel = text


def ylabel(text, **_):
    gca().ylabel = text


def _write(fig, path):
    with open(path, "w") as f:
        for ax in fig.axes:
            f.write("# xlabel=%s ylabel=%s\n" % (ax.xlabel, ax.ylabel))
            for line in ax.lines:
                f.write("# curve label=%s color=%s points=%d\n" % (line.label, line.color, line.x.size))
                for a, b in zip(line.x, line.y):
                    f.write("%.17g %.17g\n" % (a, b))


def savefig(path, **_):
    if _figures:
        _write(_figures[-1], path)


def show(*_, **__):
    out = os.environ.get("B200L_PLOT_DIR")
    for n, fig in enumerate(_figures):
        for ax in fig.axes:
            for line in ax.lines:
                if line.x.size:
                    print("[headless plot] figure %d curve %-12s %d points, x %.4g..%.4g (%s), y %.4g..%.4g (%s)"
                          % (n, line.label, line.x.size, line.x.min(), line.x.max(), ax.xlabel,
                             line.y.min(), line.y.max(), ax.ylabel))
        if out:
            os.makedirs(out, exist_ok=True)
            _write(fig, os.path.join(out, "figure_%d.txt" % n))
    del _figures[:]


def close(*_, **__):
    del _figures[:]
