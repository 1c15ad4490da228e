# -*- coding: utf-8 -*-
"""Path configuration (mirror of the reference's settings.py:11-19): ``init()`` sets the
module globals ``HOME`` and ``Dir_PERFORMANCE`` used by parameters.py and compare.py.
``B200L_HOME`` overrides the home directory (useful on read-only hosts)."""
import os
import socket
from pathlib import Path

HOME = None
Dir_PERFORMANCE = None


def init():
    global HOME, Dir_PERFORMANCE
    override = os.environ.get("B200L_HOME")
    if override:
        HOME = override
    elif socket.gethostname() == "Xng-PC":          # reference settings.py:15
        HOME = "/home/xng"
    else:
        HOME = str(Path.home())
    Dir_PERFORMANCE = HOME + "/Documents/convex_optimization/Performance"
