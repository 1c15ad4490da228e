# -*- coding: utf-8 -*-
"""``import pycuda.autoinit`` (lasso.py:14, gpu_calculation.py:4): PyCUDA creates a context on
device 0 at import.  Here the CUDA runtime context is created lazily by the first library
call; ``device`` / ``context`` exist for code that only stores them."""


class _Device:
    def __init__(self, index=0):
        self.index = index

    def name(self):
        import torch
        return torch.cuda.get_device_name(self.index)


class _Context:
    @staticmethod
    def synchronize():
        import torch
        torch.cuda.synchronize()

    def pop(self):
        pass

    def push(self):
        pass


device = _Device(0)
context = _Context()
