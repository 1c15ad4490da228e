# -*- coding: utf-8 -*-
"""``pycuda.gpuarray`` names used by the reference's solver classes (lasso.py:324-331,
371-389,485-500): ``zeros``, ``empty``, ``empty_like``, ``zeros_like``, ``to_gpu`` and a
``GPUArray`` with ``gpudata`` / ``shape`` / ``size`` / ``dtype`` / ``get`` / ``set`` /
``fill``.  Device memory is a torch tensor."""
import numpy as np

_NP2T = None


def _torch():
    import torch
    return torch


def _tdtype(np_dtype):
    global _NP2T
    torch = _torch()
    if _NP2T is None:
        _NP2T = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32,
                 np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64,
                 np.dtype(np.uint8): torch.uint8}
    return _NP2T[np.dtype(np_dtype)]


class DevicePointer(int):
    """An int device address (what cuBLAS-style calls take) that keeps its allocation alive and
    can carry what the address alone does not say: ``ld`` (padded leading dimension) and, for
    a column block of a ``GPU_Calculation``, ``block = (gpu_cal, m)``."""
    _owner = None
    ld = None
    block = None

    @classmethod
    def wrap(cls, addr, owner, ld=None, block=None):
        p = cls(int(addr))
        p._owner = owner
        p.ld = ld
        p.block = block
        return p


class GPUArray:
    def __init__(self, shape, dtype=np.float64, _tensor=None):
        torch = _torch()
        if isinstance(shape, (int, np.integer)):
            shape = (int(shape),)
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        self._t = _tensor if _tensor is not None else torch.empty(self.shape, dtype=_tdtype(self.dtype), device="cuda")

    @property
    def size(self):
        return int(np.prod(self.shape)) if self.shape else 1

    @property
    def nbytes(self):
        return self.size * self.dtype.itemsize

    @property
    def gpudata(self):
        return DevicePointer.wrap(self._t.data_ptr(), self._t)

    @property
    def ptr(self):
        return self._t.data_ptr()

    @property
    def tensor(self):
        return self._t

    def get(self, ary=None):
        host = self._t.cpu().numpy()
        if ary is None:
            return host
        ary[...] = host.reshape(ary.shape)
        return ary

    def set(self, ary):
        torch = _torch()
        src = np.ascontiguousarray(ary, dtype=self.dtype)
        if src.size != self.size:
            raise ValueError("GPUArray.set: size mismatch (%d vs %d)" % (src.size, self.size))
        self._t.copy_(torch.from_numpy(src.reshape(self.shape)))
        return self

    def fill(self, value):
        self._t.fill_(value)
        return self

    def copy(self):
        return GPUArray(self.shape, self.dtype, self._t.clone())

    def __getitem__(self, idx):
        t = self._t[idx]
        return GPUArray(tuple(t.shape), self.dtype, t)

    def __len__(self):
        return self.shape[0]

    def __repr__(self):
        return "GPUArray(shape=%s, dtype=%s)" % (self.shape, self.dtype)


def empty(shape, dtype=np.float64, **_):
    return GPUArray(shape, dtype)


def zeros(shape, dtype=np.float64, **_):
    return GPUArray(shape, dtype).fill(0)


def empty_like(other):
    return GPUArray(other.shape, other.dtype)


def zeros_like(other):
    return GPUArray(other.shape, other.dtype).fill(0)


def to_gpu(ary):
    ary = np.ascontiguousarray(ary)
    return GPUArray(ary.shape, ary.dtype).set(ary)
