# -*- coding: utf-8 -*-
"""``pycuda.driver`` names used by the reference: ``stop_profiler`` (cpu_vs_gpu.py:201),
``mem_alloc`` / ``memcpy_htod`` / ``memcpy_dtoh`` (gpu_calculation.py:227-228,265,281),
``Context.synchronize``."""
import numpy as np

from .gpuarray import DevicePointer, _torch


def init(flags=0):
    pass


def start_profiler():
    _torch().cuda.cudart().cudaProfilerStart()


def stop_profiler():
    torch = _torch()
    if torch.cuda.is_available():
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()


class DeviceAllocation(DevicePointer):
    """result of ``mem_alloc``: an int-like device pointer that owns its bytes"""

    def free(self):
        self._owner = None


def mem_alloc(nbytes):
    torch = _torch()
    buf = torch.empty(int(nbytes), dtype=torch.uint8, device="cuda")
    return DeviceAllocation.wrap(buf.data_ptr(), buf)


def _as_u8(ptr, nbytes):
    owner = getattr(ptr, "_owner", None)
    torch = _torch()
    if owner is not None and isinstance(owner, torch.Tensor):
        flat = owner.reshape(-1).view(torch.uint8)
        off = int(ptr) - owner.data_ptr()
        return flat[off:off + nbytes]
    raise TypeError("memcpy needs a pointer from mem_alloc() or GPUArray.gpudata")


def memcpy_htod(dest, src):
    torch = _torch()
    src = np.ascontiguousarray(src)
    _as_u8(dest, src.nbytes).copy_(torch.from_numpy(src.reshape(-1).view(np.uint8)))


def memcpy_dtoh(dest, src):
    torch = _torch()
    out = _as_u8(src, dest.nbytes).cpu().numpy()
    dest.reshape(-1).view(np.uint8)[:] = out


class Context:
    @staticmethod
    def synchronize():
        _torch().cuda.synchronize()
