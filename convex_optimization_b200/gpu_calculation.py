# -*- coding: utf-8 -*-
"""``GPU_Calculation`` -- the device side of the lasso hot path on B200.

Mirror of the reference class (gpu_calculation.py:141-292): same constructor, class
attributes, instance attributes and the three compute entry points, backed by the
sm_100a kernels of libb200lasso.so through ctypes.  torch is used only to own the
device copy of ``A`` (the reference's ``A_b_gpu``, gpu_calculation.py:224).

Additions (not in the reference): ``LAYOUT`` selects the device layout of the blocks
('row' = the reference's (BLOCK, N, w), 'transposed' = the pre-transposed (BLOCK, w, N)
copy the reference only sketched, gpu_calculation.py:94-113,174); ``TYPE`` may be
'float' as well as 'double'; ``A`` may be a torch tensor already on the device.
"""
import ctypes

import numpy as np

from . import _lib


class DeviceBlock:
    """``gpu_cal.A_b_gpu[m]``: column block m on the device, presented the way the reference's
    solver classes reach through it (lasso.py:336,349-351,546): ``.shape == (N, w)`` and
    ``.gpudata`` (an int device address).  The address also carries what a bare pointer cannot
    say -- the padded leading dimension, the storage type / layout and the owning
    ``GPU_Calculation`` -- so the ``skcuda.cublas`` shim can run a ``cublasDgemv`` on this
    block with the library's own kernels.  ``.tensor`` is the torch view, ``.get()`` a host copy."""

    def __init__(self, gpu_cal, m):
        self._cal = gpu_cal
        self._m = int(m)
        self.shape = (gpu_cal.MAT_HEIGHT, gpu_cal.MAT_WIDTH)
        self.dtype = np.dtype(np.float64 if gpu_cal.TYPE == 'double' else np.float32)
        self.size = self.shape[0] * self.shape[1]

    @property
    def tensor(self):
        cal = self._cal
        store = cal._A_store[self._m]
        if cal.LAYOUT == 'row':
            return store[:, :cal.MAT_WIDTH]
        return store[:, :cal.MAT_HEIGHT].t()

    @property
    def gpudata(self):
        from .dropin.pycuda.gpuarray import DevicePointer
        cal = self._cal
        return DevicePointer.wrap(cal._A_store[self._m].data_ptr(), cal._A_store, ld=cal.ld,
                                  block=(cal, self._m))

    def get(self):
        return self.tensor.cpu().numpy()

    def cpu(self):
        return self.tensor.cpu()


class DeviceBlocks:
    """sequence of ``DeviceBlock`` (the reference's ``A_b_gpu``, gpu_calculation.py:224)"""

    def __init__(self, gpu_cal):
        self._cal = gpu_cal

    def __len__(self):
        return self._cal.Block

    def __getitem__(self, m):
        if not -self._cal.Block <= m < self._cal.Block:
            raise IndexError(m)
        return DeviceBlock(self._cal, m % self._cal.Block)

    @property
    def shape(self):
        return (self._cal.Block, self._cal.MAT_HEIGHT, self._cal.MAT_WIDTH)


class GPU_Calculation:
    # launch-shape knobs of the reference kernels (gpu_calculation.py:143-145).  They are
    # accepted and stored so drivers that set them (cpu_vs_gpu.py:104-105) run unchanged;
    # the B200 kernels size their own tiles from the problem shape.
    T_WIDTH_TRANS = 64
    T_WIDTH = 64
    T_HEIGHT = 512
    # element type of the device matrix (gpu_calculation.py:146): 'double' or 'float'
    TYPE = 'double'
    LAYOUT = 'row'
    DEVICE = 0

    def __init__(self, A, Block):
        self._setup(int(A.shape[0]), int(A.shape[1]), Block)
        self.init_gpu_array(A)

    @classmethod
    def from_device_blocks(cls, store, N, K, Block):
        """Adopt an already laid-out device tensor (Block, rows, ld) -- used for instances
        generated on the device that never exist on the host."""
        self = cls.__new__(cls)
        self._setup(int(N), int(K), Block)
        rows = self.MAT_HEIGHT if self.LAYOUT == 'row' else self.MAT_WIDTH
        if (tuple(store.shape) != (self.Block, rows, self.ld) or store.dtype != self.torch_dtype
                or not store.is_contiguous() or store.device != self.device):
            raise ValueError("device blocks must be a contiguous %s tensor of shape %s on %s"
                             % (self.torch_dtype, (self.Block, rows, self.ld), self.device))
        self._adopt(store)
        return self

    @classmethod
    def padded_ld(cls, N, K, Block):
        """leading dimension the library uses for this shape / TYPE / LAYOUT"""
        V = 2 if cls.TYPE == 'double' else 4
        n = (K // Block) if cls.LAYOUT == 'row' else N
        return (n + V - 1) // V * V

    def _setup(self, N, K, Block):
        import torch
        self._torch = torch
        lib = _lib.load()
        self._lib = lib
        if self.TYPE not in ('double', 'float'):
            raise ValueError("GPU_Calculation.TYPE must be 'double' or 'float'")
        if self.LAYOUT not in ('row', 'transposed'):
            raise ValueError("GPU_Calculation.LAYOUT must be 'row' or 'transposed'")
        self.Block = int(Block)
        self.MAT_HEIGHT = N
        self.MAT_WIDTH_ALL = K
        if self.MAT_WIDTH_ALL % self.Block != 0:
            raise ValueError("K=%d is not divisible by Block=%d"
                             % (self.MAT_WIDTH_ALL, self.Block))
        self.MAT_WIDTH = self.MAT_WIDTH_ALL // self.Block
        self.dtype_code = _lib.F64 if self.TYPE == 'double' else _lib.F32
        self.layout_code = _lib.ROWMAJOR if self.LAYOUT == 'row' else _lib.TRANSPOSED
        self.torch_dtype = torch.float64 if self.TYPE == 'double' else torch.float32
        self.device = torch.device('cuda', self.DEVICE)

        ctx = ctypes.c_void_p()
        _lib.check(lib.b200l_ctx_create(ctypes.byref(ctx), self.dtype_code, self.layout_code,
                                        self.MAT_HEIGHT, self.MAT_WIDTH_ALL, self.Block,
                                        self.DEVICE))
        self.ctx = ctx
        ld = ctypes.c_int64()
        _lib.check(lib.b200l_ctx_ld(ctx, ctypes.byref(ld)))
        self.ld = int(ld.value)

    # -- device copy of A (gpu_calculation.py:171-175,222-225) ---------------------------
    def init_gpu_array(self, A):
        torch = self._torch
        N, w, B = self.MAT_HEIGHT, self.MAT_WIDTH, self.Block
        rows = N if self.LAYOUT == 'row' else w
        with torch.cuda.device(self.device):
            store = torch.zeros((B, rows, self.ld), dtype=self.torch_dtype, device=self.device)
            is_torch = isinstance(A, torch.Tensor)
            for m in range(B):
                if is_torch:
                    blk = A[:, m * w:(m + 1) * w].to(device=self.device, dtype=self.torch_dtype)
                else:
                    blk = torch.from_numpy(np.ascontiguousarray(A[:, m * w:(m + 1) * w])).to(
                        device=self.device, dtype=self.torch_dtype)
                if self.LAYOUT == 'row':
                    store[m, :, :w] = blk
                else:
                    store[m, :, :N] = blk.t()
                del blk
            torch.cuda.synchronize(self.device)
        self._adopt(store)

    def _adopt(self, store):
        N, w = self.MAT_HEIGHT, self.MAT_WIDTH
        self._A_store = store
        # the reference exposes A_b_gpu[m] with shape (N, w) and .gpudata (lasso.py:336,349)
        self.A_b_gpu = DeviceBlocks(self)
        _lib.check(self._lib.b200l_ctx_bind_A(self.ctx, ctypes.c_void_p(store.data_ptr())))

    def __del__(self):
        ctx = getattr(self, 'ctx', None)
        if ctx is not None and ctx.value:
            try:
                self._lib.b200l_ctx_destroy(ctx)
            except Exception:
                pass
            self.ctx = None

    # -- diag(A^T A) per block, (Block, w, 1) float64 (gpu_calculation.py:246-261) -------
    @property
    def diag_ATA(self):
        # computed on the device once per matrix (one pass over A) and kept there; every access
        # returns a fresh host array like the reference's .get() (gpu_calculation.py:261)
        self._use_own_diag()
        out = np.empty((self.Block, self.MAT_WIDTH, 1), np.float64)
        _lib.check(self._lib.b200l_diag_ata(self.ctx, _lib.dptr(out)))
        self._diag_host = out.copy()
        return out

    def _use_own_diag(self):
        if getattr(self, '_diag_is_custom', False):
            _lib.check(self._lib.b200l_set_diag(self.ctx, None))
            self._diag_is_custom = False

    def _use_custom_diag(self, d_ATA):
        """the solver was constructed with a diagonal that is not this matrix's own (the d_ATA
        argument of the reference's solver classes, lasso.py:26-30): upload it for the fused path"""
        d = np.ascontiguousarray(np.asarray(d_ATA, dtype=np.float64).reshape(self.Block, self.MAT_WIDTH))
        _lib.check(self._lib.b200l_set_diag(self.ctx, _lib.dptr(d)))
        self._diag_is_custom = True

    def is_own_diag(self, d_ATA):
        """True when ``d_ATA`` equals diag(A^T A) of the matrix on the device"""
        own = getattr(self, '_diag_host', None)
        if own is None:
            own = self.diag_ATA
        d = np.asarray(d_ATA)
        return d.size == own.size and np.array_equal(d.reshape(own.shape), own)

    # -- s13 <- A_m^T s11 in place (gpu_calculation.py:264-277) ---------------------------
    def mat_tMulVec_DiffSize(self, s13, index_m, s11):
        s11c = np.ascontiguousarray(s11, dtype=np.float64).reshape(-1)
        if s11c.size != self.MAT_HEIGHT:
            raise ValueError("s11 must have %d entries" % self.MAT_HEIGHT)
        out = self._out_view(s13, self.MAT_WIDTH)
        _lib.check(self._lib.b200l_gemv_t(self.ctx, int(index_m), _lib.dptr(s11c), _lib.dptr(out)))
        if out is not s13:
            s13[...] = out.reshape(s13.shape)

    # -- s23 <- A_m d in place (gpu_calculation.py:280-292) -------------------------------
    def matMulVec_DiffSize(self, s23, index_m, descent_d):
        dc = np.ascontiguousarray(descent_d, dtype=np.float64).reshape(-1)
        if dc.size != self.MAT_WIDTH:
            raise ValueError("descent_d must have %d entries" % self.MAT_WIDTH)
        out = self._out_view(s23, self.MAT_HEIGHT)
        _lib.check(self._lib.b200l_gemv_n(self.ctx, int(index_m), _lib.dptr(dc), _lib.dptr(out)))
        if out is not s23:
            s23[...] = out.reshape(s23.shape)

    @staticmethod
    def _out_view(arr, n):
        if (isinstance(arr, np.ndarray) and arr.dtype == np.float64 and arr.flags.c_contiguous
                and arr.size == n):
            return arr
        if not isinstance(arr, np.ndarray) or arr.size != n:
            raise ValueError("output must be a numpy array with %d entries" % n)
        return np.empty(n, np.float64)

    # -- fused-path helpers (used by lasso.py) ---------------------------------------------
    def run_config(self):
        vals = [ctypes.c_int32() for _ in range(6)]
        _lib.check(self._lib.b200l_run_config(self.ctx, *[ctypes.byref(v) for v in vals]))
        keys = ('grid', 'threads', 'smem_bytes', 'tile_rows', 'ring_slots', 'tiles_per_slab')
        return dict(zip(keys, [int(v.value) for v in vals]))

    def set_tuning(self, slot_bytes=0, inflight=0):
        _lib.check(self._lib.b200l_set_tuning(self.ctx, int(slot_bytes), int(inflight)))
