# -*- coding: utf-8 -*-
"""Multi-GPU lasso: column shards, one process per GPU.

The reference parallelises a block step by splitting the columns of block ``m`` over ``P``
workers (``cpu_calculation.A_bp_get``, cpu_calculation.py:23-27): each worker computes its
slice of ``A_m^T r`` and of ``Bx``/``D`` and a partial ``A_{m,p} D_p``; the partials are summed
(lasso.py:126).  Here a worker is a GPU: rank ``g`` of ``world`` holds slice ``g`` of every
block, owns those columns of ``x``, and the sum of the partial products happens inside the
fused kernel through NVLink peer memory (``b200l_comm_*``).  ``torch.distributed`` is used for
the rendezvous only (exchange of the CUDA IPC handles, gathering ``x``).

Usage on every rank (after ``torch.distributed.init_process_group``)::

    A_loc = shard_columns(A, BLOCK, rank, world)          # or generate the shard on the device
    cal = GPU_Calculation(A_loc, BLOCK)                   # class attr DEVICE = local rank
    connect(cal)                                          # peers' inboxes mapped
    solver = ClassLasso(cal, cal.diag_ATA, A_loc, b, mu, BLOCK, ITER_MAX)
    solver.run(ERR_BOUND)                                 # same arguments on every rank
    x = gather_x(solver.x, BLOCK)                         # global column order
"""
import ctypes

import numpy as np

from . import _lib


def local_columns(K, BLOCK, rank, world):
    """global column indices held by ``rank``, in local order (block-major)."""
    if K % BLOCK:
        raise ValueError("K=%d is not divisible by BLOCK=%d" % (K, BLOCK))
    w = K // BLOCK
    if w % world:
        raise ValueError("block width %d is not divisible by the world size %d "
                         "(the reference requires K %% (BLOCK*P) == 0, cpu_calculation.py:27)" % (w, world))
    wl = w // world
    base = np.arange(BLOCK)[:, None] * w + rank * wl
    return (base + np.arange(wl)[None, :]).reshape(-1)


def shard_columns(A, BLOCK, rank, world):
    """the (N, K/world) matrix of ``rank``: slice ``rank`` of every block, block-major."""
    return np.ascontiguousarray(A[:, local_columns(A.shape[1], BLOCK, rank, world)])


def unshard_x(parts, BLOCK):
    """inverse of the sharding for the solution: ``parts[g]`` is the local x of rank g."""
    world = len(parts)
    kl = parts[0].size
    K = kl * world
    x = np.zeros((K, 1), dtype=np.float64)
    for g, xg in enumerate(parts):
        x[local_columns(K, BLOCK, g, world), 0] = np.asarray(xg, dtype=np.float64).reshape(-1)
    return x


def _dist():
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised")
    return dist


def gather_x(x_local, BLOCK, group=None):
    """all-gather the local solutions and return x in global column order on every rank."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(x_local, dtype=np.float64).reshape(-1))
    dev = None
    if dist.get_backend(group) == "nccl":
        dev = torch.device("cuda", torch.cuda.current_device())
        t = t.to(dev)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    return unshard_x([p.cpu().numpy() for p in parts], BLOCK)


def connect(gpu_cal, group=None):
    """export this rank's inbox, all-gather the IPC handles, map the peers (collective)."""
    import torch
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return
    lib = gpu_cal._lib
    hb = _lib.IPC_HANDLE_BYTES
    mine = (ctypes.c_ubyte * hb)()
    _lib.check(lib.b200l_comm_export(gpu_cal.ctx, rank, world, ctypes.cast(mine, ctypes.c_void_p), hb))
    t = torch.tensor(list(bytes(mine)), dtype=torch.uint8)
    if dist.get_backend(group) == "nccl":
        t = t.to(gpu_cal.device)
    parts = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(parts, t, group=group)
    blob = b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts)
    buf = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
    _lib.check(lib.b200l_comm_connect(gpu_cal.ctx, ctypes.cast(buf, ctypes.c_void_p), hb))
    dist.barrier(group=group)


def disconnect(gpu_cal, group=None):
    """collective teardown: every rank unmaps the peers' inboxes, and only when all have done so
    does a rank free its own (an exported allocation must outlive its importers' mappings)"""
    dist = _dist()
    dist.barrier(group=group)                                        # nobody is still running a solve
    _lib.check(gpu_cal._lib.b200l_comm_close_peers(gpu_cal.ctx))
    dist.barrier(group=group)                                        # every mapping of my inbox is closed
    _lib.check(gpu_cal._lib.b200l_comm_destroy(gpu_cal.ctx))
