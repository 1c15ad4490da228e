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


def connect(gpu_cal, group=None, transport=None, multicast=None):
    """Map the ranks' inboxes into each other (collective).

    ``transport``: ``"symm"`` allocates the inbox as torch symmetric memory
    (``torch.distributed._symmetric_memory``: peer mappings plus, on NVSwitch systems, a multicast
    (NVLS) mapping, so the kernel sends a row with ONE ``multimem.st`` instead of ``world - 1`` peer
    stores) and hands the plain addresses to ``b200l_comm_attach``; ``"ipc"`` uses the library's own
    CUDA IPC export (``b200l_comm_export`` / ``_connect``).  Default (or ``B200L_COMM``): ``"auto"`` =
    symmetric memory when it is available, else IPC.  ``multicast=False`` (or ``B200L_MULTICAST=0``)
    keeps the unicast peer stores on a symmetric-memory inbox."""
    import os
    dist = _dist()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    if world == 1:
        return "single"
    transport = transport or os.environ.get("B200L_COMM", "auto")
    if multicast is None:
        multicast = os.environ.get("B200L_MULTICAST", "1") != "0"
    used = None
    if transport in ("auto", "symm") and dist.get_backend(group) == "nccl":
        try:
            used = _connect_symm(gpu_cal, dist, group, rank, world, multicast)
        except Exception as e:                      # no symmetric memory on this system / build
            if transport == "symm":
                raise
            import warnings
            warnings.warn("symmetric memory unavailable (%r); using CUDA IPC inboxes" % (e,))
    if used is None:
        _connect_ipc(gpu_cal, dist, group, rank, world)
        used = "ipc"
    gpu_cal._world, gpu_cal._rank, gpu_cal._group = world, rank, group
    gpu_cal._transport = used
    dist.barrier(group=group)
    return used


def _connect_symm(gpu_cal, dist, group, rank, world, multicast):
    import torch
    import torch.distributed._symmetric_memory as symm
    lib = gpu_cal._lib
    nbytes = ctypes.c_int64()
    _lib.check(lib.b200l_comm_inbox_bytes(gpu_cal.ctx, world, ctypes.byref(nbytes)))
    grp = group if group is not None else dist.group.WORLD
    try:
        symm.enable_symm_mem_for_group(grp.group_name)
    except Exception:
        pass
    with torch.cuda.device(gpu_cal.device):
        buf = symm.empty(int(nbytes.value), dtype=torch.uint8, device=gpu_cal.device)
        buf.zero_()
        hdl = symm.rendezvous(buf, grp)
        torch.cuda.synchronize(gpu_cal.device)
    ptrs = [int(a) for a in hdl.buffer_ptrs]
    mc = int(getattr(hdl, "multicast_ptr", 0) or 0) if multicast else 0
    dist.barrier(group=group)                        # every rank's inbox is zero-filled before anyone sends
    arr = (ctypes.c_void_p * world)(*ptrs)
    _lib.check(lib.b200l_comm_attach(gpu_cal.ctx, rank, world, arr, ctypes.c_void_p(mc) if mc else None))
    gpu_cal._symm = (buf, hdl)                       # the memory must outlive the connection
    return "symm+multicast" if mc else "symm"


def _connect_ipc(gpu_cal, dist, group, rank, world):
    """export this rank's inbox, all-gather the IPC handles, map the peers"""
    import torch
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


def disconnect(gpu_cal, group=None):
    """collective teardown: every rank unmaps the peers' inboxes, and only when all have done so
    does a rank free its own (an exported allocation must outlive its importers' mappings)"""
    dist = _dist()
    dist.barrier(group=group)                                        # nobody is still running a solve
    _lib.check(gpu_cal._lib.b200l_comm_close_peers(gpu_cal.ctx))
    dist.barrier(group=group)                                        # every mapping of my inbox is closed
    _lib.check(gpu_cal._lib.b200l_comm_destroy(gpu_cal.ctx))
    gpu_cal._world, gpu_cal._rank, gpu_cal._group = 1, 0, None
    gpu_cal._symm = None


def objective(gpu_cal, mu, group=None):
    """0.5 |Ax - b|^2 + mu |x|_1 of the whole instance: the residual is replicated on the ranks,
    the l1 term is summed over the column shards"""
    import torch
    rss, l1 = ctypes.c_double(), ctypes.c_double()
    _lib.check(gpu_cal._lib.b200l_objective_terms(gpu_cal.ctx, ctypes.byref(rss), ctypes.byref(l1)))
    total_l1 = l1.value
    if getattr(gpu_cal, '_world', 1) > 1:
        dist = _dist()
        grp = group if group is not None else gpu_cal._group
        t = torch.tensor([l1.value], dtype=torch.float64)
        if dist.get_backend(grp) == "nccl":
            t = t.to(gpu_cal.device)
        dist.all_reduce(t, group=grp)
        total_l1 = float(t.item())
    return 0.5 * rss.value + float(mu) * total_l1
