"""Host-side logic of the multi-GPU path on CPU: the column sharding / un-sharding index maps,
and a world_size-2 gloo run that plays the per-rank iteration with NumPy (local block gradient
and prox, all-reduced partial A_m D, redundant line search) and must reproduce the oracle with
the reference's P-way split (lasso.py:107-126)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT
from oracle import lasso_oracle as orc
from convex_optimization_b200 import distributed as dd


def test_shard_index_maps_roundtrip():
    K, BLOCK, world = 48, 4, 3
    seen = np.concatenate([dd.local_columns(K, BLOCK, g, world) for g in range(world)])
    assert sorted(seen.tolist()) == list(range(K))
    w, wl = K // BLOCK, K // BLOCK // world
    cols = dd.local_columns(K, BLOCK, 1, world)
    assert cols[:wl].tolist() == list(range(wl, 2 * wl))                 # slice 1 of block 0
    assert cols[wl:2 * wl].tolist() == list(range(w + wl, w + 2 * wl))   # slice 1 of block 1
    A = np.arange(5 * K, dtype=np.float64).reshape(5, K)
    parts = [dd.shard_columns(A, BLOCK, g, world) for g in range(world)]
    assert all(p.shape == (5, K // world) and p.flags.c_contiguous for p in parts)
    x = np.arange(K, dtype=np.float64).reshape(K, 1)
    assert np.array_equal(dd.unshard_x([x[dd.local_columns(K, BLOCK, g, world)] for g in range(world)], BLOCK), x)
    with pytest.raises(ValueError):
        dd.local_columns(50, 4, 0, 2)
    with pytest.raises(ValueError):
        dd.local_columns(48, 4, 0, 5)


def _rank_main(rank, world, port, N, K, BLOCK, den, seed, out):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A, _, b, mu = orc.make_problem(N, K, den, seed=seed)
    A_loc = dd.shard_columns(A, BLOCK, rank, world)
    wl = K // BLOCK // world
    d = np.sum(A_loc * A_loc, axis=0).reshape(BLOCK, wl, 1)
    x = np.zeros((BLOCK, wl, 1))
    r = -b.copy()
    cnt, gamma, iters, stopped = 0, 0.0, 0, False
    ITER_MAX = 80 * BLOCK
    for t in range(ITER_MAX):
        m = t % BLOCK
        Am = A_loc[:, m * wl:(m + 1) * wl]
        g = Am.T @ r                                                    # my slice of the block gradient
        Bx = orc.soft_thresholding(d[m] * x[m] - g, mu) / d[m]
        D = Bx - x[m]
        part = np.concatenate([(Am @ D).reshape(-1), [np.abs(Bx).sum() - np.abs(x[m]).sum()]])
        tt = torch.from_numpy(part)
        dist.all_reduce(tt)                                             # the reduce of lasso.py:126
        q = tt.numpy()[:-1].reshape(-1, 1)
        l1 = float(tt.numpy()[-1])
        e = torch.tensor([float(orc.error_crit(g, x[m], mu))], dtype=torch.float64)
        dist.all_reduce(e, op=dist.ReduceOp.MAX)
        rq, qq = float(r.T @ q), float(q.T @ q)
        if qq != 0.0:
            gamma = min(max(-(rq + mu * l1) / qq, 0.0), 1.0)
        iters = t + 1
        if float(e) < 1e-4:
            cnt += 1
        if m == BLOCK - 1:
            if cnt == BLOCK:
                stopped = True
                break
            cnt = 0
        x[m] += gamma * D
        r = r + gamma * q
    xg = dd.gather_x(x.reshape(-1, 1), BLOCK)
    if rank == 0:
        np.savez(out, x=xg, iters=iters, stopped=stopped)
    dist.destroy_process_group()


def test_gloo_world2_sharded_iteration_matches_oracle(tmp_path):
    import torch.multiprocessing as mp
    N, K, BLOCK, den, seed, world = 120, 960, 4, 0.05, 21, 2
    out = str(tmp_path / "res.npz")
    mp.spawn(_rank_main, args=(world, 29631, N, K, BLOCK, den, seed, out), nprocs=world, join=True)
    res = np.load(out)
    A, _, b, mu = orc.make_problem(N, K, den, seed=seed)
    o = orc.lasso_oracle(A, b, mu, BLOCK, 80 * BLOCK, 1e-4, P=world, faithful=False)
    assert int(res["iters"]) == o["iters"] and bool(res["stopped"]) == o["stopped"]
    assert np.array_equal(res["x"] != 0, o["x"] != 0)
    assert np.abs(res["x"] - o["x"]).max() / np.abs(o["x"]).max() < 1e-10
    o1 = orc.lasso_oracle(A, b, mu, BLOCK, 80 * BLOCK, 1e-4, P=1, faithful=False)
    assert np.abs(o1["x"] - o["x"]).max() / np.abs(o["x"]).max() < 1e-10   # P only changes summation order
