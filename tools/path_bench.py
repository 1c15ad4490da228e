#!/usr/bin/env python
"""Regularisation path with warm starts (BASELINE.json configs[4]: 20 decreasing lambdas on a 20,000 x
500,000 fp32 instance across 8 B200), on 1 GPU or under torchrun on several:

    python tools/path_bench.py                                   # 1 GPU (the 40 GB matrix fits in HBM)
    torchrun --nproc-per-node 8 ... tools/path_bench.py          # column-sharded over 8 GPUs

The instance is generated on the devices with the library's Philox generator (parameters_device): the
SAME global matrix, b and lambda grid for every world size, so the per-lambda iteration counts and
objectives of the 1- and the 8-GPU run can be compared (--check reduced.json does that against a file
written by an earlier run with --out).  20 log-spaced lambdas from 0.9 to 0.009 lambda_max, each solve
warm-started from the previous one (x and the running residual stay on the devices), ERR_BOUND 1e-4;
--cold also times every lambda from x = 0.  Prints one JSON line (rank 0).  The reference has no
counterpart: it always starts from x = 0 with one fixed mu (lasso.py:34,89)."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from convex_optimization_b200 import distributed as dd
    from convex_optimization_b200 import parameters as pm
    from convex_optimization_b200 import path as bpath
    from convex_optimization_b200.gpu_calculation import GPU_Calculation
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=20000)
    ap.add_argument("--K", type=int, default=500000, help="GLOBAL column count")
    ap.add_argument("--block", type=int, default=250)
    ap.add_argument("--nlambda", type=int, default=20)
    ap.add_argument("--max-sweeps", type=int, default=400)
    ap.add_argument("--seed", type=int, default=5)
    ap.add_argument("--cold", action="store_true", help="also time the path from cold starts")
    ap.add_argument("--out", default="", help="write the result (rank 0) to this JSON file")
    ap.add_argument("--check", default="", help="compare iterations / objectives per lambda with this earlier result")
    args = ap.parse_args()
    N, K, BLOCK = args.N, args.K, args.block
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    class Cal(GPU_Calculation):
        TYPE = "float"
        DEVICE = local
    t0 = time.time()
    cal, _, b, mu = pm.parameters_device(N, K, BLOCK, 0.01, args.seed, gpu_cal_cls=Cal)
    gen_s = time.time() - t0
    transport = dd.connect(cal) if world > 1 else "single"
    mus = bpath.lambda_grid(mu / 0.1, n=args.nlambda)
    if world > 1:
        dist.barrier()
    t0 = time.time()
    res = bpath.lasso_path(cal, b, mus, BLOCK, BLOCK * args.max_sweeps, 1e-4)
    torch.cuda.synchronize()
    wall = time.time() - t0
    nnz = np.array([int(np.count_nonzero(r["x"])) for r in res], dtype=np.int64)
    kms = np.array([r["kernel_ms"] for r in res])
    if world > 1:
        t = torch.from_numpy(nnz).cuda()
        dist.all_reduce(t)
        nnz = t.cpu().numpy()
        t = torch.from_numpy(kms).cuda()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kms = t.cpu().numpy()
    out = {"config": "lambda path %dx%d fp32, %d blocks, %d lambdas 0.9..0.009 lambda_max, eps 1e-4, warm starts, "
                     "%d GPU(s), transport %s, Philox instance seed %d" % (N, K, BLOCK, args.nlambda, world, transport, args.seed),
           "n_gpus": world, "generate_s": gen_s, "wall_s": wall, "kernel_s": float(kms.sum()) * 1e-3,
           "sweeps": [r["iters"] // BLOCK for r in res], "iters": [r["iters"] for r in res],
           "stopped": [r["stopped"] for r in res], "time_to_eps_ms": [round(float(v), 2) for v in kms],
           "nnz": [int(v) for v in nnz], "objective": [r["objective"] for r in res], "mu": [float(m) for m in mus],
           "launch": cal.run_config()}
    if args.cold:
        cold_ms, cold_sweeps = [], []
        for m in mus:
            r = bpath.lasso_path(cal, b, [m], BLOCK, BLOCK * args.max_sweeps, 1e-4, collect_x=False)[0]
            cold_ms.append(round(r["kernel_ms"], 2))
            cold_sweeps.append(r["iters"] // BLOCK)
        out["cold_time_to_eps_ms"] = cold_ms
        out["cold_sweeps"] = cold_sweeps
        out["cold_kernel_s"] = sum(cold_ms) * 1e-3
    rc = 0
    if args.check and rank == 0:
        with open(args.check) as f:
            ref = json.load(f)
        same_iters = ref["iters"] == out["iters"]
        rel_obj = max(abs(a - c) / abs(c) for a, c in zip(out["objective"], ref["objective"]))
        same_nnz = ref["nnz"] == out["nnz"]
        out["check"] = {"against": "%s (%d GPU(s))" % (args.check, ref["n_gpus"]), "same_iterations_per_lambda": same_iters,
                        "max_rel_objective_diff": rel_obj, "same_support_sizes": same_nnz,
                        "note": "the sharded run sums A_m D over the ranks in a different order than one GPU does: iteration "
                                "counts can differ by a sweep where the fp32 error sits at the threshold"}
        if rel_obj > 1e-5:
            rc = 1
    if rank == 0:
        print(json.dumps(out))
        if args.out:
            with open(args.out, "w") as f:
                json.dump(out, f)
    if world > 1:
        dd.disconnect(cal)
        dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
