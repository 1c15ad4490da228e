#!/usr/bin/env python
"""Regularisation path with warm starts at BASELINE.json config-5 size on ONE B200 (the 40 GB
matrix fits in 180 GB of HBM): 20 log-spaced lambdas from 0.9 to 0.009 lambda_max, each solve
warm-started from the previous one, ERR_BOUND 1e-4.  Prints one JSON line with the time-to-eps of
every lambda (device time of the fused kernel), the sweeps it took and the support size; also runs
the same path from cold starts for the comparison.  Reported in DESIGN.md; not the bench line."""
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
    from bench import make_device_instance
    from convex_optimization_b200 import path as bpath
    from convex_optimization_b200.gpu_calculation import GPU_Calculation
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=20000)
    ap.add_argument("--K", type=int, default=500000)
    ap.add_argument("--block", type=int, default=250)
    ap.add_argument("--nlambda", type=int, default=20)
    ap.add_argument("--max-sweeps", type=int, default=400)
    ap.add_argument("--cold", action="store_true", help="also time the path from cold starts")
    args = ap.parse_args()
    N, K, BLOCK = args.N, args.K, args.block
    dev = torch.device("cuda", 0)

    class Cal(GPU_Calculation):
        TYPE = "float"
    ld = Cal.padded_ld(N, K, BLOCK)
    t0 = time.time()
    store, b, mu = make_device_instance(torch, dev, N, K, BLOCK, 0.01, 5, torch.float32, ld)
    cal = Cal.from_device_blocks(store, N, K, BLOCK)
    gen_s = time.time() - t0
    mus = bpath.lambda_grid(mu / 0.1, n=args.nlambda)
    t0 = time.time()
    res = bpath.lasso_path(cal, b, mus, BLOCK, BLOCK * args.max_sweeps, 1e-4)
    wall = time.time() - t0
    out = {"config": "lambda path %dx%d fp32, %d blocks, %d lambdas 0.9..0.009 lambda_max, eps 1e-4, warm starts, 1 GPU"
                     % (N, K, BLOCK, args.nlambda),
           "generate_s": gen_s, "wall_s": wall, "kernel_s": sum(r["kernel_ms"] for r in res) * 1e-3,
           "sweeps": [r["iters"] // BLOCK for r in res], "stopped": [r["stopped"] for r in res],
           "time_to_eps_ms": [round(r["kernel_ms"], 2) for r in res],
           "nnz": [int(np.count_nonzero(r["x"])) for r in res],
           "objective": [r["objective"] for r in res], "launch": cal.run_config()}
    if args.cold:
        cold_ms, cold_sweeps = [], []
        for m in mus:
            r = bpath.lasso_path(cal, b, [m], BLOCK, BLOCK * args.max_sweeps, 1e-4, collect_x=False)[0]
            cold_ms.append(round(r["kernel_ms"], 2))
            cold_sweeps.append(r["iters"] // BLOCK)
        out["cold_time_to_eps_ms"] = cold_ms
        out["cold_sweeps"] = cold_sweeps
        out["cold_kernel_s"] = sum(cold_ms) * 1e-3
    print(json.dumps(out))


if __name__ == "__main__":
    main()
