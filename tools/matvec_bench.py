#!/usr/bin/env python
"""Roofline of the legacy mat-vec entry points (the step-wise path of ClassLassoCB_v1, of hooked
subclasses and of DEBUG mode: gpu_calculation.py:264-292 in the reference) on C2 blocks:
b200l_gemv_t_dev (g = A_m^T r) and b200l_gemv_n_dev (q = A_m d) on device vectors, CUDA events, cycling
through all 100 blocks so that every call streams its 40 MB block from HBM (A = 4 GB >> L2), plus
diag(A^T A) over the whole matrix (one pass).  Prints one JSON line.  GPU only."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from bench import CONFIGS, make_device_instance, measured_peak
    from convex_optimization_b200 import _lib
    from convex_optimization_b200.gpu_calculation import GPU_Calculation
    ap = argparse.ArgumentParser()
    ap.add_argument("--layout", default="row", choices=["row", "transposed"])
    ap.add_argument("--dtype", default="float", choices=["float", "double"])
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--dbg", type=int, default=0, help="b200l_debug_flags (262144: load-batch kernels only)")
    args = ap.parse_args()
    cfg = CONFIGS["c2"]
    N, K, BLOCK = cfg["N"], cfg["K"], cfg["BLOCK"]
    w = K // BLOCK
    es = 4 if args.dtype == "float" else 8
    dev = torch.device("cuda", 0)

    class Cal(GPU_Calculation):
        TYPE = args.dtype
        LAYOUT = args.layout
    ld = Cal.padded_ld(N, K, BLOCK)
    tdt = torch.float32 if args.dtype == "float" else torch.float64
    store, b, mu = make_device_instance(torch, dev, N, K, BLOCK, 0.01, 2, tdt, ld, args.layout)
    cal = Cal.from_device_blocks(store, N, K, BLOCK)
    if args.dbg:
        _lib.check(cal._lib.b200l_debug_flags(cal.ctx, args.dbg))
    stream = torch.cuda.current_stream(dev)
    _lib.check(cal._lib.b200l_ctx_set_stream(cal.ctx, ctypes.c_void_p(stream.cuda_stream)))
    r = torch.randn(max(N, cal.ld) + 64, dtype=torch.float64, device=dev)
    d = torch.randn(w + 64, dtype=torch.float64, device=dev)
    g = torch.empty(w + 64, dtype=torch.float64, device=dev)
    q = torch.empty(N + 64, dtype=torch.float64, device=dev)
    peak, src = measured_peak()
    out = {"config": "C2 blocks: %s %dx%d per block, %s layout, 100 blocks cycled (HBM-resident)" % (args.dtype, N, w, args.layout),
           "peak_GBs": peak, "peak_source": src, "block_bytes": N * w * es}

    def timed(fn, ptr_in, ptr_out, fn2=None, ptr_in2=None, ptr_out2=None):
        def sweep():
            for m in range(BLOCK):
                _lib.check(fn(cal.ctx, m, ctypes.c_void_p(ptr_in), ctypes.c_void_p(ptr_out)))
                if fn2 is not None:         # the step-wise solver's pattern: A_m^T r, then A_m d, per block
                    _lib.check(fn2(cal.ctx, m, ctypes.c_void_p(ptr_in2), ctypes.c_void_p(ptr_out2)))
        sweep()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.reps):
            sweep()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (args.reps * BLOCK)
        rec = {"us_per_call_from_python": us, "GBs_from_python": N * w * es / us / 1e3,
               "note": "one ctypes call per block from a Python loop: the host issues a call every ~15 us, the GPU idles in between"}
        # the same 100 calls captured once in a CUDA graph and replayed: device time of the kernels themselves
        try:
            side = torch.cuda.Stream()
            _lib.check(cal._lib.b200l_ctx_set_stream(cal.ctx, ctypes.c_void_p(side.cuda_stream)))
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                sweep()
                side.synchronize()
                with torch.cuda.graph(graph, stream=side):
                    sweep()
            graph.replay()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.reps):
                graph.replay()
            e1.record()
            torch.cuda.synchronize()
            usg = e0.elapsed_time(e1) * 1e3 / (args.reps * BLOCK)
            rec.update({"us_per_call": usg, "GBs": N * w * es / usg / 1e3, "frac": N * w * es / usg / 1e3 / peak,
                        "how": "100 calls (one per block) captured in a CUDA graph, replayed %d times" % args.reps})
        except Exception as e:
            rec["graph"] = "capture failed: %r" % (e,)
        finally:
            _lib.check(cal._lib.b200l_ctx_set_stream(cal.ctx, ctypes.c_void_p(stream.cuda_stream)))
        return rec
    out["gemv_t (A_m^T r)"] = timed(cal._lib.b200l_gemv_t_dev, r.data_ptr(), g.data_ptr())
    out["gemv_n (A_m d)"] = timed(cal._lib.b200l_gemv_n_dev, d.data_ptr(), q.data_ptr())
    # both mat-vecs per block, alternating like the step-wise solver (the second one finds the block in L2):
    # us_per_call here is per PAIR of calls
    out["gemv_t + gemv_n per block, alternating (us per pair)"] = timed(
        cal._lib.b200l_gemv_t_dev, r.data_ptr(), g.data_ptr(), cal._lib.b200l_gemv_n_dev, d.data_ptr(), q.data_ptr())
    # diag(A^T A): one pass over the whole matrix (re-bind to invalidate the cached diagonal)
    dh = np.empty((BLOCK, w, 1))
    ts = []
    for _ in range(3):
        _lib.check(cal._lib.b200l_ctx_bind_A(cal.ctx, ctypes.c_void_p(store.data_ptr())))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(cal._lib.b200l_diag_ata(cal.ctx, _lib.dptr(dh)))
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    out["diag_ATA (whole matrix, incl. D2H of K doubles)"] = {"ms": min(ts), "GBs": N * K * es / min(ts) / 1e6,
                                                              "frac": N * K * es / min(ts) / 1e6 / peak}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
