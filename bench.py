#!/usr/bin/env python
# -*- coding: utf-8 -*-
"""bench.py -- lasso sweeps/s on B200 (BASELINE.json metric) with roofline and CPU baseline.

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path

A "step" is one sweep: BLOCK block steps, each reading the column block A_m for the block
gradient A_m^T r and again for A_m D (prox, line search, error and updates fused).  The
workload is BASELINE.json configs[1]: dense fp32 lasso 10,000 x 100,000, 100 column
blocks, synthetic Gaussian A with unit-l2 rows (reference recipe parameters.py:20-33),
generated on the device.  A (4 GB) is far larger than L2 (126 MB), so nothing is flushed
between timed steps.

`value`  : sweeps/s with everything resident in HBM (K fused launches, CUDA events).
`e2e`    : sweeps/s through the reference-facing API, ClassLasso.run(): host b -> device,
           the solve, x -> host, all inside the timed region (A is uploaded once by
           GPU_Calculation(A, BLOCK), exactly as in the reference driver cpu_vs_gpu.py:131).
`roofline`: algorithmic bytes of a sweep / kernel time vs the
           measured HBM copy peak of MEASURED_PEAKS.json.
`cpu_baseline`: the oracle port of the reference's CPU path on a bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "lasso_sweeps_per_s"
UNIT = "sweeps/s"
# BASELINE.json configs[1]
C2 = dict(N=10000, K=100000, BLOCK=100, den=0.01, seed=2, dtype="float")
FALLBACK_HBM_GBS = 6650.0     # /opt/skills/guides/B200_PROFILING.md fallback


def sweep_bytes(N, K, BLOCK, s):
    return 2 * N * K * s + 5 * BLOCK * N * s + 4 * K * s


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def ncu_traffic(config, layout):
    """dram bytes per sweep from the committed ncu --set full capture of this workload, if any"""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    key = layout if config == "c2" else "%s_%s" % (config, layout)
    try:
        with open(path) as f:
            return json.load(f).get(key)
    except Exception:
        return None


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled from a thread
    every ~4 ms (nvidia_ml_py), nvidia-smi -lms as a fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []          # (time, sm_mhz, max_mhz, [reasons])
        self.proc = None
        self.nvml = None
        self.stop_flag = False

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.rows.append((time.time(), sm, self.max_mhz, [n for n, b in bits.items() if mask & b]))
            except Exception:
                pass
            time.sleep(0.004)

    def _pump(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.strip().split(",")]
            try:
                self.rows.append((time.time(), float(f[0]), float(f[1]),
                                  [n for n, v in zip(names, f[2:6]) if v.lower().startswith("active")]))
            except Exception:
                continue

    def stop(self, t0, t1):
        if self.nvml is None and self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml / nvidia-smi unavailable"], "samples": 0}
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1.0)
        else:
            time.sleep(0.05)
            self.proc.terminate()
        # NVML queries can stall while the process sits in a stream synchronisation, so samples
        # up to 10 ms before / after the region count as "during" (clocks do not move that fast)
        inside = [r for r in self.rows if t0 - 0.010 <= r[0] <= t1 + 0.010]
        rows = inside or self.rows          # (a region shorter than the sampling period: nearest samples)
        reasons = sorted({n for r in rows for n in r[3]})
        return {"sm_mhz": float(np.median([r[1] for r in rows])) if rows else None,
                "sm_max_mhz": float(max(r[2] for r in rows)) if rows else None,
                "reasons": reasons, "samples": len(inside),
                "source": "nvml, 4 ms period" if self.nvml is not None else "nvidia-smi -lms 20"}


# ------------------------------------------------------------------------------ data
def make_device_instance(torch, device, N, K, BLOCK, den, seed, torch_dtype, ld, layout="row",
                         dist=None, rank=0):
    """reference recipe (parameters.py:20-33) on the device: Gaussian A, unit-l2 rows,
    sparse x_true, b = A x_true + e, mu = 0.1 |A^T b|_inf.  Returns (store, b, mu).
    With ``dist`` (multi-GPU) every rank generates its own column shard (K local columns,
    seed + rank); row norms, b and mu are combined over the ranks, the noise e is the same
    on every rank."""
    w = K // BLOCK
    gen = torch.Generator(device=device)
    gen.manual_seed(seed + 1000 * rank)
    gen_common = torch.Generator(device=device)
    gen_common.manual_seed(seed + 77)
    rows = N if layout == "row" else w
    store = torch.zeros((BLOCK, rows, ld), dtype=torch_dtype, device=device)
    sq = torch.zeros(N, dtype=torch.float64, device=device)
    for m in range(BLOCK):
        blk = torch.randn((N, w), generator=gen, dtype=torch_dtype, device=device)
        sq += (blk.double() ** 2).sum(dim=1)
        if layout == "row":
            store[m, :, :w] = blk
        else:
            store[m, :, :N] = blk.t()
    if dist is not None:
        dist.all_reduce(sq)
    inv = (1.0 / sq.sqrt()).to(torch_dtype)
    x_true = torch.randn(K, generator=gen, dtype=torch.float64, device=device)
    x_true *= (torch.rand(K, generator=gen, dtype=torch.float64, device=device) < den)
    b = torch.zeros(N, dtype=torch.float64, device=device)
    for m in range(BLOCK):
        if layout == "row":
            store[m, :, :w] *= inv[:, None]
            b += store[m, :, :w].double() @ x_true[m * w:(m + 1) * w]
        else:
            store[m, :, :N] *= inv[None, :]
            b += store[m, :, :N].double().t() @ x_true[m * w:(m + 1) * w]
    if dist is not None:
        dist.all_reduce(b)
    b += 1e-2 * torch.randn(N, generator=gen_common, dtype=torch.float64, device=device)
    gmax = 0.0
    for m in range(BLOCK):
        if layout == "row":
            g = store[m, :, :w].double().t() @ b
        else:
            g = store[m, :, :N].double() @ b
        gmax = max(gmax, float(g.abs().max()))
    if dist is not None:
        t = torch.tensor([gmax], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gmax = float(t.item())
    return store, b.cpu().numpy().reshape(-1, 1), 0.1 * gmax


# ------------------------------------------------------------------------------ CPU arm
def cpu_sample(threads_note=True):
    """the oracle port (NumPy/BLAS, all host threads) on a bounded C2-shaped sample:
    10,000 x 10,000 fp64, 10 blocks of w=1000 (1/10 of C2's columns)."""
    from oracle import lasso_oracle as orc
    N, K, BLOCK = 10000, 10000, 10
    rng = np.random.RandomState(2)
    A = rng.standard_normal((N, K))
    A /= np.linalg.norm(A, axis=1, keepdims=True)
    xt = rng.standard_normal((K, 1)) * (rng.rand(K, 1) < 0.01)
    b = A @ xt + 1e-2 * rng.standard_normal((N, 1))
    mu = 0.1 * np.max(np.abs(A.T @ b))
    return orc, A, b, mu, N, K, BLOCK


def run_cpu(orc, A, b, mu, BLOCK, sweeps):
    """seconds spent in the iteration loop itself (the oracle times it like the reference's
    run(), lasso.py:98,164); the one-off column norms and the final objective are outside"""
    o = orc.lasso_oracle(A, b, mu, BLOCK, BLOCK * sweeps, None, faithful=False)
    assert o["iters"] == BLOCK * sweeps
    return o["elapsed"]


def cpu_baseline_block(seconds=12.0):
    """about `seconds` of CPU work: as many sweeps of the sample as fit (oracle time_limit)"""
    orc, A, b, mu, N, K, BLOCK = cpu_sample()
    run_cpu(orc, A, b, mu, BLOCK, 1)
    o = orc.lasso_oracle(A, b, mu, BLOCK, BLOCK * 100000, None, faithful=False, time_limit=seconds)
    sweeps, dt = o["iters"] / float(BLOCK), o["elapsed"]
    frac = K / C2["K"]
    sample_sweeps_per_s = sweeps / dt
    return {"value": sample_sweeps_per_s * frac, "unit": UNIT, "cores": os.cpu_count(),
            "kind": "port",
            "sample": "oracle port (NumPy fp64, BLAS threads) on 10000x10000, 10 blocks of w=1000 "
                      "(1/10 of the C2 columns), %.1f sweeps in %.2f s = %.2f sample-sweeps/s; value is "
                      "scaled by bytes (x%.2f) to the full 10000x100000 sweep" % (sweeps, dt, sample_sweeps_per_s, frac)}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    orc, A, b, mu, N, K, BLOCK = cpu_sample()
    for _ in range(max(args.warmup, 1)):
        run_cpu(orc, A, b, mu, BLOCK, 1)
    dt = 0.0
    for _ in range(args.steps):
        dt += run_cpu(orc, A, b, mu, BLOCK, 1)
    frac = K / C2["K"]
    value = args.steps / dt * frac
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps / frac * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": "dense lasso 10000x100000, 100 column blocks (C2); each step = one sweep "
                               "of a 10000x10000 / 10-block sample, scaled by bytes"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                         "sample": "one sweep of a 10000x10000, 10-block fp64 sample per step; NumPy/BLAS "
                                   "with all host threads; scaled x%.2f by bytes to C2" % frac},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def c1_time_to_eps(device_index):
    """BASELINE.json configs[0]: the reference's default instance (cpu_vs_gpu.py:57-74: N=1024, K=4096,
    BLOCK=2, den=0.4, fp64, ERR_BOUND=1e-4, seed 1234) solved through ClassLasso.run() from host
    arrays; reported beside the 52.45 s the unmodified ClassLassoCPU took in the survey
    (BASELINE.md section 2).  Never fatal for the bench line."""
    try:
        from convex_optimization_b200 import lasso, parameters
        from convex_optimization_b200.gpu_calculation import GPU_Calculation

        class Cal(GPU_Calculation):
            TYPE = "double"
            LAYOUT = "row"
            DEVICE = device_index
        A, _, b, mu = parameters.parameters(1024, 4096, 0.4, False, False, SILENCE=True, seed=1234)
        cal = Cal(A, 2)
        solver = lasso.ClassLasso(cal, cal.diag_ATA, A, b, mu, 2, 1000)
        solver.run(1e-4, SILENCE=True)                      # warm-up (module load, first launch)
        t0 = time.time()
        solver.run(1e-4, SILENCE=True)
        dt = time.time() - t0
        return {"workload": "default instance 1024x4096, BLOCK=2, fp64, ERR_BOUND=1e-4 (cpu_vs_gpu.py:57-74)",
                "seconds": dt, "iterations": int(solver.iters), "nnz_x": int(np.count_nonzero(solver.x)),
                "api": "ClassLasso.run() (host b in, x out)",
                "reference_cpu_seconds_survey": 52.45, "reference_iterations_survey": 128}
    except Exception as e:                                  # pragma: no cover
        sys.stderr.write("c1_time_to_eps skipped: %r\n" % (e,))
        return None


# ------------------------------------------------------------------------------ GPU arm
def main_gpu(args):
    import torch
    import torch.distributed as dist
    from convex_optimization_b200 import _lib, lasso
    from convex_optimization_b200.gpu_calculation import GPU_Calculation
    import ctypes

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)

    cfg = dict(C2)
    if args.small:
        cfg.update(N=2000, K=20000, BLOCK=20)
    if args.config == "c3":       # BASELINE.json configs[2]: fp64 20,000 x 200,000 (32 GB), not the bench line
        cfg.update(N=20000, K=200000, BLOCK=100, dtype="double", seed=3)
    elif args.config == "c4shard":  # one GPU's share of configs[3] on 8 GPUs: 50,000 x 125,000 fp32 (25 GB)
        cfg.update(N=50000, K=125000, BLOCK=100, dtype="float", seed=4)
    if args.shape:              # diagnostics only: N,K,BLOCK[,dtype]
        f = args.shape.split(",")
        cfg.update(N=int(f[0]), K=int(f[1]), BLOCK=int(f[2]))
        if len(f) > 3:
            cfg["dtype"] = f[3]
    N, K, BLOCK = cfg["N"], cfg["K"], cfg["BLOCK"]
    s = 4 if cfg["dtype"] == "float" else 8
    layout = args.layout

    class Cal(GPU_Calculation):
        TYPE = cfg["dtype"]
        LAYOUT = layout
        DEVICE = local_rank
    ld = Cal.padded_ld(N, K, BLOCK)
    tdt = torch.float32 if cfg["dtype"] == "float" else torch.float64
    # N GPUs: the instance grows with N (weak scaling): K = 100000*N columns, 100 blocks of width
    # 1000*N, column slice `rank` of every block per GPU, i.e. one C2-sized shard per GPU
    if args.instance == "philox":
        # the library's generator: the same global matrix for every world size
        from convex_optimization_b200 import parameters as pm
        cal, _, b, mu = pm.parameters_device(N, K * world, BLOCK, cfg["den"], cfg["seed"], gpu_cal_cls=Cal)
    else:
        store, b, mu = make_device_instance(torch, device, N, K, BLOCK, cfg["den"], cfg["seed"], tdt, ld, layout,
                                            dist if world > 1 else None, rank)
        cal = Cal.from_device_blocks(store, N, K, BLOCK)
    if world > 1:
        from convex_optimization_b200 import distributed as dd
        dd.connect(cal)
    if args.slot_bytes or args.inflight:
        cal.set_tuning(args.slot_bytes, args.inflight)
    lib, ctx = cal._lib, cal.ctx
    if args.dbg:
        _lib.check(lib.b200l_debug_flags(ctx, args.dbg))
    # the library launches on the stream the CUDA events below are recorded on
    stream = torch.cuda.current_stream(device)
    _lib.check(lib.b200l_ctx_set_stream(ctx, ctypes.c_void_p(stream.cuda_stream)))
    d_ATA = cal.diag_ATA
    geo = cal.run_config()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # ---- value: K sweeps, resident, device-timed ------------------------------------
    bb = np.ascontiguousarray(b.reshape(-1))
    _lib.check(lib.b200l_set_problem(ctx, _lib.dptr(bb)))

    # One launch of the persistent kernel per timed sweep (the unit the roofline and the committed
    # ncu launch list refer to); --single-launch runs the K sweeps in one launch, like a solve.
    def sweeps(n):
        if not args.single_launch:
            for _ in range(n):
                _lib.check(lib.b200l_run(ctx, None, BLOCK, float(mu), -1.0, None, None, None, None, None))
        elif n > 0:
            _lib.check(lib.b200l_run(ctx, None, BLOCK * n, float(mu), -1.0, None, None, None, None, None))

    # (the sampler starts before the warm-up: its start-up must not delay rank 0 behind the others)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    sweeps(args.warmup)
    barrier()
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    ev0.record()
    sweeps(args.steps)
    ev1.record()
    barrier()
    t_wall1 = time.time()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * 1e3 / ms_per_step

    # the timed region is K launches of the one kernel on this stream, so its average launch
    # duration is ms_per_step; the same launch timed alone (events inside the library, a host
    # synchronisation after each) is reported beside it
    kms = ctypes.c_double()
    ktimes = []
    for _ in range(min(args.steps, 10)):
        _lib.check(lib.b200l_run(ctx, None, BLOCK, float(mu), -1.0, None, None, None, None,
                                 ctypes.byref(kms)))
        ktimes.append(kms.value)
    sweeps_per_launch = args.steps if args.single_launch else 1
    n_launches = args.steps // sweeps_per_launch
    kernel_ms = ms / n_launches                  # average duration of the launches of the timed region
    kernel_ms_alone = float(np.mean(ktimes))
    obj = ctypes.c_double()
    _lib.check(lib.b200l_objective(ctx, float(mu), ctypes.byref(obj)))
    obj_bench = obj.value

    # ---- e2e: through ClassLasso.run(), host b in, host x out -------------------------
    e2e_sweeps = args.e2e_sweeps
    class HostShape:            # the solver only needs A.shape on the fused path (lasso.py:32)
        shape = (N, K)
    solver = lasso.ClassLasso(cal, d_ATA, HostShape, b, mu, BLOCK, BLOCK * e2e_sweeps)
    for _ in range(2):
        solver.run(SILENCE=True)
    barrier()
    t0 = time.time()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        solver.run(SILENCE=True)
    barrier()
    e2e_dt = time.time() - t0
    if world > 1:
        t = torch.tensor([e2e_dt], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_dt = float(t.item())
    e2e_value = world * e2e_steps * e2e_sweeps / e2e_dt
    nnz = int(np.count_nonzero(solver.x))

    # ---- time-to-eps: cold start (x = 0) to the reference's stop rule (lasso.py:141-150) ----
    tte = None
    if args.eps > 0:
        _lib.check(lib.b200l_reset(ctx))
        steps_done, stopped = ctypes.c_int64(), ctypes.c_int32()
        _lib.check(lib.b200l_run(ctx, None, BLOCK * args.eps_max_sweeps, float(mu), float(args.eps), None, None,
                                 ctypes.byref(steps_done), ctypes.byref(stopped), ctypes.byref(kms)))
        _lib.check(lib.b200l_objective(ctx, float(mu), ctypes.byref(obj)))
        tte = {"eps": args.eps, "ms": kms.value, "sweeps": steps_done.value / BLOCK,
               "reached": bool(stopped.value), "objective": obj.value,
               "note": "device time of one launch from x = 0 until every block of a sweep has error_crit < eps"}

    if rank == 0:
        W = sweep_bytes(N, K, BLOCK, s)
        peak, peak_src = measured_peak()
        achieved = W * sweeps_per_launch / (kernel_ms * 1e-3) / 1e9
        traffic_sweep = None if args.small else ncu_traffic(args.config, layout)
        traffic = None if traffic_sweep is None else traffic_sweep * sweeps_per_launch
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if s == 4 else "f64",
            "data": "synthetic",
            "config": {"workload": "dense %s lasso %dx%d, %d column blocks, %s A layout; 1 step = 1 sweep, the timed "
                                   "sweeps run in %d launch(es) of the persistent kernel; "
                                   "A=%.1f GB per GPU >> L2 so no flush between steps%s"
                                   % ("fp32" if s == 4 else "fp64", N, K * world, BLOCK, layout, n_launches,
                                      N * K * s / 1e9,
                                      ("; column-sharded over %d GPUs (slice g of every block on GPU g, partial "
                                       "A_m D summed in-kernel over NVLink peer memory); value counts C2-sized "
                                       "shard sweeps: %d per sweep of the %dx%d instance"
                                       % (world, world, N, K * world)) if world > 1 else ""),
                       "launch": geo, "objective_after_bench": obj_bench},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic,
                         "kernel": "lasso_fused<%s,1,%s>" % ("float" if s == 4 else "double",
                                                             "TRANS" if layout == "transposed" else "ROWMAJOR"),
                         "kernel_ms_per_launch": kernel_ms, "sweeps_per_launch": sweeps_per_launch,
                         "ms_per_single_sweep_launch_alone": kernel_ms_alone,
                         "algorithmic_bytes_per_launch": W * sweeps_per_launch,
                         "algorithmic_bytes_per_sweep": W, "traffic_per_sweep": traffic_sweep,
                         "peak_source": peak_src,
                         "frac_of_nominal_8TBs": achieved / 8000.0,
                         "dram_single_pass_GBs": (W - N * K * s) * sweeps_per_launch / (kernel_ms * 1e-3) / 1e9,
                         "note": "algorithmic bytes count A twice per sweep (SURVEY 8d); the second pass of a "
                                 "40 MB block is served by L2, so DRAM traffic is about half of it (see traffic)"},
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": int(N * 8 + 4 * BLOCK * e2e_sweeps),
                    "d2h_bytes_per_step": int(K * 8 + 24),
                    "sweeps_per_call": e2e_sweeps, "calls": e2e_steps, "nnz_x": nnz,
                    "api": "ClassLasso.run() (host b -> device, fused solve, x -> host)"},
            "gpu_launches": n_launches,
            "clocks": clocks,
        }
        if tte:
            line["time_to_eps"] = tte
        if world == 1 and not args.small and args.eps > 0:
            c1 = c1_time_to_eps(local_rank)
            if c1:
                line["c1_time_to_eps"] = c1
        if not args.no_cpu and world == 1:
            line["cpu_baseline"] = cpu_baseline_block()
        print(json.dumps(line))
    if world > 1:
        dd.disconnect(cal)
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--layout", default="row", choices=["row", "transposed"])
    ap.add_argument("--small", action="store_true", help="2000x20000 debug size (not a bench value)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--e2e-sweeps", type=int, default=5)
    ap.add_argument("--instance", default="torch", choices=["torch", "philox"],
                    help="how the synthetic instance is generated on the device: torch RNG per rank (default) or "
                         "b200l_gen_gaussian (Philox keyed by seed/row/global column)")
    ap.add_argument("--single-launch", action="store_true",
                    help="run the K timed sweeps in one kernel launch instead of one launch per sweep")
    ap.add_argument("--eps", type=float, default=1e-4, help="ERR_BOUND of the time-to-eps leg (0 = skip)")
    ap.add_argument("--eps-max-sweeps", type=int, default=2000)
    ap.add_argument("--slot-bytes", type=int, default=0)
    ap.add_argument("--inflight", type=int, default=0)
    ap.add_argument("--dbg", type=int, default=0, help="diagnostic flags of b200l_debug_flags (not for bench values)")
    ap.add_argument("--shape", default="", help="diagnostics: N,K,BLOCK[,float|double] instead of a named config")
    ap.add_argument("--config", default="c2", choices=["c2", "c3", "c4shard"],
                    help="c2 is the bench workload; the others are reported in DESIGN.md only")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return main_reference(args)
    return main_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
