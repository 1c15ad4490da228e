# -*- coding: utf-8 -*-
"""Run the UNMODIFIED reference ``ClassLassoCPU`` (build container only) and mint
golden vectors -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

``/root/reference`` exists only in the build container, so this script is run
there once and its outputs are committed under ``tests/golden/``:

    python oracle/run_reference.py            # writes tests/golden/*.npz

How the reference is executed without touching it:
 * ``lasso.py`` imports pycuda/skcuda at module top (lasso.py:13-16), which are
   absent here; empty stub modules are placed in ``sys.modules`` first.
   ``ClassLassoCPU`` (lasso.py:25-169) uses none of them.
 * the data seed is ``int(time())`` (parameters.py:17); ``parameters.time`` is
   patched to return the pinned seed.
 * ``run()`` discards x (lasso.py:167-169).  ``x_block[m]`` handed to the
   ``err_record`` hook is a view of the (BLOCK,w,1) array that is updated in
   place (lasso.py:90,139,153), so its ``.base`` after ``run()`` is the final x.
   The ``debug`` hook (lasso.py:138) sees the step size of every iteration.
"""
import os
import sys
import types

import numpy as np

REF = os.environ.get("B200L_REFERENCE_DIR", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def import_reference():
    if not os.path.isdir(REF):
        raise RuntimeError("reference tree %s is not present" % REF)
    for name in ("pycuda", "pycuda.gpuarray", "pycuda.autoinit",
                 "pycuda.elementwise", "pycuda.driver", "pycuda.compiler",
                 "skcuda", "skcuda.cublas"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["pycuda"].gpuarray = sys.modules["pycuda.gpuarray"]
    sys.modules["pycuda.elementwise"].ElementwiseKernel = object
    sys.modules["skcuda"].cublas = sys.modules["skcuda.cublas"]
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import lasso as ref_lasso
    import parameters as ref_parameters
    import cpu_calculation as ref_cpu
    assert os.path.dirname(os.path.abspath(ref_lasso.__file__)) == REF
    return ref_lasso, ref_parameters, ref_cpu


def run_reference(N, K, BLOCK, P, den, seed, ITER_MAX, ERR_BOUND):
    ref_lasso, ref_parameters, ref_cpu = import_reference()
    ref_parameters.time = lambda: seed
    A, x_true, b, mu = ref_parameters.parameters(N, K, den, False, False,
                                                 SILENCE=True)
    A_block_p = ref_cpu.A_bp_get(A, BLOCK, P)
    d_ATA = ref_cpu.fun_diag_ATA(A_block_p)

    class Hooked(ref_lasso.ClassLassoCPU):
        x_base = None
        gammas = []

        def err_record(self, err_iter, s13, x_block_m, t):
            Hooked.x_base = x_block_m.base
            ref_lasso.ClassLassoCPU.err_record(self, err_iter, s13,
                                               x_block_m, t)

        def debug(self, s13, x_block_m, x, t, m, r):
            Hooked.gammas.append(float(r))

    Hooked.gammas = []
    solver = Hooked(A_block_p, d_ATA, A, b, mu, BLOCK, P, ITER_MAX)
    err_iter = np.zeros(ITER_MAX)
    elapsed = solver.run(ERR_BOUND, err_iter=err_iter, SILENCE=True)
    iters = len(Hooked.gammas)
    x = np.array(Hooked.x_base).reshape(K, 1)
    stopped = iters < ITER_MAX or False
    obj = 0.5 * float(np.sum((A @ x - b) ** 2)) + mu * float(np.sum(np.abs(x)))
    return dict(A=A, b=np.asarray(b), mu=float(mu), x=x,
                err=err_iter[:iters].copy(), gamma=np.array(Hooked.gammas),
                iters=iters, elapsed=float(elapsed), objective=obj,
                d_ATA=np.asarray(d_ATA), x_true=np.asarray(x_true.todense()))


# name -> (N, K, BLOCK, P, den, seed, ITER_MAX, ERR_BOUND)
CASES = {
    "g_64x256_b1_p1": (64, 256, 1, 1, 0.1, 11, 300, 1e-4),
    "g_128x512_b2_p4": (128, 512, 2, 4, 0.1, 12, 400, 1e-4),
    "g_256x1024_b8_p4": (256, 1024, 8, 4, 0.05, 13, 800, 1e-4),
    "g_200x1200_b4_p2": (200, 1200, 4, 2, 0.05, 14, 600, 1e-4),
    # the reference driver's default instance (cpu_vs_gpu.py:57-74,95-101)
    "c1_1024x4096_b2_p4": (1024, 4096, 2, 4, 0.4, 1234, 1000, 1e-4),
}


def main(argv):
    names = argv[1:] or list(CASES)
    os.makedirs(GOLDEN, exist_ok=True)
    for name in names:
        N, K, BLOCK, P, den, seed, ITER_MAX, ERR_BOUND = CASES[name]
        out = run_reference(N, K, BLOCK, P, den, seed, ITER_MAX, ERR_BOUND)
        # A is NOT stored (regenerated from the seed by make_problem); a few
        # probe entries pin the regeneration.
        np.savez_compressed(
            os.path.join(GOLDEN, name + ".npz"),
            N=N, K=K, BLOCK=BLOCK, P=P, den=den, seed=seed,
            ITER_MAX=ITER_MAX, ERR_BOUND=ERR_BOUND,
            b=out["b"], mu=out["mu"], x=out["x"], err=out["err"],
            gamma=out["gamma"], iters=out["iters"],
            objective=out["objective"], d_ATA=out["d_ATA"],
            A_probe=out["A"][:4, :8].copy(),
            A_checksum=float(np.sum(out["A"])),
            ref_elapsed=out["elapsed"])
        print("%-22s iters=%4d nnz=%5d obj=%.14g mu=%.17g last_err=%.6e (%.1fs)"
              % (name, out["iters"], int(np.count_nonzero(out["x"])),
                 out["objective"], out["mu"], out["err"][-1], out["elapsed"]))


if __name__ == "__main__":
    main(sys.argv)
