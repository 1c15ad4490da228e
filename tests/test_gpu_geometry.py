"""Oracle parity at the geometries the bench runs ("missing" 4): the C2 block shape
(N = 10,000 rows, w = 1,000 columns per block: 67-68 rows per CTA, one column group per thread, cs = 8)
in fp32 and fp64 and both layouts, and the C3 block shape (fp64, N = 20,000, w = 2,000: CPT = 4, D from
shared memory), plus device-side edge cases of the prox (cpu_calculation.py:5-6, lasso.py:114-136)."""
import ctypes

import numpy as np
import pytest

from conftest import assert_support
from oracle import lasso_oracle as orc
from test_gpu_parity import TOL, make_gpu_cal, rel

pytestmark = pytest.mark.gpu


def _instance(N, K, den, seed, TYPE):
    rng = np.random.RandomState(seed)
    A = rng.standard_normal((N, K))
    A /= np.linalg.norm(A, axis=1, keepdims=True)
    if TYPE == "float":
        A = A.astype(np.float32).astype(np.float64)          # the same matrix on both sides
    xt = rng.standard_normal((K, 1)) * (rng.rand(K, 1) < den)
    b = A @ xt + 1e-2 * rng.standard_normal((N, 1))
    mu = 0.1 * float(np.max(np.abs(A.T @ b)))
    return A, b, mu


_cache = {}


def _oracle(N, K, BLOCK, den, seed, TYPE, sweeps):
    key = (N, K, BLOCK, den, seed, TYPE, sweeps)
    if key not in _cache:
        A, b, mu = _instance(N, K, den, seed, TYPE)
        o = orc.lasso_oracle(A, b, mu, BLOCK, BLOCK * sweeps, None, faithful=False)
        _cache.clear()                                       # one 800 MB matrix at a time
        _cache[key] = (A, b, mu, o)
    return _cache[key]


@pytest.mark.parametrize("LAYOUT", ["row", "transposed"])
@pytest.mark.parametrize("TYPE", ["float", "double"])
def test_c2_block_geometry_vs_oracle(TYPE, LAYOUT):
    """10,000 x 10,000, BLOCK = 10: the sample the CPU arm of bench.py is built from"""
    from convex_optimization_b200 import lasso
    N, K, BLOCK, sweeps = 10000, 10000, 10, 4
    A, b, mu, o = _oracle(N, K, BLOCK, 0.01, 2, TYPE, sweeps)
    cal = make_gpu_cal(A, BLOCK, TYPE, LAYOUT)
    geo = cal.run_config()
    assert geo["grid"] >= 100 and geo["threads"] == 384
    solver = lasso.ClassLasso(cal, cal.diag_ATA, A, b, mu, BLOCK, BLOCK * sweeps)
    err = np.zeros(BLOCK * sweeps)
    solver.run(err_iter=err, SILENCE=True)
    assert solver.iters == o["iters"] == BLOCK * sweeps
    assert_support(solver.x, o["x"], TYPE)
    assert rel(solver.x, o["x"]) < TOL[TYPE]
    assert abs(orc.objective(A, b, solver.x, mu) - o["objective"]) / o["objective"] < TOL[TYPE]
    assert np.abs(err - o["err"]).max() / np.abs(o["err"]).max() < max(TOL[TYPE], 1e-10)


def test_c3_block_geometry_vs_oracle():
    """fp64, N = 20,000, w = 2,000 (BLOCK = 2): the wide-block instantiation (CPT = 4, TR = 4)"""
    from convex_optimization_b200 import lasso
    N, K, BLOCK, sweeps = 20000, 4000, 2, 6
    A, b, mu, o = _oracle(N, K, BLOCK, 0.01, 3, "double", sweeps)
    cal = make_gpu_cal(A, BLOCK, "double", "row")
    solver = lasso.ClassLasso(cal, cal.diag_ATA, A, b, mu, BLOCK, BLOCK * sweeps)
    solver.run(SILENCE=True)
    assert solver.iters == o["iters"]
    assert_support(solver.x, o["x"], "double")
    assert rel(solver.x, o["x"]) < TOL["double"]
    assert abs(orc.objective(A, b, solver.x, mu) - o["objective"]) / o["objective"] < TOL["double"]


# ------------------------------------------------------------------------ prox edge cases on the device
def _solve(A, b, mu, BLOCK, iters, TYPE="double", d=None, bound=None):
    from convex_optimization_b200 import lasso
    cal = make_gpu_cal(A, BLOCK, TYPE)
    solver = lasso.ClassLasso(cal, cal.diag_ATA if d is None else d, A, b, mu, BLOCK, iters)
    err = np.zeros(iters)
    solver.run(bound, err_iter=err, SILENCE=True)
    return solver, err


def test_prox_threshold_exactly_at_mu_and_signed_zero():
    """u = +mu and u = -mu exactly give Bx = 0 (strict inequality of the soft threshold,
    cpu_calculation.py:5-6), u = 2 mu / -3 mu shrink by mu: A = I, x0 = 0, so u = -g = b"""
    mu = 0.25
    A = np.eye(8)
    b = np.array([mu, -mu, 2 * mu, -3 * mu, 0.0, -0.0, np.nextafter(mu, 1.0), -np.nextafter(mu, 1.0)]).reshape(-1, 1)
    o = orc.lasso_oracle(A, b, mu, 1, 3, None, faithful=False)
    solver, err = _solve(A, b, mu, 1, 3)
    assert np.array_equal(solver.x != 0, o["x"] != 0)
    assert np.array_equal(solver.x != 0, np.array([0, 0, 1, 1, 0, 0, 1, 1], bool).reshape(-1, 1))
    assert rel(solver.x, o["x"]) < 1e-15 and np.abs(err - o["err"]).max() < 1e-15
    assert not np.signbit(solver.x[[0, 1, 4, 5]]).any()             # exact +0.0, never -0.0


def test_first_step_with_zero_direction_keeps_gamma_and_iterates():
    """|A^T b| <= mu everywhere: D = 0, q = 0, r_2 = 0 -- the reference prints a warning and keeps the
    step size (lasso.py:133-136); no division by zero, x stays 0, the error trace is finite"""
    rng = np.random.RandomState(3)
    A = rng.standard_normal((40, 60))
    A /= np.linalg.norm(A, axis=1, keepdims=True)
    b = 1e-3 * rng.standard_normal((40, 1))
    mu = 10.0 * float(np.abs(A.T @ b).max())
    solver, err = _solve(A, b, mu, 3, 9)
    assert solver.iters == 9 and not np.any(solver.x) and np.all(np.isfinite(err))
    o = orc.lasso_oracle(A, b, mu, 3, 9, None, faithful=False)
    assert np.abs(err - o["err"]).max() < 1e-14


def test_zero_column_keeps_its_coordinate_at_zero():
    """d_j = 0 (an all-zero column): the reference divides by it (lasso.py:29-30 gives inf, then nan);
    the device treats 1/d_j as 0 -- the coordinate stays 0 and nothing else is affected"""
    A, _, b, mu = orc.make_problem(120, 360, 0.05, seed=21)
    A[:, [5, 200]] = 0.0
    keep = np.ones(360, bool)
    keep[[5, 200]] = False
    solver, err = _solve(A, b, mu, 3, 300, bound=1e-4)
    assert solver.stopped and np.all(np.isfinite(solver.x)) and np.all(np.isfinite(err))
    assert solver.x[5, 0] == 0.0 and solver.x[200, 0] == 0.0
    # the remaining coordinates solve the problem without those columns: a zero column never changes
    # r, so the iterates of the other columns of its block are those of the matrix with the column
    # replaced by any column the prox leaves at zero; compare the objective with the oracle on the
    # matrix where the two columns are deleted from the model (x_j fixed at 0)
    A2 = A.copy()
    A2[:, [5, 200]] = 1e-150                                  # d_j = 1.2e-298 > 0, and |u_j| << mu: Bx_j = 0
    o = orc.lasso_oracle(A2, b, mu, 3, 300, 1e-4, faithful=False)
    assert o["x"][5, 0] == 0.0 and o["x"][200, 0] == 0.0
    assert solver.iters == o["iters"]
    assert rel(solver.x, o["x"]) < 1e-10 and np.array_equal(solver.x != 0, o["x"] != 0)


def test_caller_supplied_diagonal_is_used_by_the_fused_path():
    """the d_ATA argument of the solver classes (lasso.py:26-30): a diagonal that is not diag(A^T A)
    must give the same iterates on the fused and on the step-wise path"""
    from convex_optimization_b200 import lasso
    A, _, b, mu = orc.make_problem(200, 600, 0.05, seed=4)
    cal = make_gpu_cal(A, 3)
    d = cal.diag_ATA * 1.5
    fused = lasso.ClassLasso(cal, d, A, b, mu, 3, 45)
    fused.run(SILENCE=True)
    step = lasso.ClassLassoCB_v1(None, cal, d, A, b, mu, 3, 45)
    step.run(SILENCE=True)
    assert rel(fused.x, step.x) < 1e-10 and np.array_equal(fused.x != 0, step.x != 0)
    own = lasso.ClassLasso(cal, cal.diag_ATA, A, b, mu, 3, 45)       # and back to the matrix's own diagonal
    own.run(SILENCE=True)
    o = orc.lasso_oracle(A, b, mu, 3, 45, None, faithful=False)
    assert rel(own.x, o["x"]) < 1e-10 and rel(fused.x, o["x"]) > 1e-6


def test_device_vector_matvecs_and_objective_terms():
    """b200l_gemv_t_dev / _n_dev (device vectors, no copies) and b200l_objective_terms"""
    import torch
    from convex_optimization_b200 import _lib
    rng = np.random.RandomState(5)
    for TYPE, LAYOUT, (N, K, BLOCK) in [("double", "row", (257, 1002, 3)), ("float", "transposed", (300, 1201 * 2, 2)),
                                        ("float", "row", (1000, 4000, 4))]:
        A = rng.standard_normal((N, K))
        if TYPE == "float":
            A = A.astype(np.float32).astype(np.float64)
        cal = make_gpu_cal(A, BLOCK, TYPE, LAYOUT)
        w = K // BLOCK
        r = rng.standard_normal(N)
        dvec = rng.standard_normal(w)
        rd = torch.from_numpy(r).cuda()
        dd = torch.from_numpy(dvec).cuda()
        g = torch.empty(w, dtype=torch.float64, device="cuda")
        q = torch.empty(N, dtype=torch.float64, device="cuda")
        m = BLOCK - 1
        _lib.check(cal._lib.b200l_gemv_t_dev(cal.ctx, m, ctypes.c_void_p(rd.data_ptr()), ctypes.c_void_p(g.data_ptr())))
        _lib.check(cal._lib.b200l_gemv_n_dev(cal.ctx, m, ctypes.c_void_p(dd.data_ptr()), ctypes.c_void_p(q.data_ptr())))
        torch.cuda.synchronize()
        Am = A[:, m * w:(m + 1) * w]
        assert rel(g.cpu().numpy(), Am.T @ r) < 1e-13 and rel(q.cpu().numpy(), Am @ dvec) < 1e-13
        with pytest.raises(_lib.B200LassoError):
            _lib.check(cal._lib.b200l_gemv_t_dev(cal.ctx, m, ctypes.c_void_p(r.ctypes.data), ctypes.c_void_p(g.data_ptr())))
    A, _, b, mu = orc.make_problem(100, 300, 0.1, seed=9)
    cal = make_gpu_cal(A, 3)
    bb = np.ascontiguousarray(b.reshape(-1))
    _lib.check(cal._lib.b200l_set_problem(cal.ctx, _lib.dptr(bb)))
    _lib.check(cal._lib.b200l_run(cal.ctx, None, 30, mu, -1.0, None, None, None, None, None))
    x = np.empty((300, 1))
    _lib.check(cal._lib.b200l_get_x(cal.ctx, _lib.dptr(x)))
    rss, l1, obj = ctypes.c_double(), ctypes.c_double(), ctypes.c_double()
    _lib.check(cal._lib.b200l_objective_terms(cal.ctx, ctypes.byref(rss), ctypes.byref(l1)))
    _lib.check(cal._lib.b200l_objective(cal.ctx, mu, ctypes.byref(obj)))
    assert abs(rss.value - float(np.sum((A @ x - b) ** 2))) < 1e-10 * rss.value
    assert abs(l1.value - float(np.abs(x).sum())) < 1e-12 * max(l1.value, 1.0)
    assert abs(obj.value - (0.5 * rss.value + mu * l1.value)) < 1e-12 * obj.value
