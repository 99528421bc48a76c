"""Operator-level parity: every CUDA step of the hot path against the oracle at identical inputs.

Bit-exact where the north star says so (active sets, entering/leaving lists, breakpoint ordering, and the
element-wise residual vectors, Ruiz scaling); 1e-10..1e-8 relative for floating-point reductions/factorisations.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from qpalm_b200 import problems
from qpalm_b200.abi import CSC

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b))) if a.size else 0.0


SHAPES = [(4, 5, 0.5), (30, 50, 0.3), (200, 333, 0.05), (129, 1000, 0.02), (300, 100, 1.0), (1000, 2000, 0.05)]


@pytest.mark.parametrize("n,m,dens", SHAPES)
def test_mat_vec_trio(gpu_ops, oracle_ops, n, m, dens):
    """mat_vec / mat_tpose_vec (solver_interface.c:252-274) for A (CSC or dense) and Q (lower triangle)."""
    p = problems.random_qp(n, m, dens, min(1.0, dens * 2), seed=n + m)
    rng = np.random.default_rng(1)
    x, y = rng.standard_normal(n), rng.standard_normal(m)
    assert _rel(gpu_ops.mat_vec(p.A, x), oracle_ops.mat_vec(p.A, x)) < 1e-13
    assert _rel(gpu_ops.mat_tpose_vec(p.A, y), oracle_ops.mat_tpose_vec(p.A, y)) < 1e-13
    assert _rel(gpu_ops.mat_vec(p.Q, x), oracle_ops.mat_vec(p.Q, x)) < 1e-13


def test_solver_interface_known_answers_on_the_gpu(gpu_ops):
    """Known answers of the reference's own tests/src/test_solver_interface.c:106-159 (the 3 x 2 A, the 2 x 2 Q with both
    triangles stored): mat_vec, mat_tpose_vec, mat_inf_norm_cols / rows, ldlchol + ldlsolveLD_neg_dphi with and without
    the proximal term -- through the CUDA operator ABI (test_oracle.py runs the same vectors through the oracle)."""
    f = problems.solver_interface_fixture()
    np.testing.assert_allclose(gpu_ops.mat_vec(f["A"], f["Qd"]), f["A_Qd"], atol=1e-8)
    np.testing.assert_allclose(gpu_ops.mat_vec(f["Q"], f["Qd"]), f["Q_Qd"], atol=1e-8)
    np.testing.assert_allclose(gpu_ops.mat_tpose_vec(f["A"], f["Ad"]), f["At_Ad"], atol=1e-8)
    np.testing.assert_allclose(gpu_ops.norm_cols(f["A"]), f["col_norms"], atol=1e-8)
    np.testing.assert_allclose(gpu_ops.norm_rows(f["A"]), f["row_norms"], atol=1e-8)
    d, _ = gpu_ops.newton_solve(f["Q"], None, None, None, 0.0, -f["neg_rhs"], want_L=False)
    np.testing.assert_allclose(d, f["d_noprox"], atol=1e-8)
    d, _ = gpu_ops.newton_solve(f["Q"], None, None, None, 1.0 / f["gamma"], -f["neg_rhs"], want_L=False)
    np.testing.assert_allclose(d, f["d_prox"], atol=1e-8)


def test_q_upper_triangle_ignored(gpu_ops, oracle_ops):
    """Callers may store both triangles with stype -1; the upper entries are ignored."""
    rng = np.random.default_rng(3)
    M = rng.standard_normal((6, 6))
    S = sp.csc_matrix(M + M.T)
    junk = S.copy()
    junk.data = junk.data.copy()
    Qfull = CSC(6, 6, junk.indptr, junk.indices, junk.data, -1)
    # corrupt the strictly upper entries
    for j in range(6):
        for k in range(Qfull.p[j], Qfull.p[j + 1]):
            if Qfull.i[k] < j:
                Qfull.x[k] = 1e9
    x = rng.standard_normal(6)
    np.testing.assert_allclose(gpu_ops.mat_vec(Qfull, x), (M + M.T) @ x, rtol=1e-13)
    np.testing.assert_allclose(oracle_ops.mat_vec(Qfull, x), (M + M.T) @ x, rtol=1e-13)


@pytest.mark.parametrize("n,m,dens", SHAPES)
def test_inf_norms_bit_exact(gpu_ops, oracle_ops, n, m, dens):
    p = problems.random_qp(n, m, dens, 0.1, seed=7)
    assert np.array_equal(gpu_ops.norm_cols(p.A), oracle_ops.norm_cols(p.A))
    assert np.array_equal(gpu_ops.norm_rows(p.A), oracle_ops.norm_rows(p.A))


@pytest.mark.parametrize("n,m,dens,densQ", [(4, 5, 0.5, 1.0), (40, 60, 0.2, 0.1), (150, 90, 1.0, 1.0), (300, 500, 0.05, 0.02)])
def test_ruiz_scaling_bit_exact(gpu_ops, oracle_ops, n, m, dens, densQ):
    """scale_data (scaling.c:34-113): every iterate lives in scaled space, so this must be bit-exact."""
    p = problems.random_qp(n, m, dens, densQ, seed=11)
    g = gpu_ops.scale_data(p.A, p.Q, p.q, p.bmin, p.bmax, 10)
    o = oracle_ops.scale_data(p.A, p.Q, p.q, p.bmin, p.bmax, 10)
    for k in ("D", "E", "q", "bmin", "bmax", "Ax"):
        assert np.array_equal(g[k], o[k]), k
    assert g["c"] == o["c"]
    nz = int(p.Q.p[-1])
    assert np.array_equal(g["Qx"][:nz], o["Qx"][:nz])


@pytest.mark.parametrize("n,m,dens", SHAPES)
@pytest.mark.parametrize("proximal", [0, 1])
def test_residuals_and_active_set_bit_exact(gpu_ops, oracle_ops, n, m, dens, proximal):
    """compute_residuals + set_active_constraints + set_entering_leaving_constraints at an identical iterate."""
    p = problems.random_qp(n, m, dens, 0.1, seed=5)
    rng = np.random.default_rng(9)
    Ax = rng.standard_normal(m) * 0.7
    # plant exact ties on the bounds: the comparisons are <= / >=
    Ax[::7] = p.bmin[::7]
    y = rng.standard_normal(m) * (rng.random(m) < 0.5)
    y[::7] = 0.0
    sigma = 10.0 ** rng.uniform(-2, 4, m)
    args = (p.A, Ax, y, sigma, p.bmin, p.bmax, rng.standard_normal(n), p.q, rng.standard_normal(n), proximal, 1e3,
            (rng.random(m) < 0.3).astype(np.int64))
    g, o = gpu_ops.residuals(*args), oracle_ops.residuals(*args)
    for k in ("Axys", "z", "pri_res", "yh", "df"):
        assert np.array_equal(g[k], o[k]), k
    for k in ("active", "enter", "leave"):
        assert np.array_equal(g[k], o[k]), k
    assert g["nb_active"] == o["nb_active"]
    assert _rel(g["Atyh"], o["Atyh"]) < 1e-13 and _rel(g["dphi"], o["dphi"]) < 1e-13


@pytest.mark.parametrize("m", [1, 2, 5, 64, 1000, 2049, 4096, 4097, 20000])
def test_linesearch_ordering_bit_exact(gpu_ops, oracle_ops, m):
    """exact_linesearch: the sorted breakpoint order (value, then original index) is bit-exact; tau to 1e-10.
    2m <= 8192 takes the single-CTA sort + walk (m = 4096 is its largest size), larger m the multi-launch radix sort."""
    rng = np.random.default_rng(m)
    Ad = rng.standard_normal(m)
    Ad[rng.random(m) < 0.1] = 0.0                 # delta = +-0: infinite / NaN breakpoints
    Ax = rng.standard_normal(m)
    y = rng.standard_normal(m) * (rng.random(m) < 0.6)
    sigma = 10.0 ** rng.uniform(-1, 3, m)
    bmin, bmax = -rng.random(m), rng.random(m)
    if m >= 64:                                   # identical rows => tied breakpoints (identity rows, symmetric bounds)
        Ad[10:20], Ax[10:20], y[10:20], sigma[10:20], bmin[10:20], bmax[10:20] = 0.5, 0.1, 0.0, 4.0, -1.0, 1.0
    eta, beta = 3.0, -5.0
    tg, sg, ig = gpu_ops.linesearch(eta, beta, Ad, Ax, y, sigma, bmin, bmax)
    to, so, io = oracle_ops.linesearch(eta, beta, Ad, Ax, y, sigma, bmin, bmax)
    assert np.array_equal(ig, io)
    assert np.array_equal(sg, so)
    assert abs(tg - to) <= 1e-10 * max(1.0, abs(to))


def test_linesearch_all_breakpoints_traversed(gpu_ops, oracle_ops):
    """The situation of tests/src/test_ls_qp.c: the walk runs off the end of the breakpoint list."""
    m = 7
    Ad, Ax, y = np.ones(m), np.zeros(m), np.zeros(m)
    sigma, bmin, bmax = np.ones(m), -np.ones(m), np.arange(1, m + 1, dtype=float)
    tg, _, ig = gpu_ops.linesearch(1e-3, -1e3, Ad, Ax, y, sigma, bmin, bmax)
    to, _, io = oracle_ops.linesearch(1e-3, -1e3, Ad, Ax, y, sigma, bmin, bmax)
    assert np.array_equal(ig, io) and abs(tg - to) <= 1e-12 * abs(to)


@pytest.mark.parametrize("n,m,dens,densQ", [(2, 3, 1.0, 1.0), (30, 50, 0.3, 0.2), (128, 64, 0.2, 0.1), (129, 300, 0.1, 0.05),
                                            (300, 200, 1.0, 1.0), (700, 1500, 0.05, 0.02), (1400, 600, 1.0, 1.0)])
def test_newton_factor_and_solve(gpu_ops, oracle_ops, n, m, dens, densQ):
    """(Q + A_J' Sigma_J A_J + beta I) d = rhs through the DMMA SYRK + blocked Cholesky + blocked solves.
    n = 1400 (three outer blocks) runs the two-stream look-ahead and both GEMM tile shapes; n > 256 the dataflow solves."""
    p = problems.random_qp(n, m, dens, densQ, seed=21)
    rng = np.random.default_rng(2)
    sigma = 10.0 ** rng.uniform(-1, 2, m)
    active = (rng.random(m) < 0.4).astype(np.int64)
    rhs = rng.standard_normal(n)
    beta = 1e-3
    dg, Lg = gpu_ops.newton_solve(p.Q, p.A, sigma, active, beta, rhs)
    do, Lo = oracle_ops.newton_solve(p.Q, p.A, sigma, active, beta, rhs)
    A, Q = p.A.to_scipy().toarray(), p.Q.to_scipy().toarray()
    J = active.astype(bool)
    H = Q + (A[J].T * sigma[J]) @ A[J] + beta * np.eye(n)
    assert np.max(np.abs(H @ dg - rhs)) < 1e-8 * max(1, np.max(np.abs(rhs))) * np.linalg.cond(H) ** 0.5
    assert _rel(dg, do) < 1e-8
    assert _rel(Lg, Lo) < 1e-9
    # no active constraints: Q + beta I alone (newton.c:109-111)
    dg0, _ = gpu_ops.newton_solve(p.Q, p.A, sigma, None, 1.0, rhs, want_L=False)
    do0, _ = oracle_ops.newton_solve(p.Q, p.A, sigma, None, 1.0, rhs, want_L=False)
    assert _rel(dg0, do0) < 1e-9


@pytest.mark.parametrize("n,k", [(5, 1), (40, 3), (128, 8), (200, 11), (513, 20)])
def test_rank_k_update_and_downdate(gpu_ops, oracle_ops, n, k):
    """cholmod_updown replacement: L L' +- W W' (t_cholmod_updown_numkr.c recurrence) and the round trip."""
    rng = np.random.default_rng(n)
    M = rng.standard_normal((n, n))
    H = M @ M.T + n * np.eye(n)
    L = np.linalg.cholesky(H)
    W = rng.standard_normal((n, k))
    Lu_g, Lu_o = gpu_ops.updown(L, W, 1), oracle_ops.updown(L, W, 1)
    assert _rel(Lu_g @ Lu_g.T, H + W @ W.T) < 1e-12
    assert _rel(Lu_g, Lu_o) < 1e-11
    Ld_g, Ld_o = gpu_ops.updown(Lu_g, W, 0), oracle_ops.updown(Lu_o, W, 0)
    assert _rel(Ld_g, Ld_o) < 1e-10
    assert _rel(Ld_g, L) < 1e-9           # update then downdate is the identity


@pytest.mark.parametrize("n", [4, 50, 400])
def test_lobpcg_lambda_min(gpu_ops, oracle_ops, n):
    """lobpcg (nonconvex.c:29-168): same start vector => same under-estimate of lambda_min."""
    p = problems.random_qp(n, 2 * n, 0.1, 0.1, seed=4, nonconvex_shift=1.0)
    x0 = np.random.default_rng(8).random(n)
    lg, itg = gpu_ops.lobpcg(p.Q, x0)
    lo, ito = oracle_ops.lobpcg(p.Q, x0)
    lam = np.linalg.eigvalsh(p.Q.to_scipy().toarray())[0]
    if ito < 1000:                                    # converged: deliberately an under-estimate (nonconvex.c:117-121)
        assert lg < lam and lo < lam
    assert abs(itg - ito) <= max(2, ito // 10)
    assert abs(lg - lo) < 2e-5 * max(1.0, abs(lo))
    assert abs(lg - lam) < 1e-3 * max(1.0, abs(lam))
