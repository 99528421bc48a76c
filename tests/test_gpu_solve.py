"""Solver-level parity through the drop-in C API: CUDA library vs the oracle (and the reference build when
oracle/_ref is present) on the reference's own known-answer QPs and on seeded random QPs.

Gates (BASELINE.json north_star): same status, x / y within 1e-8 relative, iteration counts within 5 %.
"""
import ctypes

import numpy as np
import pytest

from conftest import HAS_REF
from qpalm_b200 import problems
from qpalm_b200.interface import Qpalm, solve_qp

pytestmark = pytest.mark.gpu
libc = ctypes.CDLL("libc.so.6")


def _run(impl, p, **kw):
    libc.srand(1)  # the nonconvex path draws its LOBPCG start vector with rand() (nonconvex.c:41-44)
    st = dict(p.settings)
    st.update(kw)
    return solve_qp(impl, p.Q.copy(), p.A.copy(), p.q.copy(), p.bmin.copy(), p.bmax.copy(), warm_x=p.warm_x, warm_y=p.warm_y, **st)


def _rel(a, b):
    return np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b))) if a.size else 0.0


def _check(p, tol=1e-8, iter_tol=0.05, **kw):
    g = _run("b200", p, **kw)
    refs = [("oracle", _run("oracle", p, **kw))]
    if HAS_REF:
        refs.append(("reference", _run("reference", p, **kw)))
    for name, r in refs:
        assert g.status_val == r.status_val, (name, g.status, r.status)
        if r.status_val == 1:
            assert _rel(g.x, r.x) < tol, (name, _rel(g.x, r.x))
            assert _rel(g.y, r.y) < tol, (name, _rel(g.y, r.y))
            assert abs(g.objective - r.objective) <= 1e-8 * max(1.0, abs(r.objective))
        assert abs(g.iter - r.iter) <= max(1, int(np.ceil(iter_tol * r.iter))), (name, g.iter, r.iter)
        assert abs(g.iter_out - r.iter_out) <= max(1, int(np.ceil(iter_tol * r.iter_out))), (name, g.iter_out, r.iter_out)
    return g


VARIANTS = [dict(), dict(scaling=0), dict(proximal=0, scaling=2), dict(proximal=0, scaling=0)]


@pytest.mark.parametrize("kw", VARIANTS)
def test_basic_qp(kw):
    """tests/src/test_basic_qp.c: x* to 1e-5 relative, solved."""
    p = problems.basic_qp()
    g = _check(p, **kw)
    assert g.status_val == 1
    np.testing.assert_allclose(g.x, p.expect_x, rtol=1e-5)


def test_basic_qp_golden_trace():
    """solved at iteration 8 with 3 outer iterations (scaled, proximal, gamma_init 10)."""
    g = _run("b200", problems.basic_qp())
    assert (g.iter, g.iter_out) == (8, 3)
    g = _run("b200", problems.basic_qp(), scaling=0)
    assert g.iter == 12


@pytest.mark.parametrize("kw", VARIANTS)
def test_basic_qp_warm_start(kw):
    p = problems.basic_qp(warm_start=1)
    p.warm_x = np.array([2.0, -60.0, -3380.0, -6.0])
    p.warm_y = np.array([0.0, 0.0, -23.0, -0.014 if kw.get("scaling", 10) == 10 else -0.01, 0.0])
    kw = dict(kw)
    if "scaling" not in kw:
        kw["scaling"] = 2
    g = _check(p, **kw)
    assert g.iter < 12 and g.status_val == 1
    np.testing.assert_allclose(g.x, p.expect_x, rtol=1e-5)


def test_basic_qp_resolve_is_bit_reproducible():
    """test_basic_qp_warm_start_resolve: a second solve from the same start reproduces x, y to 1e-15."""
    p = problems.basic_qp()
    s = Qpalm("b200")
    for k, v in p.settings.items():
        setattr(s.settings, k, v)
    s.set_data(p.Q, p.A, p.q, p.bmin, p.bmax)
    s._allocate_work()
    x0, y0 = s.vec("x", p.n), s.vec("y", p.m)
    s._solve()
    r1 = s.result()
    s._warm_start(x0, y0)
    s._solve()
    r2 = s.result()
    assert r1.iter == r2.iter
    assert np.max(np.abs(r1.x - r2.x)) <= 1e-15 and np.max(np.abs(r1.y - r2.y)) <= 1e-15
    s.cleanup()


def test_basic_qp_limits_and_dual():
    p = problems.basic_qp()
    assert _run("b200", p, max_iter=1).status_val == -2
    assert _run("b200", p, eps_abs=1e-8, eps_rel=1e-8, inner_max_iter=2, max_iter=10).status_val == -2
    g = _check(p, enable_dual_termination=1)
    assert abs(g.objective - g.dual_objective) < 1e-5
    g = _run("b200", p, enable_dual_termination=1, dual_objective_limit=-1e9)
    assert g.status_val == 2 and g.iter_out == 0
    _check(p, sigma_max=1e3)
    assert _run("b200", p, time_limit=1e-5).status_val == -5


@pytest.mark.parametrize("kw", VARIANTS)
def test_ls_qp(kw):
    p = problems.ls_qp()
    g = _check(p, **kw)
    # the reference asserts 1e-5 absolute for its single (default scaling) variant; eps_rel = 1e-6 on |x| = 2e4
    np.testing.assert_allclose(g.x, p.expect_x, atol=1e-5 if not kw else 0.05)


@pytest.mark.parametrize("kw", VARIANTS)
def test_degen_hess(kw):
    p = problems.degen_hess_qp()
    g = _check(p, tol=1e-7, **kw)
    np.testing.assert_allclose(g.x, p.expect_x, atol=1e-5)


@pytest.mark.parametrize("kw", VARIANTS)
def test_primal_infeasible(kw):
    g = _check(problems.prim_inf_qp(), **kw)
    assert g.status_val == -3


@pytest.mark.parametrize("kw", VARIANTS)
def test_dual_infeasible(kw):
    g = _check(problems.dua_inf_qp(), **kw)
    assert g.status_val == -4


def test_nonconvex_qp():
    """tests/src/test_nonconvex_qp.c:124-125: gamma = 1/|lambda_min| within 10 % and under-estimated."""
    p = problems.nonconvex_qp()
    g = _check(p, tol=1e-6)
    lam = 0.0021544347
    assert abs(g.gamma - 1 / lam) < 0.1 / lam and 1 / g.gamma > lam


def test_update_bounds_q_settings():
    """tests/src/test_update.c: update_settings (more scaling iterations), update_bounds, update_q, re-solve."""
    p = problems.basic_qp()
    res = {}
    for impl in ("b200", "oracle"):
        s = Qpalm(impl)
        for k, v in p.settings.items():
            setattr(s.settings, k, v)
        s.settings.scaling = 2
        s.set_data(p.Q.copy(), p.A.copy(), p.q.copy(), p.bmin.copy(), p.bmax.copy())
        s._allocate_work()
        s._solve()
        out = [s.result()]
        s.settings.scaling = 5
        s._update_settings()
        s._solve()
        out.append(s.result())
        s._update_bounds(-1.5 * np.ones(p.m), 3.0 * np.ones(p.m))
        s._solve()
        out.append(s.result())
        s._update_q(p.q * 1.5 + 0.1)
        s._solve()
        out.append(s.result())
        # invalid updates -> QPALM_ERROR
        s._update_bounds(np.ones(p.m), -np.ones(p.m))
        err1 = int(s.info.status_val)
        s.settings.scaling = 1
        s._update_settings()
        err2 = int(s.info.status_val)
        res[impl] = (out, err1, err2)
        s.cleanup()
    for a, b in zip(res["b200"][0], res["oracle"][0]):
        assert a.status_val == b.status_val == 1
        assert _rel(a.x, b.x) < 1e-8 and _rel(a.y, b.y) < 1e-8
    assert res["b200"][1:] == res["oracle"][1:] == (0, 0)


def test_invalid_input_returns_null():
    p = problems.basic_qp()
    s = Qpalm("b200")
    s.set_data(p.Q, p.A, p.q, p.bmax, p.bmin)      # bmin > bmax
    assert not s._allocate_work()
    s2 = Qpalm("b200")
    s2.settings.rho = 2.0
    s2.set_data(p.Q, p.A, p.q, p.bmin, p.bmax)
    assert not s2._allocate_work()


@pytest.mark.parametrize("n,m,dA,dM,seed", [(60, 120, 0.3, 0.1, 0), (60, 120, 0.3, 0.1, 1), (300, 600, 0.1, 0.02, 5),
                                            (250, 400, 1.0, 1.0, 2), (1000, 2000, 0.05, 0.007, 1)])
def test_random_qp_matches_reference(n, m, dA, dM, seed):
    """BASELINE config 1 (n=1000, m=2000, density 0.05) and smaller/denser siblings."""
    _check(problems.random_qp(n, m, dA, dM, seed=seed))


@pytest.mark.parametrize("n,m,dA,dM,seed", [(60, 120, 0.3, 0.1, 0), (300, 600, 0.1, 0.02, 5)])
def test_random_qp_rank_update_path(n, m, dA, dM, seed):
    """max_rank_update_fraction = 1 (what every reference suite sets): the update/downdate path is exercised."""
    _check(problems.random_qp(n, m, dA, dM, seed=seed), max_rank_update_fraction=1.0)


def test_no_constraints_and_tiny():
    p = problems.random_qp(20, 0, 0.5, 0.5, seed=3)
    g = _run("b200", p)
    o = _run("oracle", p)
    assert g.status_val == o.status_val == 1 and _rel(g.x, o.x) < 1e-8


def test_nonconvex_random_qp():
    """BASELINE config 5 at a size the oracle finishes in seconds."""
    _check(problems.nonconvex_random_qp(100, 200, seed=3), tol=1e-5, iter_tol=0.1)


def test_mpc_instance():
    """One chain80w-sized instance (n=240, m=949, dense), BASELINE config 4 settings."""
    b = problems.mpc_batch(2, seed=1)
    _check(b.instance(0))


def test_appendix_e_probe_instance_on_the_gpu():
    """the xorshift64 probe instance (BASELINE config 1 shape): solved, 52 / 4, objective -4.0571425653e+01."""
    p = problems.probe_qp()
    g = _check(p)
    assert (g.iter, g.iter_out) == (52, 4) and abs(g.objective + 4.0571425653e+01) < 1e-8 * 40.6
