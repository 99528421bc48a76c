"""CPU suite (-m "not gpu"): pins the oracle (oracle/qpalm_oracle.c) against
  (i)   the known-answer vectors of the reference's own tests (transcribed with file:line in qpalm_b200/problems.py),
  (ii)  committed outputs of the unmodified reference (tests/golden/ref_outputs.json, made by tests/golden/make_golden.py),
  (iii) the reference itself, live, whenever oracle/_ref/libqpalm_ref.so travelled with the snapshot.
No CUDA call is made here.
"""
import ctypes
import json
import os

import numpy as np
import pytest

from conftest import HAS_REF
from qpalm_b200 import problems
from qpalm_b200.interface import Qpalm, solve_qp

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "ref_outputs.json")))
libc = ctypes.CDLL("libc.so.6")


def _cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


MG = _cases()
CASES = MG.cases()
FAST = [k for k in CASES if k not in ("c1_random_1000_2000_s1",)]


def _rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b)))) if a.size else 0.0


def test_golden_file_covers_every_case():
    assert set(GOLD) == set(CASES)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_committed_reference_outputs(name):
    """Same status, identical iteration counts, x / y / objective to 1e-9 relative, Ruiz D / E / c bit-exact."""
    g = GOLD[name]
    o = MG.run("oracle", CASES[name])
    assert o["status_val"] == g["status_val"]
    assert (o["iter"], o["iter_out"]) == (g["iter"], g["iter_out"])
    assert np.array_equal(np.array(o["D"]), np.array(g["D"])) and np.array_equal(np.array(o["E"]), np.array(g["E"]))
    assert o["c"] == g["c"]
    if g["status_val"] == 1:
        tol = 1e-6 if "nonconvex" in name else 1e-9
        assert _rel(o["x"], g["x"]) < tol and _rel(o["y"], g["y"]) < tol
        assert abs(o["objective"] - g["objective"]) <= tol * max(1.0, abs(g["objective"]))
    assert abs(o["gamma"] - g["gamma"]) <= 1e-6 * abs(g["gamma"])


@pytest.mark.skipif(not HAS_REF, reason="oracle/_ref/libqpalm_ref.so not present")
@pytest.mark.parametrize("name", sorted(FAST))
def test_oracle_matches_live_reference(name):
    g = MG.run("reference", CASES[name])
    o = MG.run("oracle", CASES[name])
    assert o["status_val"] == g["status_val"] and (o["iter"], o["iter_out"]) == (g["iter"], g["iter_out"])
    if g["status_val"] == 1:
        tol = 1e-6 if "nonconvex" in name else 1e-9
        assert _rel(o["x"], g["x"]) < tol and _rel(o["y"], g["y"]) < tol


@pytest.mark.parametrize("make", [problems.basic_qp, problems.medium_qp, problems.ls_qp, problems.degen_hess_qp, problems.update_qp])
def test_reference_known_answers(make):
    """x* of tests/src/test_{basic,medium,ls,update}_qp.c and test_degen_hess.c (rel 1e-5 / abs 1e-5 as asserted there)."""
    p = make()
    r = solve_qp("oracle", p.Q, p.A, p.q, p.bmin, p.bmax, **p.settings)
    assert r.status_val == 1
    np.testing.assert_allclose(r.x, p.expect_x, rtol=1e-5, atol=1e-5)


def test_reference_status_pins():
    """test_prim_inf_qp.c / test_dua_inf_qp.c / test_basic_qp.c:317,331,360-361,388."""
    for p, st in ((problems.prim_inf_qp(), -3), (problems.dua_inf_qp(), -4)):
        assert solve_qp("oracle", p.Q, p.A, p.q, p.bmin, p.bmax, **p.settings).status_val == st
    p = problems.basic_qp()
    run = lambda **kw: solve_qp("oracle", p.Q.copy(), p.A.copy(), p.q, p.bmin, p.bmax, **{**p.settings, **kw})
    assert run(max_iter=1).status_val == -2
    assert run(time_limit=1e-5).status_val == -5
    r = run(enable_dual_termination=1)
    assert r.status_val == 1 and abs(r.objective - r.dual_objective) < 1e-5 * abs(r.objective)
    assert run(enable_dual_termination=1, dual_objective_limit=-1e9).status_val == 2


def test_golden_trace_appendix_d():
    """basic_qp solved at iteration 8 (3 outer); unscaled at iteration 12."""
    assert (GOLD["basic_qp"]["iter"], GOLD["basic_qp"]["iter_out"]) == (8, 3)
    assert GOLD["basic_qp_unscaled"]["iter"] == 12


def test_nonconvex_gamma_pin():
    """tests/src/test_nonconvex_qp.c:124-125."""
    libc.srand(1)
    p = problems.nonconvex_qp()
    r = solve_qp("oracle", p.Q, p.A, p.q, p.bmin, p.bmax, **p.settings)
    lam = 0.0021544347
    assert r.status_val == 1 and abs(r.gamma - 1 / lam) < 0.1 / lam and 1 / r.gamma > lam


def test_solver_interface_known_answers(oracle_ops):
    """tests/src/test_solver_interface.c:106-159 through the oracle's operator twins."""
    f = problems.solver_interface_fixture()
    np.testing.assert_allclose(oracle_ops.mat_vec(f["A"], f["Qd"]), f["A_Qd"], atol=1e-8)
    np.testing.assert_allclose(oracle_ops.mat_vec(f["Q"], f["Qd"]), f["Q_Qd"], atol=1e-8)
    np.testing.assert_allclose(oracle_ops.mat_tpose_vec(f["A"], f["Ad"]), f["At_Ad"], atol=1e-8)
    np.testing.assert_allclose(oracle_ops.norm_cols(f["A"]), f["col_norms"], atol=1e-8)
    np.testing.assert_allclose(oracle_ops.norm_rows(f["A"]), f["row_norms"], atol=1e-8)
    d, _ = oracle_ops.newton_solve(f["Q"], None, None, None, 0.0, -f["neg_rhs"], want_L=False)
    np.testing.assert_allclose(d, f["d_noprox"], atol=1e-8)
    d, _ = oracle_ops.newton_solve(f["Q"], None, None, None, 1.0 / f["gamma"], -f["neg_rhs"], want_L=False)
    np.testing.assert_allclose(d, f["d_prox"], atol=1e-8)


def test_update_sequence_known_answers():
    """tests/src/test_update.c:86-140: update_settings, update_bounds, update_q, each followed by a re-solve."""
    p = problems.update_qp()
    s = Qpalm("oracle")
    for k, v in p.settings.items():
        setattr(s.settings, k, v)
    s.set_data(p.Q, p.A, p.q.copy(), p.bmin.copy(), p.bmax.copy())
    assert s._allocate_work()
    s._solve()
    s.settings.gamma_init *= 0.1
    s.settings.theta = 0.9
    s.settings.scaling = 10
    s._update_settings()
    assert int(s.info.status_val) != 0
    s._solve()
    np.testing.assert_allclose(s.result().x, [-0.1, 0.3], atol=1e-5)
    bmin, bmax = p.bmin.copy(), p.bmax.copy()
    bmin[0], bmax[1] = 0.0, 1.5
    s._update_bounds(bmin, bmax)
    s._solve()
    np.testing.assert_allclose(s.result().x, [0.0, 0.15], atol=1e-5)
    s._update_bounds(p.bmin, p.bmax)
    s._update_q(np.array([-0.5, -0.75]))
    s._solve()
    assert s.result().status_val == 1
    np.testing.assert_allclose(s.result().x, [0.02, 0.18], atol=1e-5)
    s.cleanup()


def test_resolve_reproducible():
    """tests/src/test_basic_qp.c:298-305: a re-solve from the same start reproduces x to 1e-15."""
    p = problems.basic_qp()
    s = Qpalm("oracle")
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
    assert np.max(np.abs(r1.x - r2.x)) <= 1e-15 and r1.iter == r2.iter
    s.cleanup()


def test_linesearch_oracle_vs_bruteforce(oracle_ops):
    """exact_linesearch (linesearch.c:14-120): tau is the root of psi'(tau) = eta*tau + beta + sum_i delta_i [delta_i tau - alpha_i]_+,
    checked against a brute-force evaluation; the returned order is (value, original index) ascending."""
    rng = np.random.default_rng(0)
    for m in (1, 5, 64, 777):
        Ad, Ax = rng.standard_normal(m), rng.standard_normal(m)
        y = rng.standard_normal(m) * (rng.random(m) < 0.6)
        sigma = 10.0 ** rng.uniform(-1, 3, m)
        bmin, bmax = -rng.random(m), rng.random(m)
        sq = np.sqrt(sigma)
        delta = np.concatenate([-sq * Ad, sq * Ad])
        alpha = np.concatenate([(y + sigma * (Ax - bmin)) / sq, (sigma * (bmax - Ax) - y) / sq])
        eta = 3.0
        beta = -(5.0 + np.sum(delta * np.maximum(-alpha, 0.0)))      # a descent direction: psi'(0) = -5
        tau, s, idx = oracle_ops.linesearch(eta, beta, Ad, Ax, y, sigma, bmin, bmax)
        dpsi = lambda t: eta * t + beta + np.sum(delta * np.maximum(delta * t - alpha, 0.0))
        assert tau > 0 and abs(dpsi(tau)) < 1e-9 * (1 + abs(beta))
        assert np.all(np.diff(s) >= 0)
        ties = np.diff(s) == 0
        assert np.all(np.diff(idx)[ties] > 0)


def test_appendix_e_probe_instance():
    """the xorshift64 + Box-Muller probe instance has nnz(A) = 99 803, nnz(Q lower) = 25 154 and the
    reference returns solved, 52 / 4 iterations, objective -4.0571425653e+01 -- reproduced here by the oracle (and by the
    compiled reference when it is present), independently of numpy's generators."""
    p = problems.probe_qp()
    assert int(p.A.p[-1]) == 99803 and int(p.Q.p[-1]) == 25154
    impls = ["oracle"] + (["reference"] if HAS_REF else [])
    for impl in impls:
        r = solve_qp(impl, p.Q.copy(), p.A.copy(), p.q.copy(), p.bmin.copy(), p.bmax.copy(), **p.settings)
        assert r.status_val == 1 and (r.iter, r.iter_out) == (52, 4), (impl, r.iter, r.iter_out)
        assert abs(r.objective - (-4.0571425653e+01)) < 1e-8 * 40.6, (impl, r.objective)
