"""Sparse Newton system on the GPU (csrc/sparse.cu): supernodal multifrontal factorization, level-scheduled solves and the
rank-k update/downdate along etree paths -- the replacement of cholmod_analyze / cholmod_factorize / cholmod_solve /
cholmod_updown (solver_interface.c:319-519) for problems whose Schur complement stays sparse.

Operator level: against numpy on the same matrix (factor reproduces P H P', solution 1e-10, update/downdate equals the
factor of the modified matrix).  Solver level: same gates as test_gpu_solve.py against the oracle / the reference build
(status, x / y 1e-8 relative, iterations within 5 %), with the sparse path forced or chosen by the density heuristic.
"""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import HAS_REF
from qpalm_b200 import abi, problems
from qpalm_b200.interface import Qpalm, load_library, solve_qp
from qpalm_b200.sparse import sparse_newton

pytestmark = pytest.mark.gpu


def _dense_H(p, sigma, act, beta):
    Q = p.Q.to_scipy().toarray()
    A = p.A.to_scipy().toarray() if p.m else np.zeros((0, p.n))
    a = np.asarray(act, dtype=bool)
    return Q + A[a].T @ (sigma[a, None] * A[a]) + beta * np.eye(p.n)


OP_CASES = [("basic", lambda: problems.basic_qp()), ("medium", lambda: problems.medium_qp()),
            ("grid9", lambda: problems.grid_qp(9, seed=1)), ("grid31", lambda: problems.grid_qp(31, seed=2)),
            ("grid70", lambda: problems.grid_qp(70, seed=6)),
            ("rand_sparse", lambda: problems.random_qp(150, 260, 0.02, 0.01, seed=3)),
            ("rand_one_front", lambda: problems.random_qp(300, 500, 0.2, 0.1, seed=4))]


@pytest.mark.parametrize("name,make", OP_CASES, ids=[c[0] for c in OP_CASES])
def test_factor_and_solve_match_numpy(name, make):
    p = make()
    rng = np.random.default_rng(11)
    sigma = 0.5 + 20 * rng.random(max(p.m, 1))
    act = (rng.random(max(p.m, 1)) < 0.5).astype(np.int64)
    beta = 1e-3 if name != "basic" else 1.0
    rhs = rng.standard_normal(p.n)
    d, L, perm, bound = sparse_newton(p.Q, p.A, sigma, act, beta, rhs, want_factor=True, want_bound=True)
    H = _dense_H(p, sigma[:p.m], act[:p.m], beta)
    dref = np.linalg.solve(H, rhs)
    assert np.max(np.abs(d - dref)) <= 1e-10 * max(1.0, np.max(np.abs(dref))) * np.linalg.cond(H) ** 0.5
    Hp = H[np.ix_(perm, perm)]
    assert np.max(np.abs(L @ L.T - Hp)) <= 1e-12 * np.max(np.abs(Hp))
    assert np.max(np.abs(np.triu(L, 1))) == 0.0
    G = _dense_H(p, sigma[:p.m], act[:p.m], 0.0) - p.Q.to_scipy().toarray()
    assert abs(bound - np.max(np.sum(np.abs(G), axis=1))) <= 1e-12 * max(1.0, bound)


def test_factorization_is_bit_reproducible():
    p = problems.grid_qp(40, seed=3)
    rng = np.random.default_rng(5)
    sigma = 1 + rng.random(p.m); act = (rng.random(p.m) < 0.4).astype(np.int64); rhs = rng.standard_normal(p.n)
    d1, L1, _, _ = sparse_newton(p.Q, p.A, sigma, act, 1e-4, rhs, want_factor=True)
    d2, L2, _, _ = sparse_newton(p.Q, p.A, sigma, act, 1e-4, rhs, want_factor=True)
    assert np.array_equal(L1, L2) and np.array_equal(d1, d2)


@pytest.mark.parametrize("cluster", ["1", "2", "8", "16"])
@pytest.mark.parametrize("name,make", [("grid70", lambda: problems.grid_qp(70, seed=6)),
                                       ("rand_one_front", lambda: problems.random_qp(300, 500, 0.2, 0.1, seed=4)),
                                       ("rand_one_front_700", lambda: problems.random_qp(700, 900, 0.2, 0.1, seed=5))],
                         ids=["grid70", "rand_one_front", "rand_one_front_700"])
def test_cluster_front_kernel_equals_per_block_launches(name, make, cluster):
    """Fronts above the shared-memory size are factored by ONE launch per level (a thread-block cluster per front, chain CTA with
    look-ahead, mfc::k_mf_front); the per-block launches (k_mf_extend / k_mf_diag / k_mf_trsm / k_mf_syrk) remain behind
    QPALM_B200_MF_PER_BLOCK=1.  Same operations in the same order: the factors must be bit-identical for every cluster size."""
    p = make()
    rng = np.random.default_rng(21)
    sigma = 0.5 + 20 * rng.random(p.m); act = (rng.random(p.m) < 0.5).astype(np.int64); rhs = rng.standard_normal(p.n)
    out = {}
    for mode, env in (("block", {"QPALM_B200_MF_PER_BLOCK": "1"}), ("cluster", {"QPALM_B200_MF_PER_BLOCK": "0", "QPALM_B200_MF_CLUSTER": cluster})):
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            out[mode] = sparse_newton(p.Q, p.A, sigma, act, 1e-3, rhs, want_factor=True)
        finally:
            for k, v in old.items():
                os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
    assert np.array_equal(out["block"][1], out["cluster"][1])
    assert np.array_equal(out["block"][0], out["cluster"][0])
    H = _dense_H(p, sigma, act, 1e-3)
    dref = np.linalg.solve(H, rhs)
    assert np.max(np.abs(out["cluster"][0] - dref)) <= 1e-9 * max(1.0, np.max(np.abs(dref)))


@pytest.mark.parametrize("name,make", [("grid70", lambda: problems.grid_qp(70, seed=6)),
                                       ("rand_one_front_700", lambda: problems.random_qp(700, 900, 0.2, 0.1, seed=5))],
                         ids=["grid70", "rand_one_front_700"])
def test_pipelined_solves_equal_generic_kernels(name, make):
    """Levels of large fronts take k_mf_fwd_big / k_mf_bwd_big (factor loads issued ahead of the substitution, 512 threads); the generic kernels stay
    behind QPALM_B200_MF_FWD_GENERIC=1.  Same arithmetic per entry in the same order: the solutions must be bit-identical."""
    p = make()
    rng = np.random.default_rng(33)
    sigma = 0.5 + 20 * rng.random(p.m); act = (rng.random(p.m) < 0.5).astype(np.int64); rhs = rng.standard_normal(p.n)
    out = {}
    for mode, flag in (("generic", "1"), ("pipelined", "0")):
        old = os.environ.get("QPALM_B200_MF_FWD_GENERIC")
        os.environ["QPALM_B200_MF_FWD_GENERIC"] = flag
        try:
            out[mode] = sparse_newton(p.Q, p.A, sigma, act, 1e-3, rhs)[0]
        finally:
            os.environ.pop("QPALM_B200_MF_FWD_GENERIC", None) if old is None else os.environ.__setitem__("QPALM_B200_MF_FWD_GENERIC", old)
    assert np.array_equal(out["generic"], out["pipelined"])
    dref = np.linalg.solve(_dense_H(p, sigma, act, 1e-3), rhs)
    assert np.max(np.abs(out["pipelined"] - dref)) <= 1e-9 * max(1.0, np.max(np.abs(dref)))


@pytest.mark.parametrize("name,make", [("grid31", lambda: problems.grid_qp(31, seed=2)),
                                       ("rand_sparse", lambda: problems.random_qp(150, 260, 0.02, 0.01, seed=3)),
                                       ("grid70", lambda: problems.grid_qp(70, seed=6))], ids=["grid31", "rand_sparse", "grid70"])
@pytest.mark.parametrize("k_enter,k_leave", [(1, 0), (0, 1), (5, 3), (8, 8), (13, 11)])
def test_rank_update_downdate_equals_refactorization(name, make, k_enter, k_leave):
    """cholmod_updown semantics: L L' <- L L' + sum_enter sigma a a' - sum_leave sigma a a' (multiples of 8 per sweep)."""
    p = make()
    rng = np.random.default_rng(100 * k_enter + k_leave)
    sigma = 0.5 + 5 * rng.random(p.m)
    act = (rng.random(p.m) < 0.5).astype(np.int64)
    on, off = np.flatnonzero(act == 1), np.flatnonzero(act == 0)
    enter = np.sort(rng.choice(off, k_enter, replace=False)) if k_enter else np.array([], dtype=np.int64)
    leave = np.sort(rng.choice(on, k_leave, replace=False)) if k_leave else np.array([], dtype=np.int64)
    rhs = rng.standard_normal(p.n)
    beta = 1e-2
    d, L, perm, _ = sparse_newton(p.Q, p.A, sigma, act, beta, rhs, enter=enter, leave=leave, want_factor=True)
    act2 = act.copy(); act2[enter] = 1; act2[leave] = 0
    H2 = _dense_H(p, sigma, act2, beta)
    Lref = np.linalg.cholesky(H2[np.ix_(perm, perm)])
    assert np.max(np.abs(L - Lref)) <= 1e-9 * np.max(np.abs(Lref))
    dref = np.linalg.solve(H2, rhs)
    assert np.max(np.abs(d - dref)) <= 1e-8 * max(1.0, np.max(np.abs(dref)))


# ---------------------------------------------------------------------------------------------
# solver level
# ---------------------------------------------------------------------------------------------
def _rel(a, b):
    return np.max(np.abs(a - b)) / max(1.0, np.max(np.abs(b))) if a.size else 0.0


def _solve_with_stats(p, env=None, **kw):
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update(env or {})
    try:
        lib = load_library("b200")
        lib.qpalm_b200_get_stats.argtypes = [C.POINTER(abi.QPALMWorkspace), C.POINTER(abi.QPALMB200Stats)]
        s = Qpalm("b200")
        st = dict(p.settings); st.update(kw)
        for k, v in st.items():
            setattr(s.settings, k, v)
        s.set_data(p.Q.copy(), p.A.copy(), p.q.copy(), p.bmin.copy(), p.bmax.copy())
        assert s._allocate_work()
        s._solve()
        r = s.result()
        stats = abi.QPALMB200Stats()
        lib.qpalm_b200_get_stats(s._work, C.byref(stats))
        s.cleanup()
        return r, stats
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def _ref(p, **kw):
    st = dict(p.settings); st.update(kw)
    impl = "reference" if HAS_REF else "oracle"
    return solve_qp(impl, p.Q.copy(), p.A.copy(), p.q.copy(), p.bmin.copy(), p.bmax.copy(), **st)


def _gates(g, r, tol=1e-8, iter_tol=0.05):
    assert g.status_val == r.status_val, (g.status, r.status)
    if r.status_val == 1:
        assert _rel(g.x, r.x) < tol, _rel(g.x, r.x)
        assert _rel(g.y, r.y) < tol, _rel(g.y, r.y)
    assert abs(g.iter - r.iter) <= max(1, int(np.ceil(iter_tol * r.iter))), (g.iter, r.iter)
    assert abs(g.iter_out - r.iter_out) <= max(1, int(np.ceil(iter_tol * r.iter_out))), (g.iter_out, r.iter_out)


@pytest.mark.parametrize("make", [problems.basic_qp, problems.medium_qp, problems.ls_qp, problems.degen_hess_qp,
                                  problems.prim_inf_qp, problems.dua_inf_qp],
                         ids=["basic", "medium", "ls", "degen_hess", "prim_inf", "dua_inf"])
@pytest.mark.parametrize("kw", [dict(), dict(proximal=0, scaling=2)], ids=["default", "noprox"])
def test_known_answer_qps_through_the_sparse_factor(make, kw):
    """The reference's own known-answer problems with the supernodal path forced (QPALM_B200_NEWTON=sparse)."""
    p = make()
    g, stats = _solve_with_stats(p, env={"QPALM_B200_NEWTON": "sparse"}, **kw)
    assert stats.sparse_factor_nnz > 0
    r = _ref(p, **kw)
    if p.name == "dua_inf":
        # tests/src/test_dua_inf_qp.c pins the STATUS only: with Q = 1e-10 I the iteration at which the certificate fires depends on
        # the last bits of the factor (round 1 matched 7 iterations only through a timing-dependent refactor-vs-update choice; the
        # choice is a static rule now and this 2-variable problem forced through the supernodal path fires at iteration 5)
        assert g.status_val == r.status_val == -4
    else:
        _gates(g, r)
    if p.expect_x is not None and g.status_val == 1:
        np.testing.assert_allclose(g.x, p.expect_x, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("g_,seed", [(12, 0), (25, 1), (40, 2)])
def test_grid_qp_matches_reference(g_, seed):
    """Config-2 stand-in (synthetic, see problems.grid_qp).  g = 25, 40 take the sparse path by the density heuristic."""
    p = problems.grid_qp(g_, seed=seed)
    env = {"QPALM_B200_NEWTON": "sparse"} if p.n < 512 else {}
    g, stats = _solve_with_stats(p, env=env)
    assert stats.sparse_factor_nnz > 0, "expected the supernodal path"
    _gates(g, _ref(p))
    assert g.status_val == 1


def test_grid_qp_rank_update_path_and_dense_agree():
    """Same problem through (a) sparse factor with forced rank updates, (b) sparse factor refactorizing, (c) dense factor."""
    p = problems.grid_qp(25, seed=4)
    r = _ref(p, max_rank_update_fraction=1.0)
    a, sa = _solve_with_stats(p, env={"QPALM_B200_NEWTON": "sparse", "QPALM_B200_UPDOWN_FORCE": "1"}, max_rank_update_fraction=1.0)
    b, sb = _solve_with_stats(p, env={"QPALM_B200_NEWTON": "sparse", "QPALM_B200_UPDOWN_MAX_RANK": "0"}, max_rank_update_fraction=1.0)
    c, sc = _solve_with_stats(p, env={"QPALM_B200_NEWTON": "dense"}, max_rank_update_fraction=1.0)
    assert sa.sparse_factor_nnz > 0 and sb.sparse_factor_nnz > 0 and sc.sparse_factor_nnz == 0
    assert sa.updown_calls > 0 and sb.updown_calls == 0
    for g in (a, b, c):
        _gates(g, r)


def test_sparse_grid_qp_dual_termination():
    """enable_dual_termination: the second factor (Q alone, LD_Q of iteration.c:272-299) also lives in supernodal panels."""
    p = problems.grid_qp(20, seed=9)
    kw = dict(enable_dual_termination=1, dual_objective_limit=1e30)
    g2, stats = _solve_with_stats(p, env={"QPALM_B200_NEWTON": "sparse"}, **kw)
    assert stats.sparse_factor_nnz > 0
    r2 = _ref(p, **kw)
    _gates(g2, r2)
    assert abs(g2.dual_objective - r2.dual_objective) <= 1e-6 * max(1.0, abs(r2.dual_objective))


@pytest.mark.parametrize("shift", [0.3, 0.6])
def test_nonconvex_through_the_sparse_factor(shift):
    libc = C.CDLL("libc.so.6")
    p = problems.grid_qp(18, seed=12, diag_shift=shift, nonconvex=1)
    libc.srand(1)
    g, stats = _solve_with_stats(p, env={"QPALM_B200_NEWTON": "sparse"})
    libc.srand(1)
    r = _ref(p)
    assert stats.sparse_factor_nnz > 0
    _gates(g, r, tol=1e-5, iter_tol=0.1)


def test_sparse_resolve_is_bit_reproducible():
    """tests/src/test_basic_qp.c:298-305 on the supernodal sparse path: a second solve from the same start reproduces x, y exactly
    (the refactor-vs-update choice is a static rule, not a timing comparison)."""
    p = problems.grid_qp(26, seed=4)
    s = Qpalm("b200")
    for k, v in p.settings.items():
        setattr(s.settings, k, v)
    s.set_data(p.Q, p.A, p.q, p.bmin, p.bmax)
    assert s._allocate_work()
    x0, y0 = s.vec("x", p.n), s.vec("y", p.m)
    outs = []
    for _ in range(3):
        s._warm_start(x0, y0)
        s._solve()
        outs.append(s.result())
    s.cleanup()
    for r in outs[1:]:
        assert r.iter == outs[0].iter and r.iter_out == outs[0].iter_out
        assert np.array_equal(r.x, outs[0].x) and np.array_equal(r.y, outs[0].y)
