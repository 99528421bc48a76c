"""Problem instances: the reference's own known-answer QPs and the synthetic BASELINE.json configs.

Known-answer data is transcribed from the reference test-suite (file:line given per problem); the
synthetic generators follow SURVEY.md 8(d) (distribution of /root/reference/simulations/randomQP.m:32-38)
with numpy's PCG64 so that they are fast at the full BASELINE sizes.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

from .abi import CSC


@dataclass
class QP:
    name: str
    Q: CSC            # n x n, stype -1 (lower triangle)
    A: CSC            # m x n
    q: np.ndarray
    bmin: np.ndarray
    bmax: np.ndarray
    c: float = 0.0
    settings: dict = field(default_factory=dict)
    expect_x: np.ndarray | None = None
    expect_status: int | None = None
    warm_x: np.ndarray | None = None
    warm_y: np.ndarray | None = None

    @property
    def n(self):
        return self.Q.ncol

    @property
    def m(self):
        return self.A.nrow


def _csc(nrow, ncol, p, i, x, stype=0):
    return CSC(nrow, ncol, np.array(p), np.array(i), np.array(x, dtype=float), stype)


# ---------------------------------------------------------------------------------------------
# the reference's known-answer problems
# ---------------------------------------------------------------------------------------------
def basic_qp(**settings) -> QP:
    """tests/src/test_basic_qp.c:14-88 (n=4, m=5)."""
    A = _csc(5, 4, [0, 1, 2, 3, 4], [3, 4, 0, 2], [-1.0, 0.025431136, -0.0001, 0.33066985])
    Q = _csc(4, 4, [0, 1, 2, 3, 4], [0, 1, 2, 3], [1.0, 0.046415888, 0.0021544347, 0.0001], -1)
    st = dict(eps_abs=1e-6, eps_rel=1e-6, gamma_init=1e1, max_rank_update_fraction=1.0, verbose=0)
    st.update(settings)
    return QP("basic_qp", Q, A, np.array([-2.0146781, 2.9613971, 7.2865370, 7.8925204]),
              -2 * np.ones(5), 2 * np.ones(5), 0.0, st,
              expect_x=np.array([2.0, -63.801365, -3382.1109, -6.0483288]), expect_status=1)


def nonconvex_qp(**settings) -> QP:
    """tests/src/test_nonconvex_qp.c:14-125: basic_qp with Q[2][2] negated."""
    p = basic_qp()
    p.name = "nonconvex_qp"
    p.Q.x[2] = -0.0021544347
    p.settings = dict(eps_abs=1e-6, eps_rel=1e-6, nonconvex=1, scaling=0, max_rank_update_fraction=1.0, verbose=0)
    p.settings.update(settings)
    p.expect_x = None
    return p


def ls_qp(**settings) -> QP:
    """tests/src/test_ls_qp.c:17-108 (all breakpoints of the line search are traversed)."""
    A = _csc(2, 2, [0, 1, 2], [1, 0], [-1.0, 0.0001])
    Q = _csc(2, 2, [0, 1, 2], [0, 1], [1.0, 0.0001], -1)
    st = dict(eps_abs=1e-6, eps_rel=1e-6, gamma_max=1e3, gamma_init=1e1, max_rank_update_fraction=1.0, verbose=0)
    st.update(settings)
    return QP("ls_qp", Q, A, np.array([2.5150105, 16.259589]), -2 * np.ones(2), 2 * np.ones(2), 0.0, st,
              expect_x=np.array([-2.0, -2.0e4]), expect_status=1)


def degen_hess_qp(**settings) -> QP:
    """tests/src/test_degen_hess.c:17-104 (singular Hessian, one equality row)."""
    A = _csc(4, 3, [0, 2, 4, 6], [0, 1, 0, 2, 0, 3], [1.0] * 6)
    Q = _csc(3, 3, [0, 2, 4, 4], [0, 1, 0, 1], [1.0, -1.0, -1.0, 2.0], -1)
    st = dict(eps_abs=1e-6, eps_rel=1e-6, max_rank_update_fraction=1.0, verbose=0)
    st.update(settings)
    return QP("degen_hess", Q, A, np.array([-2.0, -6.0, 1.0]),
              np.array([0.5, -10, -10, -10.0]), np.array([0.5, 10, 10, 10.0]), 0.0, st,
              expect_x=np.array([5.5, 5.0, -10.0]), expect_status=1)


def prim_inf_qp(**settings) -> QP:
    """tests/src/test_prim_inf_qp.c:20-124."""
    A = _csc(3, 2, [0, 2, 4], [0, 2, 1, 2], [1.0] * 4)
    Q = _csc(2, 2, [0, 1, 2], [0, 1], [1.0, 1.5], -1)
    st = dict(eps_abs=1e-6, eps_rel=1e-6, max_rank_update_fraction=1.0, verbose=0)
    st.update(settings)
    return QP("prim_inf", Q, A, np.array([1.0, -2.0]), np.array([-5.0, -10, 16]), np.array([5.0, 10, 20]),
              0.0, st, expect_status=-3)


def dua_inf_qp(**settings) -> QP:
    """tests/src/test_dua_inf_qp.c:19-134."""
    A = _csc(3, 2, [0, 3, 6], [0, 1, 2, 0, 1, 2], [1.0] * 6)
    Q = _csc(2, 2, [0, 1, 2], [0, 1], [1e-10, 1e-10], -1)
    st = dict(eps_abs=1e-6, eps_rel=1e-6, max_rank_update_fraction=1.0, verbose=0)
    st.update(settings)
    return QP("dua_inf", Q, A, np.array([1.0, -2.0]), np.array([-5.0, -10, -20]), np.array([5.0, 10, 20]),
              0.0, st, expect_status=-4)


def medium_qp(**settings) -> QP:
    """tests/src/test_medium_qp.c:14-135 (n = m = 15, diagonal Q, 16-digit golden x)."""
    Ap = [0, 1, 2, 5, 8, 9, 11, 12, 13, 16, 18, 21, 22, 23, 24, 25]
    Ai = [8, 2, 1, 4, 14, 1, 4, 13, 5, 0, 7, 10, 6, 1, 4, 14, 0, 7, 1, 4, 13, 3, 9, 11, 12]
    Ax = [0.3256021467039615, -0.2129201224283822, -0.03904780212604003, -0.01097664622926547, 8.93509853157044e-05,
          0.1107958814061373, -0.394140028125563, -0.03422661790473164, -0.2077231940491557, 0.2961057917719591,
          0.02901671645955232, -0.2412937540712519, 0.2180403659113273, -0.07769757105018442, -0.02184140217516474,
          -4.490435862043659e-05, -0.007144833411941969, 0.07291061197330474, 0.01354927131911815, -0.04819953694147238,
          0.2798798702152373, -0.316687763261202, 0.4390581348235377, -0.3143332085622074, -1.0]
    Qx = [1.0, 0.5179474679231212, 0.2682695795279726, 0.1389495494373138, 0.07196856730011525, 0.03727593720314943,
          0.01930697728883252, 0.01000000000000001, 0.005179474679231217, 0.002682695795279729, 0.00138949549437314,
          0.0007196856730011531, 0.0003727593720314947, 0.0001930697728883254, 0.0001000000000000002]
    q = [4.258643191312094, -12.7004345059705, -4.852188357430427, 5.943076168298481, -2.764649066392558,
         -18.57582885927374, 0.4073081174942876, 2.8297017716199, 0.6356121930249937, 4.334300651115951,
         4.228603644876851, 12.99528296551999, -10.49793234475067, -17.86411722110915, 8.16043081031918]
    x = [-4.258643191312046e+00, 9.393193922630394e+00, 1.888905966442421e+01, -2.469934088388301e+00,
         9.628197800226003e+00, 6.034505999261726e+00, -8.288652177085156e+00, -9.172613482098816e+00,
         -4.005465476438092e+01, -2.983244126863757e+01, -7.447972191390734e+00, -6.315368738609618e+00,
         4.555205430378418e+00, 6.362674847968517e+00, -2.000000000000000e+00]
    A = _csc(15, 15, Ap, Ai, Ax)
    Q = _csc(15, 15, list(range(16)), list(range(15)), Qx, -1)
    st = dict(eps_abs=1e-6, eps_rel=1e-6, max_rank_update_fraction=1.0, verbose=0)
    st.update(settings)
    return QP("medium_qp", Q, A, np.array(q), -2 * np.ones(15), 2 * np.ones(15), 0.0, st,
              expect_x=np.array(x), expect_status=1)


def update_qp(**settings) -> QP:
    """tests/src/test_update.c:17-97 (n=2, m=3): x* = {-0.1, 0.3}; after update_bounds {0, 0.15}; after update_q {0.02, 0.18}."""
    A = _csc(3, 2, [0, 2, 4], [0, 2, 1, 2], [10.0, 1.0, 10.0, 1.0])
    Q = _csc(2, 2, [0, 1, 2], [0, 1], [1.0, 1.5], -1)
    st = dict(eps_abs=1e-6, eps_rel=1e-6, scaling=2, proximal=1, verbose=0)
    st.update(settings)
    return QP("update_qp", Q, A, np.array([1.0, -2.0]), np.array([-1.0, -3.0, -0.2]), np.array([1.0, 3.0, 0.2]), 0.0, st,
              expect_x=np.array([-0.1, 0.3]), expect_status=1)


def solver_interface_fixture():
    """tests/src/test_solver_interface.c:43-59,91-159: the 3 x 2 A, 2 x 2 Q (both triangles stored, stype -1) and the
    known answers of mat_vec / mat_tpose_vec / mat_inf_norm_* / ldlchol + ldlsolveLD_neg_dphi."""
    A = _csc(3, 2, [0, 3, 5], [0, 1, 2, 0, 1], [1.0, 3.0, 5.0, 2.0, 4.0])
    Q = _csc(2, 2, [0, 2, 4], [0, 1, 0, 1], [1.0, -1.0, -1.0, 2.0], -1)
    return dict(A=A, Q=Q, Qd=np.array([1.1, -0.5]), Ad=np.array([1.1, -0.5, 20.0]),
                A_Qd=np.array([0.1, 1.3, 5.5]), Q_Qd=np.array([1.6, -2.1]), At_Ad=np.array([99.6, 0.2]),
                col_norms=np.array([5.0, 4.0]), row_norms=np.array([2.0, 4.0, 5.0]),
                neg_rhs=np.array([-1.0, -2.0]), d_noprox=np.array([4.0, 3.0]), gamma=1e3,
                d_prox=np.array([3.989028924198480, 2.993017953122679]))


# ---------------------------------------------------------------------------------------------
# synthetic BASELINE.json configurations
# ---------------------------------------------------------------------------------------------
def random_qp(n, m, dens_A=0.05, dens_M=0.007, seed=0, nonconvex_shift=0.0, name=None, **settings) -> QP:
    """C1 / C5 style: A Bernoulli(dens_A) pattern with N(0,1) values; Q = M M' (lower stored,
    diagonal always present) with M Bernoulli(dens_M) * N(0,1); q ~ N(0,1); bmin=-U, bmax=U."""
    rng = np.random.default_rng(seed)
    if dens_A >= 1.0:
        A = CSC.from_dense(rng.standard_normal((m, n)))
    else:
        A = CSC.from_scipy(sp.random(m, n, density=dens_A, format="csc", random_state=rng,
                                     data_rvs=rng.standard_normal))
    if dens_M >= 1.0:
        M = rng.standard_normal((n, n))
        if n >= 2000:
            # BASELINE-sized dense instances: M is rounded to a 2^-10 grid, so every product and partial sum of
            # M M' is an integer below 2^53 and Q comes out bit-identical under any BLAS / GPU / summation order.
            # The committed reference records (tests/golden/*_n8000*) are therefore valid on every machine.
            Qd = _gram(np.round(M * 1024.0)) / 1048576.0
        else:
            Qd = _gram(M)
        if nonconvex_shift:
            Qd[np.diag_indices(n)] -= nonconvex_shift
        Q = CSC.from_dense(Qd, stype=-1)
    else:
        M = sp.random(n, n, density=dens_M, format="csc", random_state=rng, data_rvs=rng.standard_normal)
        Qs = (M @ M.T).tocsc() + sp.eye(n, format="csc") * 0.0
        Qs = Qs + sp.diags(np.zeros(n) - nonconvex_shift)
        Ql = sp.tril(Qs, format="csc")
        # keep the diagonal structurally present, as the survey generator does
        Ql = (Ql + sp.diags(np.full(n, 1e-300))).tocsc()
        Ql.data[np.abs(Ql.data) <= 1e-299] = 0.0
        Ql.sort_indices()
        Q = CSC(n, n, Ql.indptr, Ql.indices, Ql.data, -1)
    q = rng.standard_normal(n)
    bmin = -rng.random(m)
    bmax = rng.random(m)
    st = dict(eps_abs=1e-6, eps_rel=1e-6, verbose=0)
    st.update(settings)
    return QP(name or f"random_qp_n{n}_m{m}", Q, A, q, bmin, bmax, 0.0, st)


def grid_qp(g, seed=0, box=1.0, diag_shift=0.0, name=None, **settings) -> QP:
    """C2 stand-in (the Maros-Meszaros files are absent -- SYNTHETIC, structurally similar): variables on a
    g x g grid, Q = weighted 5-point Laplacian + 0.05 I (lower triangle stored), A = [edge differences (2 g (g-1) rows);
    cell sums over the 4 corners of every cell ((g-1)^2 rows); identity box rows (g^2)].  Q + A'A is a 9-point stencil, so the
    Newton system stays sparse (CONT-300 / AUG2DCQP class) and goes through the supernodal factorization."""
    rng = np.random.default_rng(seed)
    n = g * g
    idx = np.arange(n).reshape(g, g)
    rows, cols, vals = [], [], []
    r = 0
    edges = []
    for (a, b) in ((idx[:, :-1], idx[:, 1:]), (idx[:-1, :], idx[1:, :])):
        a, b = a.ravel(), b.ravel()
        k = a.size
        w = 0.5 + rng.random(k)
        rows += [np.arange(r, r + k)] * 2
        cols += [a, b]
        vals += [w, -w]
        edges.append((a, b))
        r += k
    c00, c01, c10, c11 = idx[:-1, :-1].ravel(), idx[:-1, 1:].ravel(), idx[1:, :-1].ravel(), idx[1:, 1:].ravel()
    k = c00.size
    for cc in (c00, c01, c10, c11):
        rows.append(np.arange(r, r + k)); cols.append(cc); vals.append(0.25 * (0.5 + rng.random(k)))
    r += k
    rows.append(np.arange(r, r + n)); cols.append(np.arange(n)); vals.append(np.ones(n))
    m = r + n
    A = sp.csc_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(m, n))
    A.sort_indices()
    ea = np.concatenate([e[0] for e in edges]); eb = np.concatenate([e[1] for e in edges])
    w = 0.2 + rng.random(ea.size)
    Lap = sp.csc_matrix((np.concatenate([-w, -w]), (np.concatenate([ea, eb]), np.concatenate([eb, ea]))), shape=(n, n))
    Lap = Lap + sp.diags(-np.asarray(Lap.sum(axis=1)).ravel() + 0.05 - diag_shift)   # diag_shift > 0.05: indefinite Q
    Ql = sp.tril(Lap, format="csc"); Ql.sort_indices()
    q = 2.0 * rng.standard_normal(n)
    ne = m - n
    bmin = np.concatenate([-0.2 * rng.random(ne) - 0.05, -box * (0.2 + rng.random(n))])
    bmax = np.concatenate([0.2 * rng.random(ne) + 0.05, box * (0.2 + rng.random(n))])
    st = dict(eps_abs=1e-6, eps_rel=1e-6, verbose=0)
    st.update(settings)
    return QP(name or f"grid_qp_g{g}", CSC(n, n, Ql.indptr, Ql.indices, Ql.data, -1), CSC.from_scipy(A), q, bmin, bmax, 0.0, st)


def probe_qp(n=1000, m=2000, dens_A=0.05, dens_M=0.007, nonconvex=False, **settings) -> QP:
    """The survey's probe instance, bit for bit: xorshift64 (s ^= s << 13; s ^= s >> 7; s ^= s << 17, start
    88172645463325252), uniform u = (s >> 11) / 2^53, normal = sqrt(-2 ln(u1 + 1e-300)) cos(2 pi u2); A drawn column-major (one
    uniform per entry, a normal when it is < dens_A), then M the same way, Q = M M' accumulated in k-outer order (optionally
    Q_ii -= 1), stored lower with exact-zero off-diagonals dropped and the diagonal always kept, then q (n normals), then per row
    bmin = -u, bmax = u.  With the defaults: nnz(A) = 99 803, nnz(Q lower) = 25 154 and the reference returns solved, 52 / 4
    iterations, objective -4.0571425653e+01 (appendix D) -- an oracle cross-check that does not depend on numpy's generators."""
    import math
    mask = (1 << 64) - 1
    state = [88172645463325252]

    def uni():
        s = state[0]
        s ^= (s << 13) & mask
        s ^= s >> 7
        s ^= (s << 17) & mask
        state[0] = s
        return (s >> 11) / 9007199254740992.0

    def normal():
        u1 = uni()
        u2 = uni()
        return math.sqrt(-2.0 * math.log(u1 + 1e-300)) * math.cos(2.0 * math.pi * u2)

    def draw(nrow, ncol, dens):
        p, idx, val = [0], [], []
        for _ in range(ncol):
            for i in range(nrow):
                if uni() < dens:
                    idx.append(i)
                    val.append(normal())
            p.append(len(idx))
        return p, idx, val

    Ap, Ai, Ax = draw(m, n, dens_A)
    Mp, Mi, Mx = draw(n, n, dens_M)
    Qd = np.zeros((n, n))
    for k in range(n):                                   # k-outer accumulation: Q += M[:, k] M[:, k]'
        rows = Mi[Mp[k]:Mp[k + 1]]
        vals = Mx[Mp[k]:Mp[k + 1]]
        for a, va in zip(rows, vals):
            for b, vb in zip(rows, vals):
                Qd[a, b] += va * vb
    if nonconvex:
        Qd[np.diag_indices(n)] -= 1.0
    Qp, Qi, Qx = [0], [], []
    for j in range(n):
        col = Qd[j:, j]
        keep = np.nonzero(col)[0]
        if keep.size == 0 or keep[0] != 0:
            keep = np.concatenate([[0], keep])           # the diagonal is always stored
        Qi.extend((keep + j).tolist())
        Qx.extend(col[keep].tolist())
        Qp.append(len(Qi))
    q = np.array([normal() for _ in range(n)])
    bmin, bmax = np.empty(m), np.empty(m)
    for i in range(m):
        bmin[i] = -uni()
        bmax[i] = uni()
    st = dict(eps_abs=1e-6, eps_rel=1e-6, verbose=0)
    if nonconvex:
        st["nonconvex"] = 1
    st.update(settings)
    return QP(f"probe_qp_n{n}_m{m}", _csc(n, n, Qp, Qi, Qx, -1), _csc(m, n, Ap, Ai, Ax), q, bmin, bmax, 0.0, st)


def kkt_standin_qp(n=2500, seed=0, dense_rows=3, name=None, **settings) -> QP:
    """AUG2DCQP-class stand-in (SYNTHETIC: the Maros-Meszaros files are absent): a sparse QP whose Schur complement
    Q + A'A FILLS IN.  Q = SPD tridiagonal, A = [banded difference rows (n - 1); `dense_rows` coupling rows that touch every
    variable (budget-type constraints); identity box rows (n)].  One active coupling row makes A_J' Sigma A_J completely dense,
    while the KKT matrix [Q + I/gamma, A_J'; A_J, -inv(Sigma_J)] keeps nnz(A) + nnz(Q) entries -- the case
    qpalm_set_factorization_method (src/solver_interface.c:20-66) sends to FACTORIZE_KKT."""
    rng = np.random.default_rng(seed)
    main = 2.0 + rng.random(n)
    off = -0.5 * rng.random(n - 1)
    Ql = sp.diags([main, off], [0, -1], format="csc")
    Ql.sort_indices()
    rows = [sp.diags([np.ones(n - 1), -np.ones(n - 1)], [0, 1], shape=(n - 1, n), format="csr")]
    rows.append(sp.csr_matrix(0.5 + rng.random((dense_rows, n))))
    rows.append(sp.eye(n, format="csr"))
    A = sp.vstack(rows).tocsc()
    A.sort_indices()
    m = A.shape[0]
    q = rng.standard_normal(n)
    bmin = np.concatenate([-0.3 * np.ones(n - 1), -np.ones(dense_rows), -0.5 - rng.random(n)])
    bmax = np.concatenate([0.3 * np.ones(n - 1), np.ones(dense_rows), 0.5 + rng.random(n)])
    st = dict(eps_abs=1e-6, eps_rel=1e-6, verbose=0)
    st.update(settings)
    return QP(name or f"kkt_standin_n{n}", CSC(n, n, Ql.indptr, Ql.indices, Ql.data, -1), CSC.from_scipy(A), q, bmin, bmax, 0.0, st)


def _gram(M: np.ndarray) -> np.ndarray:
    """M M' in fp64; uses the GPU through torch when one is visible (data generation only)."""
    try:
        import torch
        if torch.cuda.is_available() and M.shape[0] >= 2000:
            t = torch.from_numpy(M).cuda()
            return (t @ t.T).cpu().numpy()
    except Exception:
        pass
    return M @ M.T


def dense_qp(n, m, seed=0, **settings) -> QP:
    """C3: both densities 1.0 (A and Q dense, still passed in CSC as the reference requires)."""
    return random_qp(n, m, 1.0, 1.0, seed, name=f"dense_qp_n{n}_m{m}", **settings)


def nonconvex_random_qp(n, m, seed=0, **settings) -> QP:
    """C5: C1-style with Q <- Q - I and nonconvex=1."""
    settings.setdefault("nonconvex", 1)
    return random_qp(n, m, 0.05, 0.007, seed, nonconvex_shift=1.0, name=f"nonconvex_qp_n{n}", **settings)


@dataclass
class BatchQP:
    """C4: one shared (Q, A) and nb instances differing in q, bmin, bmax."""
    Q: CSC
    A: CSC
    q: np.ndarray      # nb x n
    bmin: np.ndarray   # nb x m
    bmax: np.ndarray   # nb x m
    settings: dict

    def instance(self, k) -> QP:
        return QP(f"mpc_{k}", self.Q, self.A, self.q[k].copy(), self.bmin[k].copy(), self.bmax[k].copy(),
                  0.0, dict(self.settings))


def mpc_batch(nb, n=240, m0=709, seed=0, **settings) -> BatchQP:
    """chain80w-sized sweep (simulations/chain80w/info.txt:20-22, simulations/chain80w.m:39-51,90-104):
    Q 240x240 SPD dense, A = [A0 (709 x 240 dense); I_240]  =>  m = 949; instances are perturbations
    of a base instance, one PRNG stream per instance (seed + k).  Settings follow chain80w.m:
    scaling=2, proximal=FALSE, eps_abs_in=eps_rel_in=1, eps_prim_inf=eps_dual_inf=1e-6."""
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((n, n)) / np.sqrt(n)
    Qd = M @ M.T + 0.1 * np.eye(n)
    A0 = rng.standard_normal((m0, n)) / np.sqrt(n)
    Ad = np.vstack([A0, np.eye(n)])
    m = m0 + n
    xfeas = rng.standard_normal(n) * 0.5
    base_q = rng.standard_normal(n)
    q = np.empty((nb, n))
    bmin = np.empty((nb, m))
    bmax = np.empty((nb, m))
    for k in range(nb):
        r = np.random.default_rng(seed * 100003 + k + 1)
        xk = xfeas + 0.1 * r.standard_normal(n)
        ax = Ad @ xk
        q[k] = base_q + 0.3 * r.standard_normal(n)
        bmin[k] = ax - r.random(m) * 0.5
        bmax[k] = ax + r.random(m) * 0.5
    st = dict(eps_abs=1e-6, eps_rel=1e-6, verbose=0, scaling=2, proximal=0, eps_abs_in=1.0, eps_rel_in=1.0,
              eps_prim_inf=1e-6, eps_dual_inf=1e-6)
    st.update(settings)
    return BatchQP(CSC.from_dense(Qd, -1), CSC.from_dense(Ad, 0), q, bmin, bmax, st)
