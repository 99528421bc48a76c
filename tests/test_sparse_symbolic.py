"""Host-side symbolic analysis of the supernodal sparse Cholesky (csrc/sparse_sym.cu): runs without a GPU.

The structure arrays are validated two ways: (i) the true Cholesky pattern of the permuted matrix lies inside the predicted
supernodal structure, (ii) a numpy emulation of the device's multifrontal algorithm (level order, pull-based extend-add through
`rel`, partial factorization per front) driven ONLY by those arrays reproduces numpy's Cholesky factor.  The emulation is test
infrastructure for the host logic, not a product path.
"""
import numpy as np
import pytest
import scipy.sparse as sp

from qpalm_b200 import problems
from qpalm_b200.abi import CSC
from qpalm_b200.sparse import symbolic


def union_matrix(p, rng, active_frac=0.6):
    Q = p.Q.to_scipy().toarray()
    A = p.A.to_scipy().toarray()
    act = rng.random(p.m) < active_frac
    sig = 0.5 + rng.random(p.m)
    H = Q + A[act].T @ (sig[act, None] * A[act]) + 1e-3 * np.eye(p.n)
    return H


def emulate_multifrontal(sym, H):
    """Numeric multifrontal Cholesky driven by the symbolic arrays; returns the dense permuted factor."""
    n, ns_ = sym["n"], sym["nsuper"]
    perm = sym["perm"]
    Hp = H[np.ix_(perm, perm)]
    first, ro, rowidx, rel = sym["sn_first"], sym["rows_off"], sym["rowidx"], sym["rel"]
    L = np.zeros((n, n))
    U = [None] * ns_
    for lvl in range(sym["nlevels"]):
        for s in sym["lvl_sn"][sym["lvl_ptr"][lvl]:sym["lvl_ptr"][lvl + 1]]:
            f, l = first[s], first[s + 1]
            nsz = l - f
            rows = rowidx[ro[s]:ro[s + 1]]
            idx = np.concatenate([np.arange(f, l), rows])
            nf = idx.size
            F = np.zeros((nf, nf))
            F[:, :nsz] = Hp[np.ix_(idx, np.arange(f, l))]
            F[:nsz, :nsz] = np.tril(F[:nsz, :nsz])
            for c in sym["child_idx"][sym["child_ptr"][s]:sym["child_ptr"][s + 1]]:
                r = rel[ro[c]:ro[c + 1]]
                Uc = np.tril(U[c])
                F[np.ix_(r, r)] += Uc
                U[c] = None
            F = np.tril(F)
            L11 = np.linalg.cholesky(F[:nsz, :nsz] + np.tril(F[:nsz, :nsz], -1).T)
            L21 = np.linalg.solve(L11, F[nsz:, :nsz].T).T
            U[s] = F[nsz:, nsz:] - np.tril(L21 @ L21.T)
            L[np.ix_(idx[:nsz], np.arange(f, l))] = L11
            L[np.ix_(rows, np.arange(f, l))] = L21
    return L, Hp


CASES = [("grid6", lambda: problems.grid_qp(6, seed=1)), ("grid13", lambda: problems.grid_qp(13, seed=2)),
         ("rand_sparse", lambda: problems.random_qp(80, 120, 0.03, 0.02, seed=3)),
         ("rand_dense_pattern", lambda: problems.random_qp(40, 60, 0.3, 0.2, seed=4)),
         ("basic", lambda: problems.basic_qp()), ("medium", lambda: problems.medium_qp())]


@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_symbolic_structure_and_emulated_factorization(name, make):
    p = make()
    sym = symbolic(p.Q, p.A)
    n = p.n
    assert sorted(sym["perm"].tolist()) == list(range(n))
    assert np.array_equal(sym["iperm"][sym["perm"]], np.arange(n))
    first = sym["sn_first"]
    assert first[0] == 0 and first[-1] == n and np.all(np.diff(first) > 0)
    # parents come later; levels respect the tree; rel is monotone and in range
    level = np.zeros(sym["nsuper"], dtype=int)
    for l in range(sym["nlevels"]):
        level[sym["lvl_sn"][sym["lvl_ptr"][l]:sym["lvl_ptr"][l + 1]]] = l
    ro = sym["rows_off"]
    for s in range(sym["nsuper"]):
        par = sym["sn_parent"][s]
        rows = sym["rowidx"][ro[s]:ro[s + 1]]
        assert np.all(np.diff(rows) > 0) and (rows.size == 0 or rows[0] >= first[s + 1])
        if rows.size:
            assert par > s and level[par] > level[s]
            pidx = np.concatenate([np.arange(first[par], first[par + 1]), sym["rowidx"][ro[par]:ro[par + 1]]])
            assert np.array_equal(pidx[sym["rel"][ro[s]:ro[s + 1]]], rows)
        else:
            assert par == -1
    rng = np.random.default_rng(0)
    H = union_matrix(p, rng)
    L, Hp = emulate_multifrontal(sym, H)
    Lref = np.linalg.cholesky(Hp)
    # (i) the true factor lives inside the predicted structure: whatever the emulation left at zero is zero in the factor
    assert np.max(np.abs(Lref[L == 0])) < 1e-12
    # (ii) and the emulation driven by the arrays reproduces it
    assert np.max(np.abs(L - Lref)) < 1e-9 * max(1.0, np.max(np.abs(Lref)))
    # storage accounting
    nf = np.diff(first) + np.diff(ro)
    assert sym["nnzL"] == int(np.sum(nf * np.diff(first)))
    assert sym["upd_entries"] == int(np.sum(np.diff(ro) ** 2))


def test_fill_and_tree_shape_on_a_grid():
    """A 60 x 60 nine-point grid: the factor stays a few percent of the dense triangle and the assembly tree is bushy
    (few levels = few launches in sequence on the device)."""
    p = problems.grid_qp(60, seed=5)
    sym = symbolic(p.Q, p.A)
    n = p.n
    assert sym["nnzL"] < 0.03 * n * (n + 1) / 2, sym["nnzL"]
    assert sym["nnzL"] < 1.2 * 139570          # true fill of the natural (banded) ordering on this matrix
    assert sym["nlevels"] <= 40 and sym["max_nf"] <= 4 * 60
