"""The algebra behind csrc/updown_gen.cu, restated in numpy (no GPU): the rank-k update / downdate of a Cholesky factor as
L_new = L K, K = chol(I + What S What'), What = inv(L) W, with the closed-form generators the kernels compute:

    M_j = S + sum_{p < j} what_p what_p'        d_j = sqrt(1 + what_j' inv(M_j) what_j)        z_j = inv(M_j) what_j / d_j
    L_new(r, j) = d_j L(r, j) + t_r' z_j,       t_r = W_r - sum_{i <= j} L(r, i) what_i         inv(M_{j+1}) = inv(M_j) - z_j z_j'

restarted from the Gram prefix every 32 rows exactly as k_gen_gram / k_gen_scan / k_gen_rows do, and the row recurrence cut into
128-column tiles that each start from the stored running sum of the triangular solve (k_fwd_multi's Tdump -> k_gen_apply).
This replaces cholmod_updown (Modify/t_cholmod_updown_numkr.c:289-376) as called from src/solver_interface.c:407-503."""
import numpy as np
import pytest


def generator_update(L, W, kpos, block=32, tile=128):
    n, k = W.shape
    S = np.diag([1.0] * kpos + [-1.0] * (k - kpos))
    What = np.linalg.solve(L, W)                 # k_fwd_multi (forward substitution, k right-hand sides)
    D, Z = np.zeros(n), np.zeros((n, k))
    for b0 in range(0, n, block):                # k_gen_gram + k_gen_scan: M at the start of each 32-row block
        P = np.linalg.inv(S + What[:b0].T @ What[:b0])     # k_gen_rows: Gauss-Jordan, then Sherman-Morrison steps
        for j in range(b0, min(n, b0 + block)):
            g = P @ What[j]
            d2 = 1.0 + What[j] @ g
            assert d2 > 0.0
            D[j] = np.sqrt(d2)
            Z[j] = g / D[j]
            P = P - np.outer(Z[j], Z[j])
    Lnew = np.zeros_like(L)
    for r0 in range(0, n, tile):                 # k_gen_apply: tiles are independent given the running sums
        for c0 in range(0, r0 + 1, tile):
            T = W[r0:r0 + tile] - L[r0:r0 + tile, :c0] @ What[:c0]        # Tdump of tile (r0, c0)
            for r in range(r0, min(n, r0 + tile)):
                t = T[r - r0].copy()
                for j in range(c0, min(c0 + tile, r + 1)):
                    t = t - L[r, j] * What[j]
                    Lnew[r, j] = D[j] * L[r, j] + t @ Z[j]
    return Lnew


@pytest.mark.parametrize("n,k,kpos", [(70, 1, 1), (150, 5, 5), (150, 4, 0), (200, 7, 4), (260, 12, 9)])
def test_generator_form_equals_cholesky_of_the_modified_matrix(n, k, kpos):
    rng = np.random.default_rng(n + k)
    A = rng.standard_normal((n, n))
    Wn = rng.standard_normal((n, k - kpos))
    H = A @ A.T + n * np.eye(n) + Wn @ Wn.T      # the downdate is valid on its own, as in QPALM (leaving rows were active)
    Wp = 3.0 * rng.standard_normal((n, kpos))
    W = np.hstack([Wp, Wn])
    L = np.linalg.cholesky(H)
    Lnew = generator_update(L, W, kpos)
    Lref = np.linalg.cholesky(H + Wp @ Wp.T - Wn @ Wn.T)
    assert np.max(np.abs(Lnew - Lref)) <= 1e-12 * np.max(np.abs(Lref))
    assert np.all(np.diag(Lnew) > 0)


def test_a_downdate_past_definiteness_is_detected():
    L = np.linalg.cholesky(4.0 * np.eye(40))
    W = np.zeros((40, 1)); W[17, 0] = 3.0         # 4 - 9 < 0
    with pytest.raises(AssertionError):
        generator_update(L, W, 0)
