// sparse.cuh -- supernodal (multifrontal) sparse Cholesky of the Newton system  H = Q + A_J' Sigma_J A_J + I/gamma
// for problems whose Schur complement stays sparse (BASELINE config 2 class).  Replaces, for sparse problems,
// cholmod_analyze + cholmod_factorize + cholmod_solve + cholmod_updown as used by
// /root/reference/src/solver_interface.c:319-519.
//
// Design (B200-first, not CHOLMOD's):
//  * ONE host-side symbolic analysis at setup over the UNION pattern  Q + A'A  (every constraint active): approximate
//    minimum-degree ordering, elimination tree, postorder, supernodes with relaxed amalgamation, assembly-tree levels.
//    The active set only changes VALUES, never the pattern, so the device data structure is static: no reallocation,
//    no re-analysis, rank-k updates never grow columns (cholmod_updown's column reallocation disappears).
//  * factor storage: one dense column-major panel per supernode, (ns + nr) x ns, rows = [own columns ; sorted row
//    structure]; one nr x nr update (Schur) block per supernode for the multifrontal extend-add.
//  * numeric factorization: level by level over the assembly tree (leaves first); all fronts of a level are processed
//    by the same launches (pull-based extend-add in a fixed child order => bit-reproducible, no atomics).
//  * solves: level-scheduled forward / backward sweeps with per-supernode update vectors (again pull-based).
//  * rank-k update/downdate: CHOLMOD's (alpha, gamma) recurrence restated for LL', one CTA walking the union of the
//    etree paths of the modified columns (the paths are marked on the device).
#pragma once
#include "common.cuh"

namespace qb {

struct SparseChol;   // symbolic structure (host + device copies) and work buffers

struct SparseCholInfo {
  int n = 0, nsuper = 0, nlevels = 0, max_ns = 0, max_nf = 0;
  long long nnzL = 0;          // entries of the factor panels (incl. explicit zeros from amalgamation)
  long long upd_entries = 0;   // sum nr^2 of the update blocks
  long long nnzS = 0;          // entries of the union pattern (lower triangle incl. diagonal)
  double flops = 0;            // sum over supernodes of the partial-factorization flops
};

// Host CSR/CSC of A (int32) and the stored lower triangle of Q (int64 CSC, entries with row < col ignored).
// Returns 0 and *out != nullptr on success; returns 0 with *out == nullptr when the union pattern is too dense
// for the sparse path to pay off (caller takes the dense path) unless `force`.
int sparse_chol_analyze(SparseChol **out, int n, int m, const int *Acsc_p, const int *Acsc_i, const int *Acsr_p,
                        const int *Acsr_j, const long long *Qp, const long long *Qi, bool force, cudaStream_t stream);
// The same machinery for a symmetric QUASI-DEFINITE matrix (the KKT path): pattern = lower triangle in int64 CSC, N x N,
// columns >= n_pos have negative pivots.  Numeric phase: sparse_chol_assemble(with_Q = true, Q = the matrix as a full
// symmetric CSR, active = nullptr, beta = 0) + sparse_chol_factor (L S L') + sparse_chol_solve.
int sparse_ldl_analyze(SparseChol **out, int N, int n_pos, const long long *Kp, const long long *Ki, cudaStream_t stream);
void sparse_chol_destroy(SparseChol *sc);
const SparseCholInfo *sparse_chol_info(const SparseChol *sc);
size_t sparse_chol_factor_doubles(const SparseChol *sc);   // doubles of one numeric factor (panels)

// panels <- (with_Q ? Q : 0) + sum_r [active_r] sigma_r a_r a_r' + beta I, in the permuted panel layout.
// Q is the engine's full symmetric CSR; A in CSR and CSC (device, int32).  active == nullptr: no constraint terms.
int sparse_chol_assemble(SparseChol *sc, cudaStream_t s, double *panels, bool with_Q, const int *Qp, const int *Qi,
                         const double *Qx, const int *Acsc_p, const int *Acsc_i, const double *Acsc_x, const int *Acsr_p,
                         const int *Acsr_j, const double *Acsr_x, const int *active, const double *sigma, double beta);
// in-place numeric factorization of the assembled panels; *info_dev != 0 on a non-positive pivot
int sparse_chol_factor(SparseChol *sc, cudaStream_t s, double *panels, int *info_dev);
// out = (L L')^{-1} (negate ? -rhs : rhs), rhs/out of length n in the ORIGINAL ordering
int sparse_chol_solve(SparseChol *sc, cudaStream_t s, const double *panels, const double *rhs, double *out, bool negate);
// L L' <- L L' + sign * sum_c w_c w_c',  w_c = scale * A(list[c], :)', cnt <= 8
int sparse_chol_updown(SparseChol *sc, cudaStream_t s, double *panels, const int *Acsr_p, const int *Acsr_j,
                       const double *Acsr_x, const int *list, const double *scale, bool scale_by_row, int cnt, int sign,
                       int *info_dev);
// out[i] = sum_j |S_ij| of the symmetric matrix held (unfactored) in the panels, permuted order (only the max is used)
int sparse_chol_abs_rowsums(SparseChol *sc, cudaStream_t s, const double *panels, double *out_n);
// dense copy of the factor in the PERMUTED order (n x n column-major, lower) + the permutation, for the parity tests
int sparse_chol_download(SparseChol *sc, cudaStream_t s, const double *panels, double *L_host, long long *perm_host);

}  // namespace qb
