/*
 * qpalm_oracle.c -- TEST INFRASTRUCTURE ONLY (see qpalm_oracle.h).  Parity status: PINNED
 * (tests/test_oracle.py: reference known-answer vectors + live/committed outputs of oracle/_ref).
 *
 * A plain-C, single-threaded restatement of the QPALM algorithm as the reference's CHOLMOD build
 * executes it.  File:line citations are relative to /root/reference.  The reference delegates the
 * Newton system to CHOLMOD's sparse LDL'; here the same system is factorised by a dense unit-lower
 * LDL' (no pivoting, natural order -- the reference also uses NATURAL ordering,
 * src/solver_interface.c:523-541) and modified with the rank-1 recurrence of
 * suitesparse/CHOLMOD/Modify/t_cholmod_updown_numkr.c:289-376, so results agree with the reference
 * to rounding.
 */
#include "qpalm_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#define c_max(a, b) (((a) > (b)) ? (a) : (b))
#define c_min(a, b) (((a) < (b)) ? (a) : (b))
#define c_absval(x) (((x) < 0) ? -(x) : (x))
#define MIN_SCALING 1e-12

/* ---------------------------------------------------------------------------------------------
 * private state hung off work->solver->LD
 * ------------------------------------------------------------------------------------------- */
#define TRACE_MAX 20000
typedef struct {
  size_t n, m;
  c_int *Ap, *Ai; c_float *Ax;        /* scaled A, CSC (copy)                                   */
  c_int *Qp, *Qi; c_float *Qx;        /* scaled Q, CSC, stype -1 (copy)                         */
  c_float *L, *D;                     /* dense unit-lower L (n x n col-major) and D             */
  c_float *H;                         /* scratch n x n                                          */
  c_float *w;                         /* scratch n                                              */
  c_float *Lq, *Dq;                   /* LDL' of Q (+0) for the dual objective                  */
  solver_sparse Aview, Qview;
  OracleTraceEntry *trace; c_int ntrace;
  c_int last_refactor;
} Aux;

static Aux *aux_of(const QPALMWorkspace *w) { return (Aux *)w->solver->LD; }

/* ---------------------------------------------------------------------------------------------
 * vector kernels (src/lin_alg.c) -- same summation orders as the reference
 * ------------------------------------------------------------------------------------------- */
static c_float *vcopy(const c_float *a, size_t n) {
  c_float *b = (c_float *)malloc((n ? n : 1) * sizeof(c_float));
  for (size_t i = 0; i < n; i++) b[i] = a[i];
  return b;
}
static void pcopy(const c_float *a, c_float *b, size_t n) { for (size_t i = 0; i < n; i++) b[i] = a[i]; }
static void vset(c_float *a, c_float s, size_t n) { for (size_t i = 0; i < n; i++) a[i] = s; }
static c_float vprod(const c_float *a, const c_float *b, size_t n) { /* lin_alg.c:72-86 */
  c_float prod = 0.0; size_t i = 0;
  if (n >= 4) for (; i <= n - 4; i += 4) prod += (a[i]*b[i] + a[i+1]*b[i+1] + a[i+2]*b[i+2] + a[i+3]*b[i+3]);
  for (; i < n; i++) prod += a[i] * b[i];
  return prod;
}
static c_float vnorm2(const c_float *a, size_t n) { return sqrt(vprod(a, a, n)); }
static c_float vnorminf(const c_float *a, size_t n) { /* lin_alg.c:128-166; max is order independent */
  c_float mx = 0.; for (size_t i = 0; i < n; i++) { c_float s = c_absval(a[i]); mx = s > mx ? s : mx; } return mx;
}
static void vadd_scaled(const c_float *a, const c_float *b, c_float *c, c_float sc, size_t n) {
  for (size_t i = 0; i < n; i++) c[i] = a[i] + sc * b[i];
}
static void vmult_add_scaled(c_float *a, const c_float *b, c_float s1, c_float s2, size_t n) {
  for (size_t i = 0; i < n; i++) a[i] = s1 * a[i] + s2 * b[i];
}
static void vewprod(const c_float *a, const c_float *b, c_float *c, size_t n) { for (size_t i = 0; i < n; i++) c[i] = a[i]*b[i]; }
static void vscale(c_float *a, c_float s, size_t n) { for (size_t i = 0; i < n; i++) a[i] *= s; }
static void vrecip(const c_float *a, c_float *b, size_t n) { for (size_t i = 0; i < n; i++) b[i] = 1.0 / a[i]; }

/* ---------------------------------------------------------------------------------------------
 * matrix kernels (src/solver_interface.c:252-314; CHOLMOD/MatrixOps/t_cholmod_sdmult.c)
 * ------------------------------------------------------------------------------------------- */
int oracle_mat_vec(const solver_sparse *A, const c_float *x, c_float *y) {
  const c_int *Ap = (const c_int *)A->p, *Ai = (const c_int *)A->i; const c_float *Ax = (const c_float *)A->x;
  size_t ncol = A->ncol, nrow = A->nrow;
  c_float *xx = (c_float *)x, *tmp = NULL;
  if (x == y) { tmp = vcopy(x, ncol); xx = tmp; }              /* solver_interface.c:257-260 */
  for (size_t i = 0; i < nrow; i++) y[i] = 0.0;
  if (A->stype == 0) {                                         /* t_cholmod_sdmult.c:290-322 */
    for (size_t j = 0; j < ncol; j++) {
      c_float xj = xx[j];
      for (c_int p = Ap[j]; p < Ap[j+1]; p++) y[Ai[p]] += Ax[p] * xj;
    }
  } else {                                                     /* t_cholmod_sdmult.c:436-480 */
    for (size_t j = 0; j < ncol; j++) {
      c_float yj = 0.0, xj = xx[j];
      for (c_int p = Ap[j]; p < Ap[j+1]; p++) {
        c_int i = Ai[p];
        if (i == (c_int)j) y[i] += Ax[p] * xj;
        else if ((A->stype > 0 && i < (c_int)j) || (A->stype < 0 && i > (c_int)j)) {
          y[i] += Ax[p] * xj; yj += Ax[p] * xx[i];
        }
      }
      y[j] += yj;
    }
  }
  free(tmp);
  return 0;
}

int oracle_mat_tpose_vec(const solver_sparse *A, const c_float *x, c_float *y) {
  const c_int *Ap = (const c_int *)A->p, *Ai = (const c_int *)A->i; const c_float *Ax = (const c_float *)A->x;
  if (A->stype != 0) return oracle_mat_vec(A, x, y);
  for (size_t j = 0; j < A->ncol; j++) {                       /* t_cholmod_sdmult.c:131-170 */
    c_float yj = 0.0;
    for (c_int p = Ap[j]; p < Ap[j+1]; p++) yj += Ax[p] * x[Ai[p]];
    y[j] = yj;
  }
  return 0;
}

int oracle_mat_inf_norm_cols(const solver_sparse *M, c_float *E) { /* solver_interface.c:276-293 */
  const c_int *Mp = (const c_int *)M->p; const c_float *Mx = (const c_float *)M->x;
  for (size_t j = 0; j < M->ncol; j++) {
    E[j] = 0.; for (c_int k = Mp[j]; k < Mp[j+1]; k++) E[j] = c_max(c_absval(Mx[k]), E[j]);
  }
  return 0;
}
int oracle_mat_inf_norm_rows(const solver_sparse *M, c_float *E) { /* solver_interface.c:295-314 */
  const c_int *Mp = (const c_int *)M->p, *Mi = (const c_int *)M->i; const c_float *Mx = (const c_float *)M->x;
  for (size_t j = 0; j < M->nrow; j++) E[j] = 0.;
  for (size_t j = 0; j < M->ncol; j++)
    for (c_int k = Mp[j]; k < Mp[j+1]; k++) { c_int i = Mi[k]; E[i] = c_max(c_absval(Mx[k]), E[i]); }
  return 0;
}

/* Ruiz equilibration, src/scaling.c:34-113.  Qx_iter is work->Qx (all zeros at setup, so there
 * c = 1/max(1,|D q|inf); nonzero when re-entered from qpalm_update_settings, src/qpalm.c:754-771). */
static void scale_data_impl(solver_sparse *A, solver_sparse *Q, c_float *q, c_float *bmin, c_float *bmax,
                            c_int iters, c_float *D, c_float *E, c_float *c_out, c_float *Qx_iter) {
  size_t n = Q->ncol, m = A->nrow;
  c_int *Ap = (c_int *)A->p, *Ai = (c_int *)A->i; c_float *Ax = (c_float *)A->x;
  c_int *Qp = (c_int *)Q->p, *Qi = (c_int *)Q->i; c_float *Qx = (c_float *)Q->x;
  c_float *Dt = (c_float *)malloc((n + 1) * sizeof(c_float)), *Et = (c_float *)malloc((m + 1) * sizeof(c_float));
  vset(D, 1, n); vset(E, 1, m);
  for (c_int it = 0; it < iters; it++) {
    oracle_mat_inf_norm_cols(A, Dt); oracle_mat_inf_norm_rows(A, Et);
    for (size_t i = 0; i < n; i++) Dt[i] = 1.0 / sqrt(Dt[i] < MIN_SCALING ? 1.0 : Dt[i]);
    for (size_t i = 0; i < m; i++) Et[i] = 1.0 / sqrt(Et[i] < MIN_SCALING ? 1.0 : Et[i]);
    for (size_t j = 0; j < A->ncol; j++) for (c_int p = Ap[j]; p < Ap[j+1]; p++) Ax[p] *= Et[Ai[p]]; /* ROW */
    for (size_t j = 0; j < A->ncol; j++) for (c_int p = Ap[j]; p < Ap[j+1]; p++) Ax[p] *= Dt[j];     /* COL */
    vewprod(D, Dt, D, n); vewprod(E, Et, E, m);
  }
  vewprod(D, q, q, n);
  c_float c;
  if (Qx_iter) {
    vewprod(D, Qx_iter, Qx_iter, n);
    vadd_scaled(Qx_iter, q, Dt, 1, n);
    c = 1 / c_max(1.0, vnorminf(Dt, n));
  } else c = 1 / c_max(1.0, vnorminf(q, n));
  vscale(q, c, n);
  for (size_t j = 0; j < n; j++) for (c_int p = Qp[j]; p < Qp[j+1]; p++) Qx[p] *= D[j] * D[Qi[p]];   /* SYM    */
  for (size_t j = 0; j < n; j++) for (c_int p = Qp[j]; p < Qp[j+1]; p++) Qx[p] *= c;                 /* SCALAR */
  vewprod(E, bmin, bmin, m); vewprod(E, bmax, bmax, m);
  *c_out = c;
  free(Dt); free(Et);
}
int oracle_scale_data(solver_sparse *A, solver_sparse *Q, c_float *q, c_float *bmin, c_float *bmax,
                      c_int iters, c_float *D, c_float *E, c_float *c_out) {
  scale_data_impl(A, Q, q, bmin, bmax, iters, D, E, c_out, NULL);
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * dense LDL' (stand-in for cholmod_analyze + cholmod_factorize_p + cholmod_solve + cholmod_updown)
 * ------------------------------------------------------------------------------------------- */
#define Lij(L, n, i, j) (L)[(size_t)(i) + (size_t)(n) * (size_t)(j)]

/* H (lower, col-major) -> unit-lower L and D.  Right-looking, column by column. */
static void ldl_factor(size_t n, c_float *H, c_float *L, c_float *D) {
  for (size_t j = 0; j < n; j++) {
    c_float dj = Lij(H, n, j, j);
    D[j] = dj;
    Lij(L, n, j, j) = 1.0;
    for (size_t i = j + 1; i < n; i++) Lij(L, n, i, j) = Lij(H, n, i, j) / dj;
    for (size_t k = j + 1; k < n; k++) {
      c_float f = Lij(L, n, k, j) * dj;
      if (f != 0.0) for (size_t i = k; i < n; i++) Lij(H, n, i, k) -= Lij(L, n, i, j) * f;
    }
    for (size_t i = 0; i < j; i++) Lij(L, n, i, j) = 0.0;
  }
}
static void ldl_solve(size_t n, const c_float *L, const c_float *D, const c_float *b, c_float *x) {
  for (size_t i = 0; i < n; i++) x[i] = b[i];
  for (size_t j = 0; j < n; j++) { c_float xj = x[j]; if (xj != 0.0) for (size_t i = j + 1; i < n; i++) x[i] -= Lij(L, n, i, j) * xj; }
  for (size_t j = 0; j < n; j++) x[j] /= D[j];
  for (size_t jj = n; jj-- > 0;) { c_float s = x[jj]; for (size_t i = jj + 1; i < n; i++) s -= Lij(L, n, i, jj) * x[i]; x[jj] = s; }
}
/* rank-1 LDL' +- w w'; recurrence of t_cholmod_updown_numkr.c:289-318 (ALPHA_GAMMA) and :352-367 */
static void ldl_updown1(size_t n, c_float *L, c_float *D, c_float *w, int update) {
  c_float alpha = 1.0;
  for (size_t j = 0; j < n; j++) {
    c_float wj = w[j];
    if (wj == 0.0) continue;          /* CHOLMOD only walks the etree path of the nonzeros of w */
    c_float dj = D[j], a, gamma;
    if (update) { a = alpha + (wj * wj) / dj; dj *= a; gamma = -wj / dj; dj /= alpha; }
    else        { a = alpha - (wj * wj) / dj; dj *= a; gamma =  wj / dj; dj /= alpha; }
    alpha = a; D[j] = dj;
    for (size_t i = j + 1; i < n; i++) {
      w[i] -= wj * Lij(L, n, i, j);
      Lij(L, n, i, j) -= gamma * w[i];
    }
    w[j] = 0.0;
  }
}

/* H = Q(lower) + sum_{j active} sigma_j a_j a_j' + beta I   (src/solver_interface.c:372-405, 351-356) */
static void assemble_H(size_t n, size_t m, const c_int *Qp, const c_int *Qi, const c_float *Qx,
                       const c_int *Ap, const c_int *Ai, const c_float *Ax,
                       const c_float *sigma, const c_int *active, c_float beta, c_float *H) {
  memset(H, 0, n * n * sizeof(c_float));
  for (size_t j = 0; j < n; j++) for (c_int p = Qp[j]; p < Qp[j+1]; p++) if (Qi[p] >= (c_int)j) Lij(H, n, Qi[p], j) += Qx[p];
  if (active && m) {
    /* row access to A through a transposed copy (the reference keeps At_sqrt_sigma for this) */
    c_int nnz = Ap[n];
    c_int *Rp = (c_int *)calloc(m + 2, sizeof(c_int)), *Rj = (c_int *)malloc((nnz + 1) * sizeof(c_int));
    c_float *Rx = (c_float *)malloc((nnz + 1) * sizeof(c_float));
    for (c_int p = 0; p < nnz; p++) Rp[Ai[p] + 2]++;
    for (size_t i = 0; i < m; i++) Rp[i + 2] += Rp[i + 1];
    for (size_t j = 0; j < n; j++) for (c_int p = Ap[j]; p < Ap[j+1]; p++) { c_int d = Rp[Ai[p] + 1]++; Rj[d] = (c_int)j; Rx[d] = Ax[p]; }
    for (size_t r = 0; r < m; r++) if (active[r]) {
      c_float s = sigma[r];
      for (c_int a = Rp[r]; a < Rp[r+1]; a++) for (c_int b = a; b < Rp[r+1]; b++)
        Lij(H, n, Rj[b], Rj[a]) += s * Rx[a] * Rx[b];        /* Rj sorted ascending => Rj[b] >= Rj[a] */
    }
    free(Rp); free(Rj); free(Rx);
  }
  for (size_t j = 0; j < n; j++) Lij(H, n, j, j) += beta;
}

int oracle_newton_solve(const solver_sparse *Q, const solver_sparse *A, const c_float *sigma,
        const c_int *active, c_float beta, const c_float *rhs, c_float *d, c_float *L_out) {
  size_t n = Q->ncol, m = A ? A->nrow : 0;
  c_float *H = (c_float *)malloc(n * n * sizeof(c_float)), *L = (c_float *)malloc(n * n * sizeof(c_float));
  c_float *D = (c_float *)malloc(n * sizeof(c_float));
  assemble_H(n, m, (c_int *)Q->p, (c_int *)Q->i, (c_float *)Q->x, A ? (c_int *)A->p : NULL, A ? (c_int *)A->i : NULL,
             A ? (c_float *)A->x : NULL, sigma, active, beta, H);
  ldl_factor(n, H, L, D);
  ldl_solve(n, L, D, rhs, d);
  if (L_out) for (size_t j = 0; j < n; j++) { c_float s = sqrt(D[j]); for (size_t i = 0; i < n; i++) Lij(L_out, n, i, j) = Lij(L, n, i, j) * s; }
  free(H); free(L); free(D);
  return 0;
}

/* L (Cholesky, col-major lower) <- chol(L L' +- W W'); W is n x k col-major */
int oracle_updown(c_int n_, c_int k, c_float *Lc, const c_float *W, c_int update) {
  size_t n = (size_t)n_;
  c_float *L = (c_float *)malloc(n * n * sizeof(c_float)), *D = (c_float *)malloc(n * sizeof(c_float)), *w = (c_float *)malloc(n * sizeof(c_float));
  for (size_t j = 0; j < n; j++) { c_float l = Lij(Lc, n, j, j); D[j] = l * l; for (size_t i = 0; i < n; i++) Lij(L, n, i, j) = (i >= j) ? Lij(Lc, n, i, j) / l : 0.0; }
  for (c_int r = 0; r < k; r++) { pcopy(W + (size_t)r * n, w, n); ldl_updown1(n, L, D, w, (int)update); }
  for (size_t j = 0; j < n; j++) { c_float s = sqrt(D[j]); for (size_t i = 0; i < n; i++) Lij(Lc, n, i, j) = Lij(L, n, i, j) * s; }
  free(L); free(D); free(w);
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * residuals + active set (src/iteration.c:24-48, src/newton.c:122-149)
 * ------------------------------------------------------------------------------------------- */
int oracle_residuals_active_set(const solver_sparse *A,
        const c_float *Ax, const c_float *y, const c_float *sigma, const c_float *bmin, const c_float *bmax,
        const c_float *Qx, const c_float *q, const c_float *x0, c_int proximal, c_float gamma,
        const c_int *active_old,
        c_float *Axys, c_float *z, c_float *pri_res, c_float *yh, c_float *Atyh, c_float *df, c_float *dphi,
        c_int *active, c_int *nb_active, c_int *enter, c_int *nb_enter, c_int *leave, c_int *nb_leave) {
  size_t m = A->nrow, n = A->ncol;
  for (size_t i = 0; i < m; i++) {
    c_float sinv = 1.0 / sigma[i];                    /* work->sigma_inv */
    c_float t = y[i] * sinv;
    Axys[i] = Ax[i] + 1 * t;
    z[i] = c_max(bmin[i], c_min(Axys[i], bmax[i]));
    pri_res[i] = Ax[i] + (-1) * z[i];
    t = pri_res[i] * sigma[i];
    yh[i] = y[i] + 1 * t;
  }
  for (size_t i = 0; i < n; i++) df[i] = Qx[i] + 1 * q[i];
  if (proximal) { c_float sc = -1 / gamma; for (size_t i = 0; i < n; i++) df[i] = df[i] + sc * x0[i]; }
  oracle_mat_tpose_vec(A, yh, Atyh);
  for (size_t i = 0; i < n; i++) dphi[i] = df[i] + 1 * Atyh[i];
  c_int na = 0, ne = 0, nl = 0;
  for (size_t i = 0; i < m; i++) {
    active[i] = ((Axys[i] <= bmin[i]) || (Axys[i] >= bmax[i])) ? 1 : 0;
    na += active[i];
  }
  for (size_t i = 0; i < m; i++) {
    if (active[i] && !active_old[i]) enter[ne++] = (c_int)i;
    if (!active[i] && active_old[i]) leave[nl++] = (c_int)i;
  }
  *nb_active = na; *nb_enter = ne; *nb_leave = nl;
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * exact line search (src/linesearch.c:14-166)
 * ------------------------------------------------------------------------------------------- */
static void stable_sort_ae(array_element *a, array_element *tmp, size_t n) { /* glibc qsort == merge sort: stable */
  if (n < 2) return;
  size_t h = n / 2;
  stable_sort_ae(a, tmp, h); stable_sort_ae(a + h, tmp, n - h);
  size_t i = 0, j = h, k = 0;
  while (i < h && j < n) { if (a[j].x < a[i].x) tmp[k++] = a[j++]; else tmp[k++] = a[i++]; } /* compare(): ties keep order */
  while (i < h) tmp[k++] = a[i++];
  while (j < n) tmp[k++] = a[j++];
  memcpy(a, tmp, n * sizeof(array_element));
}

int oracle_linesearch(c_int m_, c_float eta, c_float beta,
        const c_float *Ad, const c_float *Ax, const c_float *y, const c_float *sigma,
        const c_float *sqrt_sigma, const c_float *bmin, const c_float *bmax,
        c_float *tau, c_float *sorted_s, c_int *sorted_idx, c_int *nL_out) {
  size_t m = (size_t)m_, m2 = 2 * m;
  c_float *delta = (c_float *)malloc((m2 + 1) * sizeof(c_float)), *alpha = (c_float *)malloc((m2 + 1) * sizeof(c_float));
  array_element *s = (array_element *)malloc((m2 + 1) * sizeof(array_element)), *tmp = (array_element *)malloc((m2 + 1) * sizeof(array_element));
  char *P = (char *)malloc(m2 + 1);
  for (size_t i = 0; i < m; i++) {
    c_float t = sqrt_sigma[i] * Ad[i];
    delta[i + m] = t; delta[i] = t * -1;
    c_float u = Ax[i] + (-1) * bmin[i]; u = sigma[i] * u; u = y[i] + 1 * u; alpha[i] = u / sqrt_sigma[i];
    u = bmax[i] + (-1) * Ax[i]; u = sigma[i] * u; u = u + (-1) * y[i]; alpha[i + m] = u / sqrt_sigma[i];
  }
  size_t nL = 0;
  c_float a = 0.0, b = 0.0;                              /* vec_prod_ind sums, ascending index */
  for (size_t i = 0; i < m2; i++) {
    c_float si = alpha[i] / delta[i];
    int inL = si > 0, inP = delta[i] > 0;
    P[i] = (char)inP;
    if (inL) { s[nL].x = si; s[nL].i = i; nL++; }
    if (inL + inP == 1) { a += delta[i] * delta[i]; b += delta[i] * alpha[i]; }
  }
  a = eta + a; b = beta - b;
  stable_sort_ae(s, tmp, nL);
  if (sorted_s) for (size_t i = 0; i < nL; i++) { sorted_s[i] = s[i].x; sorted_idx[i] = (c_int)s[i].i; }
  if (nL_out) *nL_out = (c_int)nL;
  c_float result;
  if (nL == 0 || a * s[0].x + b > 0) { result = -b / a; goto done; }
  {
    size_t i = 0, iz;
    while (i < nL - 1) {
      iz = s[i].i;
      if (P[iz]) { a = a + delta[iz]*delta[iz]; b = b - delta[iz]*alpha[iz]; }
      else       { a = a - delta[iz]*delta[iz]; b = b + delta[iz]*alpha[iz]; }
      i++;
      if (a * s[i].x + b > 0) { result = -b / a; goto done; }
    }
    iz = s[i].i;
    if (P[iz]) { a = a + delta[iz]*delta[iz]; b = b - delta[iz]*alpha[iz]; }
    else       { a = a - delta[iz]*delta[iz]; b = b + delta[iz]*alpha[iz]; }
    result = -b / a;
  }
done:
  *tau = result;
  free(delta); free(alpha); free(s); free(tmp); free(P);
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * LOBPCG (src/nonconvex.c:29-168) with small dense eigen-solvers standing in for LAPACKE
 * ------------------------------------------------------------------------------------------- */
static void jacobi_eig(int n, double A[3][3], double V[3][3], double w[3]) { /* symmetric, ascending */
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) V[i][j] = (i == j);
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0; for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) off += A[i][j] * A[i][j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++) for (int q = p + 1; q < n; q++) {
      if (A[p][q] == 0.0) continue;
      double th = (A[q][q] - A[p][p]) / (2 * A[p][q]);
      double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1)), c = 1 / sqrt(t * t + 1), s = t * c;
      for (int k = 0; k < n; k++) { double akp = A[k][p], akq = A[k][q]; A[k][p] = c*akp - s*akq; A[k][q] = s*akp + c*akq; }
      for (int k = 0; k < n; k++) { double apk = A[p][k], aqk = A[q][k]; A[p][k] = c*apk - s*aqk; A[q][k] = s*apk + c*aqk; }
      for (int k = 0; k < n; k++) { double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c*vkp - s*vkq; V[k][q] = s*vkp + c*vkq; }
    }
  }
  for (int i = 0; i < n; i++) w[i] = A[i][i];
  for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) if (w[j] < w[i]) {
    double t = w[i]; w[i] = w[j]; w[j] = t;
    for (int k = 0; k < n; k++) { t = V[k][i]; V[k][i] = V[k][j]; V[k][j] = t; }
  }
}
/* generalized B y = lambda C y (C SPD): returns smallest eigenpair, y'Cy = 1 (dsygv itype 1) */
static double gen_eig_min(int n, double B[3][3], double Cm[3][3], double y[3]) {
  double R[3][3] = {{0}}, Ri[3][3] = {{0}}, T[3][3], M[3][3], V[3][3], w[3];
  for (int j = 0; j < n; j++) {                       /* Cm = R R', R lower */
    double s = Cm[j][j]; for (int k = 0; k < j; k++) s -= R[j][k] * R[j][k];
    R[j][j] = sqrt(s);
    for (int i = j + 1; i < n; i++) { s = Cm[i][j]; for (int k = 0; k < j; k++) s -= R[i][k] * R[j][k]; R[i][j] = s / R[j][j]; }
  }
  for (int j = 0; j < n; j++) {                       /* Ri = inv(R) */
    Ri[j][j] = 1 / R[j][j];
    for (int i = j + 1; i < n; i++) { double s = 0; for (int k = j; k < i; k++) s -= R[i][k] * Ri[k][j]; Ri[i][j] = s / R[i][i]; }
  }
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { double s = 0; for (int k = 0; k < n; k++) s += Ri[i][k] * B[k][j]; T[i][j] = s; }
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) { double s = 0; for (int k = 0; k < n; k++) s += T[i][k] * Ri[j][k]; M[i][j] = s; }
  for (int i = 0; i < n; i++) for (int j = i + 1; j < n; j++) M[i][j] = M[j][i] = 0.5 * (M[i][j] + M[j][i]);
  jacobi_eig(n, M, V, w);
  for (int i = 0; i < n; i++) { double s = 0; for (int k = 0; k < n; k++) s += Ri[k][i] * V[k][0]; y[i] = s; } /* y = R^-T v */
  return w[0];
}

static c_float lobpcg_core(const solver_sparse *Q, c_float *x, c_int *iters_out) {
  size_t n = Q->ncol, i;
  c_float *Ax = (c_float *)malloc(n * 8), *w = (c_float *)malloc(n * 8), *Aw = (c_float *)malloc(n * 8);
  c_float *p = (c_float *)malloc(n * 8), *Ap = (c_float *)malloc(n * 8);
  c_float lambda, norm_w, xAw, wAw, xAp, wAp, pAp, xp, wp, p_norm_inv;
  double B[3][3], Cm[3][3] = {{1,0,0},{0,1,0},{0,0,1}}, y[3];
  oracle_mat_vec(Q, x, Ax);
  lambda = vprod(x, Ax, n);
  vadd_scaled(Ax, x, w, -lambda, n);
  vadd_scaled(w, x, w, -vprod(x, w, n), n);
  vscale(w, 1.0 / vnorm2(w, n), n);
  oracle_mat_vec(Q, w, Aw);
  xAw = vprod(Aw, x, n); wAw = vprod(Aw, w, n);
  B[0][0] = lambda; B[0][1] = B[1][0] = xAw; B[1][1] = wAw;
  { double I2[3][3] = {{1,0,0},{0,1,0},{0,0,1}}; lambda = gen_eig_min(2, B, I2, y); }
  for (i = 0; i < n; i++) { p[i] = y[1] * w[i]; Ap[i] = y[1] * Aw[i]; }
  vadd_scaled(p, x, x, y[0], n); vadd_scaled(Ap, Ax, Ax, y[0], n);
  size_t it, max_iter = 1000;
  for (it = 0; it < max_iter; it++) {
    vadd_scaled(Ax, x, w, -lambda, n);
    if (vnorminf(w, n) < 1e-5) {
      norm_w = vnorm2(w, n);
      lambda -= sqrt(2) * norm_w + 1e-6;
      if (n <= 3) lambda -= 1e-6;
      break;
    }
    vadd_scaled(w, x, w, -vprod(x, w, n), n);
    vscale(w, 1.0 / vnorm2(w, n), n);
    oracle_mat_vec(Q, w, Aw);
    xAw = vprod(Ax, w, n); wAw = vprod(w, Aw, n);
    p_norm_inv = 1.0 / vnorm2(p, n);
    vscale(p, p_norm_inv, n); vscale(Ap, p_norm_inv, n);
    xAp = vprod(Ax, p, n); wAp = vprod(Aw, p, n); pAp = vprod(Ap, p, n); xp = vprod(x, p, n); wp = vprod(w, p, n);
    B[0][0] = lambda; B[0][1] = xAw; B[0][2] = xAp; B[1][0] = xAw; B[1][1] = wAw; B[1][2] = wAp;
    B[2][0] = xAp; B[2][1] = wAp; B[2][2] = pAp;
    Cm[0][0] = Cm[1][1] = Cm[2][2] = 1.0; Cm[0][1] = Cm[1][0] = 0.0;
    Cm[0][2] = Cm[2][0] = xp; Cm[1][2] = Cm[2][1] = wp;
    lambda = gen_eig_min(3, B, Cm, y);
    vmult_add_scaled(p, w, y[2], y[1], n); vmult_add_scaled(Ap, Aw, y[2], y[1], n);
    vmult_add_scaled(x, p, y[0], 1, n);    vmult_add_scaled(Ax, Ap, y[0], 1, n);
  }
  if (iters_out) *iters_out = (c_int)it;
  free(Ax); free(w); free(Aw); free(p); free(Ap);
  return lambda;
}

int oracle_lobpcg(const solver_sparse *Q, const c_float *x0, c_float *lambda_out, c_int *iters_out) {
  size_t n = Q->ncol;
  c_float *x = vcopy(x0, n);
  vscale(x, 1.0 / vnorm2(x, n), n);
  *lambda_out = lobpcg_core(Q, x, iters_out);
  free(x);
  return 0;
}

/* =============================================================================================
 * the solver proper
 * =========================================================================================== */
void oracle_qpalm_set_default_settings(QPALMSettings *s) { /* src/qpalm.c:39-70, include/constants.h:65-116 */
  s->max_iter = 10000; s->inner_max_iter = 100; s->eps_abs = 1e-4; s->eps_rel = 1e-4;
  s->eps_abs_in = 1; s->eps_rel_in = 1; s->rho = 0.1; s->eps_prim_inf = 1e-5; s->eps_dual_inf = 1e-5;
  s->theta = 0.25; s->delta = 100; s->sigma_max = 1e9; s->sigma_init = 2e1; s->proximal = TRUE;
  s->gamma_init = 1e7; s->gamma_upd = 10; s->gamma_max = 1e7; s->scaling = 10; s->nonconvex = FALSE;
  s->verbose = TRUE; s->print_iter = 1; s->warm_start = FALSE; s->reset_newton_iter = 10000;
  s->enable_dual_termination = FALSE; s->dual_objective_limit = QPALM_INFTY; s->time_limit = QPALM_INFTY;
  s->ordering = 0; s->factorization_method = FACTORIZE_KKT_OR_SCHUR; s->max_rank_update = 160;
  s->max_rank_update_fraction = 0.1;
}

static void set_status(QPALMInfo *info, c_int v) { /* src/util.c:61-99 */
  info->status_val = v;
  const char *s = "unrecognised status value";
  switch (v) {
    case QPALM_SOLVED: s = "solved"; break;
    case QPALM_DUAL_TERMINATED: s = "dual terminated"; break;
    case QPALM_PRIMAL_INFEASIBLE: s = "primal infeasible"; break;
    case QPALM_DUAL_INFEASIBLE: s = "dual infeasible"; break;
    case QPALM_TIME_LIMIT_REACHED: s = "time limit exceeded"; break;
    case QPALM_MAX_ITER_REACHED: s = "maximum iterations reached"; break;
    case QPALM_UNSOLVED: s = "unsolved"; break;
    case QPALM_ERROR: s = "error"; break;
  }
  strcpy(info->status, s);
}

static int settings_ok(const QPALMSettings *s) { /* src/validate.c:43-221 */
  if (!s) return 0;
  if (s->max_iter <= 0 || s->inner_max_iter <= 0) return 0;
  if (s->eps_abs < 0 || s->eps_rel < 0 || (s->eps_rel == 0 && s->eps_abs == 0)) return 0;
  if (s->eps_abs_in < 0 || s->eps_rel_in < 0 || (s->eps_rel_in == 0 && s->eps_abs_in == 0)) return 0;
  if (s->rho <= 0 || s->rho >= 1) return 0;
  if (s->eps_prim_inf < 0 || s->eps_dual_inf < 0) return 0;
  if (s->theta > 1 || s->delta <= 1 || s->sigma_max <= 0 || s->sigma_init <= 0) return 0;
  if ((s->proximal != 0) && (s->proximal != 1)) return 0;
  if (s->gamma_init <= 0 || s->gamma_upd < 1 || s->gamma_max < s->gamma_init) return 0;
  if (s->scaling < 0) return 0;
  if ((s->nonconvex != 0) && (s->nonconvex != 1)) return 0;
  if ((s->warm_start != 0) && (s->warm_start != 1)) return 0;
  if ((s->verbose != 0) && (s->verbose != 1)) return 0;
  if (s->print_iter <= 0 || s->reset_newton_iter <= 0) return 0;
  if ((s->enable_dual_termination != 0) && (s->enable_dual_termination != 1)) return 0;
  if (s->max_rank_update < 0 || s->max_rank_update_fraction < 0) return 0;
  return 1;
}

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
static void tic(QPALMWorkspace *w) { double t = now_s(); w->timer->tic_sec = (int64_t)t; w->timer->tic_nsec = (int64_t)((t - (int64_t)t) * 1e9); }
static double toc(QPALMWorkspace *w) { return now_s() - ((double)w->timer->tic_sec + 1e-9 * (double)w->timer->tic_nsec); }

static solver_sparse *csc_copy(const solver_sparse *S, int lower_only) {
  solver_sparse *C = (solver_sparse *)calloc(1, sizeof(solver_sparse));
  const c_int *Sp = (const c_int *)S->p, *Si = (const c_int *)S->i; const c_float *Sx = (const c_float *)S->x;
  c_int nnz = Sp[S->ncol];
  *C = *S;
  C->p = malloc((S->ncol + 1) * sizeof(c_int)); C->i = malloc((nnz + 1) * sizeof(c_int)); C->x = malloc((nnz + 1) * sizeof(c_float));
  c_int *Cp = (c_int *)C->p, *Ci = (c_int *)C->i; c_float *Cx = (c_float *)C->x; c_int k = 0;
  for (size_t j = 0; j < S->ncol; j++) {
    Cp[j] = k;
    for (c_int p = Sp[j]; p < Sp[j+1]; p++) { if (lower_only && Si[p] < (c_int)j) continue; Ci[k] = Si[p]; Cx[k] = Sx[p]; k++; }
  }
  Cp[S->ncol] = k; C->nzmax = (size_t)(k > 0 ? k : 1); C->nz = NULL; C->z = NULL;
  return C;
}
static void csc_free(solver_sparse *S) { if (S) { free(S->p); free(S->i); free(S->x); free(S); } }

static void initialize_sigma(QPALMWorkspace *work) { /* src/iteration.c:50-84 */
  size_t n = work->data->n, m = work->data->m;
  c_float f = 0.5 * vprod(work->x, work->Qx, n) + vprod(work->data->q, work->x, n);
  for (size_t i = 0; i < m; i++) { c_float mid = c_max(work->data->bmin[i], c_min(work->Ax[i], work->data->bmax[i])); work->temp_m[i] = work->Ax[i] + (-1) * mid; }
  c_float dist2 = vprod(work->temp_m, work->temp_m, m);
  vset(work->sigma, c_max(1e-4, c_min(work->settings->sigma_init * c_max(1, c_absval(f)) / c_max(1, 0.5 * dist2), 1e4)), m);
  vrecip(work->sigma, work->sigma_inv, m);
  for (size_t i = 0; i < m; i++) work->sqrt_sigma[i] = sqrt(work->sigma[i]);
  work->sqrt_sigma_max = sqrt(work->settings->sigma_max);
}

static c_float compute_objective(QPALMWorkspace *work) { /* src/iteration.c:231-270 */
  c_float obj = 0; size_t n = work->data->n, i = 0;
  const c_float *Qx = work->Qx, *x = work->x, *q = work->data->q; c_float g = work->gamma;
  if (work->settings->proximal) {
    if (n >= 4) for (; i <= n - 4; i += 4)
      obj += (0.5*(Qx[i] - 1/g*x[i]) + q[i])*x[i] + (0.5*(Qx[i+1] - 1/g*x[i+1]) + q[i+1])*x[i+1]
           + (0.5*(Qx[i+2] - 1/g*x[i+2]) + q[i+2])*x[i+2] + (0.5*(Qx[i+3] - 1/g*x[i+3]) + q[i+3])*x[i+3];
    for (; i < n; i++) obj += (0.5*(Qx[i] - 1/g*x[i]) + q[i])*x[i];
  } else {
    if (n >= 4) for (; i <= n - 4; i += 4)
      obj += (0.5*Qx[i] + q[i])*x[i] + (0.5*Qx[i+1] + q[i+1])*x[i+1] + (0.5*Qx[i+2] + q[i+2])*x[i+2] + (0.5*Qx[i+3] + q[i+3])*x[i+3];
    for (; i < n; i++) obj += (0.5*Qx[i] + q[i])*x[i];
  }
  if (work->settings->scaling) obj *= work->scaling->cinv;
  return obj + work->data->c;
}

static c_float compute_dual_objective(QPALMWorkspace *work) { /* src/iteration.c:272-299 */
  Aux *a = aux_of(work); size_t n = work->data->n, m = work->data->m; c_float dobj = 0;
  vadd_scaled(work->Aty, work->data->q, work->neg_dphi, 1.0, n);
  ldl_solve(n, a->Lq, a->Dq, work->neg_dphi, work->D_temp);
  dobj -= 0.5 * vprod(work->neg_dphi, work->D_temp, n);
  for (size_t i = 0; i < m; i++) dobj -= work->y[i] > 0 ? work->y[i] * work->data->bmax[i] : work->y[i] * work->data->bmin[i];
  if (work->settings->scaling) dobj *= work->scaling->cinv;
  return dobj + work->data->c;
}

QPALMWorkspace *oracle_qpalm_setup(const QPALMData *data, const QPALMSettings *settings) { /* src/qpalm.c:73-319 */
  if (!data) return NULL;
  for (size_t j = 0; j < data->m; j++) if (data->bmin[j] > data->bmax[j]) return NULL;     /* validate.c:18-40 */
  if (!settings_ok(settings)) return NULL;
  QPALMWorkspace *work = (QPALMWorkspace *)calloc(1, sizeof(QPALMWorkspace));
  work->timer = (QPALMTimer *)calloc(1, sizeof(QPALMTimer)); tic(work);
  work->settings = (QPALMSettings *)malloc(sizeof(QPALMSettings)); *work->settings = *settings;
  work->sqrt_delta = sqrt(settings->delta); work->gamma = settings->gamma_init;
  size_t n = data->n, m = data->m;
  work->solver = (QPALMSolver *)calloc(1, sizeof(QPALMSolver));
  Aux *a = (Aux *)calloc(1, sizeof(Aux)); work->solver->LD = a; a->n = n; a->m = m;
  work->data = (QPALMData *)calloc(1, sizeof(QPALMData));
  work->data->n = n; work->data->m = m; work->data->c = data->c;
  work->data->bmin = vcopy(data->bmin, m); work->data->bmax = vcopy(data->bmax, m); work->data->q = vcopy(data->q, n);
  work->data->A = csc_copy(data->A, 0); work->data->A->stype = 0;
  work->data->Q = csc_copy(data->Q, 1); work->data->Q->stype = -1;
  a->Ap = (c_int *)work->data->A->p; a->Ai = (c_int *)work->data->A->i; a->Ax = (c_float *)work->data->A->x;
  a->Qp = (c_int *)work->data->Q->p; a->Qi = (c_int *)work->data->Q->i; a->Qx = (c_float *)work->data->Q->x;
  a->L = (c_float *)calloc(n * n + 1, 8); a->H = (c_float *)calloc(n * n + 1, 8); a->D = (c_float *)calloc(n + 1, 8); a->w = (c_float *)calloc(n + 1, 8);
  a->trace = (OracleTraceEntry *)calloc(TRACE_MAX, sizeof(OracleTraceEntry));
#define VN(f) work->f = (c_float *)calloc(n + 1, sizeof(c_float))
#define VM(f) work->f = (c_float *)calloc(m + 1, sizeof(c_float))
#define V2M(f) work->f = (c_float *)calloc(2 * m + 1, sizeof(c_float))
  VN(x); VM(y); VM(Ax); VN(Qx); VN(x_prev); VN(Aty); VN(x0); VM(temp_m); VN(temp_n); VM(sigma); VM(sigma_inv);
  VM(z); VM(Axys); VM(pri_res); VM(pri_res_in); VN(df); VN(xx0); VN(dphi); VN(dphi_prev); VM(sqrt_sigma);
  V2M(delta); V2M(alpha); V2M(delta2); V2M(delta_alpha); V2M(temp_2m);
  work->s = (array_element *)calloc(2 * m + 1, sizeof(array_element));
  work->index_L = (c_int *)calloc(2 * m + 1, sizeof(c_int)); work->index_P = (c_int *)calloc(2 * m + 1, sizeof(c_int)); work->index_J = (c_int *)calloc(2 * m + 1, sizeof(c_int));
  VM(delta_y); VN(Atdelta_y); VN(delta_x); VN(Qdelta_x); VM(Adelta_x);
  VN(neg_dphi); VN(d); VN(Qd); VM(Ad); VM(yh); VN(Atyh); VN(D_temp); VM(E_temp);
  work->initialized = FALSE;
  work->solver->factorization_method = FACTORIZE_SCHUR;                                     /* solver_interface.c:72-73 */
  if (settings->scaling) {
    work->scaling = (QPALMScaling *)calloc(1, sizeof(QPALMScaling));
    work->scaling->D = (c_float *)calloc(n + 1, 8); work->scaling->Dinv = (c_float *)calloc(n + 1, 8);
    work->scaling->E = (c_float *)calloc(m + 1, 8); work->scaling->Einv = (c_float *)calloc(m + 1, 8);
    scale_data_impl(work->data->A, work->data->Q, work->data->q, work->data->bmin, work->data->bmax,
                    settings->scaling, work->scaling->D, work->scaling->E, &work->scaling->c, work->Qx);
    pcopy(work->scaling->D, work->D_temp, n);
    vrecip(work->scaling->D, work->scaling->Dinv, n); vrecip(work->scaling->E, work->scaling->Einv, m);
    work->scaling->cinv = 1.0 / work->scaling->c;
  }
  work->solver->active_constraints = (c_int *)calloc(m + 1, sizeof(c_int));
  work->solver->active_constraints_old = (c_int *)calloc(m + 1, sizeof(c_int));
  work->solver->reset_newton = TRUE;
  work->solver->enter = (c_int *)calloc(m + 1, sizeof(c_int)); work->solver->leave = (c_int *)calloc(m + 1, sizeof(c_int));
  if (work->settings->nonconvex) {                                                          /* nonconvex.c:171-183 */
    for (size_t i = 0; i < n; i++) work->d[i] = (c_float)rand() / RAND_MAX;                 /* nonconvex.c:41-44   */
    vscale(work->d, 1.0 / vnorm2(work->d, n), n);
    c_float lambda = lobpcg_core(work->data->Q, work->d, NULL);
    if (lambda < 0) { work->settings->proximal = TRUE; work->settings->gamma_init = 1 / c_absval(lambda); work->settings->gamma_max = work->settings->gamma_init; work->gamma_maxed = TRUE; }
    else work->settings->nonconvex = FALSE;
  }
  work->solution = (QPALMSolution *)calloc(1, sizeof(QPALMSolution));
  work->solution->x = (c_float *)calloc(n + 1, 8); work->solution->y = (c_float *)calloc(m + 1, 8);
  work->info = (QPALMInfo *)calloc(1, sizeof(QPALMInfo));
  set_status(work->info, QPALM_UNSOLVED);
  work->info->setup_time = toc(work);
  return work;
}

void oracle_qpalm_warm_start(QPALMWorkspace *work, c_float *xw, c_float *yw) { /* src/qpalm.c:322-399 */
  work->gamma = work->settings->gamma_init;
  if (work->info->status_val != QPALM_UNSOLVED) work->info->setup_time = 0;
  tic(work);
  size_t n = work->data->n, m = work->data->m;
  if (xw != NULL) {
    pcopy(xw, work->x, n);
    if (work->settings->scaling) vewprod(work->x, work->scaling->Dinv, work->x, n);
    pcopy(work->x, work->x0, n); pcopy(work->x, work->x_prev, n); pcopy(work->x, work->neg_dphi, n);
    oracle_mat_vec(work->data->Q, work->neg_dphi, work->Qd);
    if (work->settings->proximal) vadd_scaled(work->Qd, work->x, work->Qx, 1 / work->settings->gamma_init, n);
    else pcopy(work->Qd, work->Qx, n);
    oracle_mat_vec(work->data->A, work->neg_dphi, work->Ad);
    pcopy(work->Ad, work->Ax, m);
    work->info->objective = compute_objective(work);
  } else {
    vset(work->x, 0., n); vset(work->x_prev, 0., n); vset(work->x0, 0., n); vset(work->Qx, 0., n); vset(work->Ax, 0., m);
    work->info->objective = 0.0;
  }
  if (yw != NULL) {
    pcopy(yw, work->y, m);
    if (work->settings->scaling) { vewprod(work->y, work->scaling->Einv, work->y, m); vscale(work->y, work->scaling->c, m); }
  } else vset(work->y, 0., m);
  initialize_sigma(work);
  work->initialized = TRUE;
  work->info->setup_time += toc(work);
}

/* ---- Newton direction: src/newton.c:17-120 (SCHUR branch) -------------------------------- */
static void set_active(QPALMWorkspace *work) { /* newton.c:122-149 */
  size_t m = work->data->m; QPALMSolver *s = work->solver; c_int na = 0, ne = 0, nl = 0;
  for (size_t i = 0; i < m; i++) { s->active_constraints[i] = ((work->Axys[i] <= work->data->bmin[i]) || (work->Axys[i] >= work->data->bmax[i])); na += s->active_constraints[i]; }
  for (size_t i = 0; i < m; i++) {
    if (s->active_constraints[i] && !s->active_constraints_old[i]) s->enter[ne++] = (c_int)i;
    if (!s->active_constraints[i] && s->active_constraints_old[i]) s->leave[nl++] = (c_int)i;
  }
  s->nb_active_constraints = na; s->nb_enter = ne; s->nb_leave = nl;
}
static void factor_H(QPALMWorkspace *work, int with_constraints) { /* ldlcholQAtsigmaA / ldlchol(Q): solver_interface.c:319-405 */
  Aux *a = aux_of(work);
  c_float beta = work->settings->proximal ? 1.0 / work->gamma : 0.0;
  assemble_H(a->n, a->m, a->Qp, a->Qi, a->Qx, a->Ap, a->Ai, a->Ax, work->sigma,
             with_constraints ? work->solver->active_constraints : NULL, beta, a->H);
  ldl_factor(a->n, a->H, a->L, a->D);
  a->last_refactor = 1;
}
static void column_of_At(QPALMWorkspace *work, c_int row, c_float scale, c_float *w) { /* column `row` of At_sqrt_sigma, dense */
  Aux *a = aux_of(work);
  for (size_t j = 0; j < a->n; j++) {
    w[j] = 0.0;
    for (c_int p = a->Ap[j]; p < a->Ap[j+1]; p++) if (a->Ai[p] == row) { w[j] = a->Ax[p] * scale; break; }
  }
}
static void newton_set_direction(QPALMWorkspace *work) {
  Aux *a = aux_of(work); QPALMSolver *s = work->solver; size_t n = a->n, m = a->m;
  set_active(work);
  a->last_refactor = 0;
  if ((s->reset_newton && s->nb_active_constraints) ||
      (s->nb_enter + s->nb_leave) > c_min(work->settings->max_rank_update_fraction * (n + m), work->settings->max_rank_update)) {
    factor_H(work, 1);
  } else if (s->nb_active_constraints) {
    for (c_int k = 0; k < s->nb_enter; k++) { column_of_At(work, s->enter[k], work->sqrt_sigma[s->enter[k]], a->w); ldl_updown1(n, a->L, a->D, a->w, 1); }
    for (c_int k = 0; k < s->nb_leave; k++) { column_of_At(work, s->leave[k], work->sqrt_sigma[s->leave[k]], a->w); ldl_updown1(n, a->L, a->D, a->w, 0); }
  } else {
    factor_H(work, 0);
  }
  for (size_t i = 0; i < n; i++) work->neg_dphi[i] = work->dphi[i] * -1;                    /* solver_interface.c:505-519 */
  ldl_solve(n, a->L, a->D, work->neg_dphi, work->d);
  for (size_t i = 0; i < m; i++) s->active_constraints_old[i] = s->active_constraints[i];
  s->reset_newton = FALSE;
}

static void update_sigma(QPALMWorkspace *work) { /* src/iteration.c:86-145 */
  Aux *a = aux_of(work); size_t m = a->m, n = a->n; QPALMSettings *st = work->settings;
  work->nb_sigma_changed = 0;
  c_float *At_scale = work->E_temp;   /* scratch m-vector standing in for solver->At_scale */
  c_float nrm = vnorminf(work->pri_res, m), sigma_temp, mult;
  c_int *changed = work->solver->enter;
  for (size_t k = 0; k < m; k++) {
    if ((c_absval(work->pri_res[k]) > st->theta * c_absval(work->pri_res_in[k])) && work->solver->active_constraints[k]) {
      mult = c_max(1.0, st->delta * c_absval(work->pri_res[k]) / (nrm + 1e-6));
      sigma_temp = mult * work->sigma[k];
      if (sigma_temp <= st->sigma_max) {
        if (work->sigma[k] != sigma_temp) changed[work->nb_sigma_changed++] = (c_int)k;
        work->sigma[k] = sigma_temp; work->sigma_inv[k] = 1.0 / sigma_temp;
        mult = sqrt(mult); work->sqrt_sigma[k] = mult * work->sqrt_sigma[k]; At_scale[k] = mult;
      } else {
        if (work->sigma[k] != st->sigma_max) changed[work->nb_sigma_changed++] = (c_int)k;
        work->sigma[k] = st->sigma_max; work->sigma_inv[k] = 1.0 / st->sigma_max;
        At_scale[k] = work->sqrt_sigma_max / work->sqrt_sigma[k]; work->sqrt_sigma[k] = work->sqrt_sigma_max;
      }
    } else At_scale[k] = 1.0;
  }
  /* first_factorization is never set in the CHOLMOD build */
  if ((st->proximal && work->gamma < st->gamma_max) ||
      (work->nb_sigma_changed > c_min(st->max_rank_update_fraction * (n + m), 0.25 * st->max_rank_update))) {
    work->solver->reset_newton = TRUE;
  } else if (work->nb_sigma_changed == 0) {
  } else {                                                                                 /* ldlupdate_sigma_changed, solver_interface.c:443-503 */
    for (c_int k = 0; k < work->nb_sigma_changed; k++) {
      c_int row = changed[k]; c_float f = At_scale[row]; f = f * f; f = sqrt(1 - 1 / f);
      column_of_At(work, row, work->sqrt_sigma[row] * f, a->w);
      ldl_updown1(n, a->L, a->D, a->w, 1);
    }
  }
}
static void update_gamma(QPALMWorkspace *work) { /* src/iteration.c:147-157 */
  if (work->gamma < work->settings->gamma_max) {
    c_float prev = work->gamma;
    work->gamma = c_min(work->gamma * work->settings->gamma_upd, work->settings->gamma_max);
    work->solver->reset_newton = TRUE;
    vadd_scaled(work->Qx, work->x, work->Qx, 1 / work->gamma - 1 / prev, work->data->n);
  }
}
static void boost_gamma(QPALMWorkspace *work) { /* src/iteration.c:159-211; gershgorin_max: nonconvex.c:185-210 */
  Aux *a = aux_of(work); size_t n = a->n, m = a->m; c_float prev = work->gamma;
  if (work->solver->nb_active_constraints) {
    c_int na = 0; for (size_t i = 0; i < m; i++) if (work->solver->active_constraints[i]) work->solver->enter[na++] = (c_int)i;
    assemble_H(n, m, a->Qp, a->Qi, a->Qx, a->Ap, a->Ai, a->Ax, work->sigma, work->solver->active_constraints, 0.0, a->H);
    /* subtract Q again: the reference takes Gershgorin of A_J' Sigma A_J alone (cholmod_aat result, both triangles) */
    for (size_t j = 0; j < n; j++) for (c_int p = a->Qp[j]; p < a->Qp[j+1]; p++) if (a->Qi[p] >= (c_int)j) Lij(a->H, n, a->Qi[p], j) -= a->Qx[p];
    c_float ub = 0;
    for (size_t i = 0; i < n; i++) {
      c_float center = Lij(a->H, n, i, i), radius = 0;
      for (size_t j = 0; j < n; j++) if (j != i) radius += c_absval(j < i ? Lij(a->H, n, i, j) : Lij(a->H, n, j, i));
      ub = (i == 0) ? center + radius : c_max(ub, center + radius);
    }
    work->gamma = c_max(work->settings->gamma_max, 1e14 / ub);
    work->gamma_maxed = TRUE;
  } else work->gamma = 1e12;
  if (prev != work->gamma) {
    vadd_scaled(work->Qx, work->x, work->Qx, 1.0 / work->gamma - 1.0 / prev, n);
    vadd_scaled(work->Qd, work->d, work->Qd, work->tau / work->gamma - work->tau / prev, n);
    work->solver->reset_newton = TRUE;
  }
}

static void compute_residuals(QPALMWorkspace *work) { /* src/iteration.c:24-48 */
  size_t n = work->data->n, m = work->data->m;
  vewprod(work->y, work->sigma_inv, work->temp_m, m);
  vadd_scaled(work->Ax, work->temp_m, work->Axys, 1, m);
  for (size_t i = 0; i < m; i++) work->z[i] = c_max(work->data->bmin[i], c_min(work->Axys[i], work->data->bmax[i]));
  vadd_scaled(work->Ax, work->z, work->pri_res, -1, m);
  vewprod(work->pri_res, work->sigma, work->temp_m, m);
  vadd_scaled(work->y, work->temp_m, work->yh, 1, m);
  vadd_scaled(work->Qx, work->data->q, work->df, 1, n);
  if (work->settings->proximal) vadd_scaled(work->df, work->x0, work->df, -1 / work->gamma, n);
  oracle_mat_tpose_vec(work->data->A, work->yh, work->Atyh);
  vadd_scaled(work->df, work->Atyh, work->dphi, 1, n);
}

/* ---- termination: src/termination.c ------------------------------------------------------- */
static void residuals_and_tolerances(QPALMWorkspace *work) { /* termination.c:44-128 */
  size_t n = work->data->n, m = work->data->m; QPALMSettings *st = work->settings; QPALMScaling *sc = work->scaling;
  if (st->scaling) { vewprod(sc->Einv, work->pri_res, work->temp_m, m); work->info->pri_res_norm = vnorminf(work->temp_m, m); }
  else work->info->pri_res_norm = vnorminf(work->pri_res, m);
  if (st->scaling) {
    if (st->proximal) {
      vadd_scaled(work->x, work->x0, work->xx0, -1, n);
      vadd_scaled(work->dphi, work->xx0, work->temp_n, -1 / work->gamma, n);
      vewprod(sc->Dinv, work->temp_n, work->temp_n, n); work->info->dua_res_norm = vnorminf(work->temp_n, n);
      vewprod(sc->Dinv, work->dphi, work->temp_n, n); work->info->dua2_res_norm = vnorminf(work->temp_n, n);
    } else {
      vewprod(sc->Dinv, work->dphi, work->temp_n, n); work->info->dua_res_norm = vnorminf(work->temp_n, n);
      work->info->dua2_res_norm = work->info->dua_res_norm;
    }
    work->info->dua_res_norm *= sc->cinv; work->info->dua2_res_norm *= sc->cinv;
  } else {
    if (st->proximal) {
      vadd_scaled(work->x, work->x0, work->xx0, -1, n);
      vadd_scaled(work->dphi, work->xx0, work->temp_n, -1 / work->gamma, n);
      work->info->dua_res_norm = vnorminf(work->temp_n, n); work->info->dua2_res_norm = vnorminf(work->dphi, n);
    } else { work->info->dua_res_norm = vnorminf(work->dphi, n); work->info->dua2_res_norm = work->info->dua_res_norm; }
  }
  if (st->scaling) {
    vewprod(sc->Einv, work->Ax, work->temp_2m, m); vewprod(sc->Einv, work->z, work->temp_2m + m, m);
    work->eps_pri = st->eps_abs + st->eps_rel * vnorminf(work->temp_2m, m);   /* sic: length m (termination.c:99) */
  } else work->eps_pri = st->eps_abs + st->eps_rel * c_max(vnorminf(work->Ax, m), vnorminf(work->z, m));
  c_float nQx, nq, nAtyh, mx;
  if (st->scaling) {
    vewprod(sc->Dinv, work->Qx, work->temp_n, n); nQx = vnorminf(work->temp_n, n);
    vewprod(sc->Dinv, work->data->q, work->temp_n, n); nq = vnorminf(work->temp_n, n);
    vewprod(sc->Dinv, work->Atyh, work->temp_n, n); nAtyh = vnorminf(work->temp_n, n);
  } else { nQx = vnorminf(work->Qx, n); nq = vnorminf(work->data->q, n); nAtyh = vnorminf(work->Atyh, n); }
  mx = c_max(nQx, c_max(nq, nAtyh));
  if (st->scaling) mx *= sc->cinv;
  work->eps_dua = st->eps_abs + st->eps_rel * mx;
  work->eps_dua_in = work->eps_abs_in + work->eps_rel_in * mx;
}
static int is_primal_infeasible(QPALMWorkspace *work) { /* termination.c:136-182 */
  size_t n = work->data->n, m = work->data->m; QPALMSettings *st = work->settings; QPALMScaling *sc = work->scaling;
  c_float eps;
  vadd_scaled(work->yh, work->y, work->delta_y, -1, m);
  if (st->scaling) { vewprod(sc->E, work->delta_y, work->temp_m, m); eps = st->eps_prim_inf * vnorminf(work->temp_m, m); }
  else eps = st->eps_prim_inf * vnorminf(work->delta_y, m);
  if (eps == 0) return 0;
  vadd_scaled(work->Atyh, work->Aty, work->Atdelta_y, -1, n);
  if (st->scaling) vewprod(sc->Dinv, work->Atdelta_y, work->Atdelta_y, n);
  c_float oob = 0;
  for (size_t i = 0; i < m; i++) {
    c_float Ei = st->scaling ? sc->E[i] : 1.0;
    if (st->scaling) {
      oob += (work->data->bmax[i] < Ei * QPALM_INFTY) ? work->data->bmax[i] * c_max(work->delta_y[i], 0) : 0;
      oob += (work->data->bmin[i] > -Ei * QPALM_INFTY) ? work->data->bmin[i] * c_min(work->delta_y[i], 0) : 0;
    } else {
      oob += (work->data->bmax[i] < QPALM_INFTY) ? work->data->bmax[i] * c_max(work->delta_y[i], 0) : 0;
      oob += (work->data->bmin[i] > -QPALM_INFTY) ? work->data->bmin[i] * c_min(work->delta_y[i], 0) : 0;
    }
  }
  return (vnorminf(work->Atdelta_y, n) <= eps) && (oob <= -eps);
}
static int is_dual_infeasible(QPALMWorkspace *work) { /* termination.c:184-240 */
  size_t n = work->data->n, m = work->data->m; QPALMSettings *st = work->settings; QPALMScaling *sc = work->scaling;
  c_float eps, dxQdx, dxdx;
  vadd_scaled(work->x, work->x_prev, work->delta_x, -1, n);
  if (st->scaling) { vewprod(sc->D, work->delta_x, work->temp_n, n); eps = st->eps_dual_inf * vnorminf(work->temp_n, n); dxdx = vprod(work->temp_n, work->temp_n, n); }
  else { eps = st->eps_dual_inf * vnorminf(work->delta_x, n); dxdx = vprod(work->delta_x, work->delta_x, n); }
  if (eps == 0) return 0;
  if (st->scaling) {
    vewprod(sc->Einv, work->Ad, work->Adelta_x, m);
    for (size_t k = 0; k < m; k++)
      if ((work->data->bmax[k] < sc->E[k] * QPALM_INFTY && work->Adelta_x[k] >= eps) || (work->data->bmin[k] > -sc->E[k] * QPALM_INFTY && work->Adelta_x[k] <= -eps)) return 0;
  } else {
    for (size_t k = 0; k < m; k++)
      if ((work->data->bmax[k] < QPALM_INFTY && work->Ad[k] >= eps) || (work->data->bmin[k] > -QPALM_INFTY && work->Ad[k] <= -eps)) return 0;
  }
  if (st->proximal) { vadd_scaled(work->Qd, work->d, work->temp_n, -work->tau / work->gamma, n); dxQdx = vprod(work->delta_x, work->temp_n, n); }
  else dxQdx = vprod(work->Qd, work->delta_x, n);
  if (st->scaling)
    return (dxQdx <= -sc->c * st->eps_dual_inf * st->eps_dual_inf * dxdx) ||
           ((dxQdx <= sc->c * st->eps_dual_inf * st->eps_dual_inf * dxdx) && (vprod(work->data->q, work->delta_x, n) <= -sc->c * eps));
  return (dxQdx <= -st->eps_dual_inf * st->eps_dual_inf * dxdx) ||
         ((dxQdx <= st->eps_dual_inf * st->eps_dual_inf * dxdx) && (vprod(work->data->q, work->delta_x, n) <= -eps));
}
static void store_solution(QPALMWorkspace *work) { /* termination.c:242-252 */
  size_t n = work->data->n, m = work->data->m;
  if (work->settings->scaling) {
    vewprod(work->x, work->scaling->D, work->solution->x, n);
    vscale(work->yh, work->scaling->cinv, m);
    vewprod(work->yh, work->scaling->E, work->solution->y, m);
  } else { pcopy(work->x, work->solution->x, n); pcopy(work->yh, work->solution->y, m); }
  work->info->objective = compute_objective(work);
}
static int check_termination(QPALMWorkspace *work) { /* termination.c:19-42 */
  residuals_and_tolerances(work);
  if ((work->info->pri_res_norm < work->eps_pri) && (work->info->dua_res_norm < work->eps_dua)) {
    set_status(work->info, QPALM_SOLVED); store_solution(work); return 1;
  } else if (is_primal_infeasible(work)) {
    set_status(work->info, QPALM_PRIMAL_INFEASIBLE);
    if (work->settings->scaling) { vscale(work->delta_y, work->scaling->cinv, work->data->m); vewprod(work->scaling->E, work->delta_y, work->delta_y, work->data->m); }
    return 1;
  } else if (is_dual_infeasible(work)) {
    set_status(work->info, QPALM_DUAL_INFEASIBLE);
    if (work->settings->scaling) vewprod(work->scaling->D, work->delta_x, work->delta_x, work->data->n);
    return 1;
  }
  return 0;
}

static c_float exact_linesearch(QPALMWorkspace *work) { /* src/linesearch.c:14-120 */
  size_t n = work->data->n, m = work->data->m;
  oracle_mat_vec(work->data->Q, work->d, work->Qd);
  if (work->settings->proximal) vadd_scaled(work->Qd, work->d, work->Qd, 1 / work->gamma, n);
  oracle_mat_vec(work->data->A, work->d, work->Ad);
  work->eta = vprod(work->d, work->Qd, n);
  work->beta = vprod(work->d, work->df, n);
  c_float tau;
  oracle_linesearch((c_int)m, work->eta, work->beta, work->Ad, work->Ax, work->y, work->sigma, work->sqrt_sigma,
                    work->data->bmin, work->data->bmax, &tau, NULL, NULL, NULL);
  return tau;
}
static void update_primal_iterate(QPALMWorkspace *work) { /* src/iteration.c:213-229 */
  size_t n = work->data->n, m = work->data->m;
  newton_set_direction(work);
  work->tau = exact_linesearch(work);
  pcopy(work->x, work->x_prev, n); pcopy(work->dphi, work->dphi_prev, n);
  vadd_scaled(work->x, work->d, work->x, work->tau, n);
  vscale(work->Qd, work->tau, n); vscale(work->Ad, work->tau, m);
  vadd_scaled(work->Qx, work->Qd, work->Qx, 1, n); vadd_scaled(work->Ax, work->Ad, work->Ax, 1, m);
}

static void trace_add(QPALMWorkspace *work, c_int iter, c_int kind) {
  Aux *a = aux_of(work);
  if (a->ntrace >= TRACE_MAX) return;
  OracleTraceEntry *e = &a->trace[a->ntrace++];
  e->iter = iter; e->kind = kind; e->nb_active = work->solver->nb_active_constraints; e->nb_enter = work->solver->nb_enter;
  e->nb_leave = work->solver->nb_leave; e->refactor = a->last_refactor; e->tau = work->tau;
  e->pri_res_norm = work->info->pri_res_norm; e->dua_res_norm = work->info->dua_res_norm; e->gamma = work->gamma;
}
c_int oracle_trace(const QPALMWorkspace *work, OracleTraceEntry *out, c_int max_entries) {
  Aux *a = aux_of(work); c_int k = a->ntrace < max_entries ? a->ntrace : max_entries;
  memcpy(out, a->trace, (size_t)k * sizeof(OracleTraceEntry)); return a->ntrace;
}

static void finish(QPALMWorkspace *work, c_int iter, c_int iter_out) {
  work->info->iter = iter; work->info->iter_out = iter_out;
  work->info->solve_time = toc(work); work->info->run_time = work->info->setup_time + work->info->solve_time;
  work->initialized = FALSE;
}

void oracle_qpalm_solve(QPALMWorkspace *work) { /* src/qpalm.c:401-736 */
  Aux *a = aux_of(work); QPALMSettings *st = work->settings;
  work->eps_abs_in = st->eps_abs_in; work->eps_rel_in = st->eps_rel_in;
  work->solver->reset_newton = TRUE; work->gamma = st->gamma_init;
  work->gamma_maxed = (FALSE || st->nonconvex);
  size_t n = work->data->n, m = work->data->m;
  for (size_t i = 0; i < m; i++) work->solver->active_constraints_old[i] = FALSE;
  if (!work->initialized) oracle_qpalm_warm_start(work, NULL, NULL);
  tic(work);
  a->ntrace = 0;
  if (st->enable_dual_termination) {                                                       /* qpalm.c:459-472 */
    if (!a->Lq) { a->Lq = (c_float *)calloc(n * n + 1, 8); a->Dq = (c_float *)calloc(n + 1, 8); }
    assemble_H(n, 0, a->Qp, a->Qi, a->Qx, NULL, NULL, NULL, NULL, NULL, 0.0, a->H);
    ldl_factor(n, a->H, a->Lq, a->Dq);
    work->info->dual_objective = compute_dual_objective(work);
  } else work->info->dual_objective = QPALM_NULL;
  c_int iter, iter_out = 0, prev_iter = 0, no_change = 0;
  c_float eps_k_abs = st->eps_abs_in, eps_k_rel = st->eps_rel_in, eps_k;
  for (iter = 0; iter < st->max_iter; iter++) {
    compute_residuals(work);
    if (check_termination(work)) { trace_add(work, iter, 3); finish(work, iter, iter_out); return; }
    else if ((work->info->dua2_res_norm <= work->eps_dua_in) || (no_change == 3)) {        /* qpalm.c:515 */
      no_change = 0;
      if (iter_out > 0 && work->info->pri_res_norm > work->eps_pri) update_sigma(work);
      pcopy(work->yh, work->y, m); pcopy(work->Atyh, work->Aty, n);
      if (st->enable_dual_termination) {
        work->info->dual_objective = compute_dual_objective(work);
        if (work->info->dual_objective > st->dual_objective_limit) {
          set_status(work->info, QPALM_DUAL_TERMINATED); store_solution(work); finish(work, iter, iter_out); return;
        }
      }
      work->eps_abs_in = c_max(st->eps_abs, st->rho * work->eps_abs_in);
      work->eps_rel_in = c_max(st->eps_rel, st->rho * work->eps_rel_in);
      if (st->nonconvex) {                                                                 /* qpalm.c:586-609 */
        if (st->scaling) {
          vewprod(work->scaling->Einv, work->Ax, work->temp_2m, m); vewprod(work->scaling->Einv, work->z, work->temp_2m + m, m);
          eps_k = eps_k_abs + eps_k_rel * vnorminf(work->temp_2m, m);
        } else eps_k = eps_k_abs + eps_k_rel * c_max(vnorminf(work->Ax, m), vnorminf(work->z, m));
        if (work->info->pri_res_norm < eps_k) {
          pcopy(work->x, work->x0, n);
          eps_k_abs = c_max(st->eps_abs, st->rho * eps_k_abs); eps_k_rel = c_max(st->eps_rel, st->rho * eps_k_rel);
        }
      } else if (st->proximal) {                                                           /* qpalm.c:612-630 */
        if (!work->gamma_maxed && iter_out > 0 && work->solver->nb_enter == 0 && work->solver->nb_leave == 0 && work->info->pri_res_norm < work->eps_pri) {
          for (size_t i = 0; i < m; i++) { work->temp_m[i] = work->y[i] / work->sigma[i]; work->Axys[i] = work->Ax[i] + 1 * work->temp_m[i]; }
          set_active(work);
          if (work->solver->nb_enter == 0 && work->solver->nb_leave == 0) boost_gamma(work); else update_gamma(work);
        } else update_gamma(work);
        pcopy(work->x, work->x0, n);
      }
      pcopy(work->pri_res, work->pri_res_in, m);
      trace_add(work, iter, 1);
      iter_out++; prev_iter = iter;
    } else if (iter == prev_iter + st->inner_max_iter) {                                    /* qpalm.c:647-660 */
      no_change = 0;
      if (iter_out > 0 && work->info->pri_res_norm > work->eps_pri) update_sigma(work);
      if (st->proximal) { update_gamma(work); if (!st->nonconvex) pcopy(work->x, work->x0, n); }
      pcopy(work->pri_res, work->pri_res_in, m);
      trace_add(work, iter, 2);
      iter_out++; prev_iter = iter;
    } else {                                                                               /* qpalm.c:662-678 */
      if (work->solver->nb_enter + work->solver->nb_leave) no_change = 0; else no_change++;
      if ((iter % st->reset_newton_iter) == 0) work->solver->reset_newton = TRUE;
      update_primal_iterate(work);
      trace_add(work, iter, 0);
    }
    c_float now = work->info->setup_time + toc(work);                                       /* qpalm.c:680-708 */
    if (now > st->time_limit) {
      set_status(work->info, QPALM_TIME_LIMIT_REACHED); store_solution(work); finish(work, iter, iter_out); return;
    }
  }
  set_status(work->info, QPALM_MAX_ITER_REACHED); store_solution(work); finish(work, iter, iter_out);
}

void oracle_qpalm_update_settings(QPALMWorkspace *work, const QPALMSettings *settings) { /* src/qpalm.c:739-791 */
  if (!settings_ok(settings)) { set_status(work->info, QPALM_ERROR); return; }
  if (work->settings->scaling > settings->scaling) { set_status(work->info, QPALM_ERROR); return; }
  else if (work->settings->scaling < settings->scaling) {
    size_t n = work->data->n, m = work->data->m;
    pcopy(work->scaling->D, work->temp_n, n); pcopy(work->scaling->E, work->temp_m, m);
    c_float c_temp = work->scaling->c;
    c_float *D = work->scaling->D, *E = work->scaling->E;
    scale_data_impl(work->data->A, work->data->Q, work->data->q, work->data->bmin, work->data->bmax,
                    settings->scaling - work->settings->scaling, D, E, &work->scaling->c, work->Qx);
    pcopy(D, work->D_temp, n);
    vewprod(D, work->temp_n, D, n); vewprod(E, work->temp_m, E, m);
    work->scaling->c *= c_temp;
    vrecip(D, work->scaling->Dinv, n); vrecip(E, work->scaling->Einv, m);
    work->scaling->cinv = 1 / work->scaling->c;
  }
  *work->settings = *settings;
  work->sqrt_delta = sqrt(work->settings->delta);
}

void oracle_qpalm_update_bounds(QPALMWorkspace *work, const c_float *bmin, const c_float *bmax) { /* src/qpalm.c:793-825 */
  size_t m = work->data->m;
  if (bmin != NULL && bmax != NULL) for (size_t j = 0; j < m; j++) if (bmin[j] > bmax[j]) { set_status(work->info, QPALM_ERROR); return; }
  if (bmin != NULL) pcopy(bmin, work->data->bmin, m);
  if (bmax != NULL) pcopy(bmax, work->data->bmax, m);
  if (work->settings->scaling) {
    if (bmin != NULL) vewprod(work->scaling->E, work->data->bmin, work->data->bmin, m);
    if (bmax != NULL) vewprod(work->scaling->E, work->data->bmax, work->data->bmax, m);
  }
}

void oracle_qpalm_update_q(QPALMWorkspace *work, const c_float *q) { /* src/qpalm.c:827-871 */
  size_t n = work->data->n;
  pcopy(q, work->data->q, n);
  if (work->settings->scaling) {
    vewprod(work->scaling->D, work->data->q, work->data->q, n);
    c_float c_old = work->scaling->c, c_ratio;
    if (work->settings->proximal) vadd_scaled(work->Qx, work->x, work->Qx, -1 / work->gamma, n);
    vadd_scaled(work->data->q, work->Qx, work->temp_n, work->scaling->cinv, n);
    work->scaling->c = 1 / c_max(1.0, vnorminf(work->temp_n, n));
    work->scaling->cinv = 1 / work->scaling->c;
    vscale(work->data->q, work->scaling->c, n);
    c_ratio = work->scaling->c / c_old;
    c_float *Qx = (c_float *)work->data->Q->x; c_int nnz = ((c_int *)work->data->Q->p)[n];
    for (c_int p = 0; p < nnz; p++) Qx[p] *= c_ratio;
    vscale(work->Qx, c_ratio, n);
    if (work->settings->proximal) { work->gamma = work->settings->gamma_init; vadd_scaled(work->Qx, work->x, work->Qx, 1 / work->gamma, n); }
  }
}

void oracle_qpalm_cleanup(QPALMWorkspace *work) { /* src/qpalm.c:874-1096 */
  if (!work) return;
  Aux *a = aux_of(work);
  if (work->data) { csc_free(work->data->Q); csc_free(work->data->A); free(work->data->q); free(work->data->bmin); free(work->data->bmax); free(work->data); }
  if (work->scaling) { free(work->scaling->D); free(work->scaling->Dinv); free(work->scaling->E); free(work->scaling->Einv); free(work->scaling); }
#define FR(f) free(work->f)
  FR(x); FR(y); FR(Ax); FR(Qx); FR(x_prev); FR(Aty); FR(x0); FR(temp_m); FR(temp_n); FR(sigma); FR(sigma_inv);
  FR(z); FR(Axys); FR(pri_res); FR(pri_res_in); FR(df); FR(xx0); FR(dphi); FR(dphi_prev); FR(sqrt_sigma);
  FR(delta); FR(alpha); FR(delta2); FR(delta_alpha); FR(temp_2m); FR(s); FR(index_L); FR(index_P); FR(index_J);
  FR(delta_y); FR(Atdelta_y); FR(delta_x); FR(Qdelta_x); FR(Adelta_x);
  FR(neg_dphi); FR(d); FR(Qd); FR(Ad); FR(yh); FR(Atyh); FR(D_temp); FR(E_temp);
  free(work->settings);
  if (work->solver) { free(work->solver->active_constraints); free(work->solver->active_constraints_old); free(work->solver->enter); free(work->solver->leave); free(work->solver); }
  if (a) { free(a->L); free(a->H); free(a->D); free(a->w); free(a->Lq); free(a->Dq); free(a->trace); free(a); }
  if (work->solution) { free(work->solution->x); free(work->solution->y); free(work->solution); }
  free(work->timer); free(work->info); free(work);
}
