// compat.cu -- the reference's INTERNAL entry points under their reference names.
//
// The reference's own test-suite (tests/run_all_tests.c, tests/src/*.c) links not only against the public API of
// include/qpalm.h but also against internals: mat_vec / mat_tpose_vec / mat_inf_norm_cols / mat_inf_norm_rows / ldlchol /
// ldlsolveLD_neg_dphi (include/solver_interface.h:33-232, src/solver_interface.c:252-519), every kernel of include/lin_alg.h
// (src/lin_alg.c) and print_final_message (src/util.c:120-200).  Exporting them here lets that suite be compiled UNMODIFIED
// and linked against libqpalm_b200.so (oracle/Makefile target `reftests`, tests/test_gpu_reference_suite.py).
//
// The matrix / factor entry points run the SAME device kernels as the solver (through the operator ABI and the workspace's
// engine); the `solver_common *` argument of the reference is accepted and ignored.  The lin_alg.h functions operate on the
// caller's HOST arrays -- in the reference they are the inner loops of the iteration, here the iteration runs in fused
// device kernels and these host twins exist only for callers (and tests) that use them as a utility library.
#include "../../include/qpalm_b200.h"
#include "engine.cuh"
#include <math.h>
#include <string.h>
#include <vector>

using namespace qb;

extern "C" {

// cholmod_dense (suitesparse/CHOLMOD/Include/cholmod_core.h:1894-1905): the reference wraps its work vectors in these
typedef struct qpalm_b200_dense { size_t nrow, ncol, nzmax, d; void *x, *z; int xtype, dtype; } solver_dense;

// ---- solver_interface.h ---------------------------------------------------------------------------------------------
void mat_vec(solver_sparse *A, solver_dense *x, solver_dense *y, void *c) {   // y = A x (solver_interface.c:252-262; x may alias y)
  (void)c;
  std::vector<double> out(A->nrow + 1);
  if (qpalm_b200_mat_vec(A, (const c_float *)x->x, out.data())) { fprintf(stderr, "ERROR in mat_vec: device product failed\n"); return; }
  memcpy(y->x, out.data(), sizeof(double) * A->nrow);
}
void mat_tpose_vec(solver_sparse *A, solver_dense *x, solver_dense *y, void *c) {   // y = A' x (solver_interface.c:264-274)
  (void)c;
  std::vector<double> out(A->ncol + 1);
  if (qpalm_b200_mat_tpose_vec(A, (const c_float *)x->x, out.data())) { fprintf(stderr, "ERROR in mat_tpose_vec: device product failed\n"); return; }
  memcpy(y->x, out.data(), sizeof(double) * A->ncol);
}
void mat_inf_norm_cols(solver_sparse *M, c_float *E) { qpalm_b200_mat_inf_norm_cols(M, E); }   // solver_interface.c:276-293
void mat_inf_norm_rows(solver_sparse *M, c_float *E) { qpalm_b200_mat_inf_norm_rows(M, E); }   // solver_interface.c:295-314

// ldlchol (solver_interface.c:319-370): factor M + I/gamma (M alone when !proximal) into the workspace's factor.
// M is the CALLER's matrix (symmetric, lower triangle read when stype = -1, upper when stype = 1).
void ldlchol(solver_sparse *M, QPALMWorkspace *work, void *c) {
  (void)c;
  Engine *e = (Engine *)work->solver->LD;
  if (!e || e->sp) { fprintf(stderr, "ERROR in ldlchol: only the dense Newton factor can take a caller-supplied matrix\n"); return; }
  const int n = e->n, npad = e->npad, ld = e->ld;
  if ((int)M->nrow != n || (int)M->ncol != n) { fprintf(stderr, "ERROR in ldlchol: matrix is not n x n\n"); return; }
  const double beta = work->settings->proximal ? 1.0 / work->gamma : 0.0;
  std::vector<double> h((size_t)ld * npad, 0.0);
  const long long *p = (const long long *)M->p, *ri = (const long long *)M->i;
  const double *xv = (const double *)M->x;
  for (int j = 0; j < n; j++)
    for (long long k = p[j]; k < p[j + 1]; k++) {
      const long long i = ri[k];
      if (M->stype < 0 ? i >= j : (M->stype > 0 ? i <= j : i >= j)) {
        const long long r = i >= j ? i : j, cc = i >= j ? j : i;   // lower-triangle position
        h[(size_t)r + (size_t)ld * cc] += xv[k];
      }
    }
  for (int i = 0; i < npad; i++) h[(size_t)i * (ld + 1)] = (i < n) ? h[(size_t)i * (ld + 1)] + beta : 1.0;
  cudaStreamSynchronize(e->stream);
  if (cudaMemcpy(e->L, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice) != cudaSuccess) { fprintf(stderr, "ERROR in ldlchol: upload failed\n"); return; }
  cudaMemsetAsync(e->info_dev, 0, sizeof(int), e->stream);
  if (potrf_lower(e->stream, npad, e->L, ld, e->invdiag, e->info_dev)) fprintf(stderr, "ERROR in ldlchol: factorization failed\n");
  cudaStreamSynchronize(e->stream);
  work->solver->reset_newton = TRUE;   // the factor no longer belongs to the iteration's Newton system
}

// ldlsolveLD_neg_dphi (solver_interface.c:505-519): d = -(L L')^{-1} dphi with the current factor; work->dphi is the input
void ldlsolveLD_neg_dphi(QPALMWorkspace *work, void *c) {
  (void)c;
  Engine *e = (Engine *)work->solver->LD;
  if (!e) return;
  const int n = e->n;
  upload(e, e->dphi, work->dphi, n);
  step_newton_solve(e);
  download(e, work->d, e->d, n);
  for (int i = 0; i < n; i++) work->neg_dphi[i] = -work->dphi[i];
}

// ---- lin_alg.h (src/lin_alg.c): host utility twins, same results as the reference loops -------------------------------
c_float *vec_copy(const c_float *a, size_t n) {
  c_float *b = (c_float *)malloc((n ? n : 1) * sizeof(c_float));
  if (b && n) memcpy(b, a, n * sizeof(c_float));
  return b;
}
void prea_vec_copy(const c_float *a, c_float *b, size_t n) { for (size_t i = 0; i < n; i++) b[i] = a[i]; }
void prea_int_vec_copy(const c_int *a, c_int *b, size_t n) { for (size_t i = 0; i < n; i++) b[i] = a[i]; }
void vec_set_scalar(c_float *a, c_float sc, size_t n) { for (size_t i = 0; i < n; i++) a[i] = sc; }
void vec_set_scalar_int(c_int *a, c_int sc, size_t n) { for (size_t i = 0; i < n; i++) a[i] = sc; }
void vec_self_mult_scalar(c_float *a, c_float sc, size_t n) { for (size_t i = 0; i < n; i++) a[i] *= sc; }
void vec_mult_scalar(const c_float *a, c_float sc, c_float *b, size_t n) { for (size_t i = 0; i < n; i++) b[i] = sc * a[i]; }
c_float vec_prod(const c_float *a, const c_float *b, size_t n) {   // four partial sums, as src/lin_alg.c:66-90
  c_float prod = 0.0;
  size_t i = 0;
  if (n >= 4) for (; i <= n - 4; i += 4) prod += (a[i] * b[i] + a[i + 1] * b[i + 1] + a[i + 2] * b[i + 2] + a[i + 3] * b[i + 3]);
  for (; i < n; i++) prod += a[i] * b[i];
  return prod;
}
c_float vec_norm_two(const c_float *a, size_t n) { return sqrt(vec_prod(a, a, n)); }
c_float vec_norm_inf(const c_float *a, size_t n) {
  c_float mx = 0.0;
  for (size_t i = 0; i < n; i++) { const c_float v = a[i] < 0 ? -a[i] : a[i]; if (v > mx) mx = v; }
  return mx;
}
void vec_add_scaled(const c_float *a, const c_float *b, c_float *c, c_float sc, size_t n) { for (size_t i = 0; i < n; i++) c[i] = a[i] + sc * b[i]; }
void vec_mult_add_scaled(c_float *a, const c_float *b, c_float sc1, c_float sc2, size_t n) { for (size_t i = 0; i < n; i++) a[i] = sc1 * a[i] + sc2 * b[i]; }
void vec_ew_recipr(const c_float *a, c_float *b, size_t n) { for (size_t i = 0; i < n; i++) b[i] = (c_float)1.0 / a[i]; }
void vec_ew_max_vec(const c_float *a, const c_float *b, c_float *c, size_t n) { for (size_t i = 0; i < n; i++) c[i] = a[i] > b[i] ? a[i] : b[i]; }
void vec_ew_min_vec(const c_float *a, const c_float *b, c_float *c, size_t n) { for (size_t i = 0; i < n; i++) c[i] = a[i] < b[i] ? a[i] : b[i]; }
void vec_ew_mid_vec(const c_float *a, const c_float *bmin, const c_float *bmax, c_float *c, size_t n) {
  for (size_t i = 0; i < n; i++) { const c_float lo = a[i] < bmax[i] ? a[i] : bmax[i]; c[i] = bmin[i] > lo ? bmin[i] : lo; }
}
void vec_ew_prod(const c_float *a, const c_float *b, c_float *c, size_t n) { for (size_t i = 0; i < n; i++) c[i] = a[i] * b[i]; }
void vec_ew_div(const c_float *a, const c_float *b, c_float *c, size_t n) { for (size_t i = 0; i < n; i++) c[i] = a[i] / b[i]; }
void vec_ew_sqrt(const c_float *a, c_float *b, size_t n) { for (size_t i = 0; i < n; i++) b[i] = sqrt(a[i]); }

// ---- util.h: print_final_message (src/util.c:120-200) ---------------------------------------------------------------
void print_final_message(QPALMWorkspace *work) {
  QPALMInfo *info = work->info;
  printf("\n\n=============================================================\n");
  switch (info->status_val) {
    case QPALM_SOLVED:
      printf("| QPALM finished successfully.                              |\n");
      printf("| primal residual: %5.4e, primal tolerance: %5.4e |\n", info->pri_res_norm, work->eps_pri);
      printf("| dual residual  : %5.4e, dual tolerance  : %5.4e |\n", info->dua_res_norm, work->eps_dua);
      printf("| objective value: %+-5.4e                              |\n", info->objective);
      break;
    case QPALM_DUAL_TERMINATED:
      printf("| QPALM has terminated because the dual objective at the    |\n");
      printf("| current iterate is higher than the value specified in     |\n");
      printf("| dual_objective_limit.                                     |\n");
      printf("| dual objective : %+-4.3e, specified limit : %+-4.3e |\n", info->dual_objective, work->settings->dual_objective_limit);
      break;
    case QPALM_PRIMAL_INFEASIBLE:
      printf("| QPALM detected a primal infeasible problem. You can check |\n");
      printf("| the certificate of this infeasiblity. If you think the    |\n");
      printf("| problem might not be infeasible, try lowering the         |\n");
      printf("| infeasiblity tolerance eps_prim_inf.                      |\n");
      break;
    case QPALM_DUAL_INFEASIBLE:
      printf("| QPALM detected a dual infeasible problem. You can check   |\n");
      printf("| the certificate of this infeasiblity. If you think the    |\n");
      printf("| problem might not be dual infeasible, try lowering the    |\n");
      printf("| infeasiblity tolerance eps_dual_inf.                      |\n");
      break;
    case QPALM_MAX_ITER_REACHED:
      printf("| QPALM hit the maximum number of iterations.               |\n");
      printf("| primal residual: %5.4e, primal tolerance: %5.4e |\n", info->pri_res_norm, work->eps_pri);
      printf("| dual residual  : %5.4e, dual tolerance  : %5.4e |\n", info->dua_res_norm, work->eps_dua);
      printf("| objective value: %+-5.4e                              |\n", info->objective);
      break;
    case QPALM_TIME_LIMIT_REACHED:
      printf("| QPALM has exceeded the specified time limit.              |\n");
      printf("| primal residual: %5.4e, primal tolerance: %5.4e |\n", info->pri_res_norm, work->eps_pri);
      printf("| dual residual  : %5.4e, dual tolerance  : %5.4e |\n", info->dua_res_norm, work->eps_dua);
      printf("| objective value: %+-5.4e                              |\n", info->objective);
      break;
    default:   // also QPALM_UNSOLVED / QPALM_ERROR: the reference rewrites the status string here (util.c:180-183)
      strcpy(work->info->status, "unrecognised status value");
      fprintf(stderr, "ERROR in print_final_message: Unrecognised final status value %ld\n", (long)info->status_val);
      return;
  }
  {
    char buf[80];
    if (info->run_time > 1.0) snprintf(buf, 80, "| runtime:         %4.2f seconds", info->run_time);
    else snprintf(buf, 80, "| runtime:         %4.2f milliseconds", info->run_time * 1000);
    printf("%s", buf);
    for (size_t k = strlen(buf); k < 59; k++) printf(" ");
    printf("|\n");
  }
  printf("=============================================================\n\n\n");
}

}  // extern "C"
