/*
 * qpalm_oracle.h -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product
 * (qpalm_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use it.
 *
 * Plain-C, single-threaded restatement of the QPALM hot path (reference: Benny44/QPALM, CHOLMOD
 * build).  It shares the public struct layouts with the product (include/qpalm_b200.h) so that one
 * ctypes harness drives the reference, this oracle and the CUDA library alike.  Every entry point of
 * the product's operator ABI `qpalm_b200_X` has a twin `oracle_X` here with the same signature.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against (i) every known-answer
 * vector the reference's own tests hold for the path, committed in
 * tests/golden/reference_tests.json, and (ii) outputs of the unmodified reference built into
 * oracle/_ref/ (committed as tests/golden/ref_outputs.json by tests/golden/make_golden.py, and
 * compared live whenever oracle/_ref/libqpalm_ref.so is present).
 */
#ifndef QPALM_ORACLE_H
#define QPALM_ORACLE_H
#include "../include/qpalm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

void            oracle_qpalm_set_default_settings(QPALMSettings *settings);
QPALMWorkspace* oracle_qpalm_setup(const QPALMData *data, const QPALMSettings *settings);
void            oracle_qpalm_warm_start(QPALMWorkspace *work, c_float *x_ws, c_float *y_ws);
void            oracle_qpalm_solve(QPALMWorkspace *work);
void            oracle_qpalm_update_settings(QPALMWorkspace *work, const QPALMSettings *settings);
void            oracle_qpalm_update_bounds(QPALMWorkspace *work, const c_float *bmin, const c_float *bmax);
void            oracle_qpalm_update_q(QPALMWorkspace *work, const c_float *q);
void            oracle_qpalm_cleanup(QPALMWorkspace *work);

int oracle_mat_vec(const solver_sparse *A, const c_float *x, c_float *y);
int oracle_mat_tpose_vec(const solver_sparse *A, const c_float *x, c_float *y);
int oracle_mat_inf_norm_cols(const solver_sparse *M, c_float *E);
int oracle_mat_inf_norm_rows(const solver_sparse *M, c_float *E);
int oracle_scale_data(solver_sparse *A, solver_sparse *Q, c_float *q, c_float *bmin, c_float *bmax,
                      c_int scaling_iters, c_float *D, c_float *E, c_float *c_out);
int oracle_residuals_active_set(const solver_sparse *A,
        const c_float *Ax, const c_float *y, const c_float *sigma, const c_float *bmin, const c_float *bmax,
        const c_float *Qx, const c_float *q, const c_float *x0, c_int proximal, c_float gamma,
        const c_int *active_old,
        c_float *Axys, c_float *z, c_float *pri_res, c_float *yh, c_float *Atyh, c_float *df, c_float *dphi,
        c_int *active, c_int *nb_active, c_int *enter, c_int *nb_enter, c_int *leave, c_int *nb_leave);
int oracle_linesearch(c_int m, c_float eta, c_float beta,
        const c_float *Ad, const c_float *Ax, const c_float *y, const c_float *sigma,
        const c_float *sqrt_sigma, const c_float *bmin, const c_float *bmax,
        c_float *tau, c_float *sorted_s, c_int *sorted_idx, c_int *nL);
int oracle_newton_solve(const solver_sparse *Q, const solver_sparse *A, const c_float *sigma,
        const c_int *active, c_float beta, const c_float *rhs, c_float *d, c_float *L_out);
int oracle_updown(c_int n, c_int k, c_float *L, const c_float *W, c_int update);
int oracle_lobpcg(const solver_sparse *Q, const c_float *x0, c_float *lambda_out, c_int *iters_out);

/* per-iteration trace of the last oracle_qpalm_solve (for trajectory comparisons in the tests) */
typedef struct {
  c_int iter, kind;          /* kind: 0 inner step, 1 outer update, 2 forced outer update, 3 terminated */
  c_int nb_active, nb_enter, nb_leave, refactor;
  c_float tau, pri_res_norm, dua_res_norm, gamma;
} OracleTraceEntry;
c_int oracle_trace(const QPALMWorkspace *work, OracleTraceEntry *out, c_int max_entries);

#ifdef __cplusplus
}
#endif
#endif
