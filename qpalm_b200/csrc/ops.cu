// ops.cu -- operator-level C ABI (include/qpalm_b200.h Part 2): host buffers in, host buffers out, each
// call runs the same device kernels the solver uses.  These are the entry points the parity tests drive
// "at identical iterates" against the oracle and the reference.
#include "../../include/qpalm_b200.h"
#include "engine.cuh"
#include "sparse_host.h"
#include <math.h>
#include <string.h>
#include <vector>
#include <string>

using namespace qb;

extern "C" int qpalm_b200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}
extern "C" const char *qpalm_b200_version(void) { return "qpalm_b200 0.1 (sm_100a)"; }

namespace {
struct EmptyCSC {
  std::vector<long long> p, i; std::vector<double> x;
  explicit EmptyCSC(size_t ncol) : p(ncol + 1, 0), i(1, 0), x(1, 0.0) {}
};
// engine from optional A (m x n) and optional Q (n x n); missing pieces are empty
int make_engine(Engine **e, int n, int m, const solver_sparse *A, const solver_sparse *Q, bool need_LQ = false, int newton_override = 0) {
  EmptyCSC ea((size_t)n), eq((size_t)n);
  std::vector<double> zn((size_t)n + 1, 0.0), zm((size_t)m + 1, 0.0);
  const long long *Ap = A ? (const long long *)A->p : ea.p.data(), *Ai = A ? (const long long *)A->i : ea.i.data();
  const double *Ax = A ? (const double *)A->x : ea.x.data();
  const long long *Qp = Q ? (const long long *)Q->p : eq.p.data(), *Qi = Q ? (const long long *)Q->i : eq.i.data();
  const double *Qx = Q ? (const double *)Q->x : eq.x.data();
  return engine_create(e, n, m, Ap, Ai, Ax, Qp, Qi, Qx, zn.data(), zm.data(), zm.data(), need_LQ, newton_override);
}
int up_int(Engine *e, int *dst, const c_int *src, int len) {
  std::vector<int> t((size_t)len + 1);
  for (int i = 0; i < len; i++) t[i] = (int)src[i];
  QB_CUDA_TRY(cudaMemcpy(dst, t.data(), sizeof(int) * (size_t)len, cudaMemcpyHostToDevice));
  (void)e;
  return 0;
}
}  // namespace

extern "C" int qpalm_b200_mat_vec(const solver_sparse *A, const c_float *x, c_float *y) {
  Engine *e = nullptr;
  const int nrow = (int)A->nrow, ncol = (int)A->ncol;
  int rc;
  if (A->stype != 0) {   // symmetric: only the lower triangle is read (stype -1)
    if ((rc = make_engine(&e, ncol, 0, nullptr, A))) return rc;
    upload(e, e->x, x, ncol); spmv_Q(e, e->x, e->Qx); rc = download(e, y, e->Qx, ncol);
  } else {
    if ((rc = make_engine(&e, ncol, nrow, A, nullptr))) return rc;
    upload(e, e->x, x, ncol); spmv_A(e, e->x, e->Ax); rc = download(e, y, e->Ax, nrow);
  }
  engine_destroy(e);
  return rc;
}
extern "C" int qpalm_b200_mat_tpose_vec(const solver_sparse *A, const c_float *x, c_float *y) {
  if (A->stype != 0) return qpalm_b200_mat_vec(A, x, y);
  Engine *e = nullptr;
  const int nrow = (int)A->nrow, ncol = (int)A->ncol;
  int rc;
  if ((rc = make_engine(&e, ncol, nrow, A, nullptr))) return rc;
  upload(e, e->y, x, nrow); spmv_At(e, e->y, e->Aty); rc = download(e, y, e->Aty, ncol);
  engine_destroy(e);
  return rc;
}

namespace qb { int ruiz_norms_public(Engine *e, double *colnorm_dev, double *rownorm_dev); }

extern "C" int qpalm_b200_scale_data(solver_sparse *A, solver_sparse *Q, c_float *q, c_float *bmin, c_float *bmax,
                                     c_int scaling_iters, c_float *D, c_float *E, c_float *c_out) {
  const int n = (int)Q->ncol, m = (int)A->nrow;
  Engine *e = nullptr;
  int rc;
  if ((rc = make_engine(&e, n, m, A, Q))) return rc;
  upload(e, e->q, q, n); upload(e, e->bmin, bmin, m); upload(e, e->bmax, bmax, m);
  double cc = 1.0;
  rc = engine_ruiz_scale(e, (int)scaling_iters, &cc);
  *c_out = cc;
  rc |= download(e, D, e->D, n); rc |= download(e, E, e->E, m);
  rc |= download(e, q, e->q, n); rc |= download(e, bmin, e->bmin, m); rc |= download(e, bmax, e->bmax, m);
  // scaled matrix values back into the caller's CSC arrays
  const long long *Ap = (const long long *)A->p, *Ai = (const long long *)A->i; double *Axv = (double *)A->x;
  if (m > 0) {
    if (e->A_dense) {
      std::vector<double> At((size_t)n * m);
      rc |= download(e, At.data(), e->At, n * m);
      for (int j = 0; j < n; j++) for (long long k = Ap[j]; k < Ap[j + 1]; k++) Axv[k] = At[(size_t)j + (size_t)n * Ai[k]];
    } else rc |= download(e, Axv, e->A_csc.x, (int)Ap[n]);
  }
  const long long *Qp = (const long long *)Q->p, *Qi = (const long long *)Q->i; double *Qxv = (double *)Q->x;
  if (e->Q_dense) {
    std::vector<double> Qd((size_t)n * n);
    rc |= download(e, Qd.data(), e->Qd, n * n);
    for (int j = 0; j < n; j++) for (long long k = Qp[j]; k < Qp[j + 1]; k++) if (Qi[k] >= j) Qxv[k] = Qd[(size_t)Qi[k] + (size_t)n * j];
  } else {
    std::vector<int> rp((size_t)n + 1);
    std::vector<double> rx((size_t)e->Q_csr.nnz + 1);
    QB_CUDA_TRY(cudaMemcpy(rp.data(), e->Q_csr.p, sizeof(int) * ((size_t)n + 1), cudaMemcpyDeviceToHost));
    rc |= download(e, rx.data(), e->Q_csr.x, (int)e->Q_csr.nnz);
    std::vector<int> fill(rp.begin(), rp.end());
    for (int j = 0; j < n; j++) for (long long k = Qp[j]; k < Qp[j + 1]; k++) { const long long i = Qi[k]; if (i < j) continue; Qxv[k] = rx[fill[i]++]; }
  }
  engine_destroy(e);
  return rc;
}

extern "C" int qpalm_b200_mat_inf_norm_cols(const solver_sparse *M, c_float *E) {
  Engine *e = nullptr; int rc;
  if ((rc = make_engine(&e, (int)M->ncol, (int)M->nrow, M, nullptr))) return rc;
  rc = ruiz_norms_public(e, e->tmp_n, e->tmp_m);
  rc |= download(e, E, e->tmp_n, (int)M->ncol);
  engine_destroy(e);
  return rc;
}
extern "C" int qpalm_b200_mat_inf_norm_rows(const solver_sparse *M, c_float *E) {
  Engine *e = nullptr; int rc;
  if ((rc = make_engine(&e, (int)M->ncol, (int)M->nrow, M, nullptr))) return rc;
  rc = ruiz_norms_public(e, e->tmp_n, e->tmp_m);
  rc |= download(e, E, e->tmp_m, (int)M->nrow);
  engine_destroy(e);
  return rc;
}

extern "C" int qpalm_b200_residuals_active_set(const solver_sparse *A,
        const c_float *Ax, const c_float *y, const c_float *sigma, const c_float *bmin, const c_float *bmax,
        const c_float *Qx, const c_float *q, const c_float *x0, c_int proximal, c_float gamma,
        const c_int *active_old,
        c_float *Axys, c_float *z, c_float *pri_res, c_float *yh, c_float *Atyh, c_float *df, c_float *dphi,
        c_int *active, c_int *nb_active, c_int *enter, c_int *nb_enter, c_int *leave, c_int *nb_leave) {
  const int m = (int)A->nrow, n = (int)A->ncol;
  Engine *e = nullptr; int rc;
  if ((rc = make_engine(&e, n, m, A, nullptr))) return rc;
  std::vector<double> sinv((size_t)m + 1);
  for (int i = 0; i < m; i++) sinv[i] = 1.0 / sigma[i];
  upload(e, e->Ax, Ax, m); upload(e, e->y, y, m); upload(e, e->sigma, sigma, m); upload(e, e->sigma_inv, sinv.data(), m);
  upload(e, e->bmin, bmin, m); upload(e, e->bmax, bmax, m); upload(e, e->Qx, Qx, n); upload(e, e->q, q, n); upload(e, e->x0, x0, n);
  cudaStreamSynchronize(e->stream);
  up_int(e, e->active_old, active_old, m);
  e->scaling = 0;
  rc = step_residuals(e, proximal != 0, gamma, 0.0);
  rc |= step_compact_lists(e);
  rc |= sync_scalars(e);
  *nb_active = (c_int)e->scal_host[S_NB_ACTIVE]; *nb_enter = (c_int)e->scal_host[S_NB_ENTER]; *nb_leave = (c_int)e->scal_host[S_NB_LEAVE];
  rc |= download(e, Axys, e->Axys, m); rc |= download(e, z, e->z, m); rc |= download(e, pri_res, e->pri_res, m);
  rc |= download(e, yh, e->yh, m); rc |= download(e, Atyh, e->Atyh, n); rc |= download(e, df, e->df, n); rc |= download(e, dphi, e->dphi, n);
  rc |= download_int(e, (long long *)active, e->active, m);
  rc |= download_int(e, (long long *)enter, e->enter, (int)*nb_enter);
  rc |= download_int(e, (long long *)leave, e->leave, (int)*nb_leave);
  engine_destroy(e);
  return rc;
}

extern "C" int qpalm_b200_linesearch(c_int m_, c_float eta, c_float beta,
        const c_float *Ad, const c_float *Ax, const c_float *y, const c_float *sigma,
        const c_float *sqrt_sigma, const c_float *bmin, const c_float *bmax,
        c_float *tau, c_float *sorted_s, c_int *sorted_idx, c_int *nL) {
  const int m = (int)m_;
  Engine *e = nullptr; int rc;
  if ((rc = make_engine(&e, 1, m, nullptr, nullptr))) return rc;
  upload(e, e->Ad, Ad, m); upload(e, e->Ax, Ax, m); upload(e, e->y, y, m); upload(e, e->sigma, sigma, m);
  upload(e, e->sqrt_sigma, sqrt_sigma, m); upload(e, e->bmin, bmin, m); upload(e, e->bmax, bmax, m);
  double eb[2] = {eta, beta};
  QB_CUDA_TRY(cudaMemcpyAsync(e->scal_dev + S_ETA, eb, sizeof(double) * 2, cudaMemcpyHostToDevice, e->stream));
  rc = linesearch_device(e, m, e->Ad, e->Ax, e->y, e->sigma, e->sqrt_sigma, e->bmin, e->bmax);
  rc |= sync_scalars(e);
  *tau = e->scal_host[S_TAU];
  const int n_l = (int)e->scal_host[S_NL];
  if (nL) *nL = n_l;
  if (sorted_s && sorted_idx && n_l > 0) {
    std::vector<unsigned long long> k((size_t)n_l); std::vector<unsigned int> v((size_t)n_l);
    QB_CUDA_TRY(cudaMemcpy(k.data(), e->ls_key[0], sizeof(unsigned long long) * (size_t)n_l, cudaMemcpyDeviceToHost));
    QB_CUDA_TRY(cudaMemcpy(v.data(), e->ls_val[0], sizeof(unsigned int) * (size_t)n_l, cudaMemcpyDeviceToHost));
    for (int i = 0; i < n_l; i++) { memcpy(&sorted_s[i], &k[i], 8); sorted_idx[i] = (c_int)v[i]; }
  }
  engine_destroy(e);
  return rc;
}

extern "C" int qpalm_b200_newton_solve(const solver_sparse *Q, const solver_sparse *A, const c_float *sigma,
        const c_int *active, c_float beta, const c_float *rhs, c_float *d, c_float *L_out) {
  const int n = (int)Q->ncol, m = A ? (int)A->nrow : 0;
  Engine *e = nullptr; int rc;
  if ((rc = make_engine(&e, n, m, A, Q, false, L_out ? 1 : 0))) return rc;   // a dense L can only be returned by the dense path
  int na = 0;
  if (active && m > 0) {
    std::vector<double> ss((size_t)m);
    for (int i = 0; i < m; i++) { ss[i] = sqrt(sigma[i]); na += active[i] != 0; }
    upload(e, e->sigma, sigma, m); upload(e, e->sqrt_sigma, ss.data(), m);
    cudaStreamSynchronize(e->stream);
    up_int(e, e->active, active, m);
  }
  std::vector<double> neg((size_t)n);
  for (int i = 0; i < n; i++) neg[i] = -rhs[i];
  upload(e, e->dphi, neg.data(), n);
  rc = step_newton_refactor(e, active != nullptr && m > 0, true, beta, na);
  rc |= step_newton_solve(e);
  rc |= download(e, d, e->d, n);
  if (L_out) {
    std::vector<double> L((size_t)e->ld * e->npad);
    QB_CUDA_TRY(cudaMemcpy(L.data(), e->L, sizeof(double) * L.size(), cudaMemcpyDeviceToHost));
    for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) L_out[(size_t)i + (size_t)n * j] = (i >= j) ? L[(size_t)i + (size_t)e->ld * j] : 0.0;
  }
  int info = 0;
  QB_CUDA_TRY(cudaMemcpy(&info, e->info_dev, sizeof(int), cudaMemcpyDeviceToHost));
  engine_destroy(e);
  return rc ? rc : (info ? 1000 + info : 0);
}

extern "C" int qpalm_b200_updown(c_int n_, c_int k_, c_float *L, const c_float *W, c_int update) {
  const int n = (int)n_, k = (int)k_;
  Engine *e = nullptr; int rc;
  if ((rc = make_engine(&e, n, 0, nullptr, nullptr, false, 1))) return rc;   // dense-factor operator: never the supernodal path
  const int ld = e->ld, npad = e->npad;
  std::vector<double> Lp((size_t)ld * npad, 0.0);
  for (int j = 0; j < npad; j++) for (int i = j; i < npad; i++)
    Lp[(size_t)i + (size_t)ld * j] = (i < n && j < n) ? L[(size_t)i + (size_t)n * j] : (i == j ? 1.0 : 0.0);
  QB_CUDA_TRY(cudaMemcpy(e->L, Lp.data(), sizeof(double) * Lp.size(), cudaMemcpyHostToDevice));
  QB_CUDA_TRY(cudaMemset(e->info_dev, 0, sizeof(int)));
  rc = 0;
  static const int flow_min = [] { const char *s = getenv("QPALM_B200_UPDOWN_FLOW_MIN"); return s ? atoi(s) : 256; }();
  bool flow = npad >= flow_min;   // same dispatch as step_newton_updown: one dataflow launch per <= 64 columns
  const int chunk = flow ? chol_updown_flow_max_rank() : 8;
  bool gen = use_updown_gen(e);   // ... preceded by the generator-form passes of <= 32 columns (updown_gen.cu)
  for (int off = 0; off < k && !rc && gen; off += chol_updown_gen_max_rank()) {
    const int kk = k - off < chol_updown_gen_max_rank() ? k - off : chol_updown_gen_max_rank();
    std::vector<double> Wp((size_t)ld * kk, 0.0);
    for (int c = 0; c < kk; c++) for (int i = 0; i < n; i++) Wp[(size_t)i + (size_t)ld * c] = W[(size_t)i + (size_t)n * (off + c)];
    QB_CUDA_TRY(cudaStreamSynchronize(e->stream));
    QB_CUDA_TRY(cudaMemcpy(e->W, Wp.data(), sizeof(double) * Wp.size(), cudaMemcpyHostToDevice));
    rc = trtri_diag_blocks(e->stream, npad, e->L, ld, e->invdiag);
    if (!rc) rc = chol_updown_gen(e->stream, npad, e->L, ld, e->invdiag, e->W, ld, kk, update ? kk : 0, e->info_dev);
    if (rc == 1 && off == 0) { gen = false; rc = 0; }
  }
  if (gen) flow = false;
  for (int off = 0; off < k && !rc && flow; off += chunk) {
    const int kk = k - off < chunk ? k - off : chunk;
    std::vector<double> Wp((size_t)ld * kk, 0.0);
    for (int c = 0; c < kk; c++) for (int i = 0; i < n; i++) Wp[(size_t)i + (size_t)ld * c] = W[(size_t)i + (size_t)n * (off + c)];
    QB_CUDA_TRY(cudaStreamSynchronize(e->stream));
    QB_CUDA_TRY(cudaMemcpy(e->W, Wp.data(), sizeof(double) * Wp.size(), cudaMemcpyHostToDevice));
    rc = chol_updown_flow(e->stream, npad, e->L, ld, e->W, ld, kk, update ? kk : 0, e->info_dev);
    if (rc == 1 && off == 0) { flow = false; rc = 0; }   // no cooperative launch here: per-panel kernels below
  }
  for (int off = 0; off < k && !rc && !flow && !gen; off += 8) {
    const int kk = k - off < 8 ? k - off : 8;
    std::vector<double> Wp((size_t)ld * 8, 0.0);
    for (int c = 0; c < kk; c++) for (int i = 0; i < n; i++) Wp[(size_t)i + (size_t)ld * c] = W[(size_t)i + (size_t)n * (off + c)];
    QB_CUDA_TRY(cudaStreamSynchronize(e->stream));   // e->stream is non-blocking: order the copy after the previous sweep
    QB_CUDA_TRY(cudaMemcpy(e->W, Wp.data(), sizeof(double) * Wp.size(), cudaMemcpyHostToDevice));
    rc = chol_updown(e->stream, npad, e->L, ld, e->W, ld, kk, update ? +1 : -1, e->ud_coef, e->info_dev);
  }
  QB_CUDA_TRY(cudaStreamSynchronize(e->stream));
  QB_CUDA_TRY(cudaMemcpy(Lp.data(), e->L, sizeof(double) * Lp.size(), cudaMemcpyDeviceToHost));
  for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) L[(size_t)i + (size_t)n * j] = (i >= j) ? Lp[(size_t)i + (size_t)ld * j] : 0.0;
  int info = 0;
  QB_CUDA_TRY(cudaMemcpy(&info, e->info_dev, sizeof(int), cudaMemcpyDeviceToHost));
  engine_destroy(e);
  return rc ? rc : (info ? 1000 : 0);
}


// ------------------------------------------------------------------------------------------------
// sparse Newton system (sparse.cuh): host-only symbolic analysis + device factor / update / solve
// ------------------------------------------------------------------------------------------------
namespace {
struct HostA { std::vector<int> cp, ci, rp, rj; };
void host_csr_csc(const solver_sparse *A, int n, HostA *h) {
  const int m = A ? (int)A->nrow : 0;
  h->cp.assign((size_t)n + 1, 0); h->rp.assign((size_t)m + 1, 0);
  if (!A || m == 0) return;
  const long long *Ap = (const long long *)A->p, *Ai = (const long long *)A->i;
  const long long nnz = Ap[n];
  h->ci.resize(nnz); h->rj.resize(nnz);
  for (int j = 0; j <= n; j++) h->cp[j] = (int)Ap[j];
  for (long long k = 0; k < nnz; k++) { h->ci[k] = (int)Ai[k]; h->rp[Ai[k] + 1]++; }
  for (int i = 0; i < m; i++) h->rp[i + 1] += h->rp[i];
  std::vector<int> fill(h->rp.begin(), h->rp.end() - 1);
  for (int j = 0; j < n; j++) for (long long k = Ap[j]; k < Ap[j + 1]; k++) h->rj[fill[Ai[k]]++] = j;
}
}  // namespace

struct QPALMB200Symbolic { SymHost h; };

extern "C" QPALMB200Symbolic *qpalm_b200_symbolic_analyze(const solver_sparse *Q, const solver_sparse *A) {
  const int n = (int)Q->ncol, m = A ? (int)A->nrow : 0;
  HostA ha; host_csr_csc(A, n, &ha);
  QPALMB200Symbolic *S = new QPALMB200Symbolic();
  if (symbolic_analyze(n, m, ha.cp.data(), ha.ci.data(), ha.rp.data(), ha.rj.data(), (const long long *)Q->p, (const long long *)Q->i,
                       true, &S->h) != 0) { delete S; return nullptr; }
  return S;
}
extern "C" int qpalm_b200_symbolic_info(const QPALMB200Symbolic *S, c_int out8[8], double *flops) {
  const SymHost &h = S->h;
  out8[0] = h.n; out8[1] = h.nsuper; out8[2] = h.nlevels; out8[3] = h.max_ns; out8[4] = h.max_nf; out8[5] = h.nnzS; out8[6] = h.nnzL;
  out8[7] = h.upd_total;
  if (flops) *flops = h.flops;
  return 0;
}
extern "C" c_int qpalm_b200_symbolic_array(const QPALMB200Symbolic *S, const char *name, c_int *out, c_int cap) {
  const SymHost &h = S->h;
  const std::vector<int> *v = nullptr; const std::vector<long long> *w = nullptr;
  const std::string nm(name);
  if (nm == "perm") v = &h.perm; else if (nm == "iperm") v = &h.iperm; else if (nm == "sn_first") v = &h.sn_first;
  else if (nm == "sn_of_col") v = &h.sn_of_col; else if (nm == "rows_off") v = &h.rows_off; else if (nm == "rowidx") v = &h.rowidx;
  else if (nm == "rel") v = &h.rel; else if (nm == "sn_parent") v = &h.sn_parent; else if (nm == "child_ptr") v = &h.child_ptr;
  else if (nm == "child_idx") v = &h.child_idx; else if (nm == "lvl_ptr") v = &h.lvl_ptr; else if (nm == "lvl_sn") v = &h.lvl_sn;
  else if (nm == "panel_off") w = &h.panel_off; else if (nm == "upd_off") w = &h.upd_off;
  else return -1;
  const c_int len = v ? (c_int)v->size() : (c_int)w->size();
  if (out) for (c_int i = 0; i < len && i < cap; i++) out[i] = v ? (c_int)(*v)[i] : (c_int)(*w)[i];
  return len;
}
extern "C" void qpalm_b200_symbolic_free(QPALMB200Symbolic *S) { delete S; }

extern "C" int qpalm_b200_sparse_newton(const solver_sparse *Q, const solver_sparse *A, const c_float *sigma, const c_int *active,
        c_float beta, const c_float *rhs, c_float *d, c_float *L_out, c_int *perm_out, const c_int *enter, c_int nb_enter,
        const c_int *leave, c_int nb_leave, c_float *rowsums_out) {
  const int n = (int)Q->ncol, m = A ? (int)A->nrow : 0;
  Engine *e = nullptr; int rc;
  setenv("QPALM_B200_NEWTON", "sparse", 1);
  rc = make_engine(&e, n, m, A, Q);
  unsetenv("QPALM_B200_NEWTON");
  if (rc) return rc;
  if (!e->sp) { engine_destroy(e); return 7; }
  int na = 0;
  if (m > 0) {
    std::vector<double> ss((size_t)m);
    std::vector<c_int> act((size_t)m, 0);
    for (int i = 0; i < m; i++) { ss[i] = sqrt(sigma[i]); if (active) act[i] = active[i]; na += act[i] != 0; }
    upload(e, e->sigma, sigma, m); upload(e, e->sqrt_sigma, ss.data(), m);
    cudaStreamSynchronize(e->stream);
    up_int(e, e->active, act.data(), m);
  }
  std::vector<double> neg((size_t)n);
  for (int i = 0; i < n; i++) neg[i] = -rhs[i];
  upload(e, e->dphi, neg.data(), n);
  QB_CUDA_TRY(cudaMemset(e->info_dev, 0, sizeof(int)));
  if (rowsums_out) {   // Gershgorin row sums of A_J' Sigma_J A_J in the ORIGINAL ordering (boost_gamma)
    double ub = 0;
    rc = step_gershgorin_AtSA(e, &ub);
    rowsums_out[0] = ub;
  }
  rc |= step_newton_refactor(e, m > 0, true, beta, na);
  if (nb_enter > 0) { up_int(e, e->enter, enter, (int)nb_enter); }
  if (nb_leave > 0) { up_int(e, e->leave, leave, (int)nb_leave); }
  if (nb_enter + nb_leave > 0) rc |= step_newton_updown(e, (int)nb_enter, (int)nb_leave);
  rc |= step_newton_solve(e);
  rc |= download(e, d, e->d, n);
  if (L_out && perm_out) rc |= sparse_chol_download(e->sp, e->stream, e->spL, L_out, (long long *)perm_out);
  int info = 0;
  QB_CUDA_TRY(cudaMemcpy(&info, e->info_dev, sizeof(int), cudaMemcpyDeviceToHost));
  engine_destroy(e);
  return rc ? rc : (info ? 1000 + info : 0);
}

extern "C" int qpalm_b200_lobpcg(const solver_sparse *Q, const c_float *x0, c_float *lambda_out, c_int *iters_out) {
  Engine *e = nullptr; int rc;
  if ((rc = make_engine(&e, (int)Q->ncol, 0, nullptr, Q))) return rc;
  long long its = 0;
  rc = lobpcg_device(e, x0, lambda_out, &its);
  if (iters_out) *iters_out = its;
  engine_destroy(e);
  return rc;
}

// ---- micro-benchmarks for the FP64 tensor roofline ---------------------------------------------------
__global__ void k_fill_rand(double *p, size_t len, unsigned long long seed) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < len; i += (size_t)gridDim.x * blockDim.x) {
    unsigned long long s = seed + i * 0x9E3779B97F4A7C15ull;
    s ^= s >> 30; s *= 0xBF58476D1CE4E5B9ull; s ^= s >> 27; s *= 0x94D049BB133111EBull; s ^= s >> 31;
    p[i] = (double)(s >> 11) * (1.0 / 9007199254740992.0) - 0.5;
  }
}
extern "C" int qpalm_b200_bench_dsyrk(c_int n_, c_int k_, c_int reps, double *ms_out) {
  const int n = round_up((int)n_, 128), k = round_up((int)k_, 16);
  double *W = nullptr, *C = nullptr;
  QB_CUDA_TRY(cudaMalloc(&W, sizeof(double) * (size_t)n * k));
  QB_CUDA_TRY(cudaMalloc(&C, sizeof(double) * (size_t)n * n));
  cudaStream_t s; QB_CUDA_TRY(cudaStreamCreate(&s));
  QB_LAUNCH(k_fill_rand, 1024, 256, 0, s, W, (size_t)n * k, 1234ull);
  QB_CUDA_TRY(cudaMemsetAsync(C, 0, sizeof(double) * (size_t)n * n, s));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  int rc = dgemm_nt(s, n, n, k, W, n, W, n, C, n, 1.0, 1.0, true);   // warm-up
  cudaEventRecord(a, s);
  for (int r = 0; r < reps && !rc; r++) rc = dgemm_nt(s, n, n, k, W, n, W, n, C, n, 1.0, 1.0, true);
  cudaEventRecord(b, s);
  QB_CUDA_TRY(cudaEventSynchronize(b));
  float ms = 0; cudaEventElapsedTime(&ms, a, b);
  *ms_out = ms / (reps > 0 ? reps : 1);
  cudaEventDestroy(a); cudaEventDestroy(b); cudaStreamDestroy(s); cudaFree(W); cudaFree(C);
  return rc;
}
__global__ void k_make_spd(int n, double *A) {   // A <- small random symmetric part + n on the diagonal (lower used)
  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
  if (i < n && i == j) A[(size_t)i + (size_t)n * j] += (double)n;
}
extern "C" int qpalm_b200_bench_potrf(c_int n_, c_int reps, double *ms_out) {
  const int n = round_up((int)n_, 128);
  double *A = nullptr, *L = nullptr, *X = nullptr; int *info = nullptr;
  QB_CUDA_TRY(cudaMalloc(&A, sizeof(double) * (size_t)n * n));
  QB_CUDA_TRY(cudaMalloc(&L, sizeof(double) * (size_t)n * n));
  QB_CUDA_TRY(cudaMalloc(&X, sizeof(double) * (size_t)n * 128));
  QB_CUDA_TRY(cudaMalloc(&info, sizeof(int)));
  QB_CUDA_TRY(cudaMemset(info, 0, sizeof(int)));
  cudaStream_t s; QB_CUDA_TRY(cudaStreamCreate(&s));
  QB_LAUNCH(k_fill_rand, 1024, 256, 0, s, A, (size_t)n * n, 99ull);
  dim3 g(cdiv(n, 256), n);
  QB_LAUNCH(k_make_spd, g, 256, 0, s, n, A);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float total = 0; int rc = 0;
  for (int r = 0; r < reps + 1 && !rc; r++) {
    QB_CUDA_TRY(cudaMemcpyAsync(L, A, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToDevice, s));
    cudaEventRecord(a, s);
    rc = potrf_lower(s, n, L, n, X, info);
    cudaEventRecord(b, s);
    QB_CUDA_TRY(cudaEventSynchronize(b));
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    if (r > 0) total += ms;
  }
  *ms_out = total / (reps > 0 ? reps : 1);
  int hinfo = 0; cudaMemcpy(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost);
  cudaEventDestroy(a); cudaEventDestroy(b); cudaStreamDestroy(s); cudaFree(A); cudaFree(L); cudaFree(X); cudaFree(info);
  return rc ? rc : hinfo;
}

// one rank-k sweep (update, then the matching downdate so the factor stays bounded) on a factored random SPD matrix:
// ms per sweep, averaged over 2 * reps sweeps
extern "C" int qpalm_b200_bench_updown(c_int n_, c_int k_, c_int reps, double *ms_out) {
  const int n = round_up((int)n_, 128), k = (int)k_;
  if (k < 1 || k > chol_updown_flow_max_rank()) return 1;
  double *L = nullptr, *X = nullptr, *W = nullptr; int *info = nullptr;
  QB_CUDA_TRY(cudaMalloc(&L, sizeof(double) * (size_t)n * n));
  QB_CUDA_TRY(cudaMalloc(&X, sizeof(double) * (size_t)n * 128));
  QB_CUDA_TRY(cudaMalloc(&W, sizeof(double) * (size_t)n * k));
  QB_CUDA_TRY(cudaMalloc(&info, sizeof(int)));
  QB_CUDA_TRY(cudaMemset(info, 0, sizeof(int)));
  cudaStream_t s; QB_CUDA_TRY(cudaStreamCreate(&s));
  QB_LAUNCH(k_fill_rand, 1024, 256, 0, s, L, (size_t)n * n, 99ull);
  QB_LAUNCH(k_fill_rand, 256, 256, 0, s, W, (size_t)n * k, 7ull);
  dim3 g(cdiv(n, 256), n);
  QB_LAUNCH(k_make_spd, g, 256, 0, s, n, L);
  int rc = potrf_lower(s, n, L, n, X, info);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float total = 0;
  const char *genv = getenv("QPALM_B200_UPDOWN_GEN");
  const bool gen = !(genv && atoi(genv) == 0);
  for (int r = 0; r < reps + 1 && !rc; r++) {
    cudaEventRecord(a, s);
    if (gen && k <= chol_updown_gen_max_rank()) {     // each pass with the refresh of the inverted diagonal blocks it needs
      rc = chol_updown_gen(s, n, L, n, X, W, n, k, k, info);
      if (!rc) rc = trtri_diag_blocks(s, n, L, n, X);
      if (!rc) rc = chol_updown_gen(s, n, L, n, X, W, n, k, 0, info);
      if (!rc) rc = trtri_diag_blocks(s, n, L, n, X);
    } else {
      rc = chol_updown_flow(s, n, L, n, W, n, k, k, info);
      if (!rc) rc = chol_updown_flow(s, n, L, n, W, n, k, 0, info);
    }
    cudaEventRecord(b, s);
    QB_CUDA_TRY(cudaEventSynchronize(b));
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    if (r > 0) total += ms;
  }
  *ms_out = total / (2.0 * (reps > 0 ? reps : 1));
  int hinfo = 0; cudaMemcpy(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost);
  chol_updown_flow_release(s); chol_updown_gen_release(s);
  cudaEventDestroy(a); cudaEventDestroy(b); cudaStreamDestroy(s); cudaFree(L); cudaFree(X); cudaFree(W); cudaFree(info);
  return rc ? rc : hinfo;
}

// sparse products at a caller-given shape (C1 / C2 of BASELINE.json): ms per A x, A' y and Q x through the solver's kernels.
// out3 = {ms A x, ms A' y, ms Q x}; algorithmic bytes per product are 12 nnz + 4 (rows + 1) + 8 (rows + cols).
extern "C" int qpalm_b200_bench_spmv(const solver_sparse *A, const solver_sparse *Q, c_int reps, double *out3, c_int *nnz3) {
  Engine *e = nullptr;
  const int n = (int)Q->ncol, m = (int)A->nrow;
  if (int rc = make_engine(&e, n, m, A, Q, false, 2)) return rc;
  if (e->A_dense || e->Q_dense) { engine_destroy(e); return 7; }
  vec_set(e, e->x, 1.0, n); vec_set(e, e->y, 1.0, m);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int which = 0; which < 3; which++) {
    auto run = [&]() { return which == 0 ? spmv_A(e, e->x, e->Ax) : (which == 1 ? spmv_At(e, e->y, e->Aty) : spmv_Q(e, e->x, e->Qx)); };
    run();
    cudaEventRecord(a, e->stream);
    for (int r = 0; r < reps; r++) run();
    cudaEventRecord(b, e->stream);
    QB_CUDA_TRY(cudaEventSynchronize(b));
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    out3[which] = ms / (reps > 0 ? reps : 1);
  }
  nnz3[0] = e->A_csr.nnz; nnz3[1] = e->A_csc.nnz; nnz3[2] = e->Q_csr.nnz;
  cudaEventDestroy(a); cudaEventDestroy(b);
  engine_destroy(e);
  return 0;
}

// stage clocks of the chain CTA of one sweep (panels 100 and 101): out32 receives up to 32 clock64 stamps
extern "C" int qpalm_b200_bench_updown_clocks(c_int n_, c_int k_, long long *out32) {
  const int n = round_up((int)n_, 128), k = (int)k_;
  double *L = nullptr, *X = nullptr, *W = nullptr; int *info = nullptr; long long *clk = nullptr;
  QB_CUDA_TRY(cudaMalloc(&L, sizeof(double) * (size_t)n * n));
  QB_CUDA_TRY(cudaMalloc(&X, sizeof(double) * (size_t)n * 128));
  QB_CUDA_TRY(cudaMalloc(&W, sizeof(double) * (size_t)n * k));
  QB_CUDA_TRY(cudaMalloc(&info, sizeof(int)));
  QB_CUDA_TRY(cudaMalloc(&clk, sizeof(long long) * 32));
  QB_CUDA_TRY(cudaMemset(info, 0, sizeof(int)));
  QB_CUDA_TRY(cudaMemset(clk, 0, sizeof(long long) * 32));
  cudaStream_t s; QB_CUDA_TRY(cudaStreamCreate(&s));
  QB_LAUNCH(k_fill_rand, 1024, 256, 0, s, L, (size_t)n * n, 99ull);
  QB_LAUNCH(k_fill_rand, 256, 256, 0, s, W, (size_t)n * k, 7ull);
  dim3 g(cdiv(n, 256), n);
  QB_LAUNCH(k_make_spd, g, 256, 0, s, n, L);
  int rc = potrf_lower(s, n, L, n, X, info);
  if (!rc) rc = chol_updown_flow(s, n, L, n, W, n, k, k, info);          // warm
  if (!rc) rc = chol_updown_flow(s, n, L, n, W, n, k, 0, info, clk);
  QB_CUDA_TRY(cudaStreamSynchronize(s));
  QB_CUDA_TRY(cudaMemcpy(out32, clk, sizeof(long long) * 32, cudaMemcpyDeviceToHost));
  chol_updown_flow_release(s);
  cudaStreamDestroy(s); cudaFree(L); cudaFree(X); cudaFree(W); cudaFree(info); cudaFree(clk);
  return rc;
}

// ---- batch API: implemented in batch.cu ---------------------------------------------------------------

// FP64 tensor-pipe issue-rate peak: register-resident mma.sync m8n8k4 chains, no memory traffic.
// MEASURED_PEAKS.json has no FP64 entry, so bench.py uses this as the denominator of the tensor roofline.
__global__ void __launch_bounds__(256) k_dmma_peak(int iters, double *sink) {
  double acc[16][2];
#pragma unroll
  for (int i = 0; i < 16; i++) { acc[i][0] = threadIdx.x * 1e-9; acc[i][1] = i * 1e-9; }
  double a = 1.0 + threadIdx.x * 1e-6, b = 1.0 - threadIdx.x * 1e-6;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(acc[i][0]), "+d"(acc[i][1]) : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i][0] + acc[i][1];
  if (s == 123.456) sink[0] = s;
}
extern "C" int qpalm_b200_bench_dmma_peak(double *tflops_out) {
  double *sink = nullptr;
  QB_CUDA_TRY(cudaMalloc(&sink, 8));
  cudaStream_t s; QB_CUDA_TRY(cudaStreamCreate(&s));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 4096, grid = 148 * 8;
  QB_LAUNCH(k_dmma_peak, grid, 256, 0, s, 64, sink);
  double best = 0;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(a, s);
    QB_LAUNCH(k_dmma_peak, grid, 256, 0, s, iters, sink);
    cudaEventRecord(b, s);
    QB_CUDA_TRY(cudaEventSynchronize(b));
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    const double flops = (double)grid * 8 /*warps*/ * iters * 16.0 * 512.0;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  *tflops_out = best;
  cudaEventDestroy(a); cudaEventDestroy(b); cudaStreamDestroy(s); cudaFree(sink);
  return 0;
}

// scalar FP64 FMA issue-rate peak (register-resident DFMA chains, 16 independent accumulators per thread): the bound of every
// kernel that does its arithmetic with plain fma() instead of DMMA.  warps_per_sm: resident warps per SM during the probe.
__global__ void __launch_bounds__(256) k_dfma_peak(int iters, double *sink) {
  double acc[16];
#pragma unroll
  for (int i = 0; i < 16; i++) acc[i] = threadIdx.x * 1e-9 + i * 1e-7;
  const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) s += acc[i];
  if (s == 123.456) sink[0] = s;
}
extern "C" int qpalm_b200_bench_dfma_peak(double *tflops_out) {
  double *sink = nullptr;
  QB_CUDA_TRY(cudaMalloc(&sink, 8));
  cudaStream_t s; QB_CUDA_TRY(cudaStreamCreate(&s));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const int iters = 8192, grid = 148 * 8;
  QB_LAUNCH(k_dfma_peak, grid, 256, 0, s, 64, sink);
  double best = 0;
  for (int r = 0; r < 3; r++) {
    cudaEventRecord(a, s);
    QB_LAUNCH(k_dfma_peak, grid, 256, 0, s, iters, sink);
    cudaEventRecord(b, s);
    QB_CUDA_TRY(cudaEventSynchronize(b));
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    const double flops = (double)grid * 256 * iters * 16.0 * 2.0;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  *tflops_out = best;
  cudaEventDestroy(a); cudaEventDestroy(b); cudaStreamDestroy(s); cudaFree(sink);
  return 0;
}

// HBM roofline probes for the matrix-vector kernels: dense At (n x m) products at the solver's shapes.
extern "C" int qpalm_b200_bench_gemv(c_int n_, c_int m_, c_int reps, double *ms_cols_out, double *ms_rows_out) {
  const int n = (int)n_, m = (int)m_;
  Engine e;   // minimal engine: only what the two dense kernels need
  QB_CUDA_TRY(cudaStreamCreate(&e.stream));
  e.n = n; e.m = m; e.A_dense = true; e.m_loc = m; e.m_cap = m;
  QB_CUDA_TRY(cudaMalloc(&e.At, sizeof(double) * (size_t)n * m));
  QB_CUDA_TRY(cudaMalloc(&e.x, sizeof(double) * n)); QB_CUDA_TRY(cudaMalloc(&e.y, sizeof(double) * m));
  QB_CUDA_TRY(cudaMalloc(&e.Ax, sizeof(double) * m)); QB_CUDA_TRY(cudaMalloc(&e.Aty, sizeof(double) * n));
  const int rowctas = cdiv(n, 128); int splits = cdiv(2048, rowctas); splits = splits < 1 ? 1 : (splits > 64 ? 64 : splits);
  e.gemv_splits = splits;
  QB_CUDA_TRY(cudaMalloc(&e.gemv_partials, sizeof(double) * (size_t)splits * n));
  QB_LAUNCH(k_fill_rand, 2048, 256, 0, e.stream, e.At, (size_t)n * m, 5ull);
  QB_LAUNCH(k_fill_rand, 64, 256, 0, e.stream, e.x, (size_t)n, 6ull);
  QB_LAUNCH(k_fill_rand, 64, 256, 0, e.stream, e.y, (size_t)m, 7ull);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  spmv_A(&e, e.x, e.Ax); spmv_At(&e, e.y, e.Aty);
  float ms = 0;
  cudaEventRecord(a, e.stream);
  for (int r = 0; r < reps; r++) spmv_A(&e, e.x, e.Ax);
  cudaEventRecord(b, e.stream); QB_CUDA_TRY(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms, a, b);
  *ms_cols_out = ms / reps;
  cudaEventRecord(a, e.stream);
  for (int r = 0; r < reps; r++) spmv_At(&e, e.y, e.Aty);
  cudaEventRecord(b, e.stream); QB_CUDA_TRY(cudaEventSynchronize(b)); cudaEventElapsedTime(&ms, a, b);
  *ms_rows_out = ms / reps;
  cudaEventDestroy(a); cudaEventDestroy(b);
  cudaFree(e.At); cudaFree(e.x); cudaFree(e.y); cudaFree(e.Ax); cudaFree(e.Aty); cudaFree(e.gemv_partials);
  chol_solve_release(e.stream);
  cudaStreamDestroy(e.stream);
  return 0;
}
