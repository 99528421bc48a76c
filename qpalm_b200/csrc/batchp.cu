// batchp.cu -- persistent batch engine: ONE CTA owns one QP instance for its whole solve (BASELINE config 4).
//
// The lock-step engine (batch.cu) advances all instances together, one launch per step and one host poll per
// iteration; every instance then pays for the slowest one and for ~30 launches per iteration.  Here a single kernel
// launch solves the whole batch: CTAs pull instance indices from an atomic work queue, run the complete QPALM loop of
// src/qpalm.c:484-711 for that instance (residuals, termination, outer updates, active set, Newton system, exact
// line search) with block-level barriers between the steps, and take the next instance when done.  There is no host
// synchronisation inside a solve and no instance waits for another.
//
// Memory plan (B200): the shared A' (n x m), A (m x n) and D Q D (n x n) are read by every CTA and stay in the 126 MB L2; the
// per-instance H / L (n x n lower, column-major) and vectors are touched by one CTA only.  At 512 instances the factors
// (2 x 230 KB each) do not fit the L2 and stream from HBM, which is why every pass over them is written as a batch of
// requests (cp.async, eight loads in flight, rows requested one block ahead) rather than load -> use.  Shared memory per CTA
// (72 KB at 3 CTAs / SM, 55 KB at 4) holds one panel of the factor (Cholesky, triangular solves, SYRK chunks) -- or, during an
// update sweep, the rows of W, the chain warp's ring of window blocks and the row owners' staging ring --, the radix-sort
// buffers of the line search (2m <= 2048 keys) and the staged GEMV operand.
//
// The Newton system (Q + A_J' Sigma_J A_J + I/gamma) d = -dphi  (src/newton.c:96-118, solver_interface.c:319-519):
//   H is kept per instance and updated incrementally (+ entering / - leaving / +- sigma changes) by an in-CTA SYRK in chunks of
//   one panel width, L <- chol(H + beta I) by an in-CTA right-looking blocked Cholesky (diagonal sub-blocks in one warp's
//   registers, panel solve one row per thread); SYRK and trailing updates run on the FP64 tensor pipe (mma.sync m8n8k4, 16 x 16
//   tiles per warp); then two blocked triangular solves.  When the active set changes by at most min(0.1 (n + m),
//   max_rank_update) rows -- and at most 40, see p_control -- and the factor is current, the factor is UPDATED in place instead
//   (cta_updown_sweep: <= 8 ranks per sweep, entering rows before leaving rows; a chain warp on the diagonal, every other thread
//   the owner of one row, per-column mbarriers), where the reference takes cholmod_updown (newton.c:98-108).  All arithmetic is
//   fp64 FMA / DMMA.
#include "batch.cuh"
#include "chol32.cuh"
#include <math.h>
#include <string.h>

using namespace qb;

// Build-time shape of a CTA; this file is compiled twice (build.sh):
//   default            8 warps, 32-column panel (16-wide sub-panels), both sort buffers in shared memory: 80 registers x 256
//                      threads, 72 KB -> 3 CTAs / SM (444 slots on a B200).  Best per-wave throughput: 444 instances in
//                      57.9 ms (7675 solves/s), 4096 instances in 488 ms (8396 solves/s) at the end of round 2.
//   -DQB_BP_VARIANT4   8 warps, 24-column panel (-DQB_BP_SW=8: 8-wide sub-panels), one sort buffer in shared memory and the other
//                      in the instance's global arrays, registers capped at 64: 55 KB -> 4 CTAs / SM (592 slots).  Slower per
//                      instance (64 registers: several loops are shaped around that, see QB_BP_SW and p_gemv_rows) but the
//                      BASELINE sweep of 512 instances per GPU fits in ONE wave: 78.8 ms against 90.3 ms with the 68-instance
//                      tail of the default shape.  batch.cu picks it when 3 * SMs < nb <= 4 * SMs.
// Other shapes measured in round 1 (512 instances): 192 threads / PW 24 / 4 per SM 138.6 ms; 192 / PW 16 / 4 per SM
// 147.7 ms; 256 / PW 16 / 64 regs / 4 per SM 133.0 ms; 192 threads / PW 32 / 3 per SM 192.5 ms (fewer threads per CTA
// cost more than the extra occupancy buys); 4-per-SM shape with 224 threads (72 registers) 142.8 ms, with 288 threads
// (56 registers) 145.9 ms.  (The update sweep of round 2 maps thread t to row t of L: NT >= n.)
#ifdef QB_BP_VARIANT4
#define QB_BP_NAMESPACE bp4
#define QB_BP_SUPPORTED batchp4_supported
#define QB_BP_SOLVE batchp4_solve
#ifndef QB_BP_NT
#define QB_BP_NT 256
#endif
#ifndef QB_BP_PW
#define QB_BP_PW 24
#endif
#define QB_BP_SORT_GLOBAL 1
#define QB_BP_MINB 4
#else
#define QB_BP_NAMESPACE bp
#define QB_BP_SUPPORTED batchp_supported
#define QB_BP_SOLVE batchp_solve
#endif
#ifndef QB_BP_NT
#define QB_BP_NT 256
#endif
#ifndef QB_BP_PW
#define QB_BP_PW 32
#endif
#ifndef QB_BP_SORT_GLOBAL
#define QB_BP_SORT_GLOBAL 0
#endif
namespace qb {
namespace QB_BP_NAMESPACE {

constexpr int NT = QB_BP_NT, NW = NT / 32;
constexpr int PW = QB_BP_PW;      // panel width (sub-panels of 16 + 8 columns)
constexpr int NMAX = 240;         // largest n of this engine (panel = PW x LDP doubles of shared memory)
constexpr int LDP = NMAX + 1;     // odd leading dimension: conflict-free row and column access
static_assert(NMAX <= NT, "the update sweep maps thread t to row t of the factor");
constexpr int SORT_MAX = 2048;    // largest 2m
constexpr int VS_LEN = (QB_BP_PW == 32) ? 1024 : 960;   // staged GEMV operand (max(n, m) doubles)
constexpr size_t kSortBytes = (size_t)SORT_MAX * (QB_BP_SORT_GLOBAL ? 1 : 2) * (8 + 4) + sizeof(unsigned) * (NW * 256 + 256 + 256);
constexpr size_t kPanelBytes = sizeof(double) * PW * LDP;
constexpr size_t kUnionBytes = kSortBytes > kPanelBytes ? kSortBytes : kPanelBytes;
constexpr size_t kSmemBytes = kUnionBytes + sizeof(double) * (VS_LEN + NMAX + 16 + 2 * PW);

#define BSC(b, slot) P.scal[(size_t)(b) * S_COUNT + (slot)]

struct Args {
  int nb, n, m, ld;
  BSet st;
  const double *At, *Am, *Qs, *D, *Dinv, *E, *Einv;   // Am: A as m x n column-major (A d row-parallel), At: A' (n x m)
  const double *q_raw, *bmin_raw, *bmax_raw;
  double *x_out, *y_out;
  double *q, *bmin, *bmax, *x, *y, *Ax, *Qx, *Aty, *x_prev, *x0, *sigma, *sigma_inv, *sqrt_sigma, *Axys, *z, *pri_res, *pri_res_in,
      *yh, *Atyh, *df, *dphi, *d, *Qd, *Ad;
  int *active, *active_old, *active_cand, *activeH, *list_pos, *list_neg;
  double *sigmaH, *w_pos, *w_neg;
  double *H, *L, *rdiag;
  long long sLL, sR;
  unsigned long long *keys; unsigned int *vals; double *ls_da, *ls_db;
  double *scal;
  BCtl *ctl;
  int *queue;
  long long *prof;   // optional [nb][16] per-phase clock64 totals (QPALM_B200_BATCH_PROF=1)
};

struct Flags { int outer, sigma, inner, refac, factor, fq, boost, done, updown; };

struct Smem {
  unsigned char *u;      // union region: sort buffers / panel
  double *panel;         // PW x LDP
  double *vs;            // staged vector (VS_LEN)
  double *v;             // Newton rhs / solution (NMAX + 16)
  double *rd;            // 1 / l_jj of the current panel (PW), spare (PW)
};

// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
#define RED_OUT(op, val, slot) { const double r_ = block_red<op>(val, scratch); if (threadIdx.x == 0) BSC(b, slot) = r_; }

// K block reductions behind ONE pair of barriers (block_red pays two barriers per value: p_res_m alone had 22).  Stage 1 is
// block_red's warp butterfly; stage 2 replays, in one thread per value, the tree that block_red's second butterfly forms over
// the per-warp partials (lanes >= NW hold the identity there), so every result is bit-identical to block_red's.
constexpr int RED_MULTI_MAX = 12;
__device__ __forceinline__ double red_dyn(int op, double a, double b) { return op == RED_SUM ? a + b : (op == RED_MAX ? fmax(a, b) : fmin(a, b)); }
template <int K>
__device__ __forceinline__ void block_red_multi(const Args &P, int b, const double (&val)[K], const int (&op)[K], const int (&slot)[K], double *part_buf) {
  static_assert(K <= RED_MULTI_MAX && NW <= 8 && RED_MULTI_MAX * 8 <= VS_LEN, "partials live in the staged-vector buffer (free between phases)");
  double (*part)[8] = reinterpret_cast<double (*)[8]>(part_buf);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();   // the previous user of part[] is done
#pragma unroll
  for (int k = 0; k < K; k++) {
    double x = val[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = red_dyn(op[k], x, __shfl_xor_sync(0xffffffffu, x, o));
    if (lane == 0) part[k][wid] = x;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; k++) {
    if (threadIdx.x == (k % NW) * 32 + k / NW) {
      double a[8];
#pragma unroll
      for (int w = 0; w < 8; w++) a[w] = (w < NW) ? part[k][w] : (op[k] == RED_SUM ? 0.0 : (op[k] == RED_MAX ? -1.0e300 : 1.0e300));
#pragma unroll
      for (int o = 4; o > 0; o >>= 1)
#pragma unroll
        for (int i = 0; i < o; i++) a[i] = red_dyn(op[k], a[i], a[i + o]);
      BSC(b, slot[k]) = a[0];
    }
  }
}

__device__ __noinline__ void p_init(const Args &P, int b, double *scratch) {
  __shared__ double cc_s, sig_s;
  const int n = P.n, m = P.m, tid = threadIdx.x;
  const BSet &st = P.st;
  const size_t on = (size_t)b * n, om = (size_t)b * m;
  double mx = 0.0;
  for (int i = tid; i < n; i += NT) {   // q <- c * (D .* q), c = 1 / max(1, |D q|inf)   (scaling.c:82-90, Qx = 0 at setup)
    double v = P.q_raw[on + i];
    if (st.scaling) v = P.D[i] * v;
    P.q[on + i] = v;
    mx = fmax(mx, fabs(v));
  }
  mx = block_red<RED_MAX>(mx, scratch);
  if (tid == 0) cc_s = st.scaling ? 1 / fmax(1.0, mx) : 1.0;
  __syncthreads();
  const double cc = cc_s;
  for (int i = tid; i < n; i += NT) {
    if (st.scaling) P.q[on + i] *= cc;
    P.x[on + i] = 0; P.Qx[on + i] = 0; P.Aty[on + i] = 0; P.x_prev[on + i] = 0; P.x0[on + i] = 0; P.Qd[on + i] = 0; P.d[on + i] = 0;
  }
  double dist2 = 0.0;
  for (int i = tid; i < m; i += NT) {
    double lo = P.bmin_raw[om + i], hi = P.bmax_raw[om + i];
    if (st.scaling) { lo = P.E[i] * lo; hi = P.E[i] * hi; }
    P.bmin[om + i] = lo; P.bmax[om + i] = hi;
    P.y[om + i] = 0; P.Ax[om + i] = 0; P.Ad[om + i] = 0; P.pri_res_in[om + i] = 0;
    P.active[om + i] = 0; P.active_old[om + i] = 0; P.activeH[om + i] = 0;
    const double t = 0.0 - fmax(lo, fmin(0.0, hi));
    dist2 += t * t;
  }
  dist2 = block_red<RED_SUM>(dist2, scratch);
  if (tid == 0) {
    double s = st.sigma_init * 1.0 / fmax(1.0, 0.5 * dist2);   // f = 0 at x = 0 (iteration.c:50-58)
    s = fmax(1e-4, fmin(s, 1e4));
    sig_s = s;
    BCtl c;
    memset(&c, 0, sizeof(c));
    c.reset_newton = 1; c.gamma = st.gamma_init; c.gamma_prev = st.gamma_init; c.eps_abs_in = st.eps_abs_in; c.eps_rel_in = st.eps_rel_in;
    c.c = cc; c.cinv = 1.0 / cc; c.status = QPALM_UNSOLVED;
    P.ctl[b] = c;
    for (int k = 0; k < S_COUNT; k++) BSC(b, k) = 0.0;
  }
  __syncthreads();
  const double s = sig_s;
  for (int i = tid; i < m; i += NT) { P.sigma[om + i] = s; P.sigma_inv[om + i] = 1.0 / s; P.sqrt_sigma[om + i] = sqrt(s); }
}

// compute_residuals (iteration.c:24-48) + candidate active set (newton.c:122-149) + m-side termination reductions
__device__ __noinline__ void p_res_m(const Args &P, int b, double *part) {
  const int m = P.m, scaling = P.st.scaling;
  const size_t om = (size_t)b * m;
  double r_pri = 0, r_raw = 0, r_ax = 0, r_z = 0, r_edy = 0, oob = 0, adx_max = -1.0e300, adx_min = 1.0e300;
  double n_act = 0, n_ent = 0, n_lea = 0;
  for (int i = threadIdx.x; i < m; i += NT) {
    const double ax = P.Ax[om + i], yi = P.y[om + i], lo = P.bmin[om + i], hi = P.bmax[om + i];
    double t = yi * P.sigma_inv[om + i];
    const double axys = ax + t;
    const double zi = fmax(lo, fmin(axys, hi));
    const double pr = ax - zi;
    t = pr * P.sigma[om + i];
    const double yhi = yi + t;
    P.Axys[om + i] = axys; P.z[om + i] = zi; P.pri_res[om + i] = pr; P.yh[om + i] = yhi;
    const int act = (axys <= lo) || (axys >= hi);
    const int old = P.active_old[om + i];
    P.active_cand[om + i] = act;
    n_act += act; n_ent += (act && !old); n_lea += (!act && old);
    const double ei = scaling ? P.E[i] : 1.0, einv = scaling ? P.Einv[i] : 1.0;
    r_pri = fmax(r_pri, fabs(einv * pr)); r_raw = fmax(r_raw, fabs(pr));
    r_ax = fmax(r_ax, fabs(einv * ax)); r_z = fmax(r_z, fabs(einv * zi));
    const double dy = yhi - yi;
    r_edy = fmax(r_edy, fabs(ei * dy));
    const bool hi_fin = hi < ei * kInf, lo_fin = lo > -ei * kInf;
    oob += hi_fin ? hi * fmax(dy, 0.0) : 0.0;
    oob += lo_fin ? lo * fmin(dy, 0.0) : 0.0;
    const double adx = einv * P.Ad[om + i];
    if (hi_fin) adx_max = fmax(adx_max, adx);
    if (lo_fin) adx_min = fmin(adx_min, adx);
  }
  const double vals[11] = {r_pri, r_raw, r_ax, r_z, r_edy, oob, adx_max, adx_min, n_act, n_ent, n_lea};
  const int ops[11] = {RED_MAX, RED_MAX, RED_MAX, RED_MAX, RED_MAX, RED_SUM, RED_MAX, RED_MIN, RED_SUM, RED_SUM, RED_SUM};
  const int slots[11] = {S_PRI_RES, S_PRI_RES_RAW, S_NORM_AX, S_NORM_Z, S_NORM_EDY, S_OOB, S_ADX_MAX, S_ADX_MIN, S_NB_ACTIVE, S_NB_ENTER, S_NB_LEAVE};
  block_red_multi<11>(P, b, vals, ops, slots, part);
}

// out[i] = scale * sum_k M[i + ld*k] v[k], i < nrows (row sums of a shared column-major matrix: A'yh, Q d and, with the
// m x n copy of A, A d).  Each thread owns up to 4 rows (i, i + NT, ...) and walks the columns with 4-way unrolling, so
// 16 independent L2 requests are in flight per thread; per-row summation order is fixed (reproducible).
template <int RP, int KU>
__device__ __forceinline__ void gemv_rows_body(int nrows, int ncols, int ld, const double *__restrict__ M, double *out, double scale,
                                               const double *vs) {
  const int tid = threadIdx.x;
  for (int i0 = tid; i0 < nrows; i0 += RP * NT) {
    int ri[RP];
#pragma unroll
    for (int a = 0; a < RP; a++) { ri[a] = i0 + a * NT; if (ri[a] >= nrows) ri[a] = nrows - 1; }
    double acc[RP][KU];
#pragma unroll
    for (int a = 0; a < RP; a++)
#pragma unroll
      for (int u = 0; u < KU; u++) acc[a][u] = 0.0;
    int k = 0;
    for (; k + KU - 1 < ncols; k += KU) {
#pragma unroll
      for (int u = 0; u < KU; u++) {
        const double vk = vs[k + u];
        const double *col = M + (size_t)(k + u) * ld;
#pragma unroll
        for (int a = 0; a < RP; a++) acc[a][u] = fma(col[ri[a]], vk, acc[a][u]);
      }
    }
    for (; k < ncols; k++) {
      const double vk = vs[k];
      const double *col = M + (size_t)k * ld;
#pragma unroll
      for (int a = 0; a < RP; a++) acc[a][0] = fma(col[ri[a]], vk, acc[a][0]);
    }
#pragma unroll
    for (int a = 0; a < RP; a++) {
      double t = 0.0;
#pragma unroll
      for (int u = KU - 1; u >= 0; u--) t += acc[a][u];
      if (i0 + a * NT < nrows) out[i0 + a * NT] = t * scale;
    }
  }
}
// The same product for an EVEN number of rows <= NT with an even leading dimension (A' yh and Q d: n rows).  A thread owns a PAIR
// of rows (one 16-byte load per column) and one of G = NT / (nrows / 2) interleaved column groups, so all NT threads have eight
// 16-byte requests in flight (the one-row-per-thread form had 240 threads x 8 x 8 bytes: the kernel's hottest line, waiting on
// the L2, profiles/r02g_ncu_source_kbp_solve.txt).  Four accumulators per row, the groups' partial sums are added in group order:
// fixed summation order, reproducible.  `part` holds G x nrows doubles of shared memory.
__device__ __forceinline__ void gemv_rowpairs_body(int nrows, int ncols, int ld, const double *__restrict__ M, double *out, double scale,
                                                   const double *vs, double *part) {
  const int tid = threadIdx.x, np = nrows >> 1, G = NT / np;
  const int g = tid / np, p = tid - g * np;
  if (g < G) {
    double2 acc[4];
#pragma unroll
    for (int u = 0; u < 4; u++) acc[u] = make_double2(0.0, 0.0);
    const double *Mp = M + 2 * p;
    int k0 = g * 8;
    for (; k0 + 7 < ncols; k0 += G * 8) {
      double2 a[8];
#pragma unroll
      for (int u = 0; u < 8; u++) a[u] = *reinterpret_cast<const double2 *>(Mp + (size_t)(k0 + u) * ld);
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const double vk = vs[k0 + u];
        acc[u & 3].x = fma(a[u].x, vk, acc[u & 3].x);
        acc[u & 3].y = fma(a[u].y, vk, acc[u & 3].y);
      }
    }
    for (int k = k0; k < ncols && k < k0 + 8; k++) {   // this group's ragged last block
      const double2 a = *reinterpret_cast<const double2 *>(Mp + (size_t)k * ld);
      const double vk = vs[k];
      acc[0].x = fma(a.x, vk, acc[0].x);
      acc[0].y = fma(a.y, vk, acc[0].y);
    }
    part[g * nrows + 2 * p] = (acc[3].x + acc[2].x) + (acc[1].x + acc[0].x);
    part[g * nrows + 2 * p + 1] = (acc[3].y + acc[2].y) + (acc[1].y + acc[0].y);
  }
  __syncthreads();
  for (int i = tid; i < nrows; i += NT) {
    double t = 0.0;
    for (int gg = G - 1; gg >= 0; gg--) t += part[gg * nrows + i];
    out[i] = t * scale;
  }
}
__device__ __forceinline__ void p_gemv_rows(int nrows, int ncols, int ld, const double *__restrict__ M, const double *__restrict__ v,
                                            double *out, double scale, const Smem &S) {
  for (int k = threadIdx.x; k < ncols; k += NT) S.vs[k] = v[k];
  __syncthreads();
  if (!(nrows & 1) && !(ld & 1) && nrows >= 2 && (nrows >> 1) <= NT)
    gemv_rowpairs_body(nrows, ncols, ld, M, out, scale, S.vs, S.panel);               // row pairs x column groups (A' yh, Q d)
  else if (nrows <= NT) gemv_rows_body<1, 8>(nrows, ncols, ld, M, out, scale, S.vs);    // one row per thread, 8 columns in flight
#ifdef QB_BP_VARIANT4
  else gemv_rows_body<4, 2>(nrows, ncols, ld, M, out, scale, S.vs);               // 64 registers: 4 x 4 accumulators + 16 loads in flight spill in the loop
#else
  else gemv_rows_body<4, 4>(nrows, ncols, ld, M, out, scale, S.vs);               // 4 rows x 4 columns in flight
#endif
}
// out[k] = sum_i M[i + ld*k] v[i], k < ncols (column dots: A d with A' as M); one warp per column
__device__ __forceinline__ void p_gemv_cols(int len, int ncols, int ld, const double *__restrict__ M, const double *__restrict__ v, double *out,
                            const Smem &S) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int k = tid; k < len; k += NT) S.vs[k] = v[k];
  __syncthreads();
  // four columns per warp step: their loads are independent, so 4 x len/32 L2 requests are in flight per lane
  for (int col0 = warp * 4; col0 < ncols; col0 += NW * 4) {
    double acc[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int col = (col0 + u < ncols) ? col0 + u : ncols - 1;
      const double *c = M + (size_t)col * ld;
      double a0 = 0, a1 = 0;
      int i = lane;
      for (; i + 32 < len; i += 64) { a0 = fma(c[i], S.vs[i], a0); a1 = fma(c[i + 32], S.vs[i + 32], a1); }
      for (; i < len; i += 32) a0 = fma(c[i], S.vs[i], a0);
      acc[u] = a0 + a1;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const double r = warp_sum(acc[u]);
      if (lane == 0 && col0 + u < ncols) out[col0 + u] = r;
    }
  }
}

__device__ __noinline__ void p_res_n(const Args &P, int b, double *part) {
  const int n = P.n;
  const BSet &st = P.st;
  const size_t on = (size_t)b * n;
  const double gamma = P.ctl[b].gamma, neg_inv_gamma = -1 / gamma;
  const double neg_tau_over_gamma = -BSC(b, S_TAU) * (1 / gamma);
  double r_dua = 0, r_dua2 = 0, r_qx = 0, r_q = 0, r_atyh = 0, r_atdy = 0, r_ddx = 0, dxdx = 0, dxqdx = 0, qdx = 0;
  for (int j = threadIdx.x; j < n; j += NT) {
    const double qx = P.Qx[on + j], qj = P.q[on + j], xj = P.x[on + j], at = P.Atyh[on + j];
    double dfj = qx + qj;
    if (st.proximal) dfj = dfj + neg_inv_gamma * P.x0[on + j];
    const double dp = dfj + at;
    P.df[on + j] = dfj; P.dphi[on + j] = dp;
    const double dinv = st.scaling ? P.Dinv[j] : 1.0;
    if (st.proximal) {
      const double xx0 = xj - P.x0[on + j];
      const double t = dp + neg_inv_gamma * xx0;
      r_dua = fmax(r_dua, fabs(dinv * t));
    } else r_dua = fmax(r_dua, fabs(dinv * dp));
    r_dua2 = fmax(r_dua2, fabs(dinv * dp));
    r_qx = fmax(r_qx, fabs(dinv * qx)); r_q = fmax(r_q, fabs(dinv * qj)); r_atyh = fmax(r_atyh, fabs(dinv * at));
    r_atdy = fmax(r_atdy, fabs(dinv * (at - P.Aty[on + j])));
    const double dx = xj - P.x_prev[on + j];
    const double ddx = st.scaling ? P.D[j] * dx : dx;
    r_ddx = fmax(r_ddx, fabs(ddx));
    dxdx += ddx * ddx;
    if (st.proximal) { const double t2 = P.Qd[on + j] + neg_tau_over_gamma * P.d[on + j]; dxqdx += dx * t2; }
    else dxqdx += P.Qd[on + j] * dx;
    qdx += qj * dx;
  }
  const double vals[10] = {r_dua, r_dua2, r_qx, r_q, r_atyh, r_atdy, r_ddx, dxdx, dxqdx, qdx};
  const int ops[10] = {RED_MAX, RED_MAX, RED_MAX, RED_MAX, RED_MAX, RED_MAX, RED_MAX, RED_SUM, RED_SUM, RED_SUM};
  const int slots[10] = {S_DUA_RES, S_DUA2_RES, S_NORM_QX, S_NORM_Q, S_NORM_ATYH, S_NORM_ATDY, S_NORM_DDX, S_DXDX, S_DXQDX, S_QDX};
  block_red_multi<10>(P, b, vals, ops, slots, part);
}

// the control flow of qpalm_solve for one iteration (src/qpalm.c:484-711), executed by one thread
__device__ __noinline__ void p_control(const Args &P, int b, Flags &f) {
  const BSet &st = P.st;
  const int n = P.n, m = P.m;
  f.outer = f.sigma = f.inner = f.refac = f.factor = f.fq = f.boost = f.done = f.updown = 0;
  BCtl c = P.ctl[b];
  const double *h = P.scal + (size_t)b * S_COUNT;
  const double cinv = st.scaling ? c.cinv : 1.0;
  c.pri_res_norm = h[S_PRI_RES];
  c.dua_res_norm = h[S_DUA_RES] * cinv;
  c.dua2_res_norm = h[S_DUA2_RES] * cinv;
  const double nrm_axz = st.scaling ? h[S_NORM_AX] : fmax(h[S_NORM_AX], h[S_NORM_Z]);   // sic, termination.c:99
  c.eps_pri = st.eps_abs + st.eps_rel * nrm_axz;
  double max_norm = fmax(h[S_NORM_QX], fmax(h[S_NORM_Q], h[S_NORM_ATYH]));
  if (st.scaling) max_norm *= cinv;
  c.eps_dua = st.eps_abs + st.eps_rel * max_norm;
  c.eps_dua_in = c.eps_abs_in + c.eps_rel_in * max_norm;
  int term = 0;
  if ((c.pri_res_norm < c.eps_pri) && (c.dua_res_norm < c.eps_dua)) term = QPALM_SOLVED;
  else {
    const double eps_pinf = st.eps_prim_inf * h[S_NORM_EDY];
    if ((eps_pinf != 0) && (h[S_NORM_ATDY] <= eps_pinf) && (h[S_OOB] <= -eps_pinf)) term = QPALM_PRIMAL_INFEASIBLE;
    else {
      const double eps_dinf = st.eps_dual_inf * h[S_NORM_DDX];
      if (eps_dinf != 0) {
        const bool blocked = (m > 0) && ((h[S_ADX_MAX] >= eps_dinf) || (h[S_ADX_MIN] <= -eps_dinf));
        if (!blocked) {
          const double cc = st.scaling ? c.c : 1.0, e2 = st.eps_dual_inf * st.eps_dual_inf;
          if ((h[S_DXQDX] <= -cc * e2 * h[S_DXDX]) || ((h[S_DXQDX] <= cc * e2 * h[S_DXDX]) && (h[S_QDX] <= -cc * eps_dinf)))
            term = QPALM_DUAL_INFEASIBLE;
        }
      }
    }
  }
  if (term) { c.status = term; c.done = 1; P.ctl[b] = c; f.done = 1; return; }
  const int na = (m > 0) ? (int)h[S_NB_ACTIVE] : 0, ne = (m > 0) ? (int)h[S_NB_ENTER] : 0, nl = (m > 0) ? (int)h[S_NB_LEAVE] : 0;
  if ((c.dua2_res_norm <= c.eps_dua_in) || (c.no_change == 3)) {   // qpalm.c:515
    c.no_change = 0;
    f.outer = 1;
    if (c.iter_out > 0 && c.pri_res_norm > c.eps_pri) f.sigma = 1;
    c.eps_abs_in = fmax(st.eps_abs, st.rho * c.eps_abs_in);
    c.eps_rel_in = fmax(st.eps_rel, st.rho * c.eps_rel_in);
    c.gamma_prev = c.gamma;
    if (st.proximal) {
      const bool try_boost = !c.gamma_maxed && c.iter_out > 0 && c.nb_enter == 0 && c.nb_leave == 0 && c.pri_res_norm < c.eps_pri;
      if (try_boost) f.boost = 1;
      else if (c.gamma < st.gamma_max) { c.gamma = fmin(c.gamma * st.gamma_upd, st.gamma_max); c.reset_newton = 1; }
    }
    c.iter_out++; c.prev_iter = c.iter;
  } else if (c.iter == c.prev_iter + st.inner_max_iter) {   // qpalm.c:647-660
    c.no_change = 0;
    f.outer = 2;
    if (c.iter_out > 0 && c.pri_res_norm > c.eps_pri) f.sigma = 1;
    c.gamma_prev = c.gamma;
    if (st.proximal && c.gamma < st.gamma_max) { c.gamma = fmin(c.gamma * st.gamma_upd, st.gamma_max); c.reset_newton = 1; }
    c.iter_out++; c.prev_iter = c.iter;
  } else {   // inner step
    if (c.nb_enter + c.nb_leave) c.no_change = 0; else c.no_change++;
    if ((c.iter % st.reset_newton_iter) == 0) c.reset_newton = 1;
    c.nb_active = na; c.nb_enter = ne; c.nb_leave = nl;
    f.inner = 1;
    c.beta = st.proximal ? 1.0 / c.gamma : 0.0;
    const double rank_limit = fmin(st.max_rank_update_fraction * (double)(n + m), (double)st.max_rank_update);
    c.scratch = 0;
    if ((c.reset_newton && na) || (double)(ne + nl) > rank_limit) { f.refac = 1; f.factor = 1; c.scratch = (c.reset_newton && !st.batch_h_incremental) || !c.H_valid; }
    else if (na) {   // newton.c:103-108: rank update of the factor (entering rows, then leaving rows)
      if (ne + nl > 0) {
        // The reference updates the factor here whatever the rank (<= min(0.1 (n + m), 160) rows).  A sweep handles 8 ranks, so
        // beyond ~5 sweeps the incremental SYRK + refactorisation is the cheaper way to the SAME matrix: 512-instance sweep
        // 91.5 ms without the cap, 88.0 ms at 40 (88.8 / 88.8 / 88.2 / 90.6 ms at 24 / 32 / 48 / 56), iteration counts unchanged.
        if (st.batch_updown && ne + nl <= st.batch_updown_max_rank) f.updown = 1;
        else { f.refac = 1; f.factor = 1; c.scratch = !c.H_valid; }
      }
    }
    else { f.fq = 1; f.factor = 1; }
    c.reset_newton = 0;
    c.n_inner++;
    if (f.factor) { c.n_refac++; c.refac_J += na; }
  }
  P.ctl[b] = c;
}

// update_sigma (iteration.c:86-145); every sigma change leads to a refactorisation (same matrix as the reference's
// rank update of solver_interface.c:443-503)
__device__ __noinline__ void p_update_sigma(const Args &P, int b, double *scratch) {
  const int m = P.m;
  const BSet &st = P.st;
  const size_t om = (size_t)b * m;
  const double nrm = BSC(b, S_PRI_RES_RAW);
  double changed = 0;
  for (int k = threadIdx.x; k < m; k += NT) {
    const double pr = fabs(P.pri_res[om + k]);
    if ((pr > st.theta * fabs(P.pri_res_in[om + k])) && P.active[om + k]) {
      double mult = fmax(1.0, st.delta * pr / (nrm + 1e-6));
      const double sg = P.sigma[om + k], stmp = mult * sg;
      if (stmp <= st.sigma_max) {
        changed += (sg != stmp);
        P.sigma[om + k] = stmp; P.sigma_inv[om + k] = 1.0 / stmp;
        mult = sqrt(mult);
        P.sqrt_sigma[om + k] = mult * P.sqrt_sigma[om + k];
      } else {
        changed += (sg != st.sigma_max);
        P.sigma[om + k] = st.sigma_max; P.sigma_inv[om + k] = 1.0 / st.sigma_max; P.sqrt_sigma[om + k] = st.sqrt_sigma_max;
      }
    }
  }
  changed = block_red<RED_SUM>(changed, scratch);
  if (threadIdx.x == 0) {
    if ((st.proximal && P.ctl[b].gamma_prev < st.gamma_max) || changed > 0) P.ctl[b].reset_newton = 1;
  }
}

// y <- yh, Aty <- Atyh, Qx += (1/gamma - 1/gamma_prev) x, x0 <- x, pri_res_in <- pri_res (qpalm.c:525-526, 629, 635, 655, 658)
__device__ void p_outer(const Args &P, int b, int kind) {
  const int n = P.n, m = P.m;
  const size_t on = (size_t)b * n, om = (size_t)b * m;
  const double g = P.ctl[b].gamma, gp = P.ctl[b].gamma_prev;
  const double dg = (g != gp) ? (1 / g - 1 / gp) : 0.0;
  for (int i = threadIdx.x; i < m; i += NT) {
    if (kind == 1) P.y[om + i] = P.yh[om + i];
    P.pri_res_in[om + i] = P.pri_res[om + i];
  }
  for (int j = threadIdx.x; j < n; j += NT) {
    if (kind == 1) P.Aty[on + j] = P.Atyh[on + j];
    if (P.st.proximal) {
      if (dg != 0.0) P.Qx[on + j] = P.Qx[on + j] + dg * P.x[on + j];
      P.x0[on + j] = P.x[on + j];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// dense building blocks on an n x n lower matrix in global memory (L2 resident), panels in shared memory
// ------------------------------------------------------------------------------------------------
// dst(i, j) = [first ? sscale * src(i, j) + (i == j) * diag_add : dst(i, j)] + sign * sum_{c < w} U[c][i - u0] U[c][j - u0]
// for r0 <= j <= i < n.  U is PW x LDP in shared memory.  Warp task = 128 rows x 4 columns, 4 x 4 register block per lane.
// The accumulators START from the old values: all 16 global loads of a task are issued up front and their latency is
// paid once per task (a read-modify-write after the loop serialises 16 dependent L2 round trips: measured 6x slower).
// FP64 tensor-core tile of the in-CTA rank-w updates: d += a (8 x 4) * b (4 x 8), mma.sync m8n8k4 (DMMA)
__device__ __forceinline__ void bp_dmma884(double &d0, double &d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// dst(i, j) = base(i, j) + sign * sum_{c < w} U[c][i] U[c][j]   for r0 <= j <= i < n   (the SYRK of the H record and the trailing
// update of the in-CTA Cholesky), base = first ? sscale * src + diag_add I : dst.  U = shared panel, column c at U + c * LDP - u0.
// Round 2: the products run on the FP64 tensor pipe.  Each warp takes 16 x 16 tiles of the lower triangle (2 x 2 DMMA tiles,
// round-robin over the warps); per 4 columns of U it loads two A and two B fragments from shared memory and issues four
// mma.sync.m8n8k4 -- a quarter of the shared-memory wavefronts and a sixth of the instructions of the 4 x 4 scalar register
// tile this replaces (which ran the FP64 pipe at 18 %).  The sign rides on the final add.
__device__ __forceinline__ void cta_rank_update_lower(double *dst, int ldd, const double *src, int lds, double sscale, double diag_add, bool first,
                                      int r0, int n, const double *U, int u0, int w, double sign) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rem = n - r0;
  if (rem <= 0) return;
  const int T = (rem + 15) >> 4, ntiles = T * (T + 1) / 2;
  const int fr = lane >> 2, fk = lane & 3;
  const int wpad = (w + 3) & ~3;
  for (int q = warp; q < ntiles; q += NW) {
    int ib = (int)((sqrtf(8.0f * (float)q + 1.0f) - 1.0f) * 0.5f);
    while (ib * (ib + 1) / 2 > q) ib--;
    while ((ib + 1) * (ib + 2) / 2 <= q) ib++;
    const int jb = q - ib * (ib + 1) / 2;
    const int i0 = r0 + 16 * ib, j0 = r0 + 16 * jb;
    int ra[2], rb[2];   // fragment rows (clamped: rows beyond n read a valid address and are never stored)
#pragma unroll
    for (int t = 0; t < 2; t++) {
      ra[t] = i0 + 8 * t + fr; if (ra[t] >= n) ra[t] = n - 1;
      rb[t] = j0 + 8 * t + fr; if (rb[t] >= n) rb[t] = n - 1;
    }
    // the old values of the tile are requested BEFORE the product loop, so their L2 / DRAM latency hides under it (a
    // read-modify-write after the loop left every warp on the long scoreboard: 15 % of the kernel's stall samples in
    // profiles/r02f_ncu_source_kbp_solve.txt); the arithmetic is unchanged (base + sign * sum)
    double acc[2][2][2], base[2][2][2];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
      for (int e = 0; e < 2; e++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
          acc[a][e][h] = 0.0;
          const int i = i0 + 8 * a + fr, j = j0 + 8 * e + 2 * fk + h;
          double bv = 0.0;
          if (i < n && j < n && i >= j) {
            if (first) { bv = sscale * src[(size_t)i + (size_t)lds * j]; if (i == j) bv += diag_add; }
            else bv = dst[(size_t)i + (size_t)ldd * j];
          }
          base[a][e][h] = bv;
        }
    // fragment addresses once per tile; the callers zero the panel columns [w, wpad), so the product loop has no predicate and no
    // branch (the first version selected `c < w ? load : 0` per fragment: ~60 integer / branch instructions per four DMMAs)
    // 32-bit shared-window addresses (four registers instead of four 64-bit generic pointers: at 64 registers per thread the
    // pointers were spilled inside this loop)
    const unsigned ub = (unsigned)__cvta_generic_to_shared(U + fk * LDP - u0);
    const unsigned sa0 = ub + 8u * (unsigned)ra[0], sa1 = ub + 8u * (unsigned)ra[1], sb0 = ub + 8u * (unsigned)rb[0], sb1 = ub + 8u * (unsigned)rb[1];
#pragma unroll 3
    for (int kc = 0; kc < wpad; kc += 4) {
      const unsigned o = (unsigned)(kc * LDP * 8);
      double fa0, fa1, fb0, fb1;
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(fa0) : "r"(sa0 + o));
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(fa1) : "r"(sa1 + o));
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(fb0) : "r"(sb0 + o));
      asm volatile("ld.shared.f64 %0, [%1];" : "=d"(fb1) : "r"(sb1 + o));
      bp_dmma884(acc[0][0][0], acc[0][0][1], fa0, fb0);
      bp_dmma884(acc[0][1][0], acc[0][1][1], fa0, fb1);
      bp_dmma884(acc[1][0][0], acc[1][0][1], fa1, fb0);
      bp_dmma884(acc[1][1][0], acc[1][1][1], fa1, fb1);
    }
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
      for (int e = 0; e < 2; e++)
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const int i = i0 + 8 * a + fr, j = j0 + 8 * e + 2 * fk + h;
          if (i < n && j < n && i >= j) dst[(size_t)i + (size_t)ldd * j] = fma(sign, acc[a][e][h], base[a][e][h]);
        }
  }
}

// cp.async (LDGSTS) 8-byte copy global -> shared.  A panel load requests ALL its elements before anything waits and holds
// no registers; the former `Pn[..] = L[..]` loop kept one or two loads in flight per thread and sat on the long scoreboard
// (the per-instance factors do not fit the L2: 15 % of the kernel's stall samples, profiles/r02f_ncu_source_kbp_solve.txt).
__device__ __forceinline__ void bp_cp_async8(double *smem_dst, const double *gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void bp_cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
// Pn[t][r] <- L(k0 + r, k0 + t) for t <= r < rows, t < w; zeros above the diagonal of the w x w block.  Rows fastest (coalesced).
// The caller's __syncthreads() publishes the panel.
__device__ __forceinline__ void panel_load_async(double *Pn, const double *L, int ld, int k0, int w, int rows) {
  for (int t = 0; t < w; t++) {
    const double *col = L + (size_t)k0 + (size_t)ld * (k0 + t);
    for (int r = threadIdx.x; r < rows; r += NT) {
      if (r >= t) bp_cp_async8(&Pn[t * LDP + r], col + r);
      else Pn[t * LDP + r] = 0.0;
    }
  }
  bp_cp_async_wait_all();
}

// The PW = 32 column panel is factorised as two 16-column sub-panels so that every per-thread register array is 16
// doubles (a 32-wide version needs 64 registers per array and spills at 3 CTAs / SM).
#ifndef QB_BP_SW
#define QB_BP_SW 16
#endif
constexpr int SW = QB_BP_SW;          // sub-panel width of the in-CTA Cholesky (the diagonal block one warp factors in registers)
static_assert(PW % SW == 0 || PW == 24, "panel = whole sub-panels (24 = 16 + 8 in the 16-wide build)");
constexpr int UPW = PW - SW;          // columns of the panel to the right of a sub-panel (at most)

// Cholesky of the 16 x 16 block at (c0, c0) of the panel (P[t][r], t = column, r = row), warp 0; lane = row, the row
// lives in registers.  The column loop is ROLLED around a rotating register file: a[0] is always the current column and
// every step shifts the row one register to the left while it applies the rank-1 update, so all register indices are
// static although the loop is not unrolled.  The body is ~100 instructions and stays in the instruction cache.  (A fully
// unrolled version is ~1k straight-line instructions executed once per call; with three CTAs per SM in different phases
// of this large kernel it ran at instruction-fetch-miss speed: 27 us per block in situ, phase profile of round 1.)  The
// arithmetic and its order are those of the unrolled form, so results are bit-identical to it.
// Columns/rows >= w (ragged last panel) are skipped.  rd[c0 + j] = 1 / l_jj.
__device__ __noinline__ void warp_factor_diag16(double *Pn, double *rd, double *colbuf, int c0, int w, int *info) {
  const int lane = threadIdx.x & 31;
  const int row = c0 + lane;
  const int wend = (w - c0 < SW) ? w - c0 : SW;
  const bool live = lane < wend;
  double a[SW];
#pragma unroll
  for (int c = 0; c < SW; c++) a[c] = (live && c < wend) ? ((c <= lane) ? Pn[(c0 + c) * LDP + row] : 0.0) : ((c == lane) ? 1.0 : 0.0);
  bool bad = false;
  __syncwarp();
#pragma unroll 1
  for (int j = 0; j < SW; j++) {
    const double pjj = __shfl_sync(0xffffffffu, a[0], j);
    if (!(pjj > 0.0)) bad = true;
    // branch-free sqrt / reciprocal (chol32.cuh): rsqrt() carries a slow-path CALL, and the values live across that call site were
    // spilled and reloaded from local memory in every column (five dependent LDL per column under a thrashed 30 KB L1)
    double ljj, inv;
    chol32::sqrt_and_rcp(pjj, ljj, inv);
    const double a0 = (lane == j) ? ljj : a[0] * inv;     // lanes < j: upper triangle, never used
    if (lane == j) rd[c0 + j] = inv;
    if (live && lane >= j && j < wend) Pn[(c0 + j) * LDP + row] = a0;
    // column j is broadcast through shared memory (two alternating 32-entry buffers, one __syncwarp per column): the 15 loads are
    // independent and issue back to back, where 15 64-bit shuffles each exposed their latency before the FMA that consumes them
    // (the loop ran at ~1270 clocks per column on an otherwise idle SM)
    double *cb = colbuf + ((j & 1) << 5);
    cb[lane] = a0;
    __syncwarp();
#pragma unroll
    for (int c = 1; c < SW; c++) {
      const double lcj = cb[(j + c) & 31];
      a[c - 1] = (lane >= j + c) ? fma(-a0, lcj, a[c]) : a[c];   // column j + c moves to register c - 1
    }
    a[SW - 1] = 0.0;
  }
  if (bad && lane == 0 && info) *info = 1;
}

// rows below the sub-panel's diagonal block: P[c0 + c][r] <- forward substitution against the factor at (c0, c0)
__device__ __forceinline__ void panel_solve16(double *Pn, const double *rd, int c0, int w, int rows) {
  const int wsub = (w - c0 < SW) ? w - c0 : SW;       // the second sub-panel of a 24-column panel has 8 columns
  for (int r = c0 + wsub + threadIdx.x; r < rows; r += NT) {
    double v[SW];
#pragma unroll
    for (int c = 0; c < SW; c++) v[c] = (c0 + c < w) ? Pn[(c0 + c) * LDP + r] : 0.0;
#pragma unroll
    for (int c = 0; c < SW; c++) {
      if (c0 + c < w) {
        double s = v[c];
#pragma unroll
        for (int t = 0; t < SW; t++) if (t < c) s = fma(-v[t], Pn[(c0 + t) * LDP + c0 + c], s);
        v[c] = s * rd[c0 + c];
      }
    }
#pragma unroll
    for (int c = 0; c < SW; c++) if (c0 + c < w) Pn[(c0 + c) * LDP + r] = v[c];
  }
}

// after sub-panel [c0, c0 + SW): the panel's columns to its right, rows r >= c0 + SW:
//   P[c0 + SW + c][r] -= sum_{t < SW} P[c0 + t][r] P[c0 + t][c0 + SW + c]   for c0 + SW + c <= min(r, w - 1)
__device__ __forceinline__ void panel_update16(double *Pn, int c0, int w, int rows) {
  for (int r = c0 + SW + threadIdx.x; r < rows; r += NT) {
    double acc[UPW];
#pragma unroll
    for (int c = 0; c < UPW; c++) acc[c] = 0.0;
#pragma unroll 4
    for (int t = 0; t < SW; t++) {
      const double lr = Pn[(c0 + t) * LDP + r];
#pragma unroll
      for (int c = 0; c < UPW; c++) if (c0 + SW + c < PW) acc[c] = fma(lr, Pn[(c0 + t) * LDP + c0 + SW + c], acc[c]);
    }
#pragma unroll
    for (int c = 0; c < UPW; c++) if (c0 + SW + c < w && c0 + SW + c <= r) Pn[(c0 + SW + c) * LDP + r] -= acc[c];
  }
}

// L <- chol(sscale * src + beta I) (lower, n x n), right-looking with PW-column panels.  rdiag_g[j] = 1 / l_jj.
// forward substitution step of one resident panel: z_k = L_kk^{-1} v_k (warp 0), v_below -= L_below,k z_k (all threads)
__device__ __forceinline__ void panel_forward(const double *Pn, double *v, const double *rdv, int k0, int w, int rows) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (warp == 0) {
    double vl = (lane < w) ? v[k0 + lane] : 0.0;
    const double rd = (lane < w) ? rdv[lane] : 1.0;
    __syncwarp();   // re-join the warp: a split warp pays the slow collective path on every shuffle
    for (int j = 0; j < w; j++) {
      const double zj = __shfl_sync(0xffffffffu, vl, j) * __shfl_sync(0xffffffffu, rd, j);
      if (lane == j) vl = zj;
      else if (lane > j && lane < w) vl = fma(-Pn[j * LDP + lane], zj, vl);
    }
    if (lane < w) v[k0 + lane] = vl;
  }
  __syncthreads();
  for (int r = w + tid; r < rows; r += NT) {
    double s0 = 0.0, s1 = 0.0;
    int t = 0;
    for (; t + 1 < w; t += 2) { s0 = fma(Pn[t * LDP + r], v[k0 + t], s0); s1 = fma(Pn[(t + 1) * LDP + r], v[k0 + t + 1], s1); }
    if (t < w) s0 = fma(Pn[t * LDP + r], v[k0 + t], s0);
    v[k0 + r] -= (s0 + s1);
  }
  __syncthreads();
}

// fwd: additionally run the forward substitution L z = v on S.v while each panel is resident (saves re-reading L).
__device__ __forceinline__ void cta_potrf(double *L, int ld, const double *src, int lds, double sscale, double beta, int n, double *rdiag_g,
                          const Smem &S, int *info, bool fwd, long long *pf) {
  const int tid = threadIdx.x;
  double *Pn = S.panel;
  long long tq = clock64();
// Phase clocks: one predicated atomic, NO divergent region.  (An `if (tid == 0) { load; add; store }` body leaves lane 0 split
// from the rest of warp 0 under independent thread scheduling; every __shfl_sync of the following single-warp phases then
// takes the WARPSYNC.COLLECTIVE slow path -- measured 131k instead of 13.6k clocks per 16 x 16 diagonal block, which made
// the profile of round 1 blame the wrong phase.)
#define PQ(k) do { if (pf) { const long long t_ = clock64(); if (tid == NT - 1) atomicAdd(reinterpret_cast<unsigned long long *>(pf + (k)), (unsigned long long)(t_ - tq)); tq = t_; } } while (0)
  for (int k0 = 0; k0 < n; k0 += PW) {
    const int w = (n - k0 < PW) ? n - k0 : PW, rows = n - k0;
    const bool first = (k0 == 0);
    if (first) {
      for (int idx = tid; idx < w * rows; idx += NT) {   // panel load, rows fastest (coalesced)
        const int t = idx / rows, r = idx - t * rows;
        double v = 0.0;
        if (r >= t) { v = sscale * src[(size_t)r + (size_t)lds * t]; if (r == t) v += beta; }
        Pn[t * LDP + r] = v;
      }
    } else panel_load_async(Pn, L, ld, k0, w, rows);
    __syncthreads();
    PQ(16);
    for (int c0 = 0; c0 < w; c0 += SW) {   // sub-panels: diagonal block (one warp) -> rows below -> the panel's columns to the right
      if (tid < 32) warp_factor_diag16(Pn, S.rd, S.vs, c0, w, info);
      __syncthreads();
      PQ(17);
      panel_solve16(Pn, S.rd, c0, w, rows);
      __syncthreads();
      PQ(18);
      if (c0 + SW < w) {
        panel_update16(Pn, c0, w, rows);
        __syncthreads();
        PQ(19);
      }
    }
    for (int idx = tid; idx < w * rows; idx += NT) {   // final columns of L
      const int t = idx / rows, r = idx - t * rows;
      if (r >= t) L[(size_t)(k0 + r) + (size_t)ld * (k0 + t)] = Pn[t * LDP + r];
    }
    if (tid < w) rdiag_g[k0 + tid] = S.rd[tid];
    if (fwd) {
      panel_forward(Pn, S.v, S.rd, k0, w, rows);   // ends with a barrier; only reads the panel
    }
    PQ(20);
    if (w & 3) {   // ragged last panel: zero the pad columns the DMMA product loop reads (no trailing matrix then, but keep it defined)
      for (int idx = tid; idx < (((w + 3) & ~3) - w) * rows; idx += NT) { const int t = w + idx / rows, r = idx - (t - w) * rows; Pn[t * LDP + r] = 0.0; }
      __syncthreads();
    }
    cta_rank_update_lower(L, ld, src, lds, sscale, beta, first, k0 + w, n, Pn, k0, w, -1.0);
    __syncthreads();
    PQ(21);
  }
#undef PQ
}

// v <- (L L')^{-1} v, v in shared memory (S.v), L n x n lower in global memory, rdiag_g = 1 / diag(L).
// skip_forward: the forward substitution was already done inside cta_potrf.
__device__ __forceinline__ void cta_chol_solve(const double *L, int ld, int n, const double *rdiag_g, const Smem &S, bool skip_forward) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double *Pn = S.panel, *v = S.v;
  if (!skip_forward) {
    for (int k0 = 0; k0 < n; k0 += PW) {   // forward: L z = v
      const int w = (n - k0 < PW) ? n - k0 : PW, rows = n - k0;
      panel_load_async(Pn, L, ld, k0, w, rows);
      if (tid < w) S.rd[tid] = rdiag_g[k0 + tid];
      __syncthreads();
      panel_forward(Pn, v, S.rd, k0, w, rows);
    }
  }
  // backward: L' d = z
  const int last = ((n - 1) / PW) * PW;
  for (int k0 = last; k0 >= 0; k0 -= PW) {
    const int w = (n - k0 < PW) ? n - k0 : PW, rows = n - k0;
    panel_load_async(Pn, L, ld, k0, w, rows);
    __syncthreads();
    for (int t = warp; t < w; t += NW) {   // v[k0 + t] -= sum_{r >= w} L(k0 + r, k0 + t) d(k0 + r)
      double s = 0.0;
      for (int r = w + lane; r < rows; r += 32) s = fma(Pn[t * LDP + r], v[k0 + r], s);
      s = warp_sum(s);
      if (lane == 0) S.rd[PW + t] = s;
    }
    __syncthreads();
    if (warp == 0) {
      double vl = (lane < w) ? v[k0 + lane] - S.rd[PW + lane] : 0.0;
      const double rd = (lane < w) ? rdiag_g[k0 + lane] : 1.0;
      __syncwarp();
      for (int j = w - 1; j >= 0; j--) {
        const double dj = __shfl_sync(0xffffffffu, vl, j) * __shfl_sync(0xffffffffu, rd, j);
        if (lane == j) vl = dj;
        else if (lane < j) vl = fma(-Pn[lane * LDP + j], dj, vl);   // L(k0 + j, k0 + lane)
      }
      if (lane < w) v[k0 + lane] = vl;
    }
    __syncthreads();
  }
}

// dst(lower) += sign * sum_{c < cnt} (wgt[c] A'[:, list[c]]) (...)'   in chunks of PW list entries staged in the panel
__device__ __forceinline__ void cta_syrk_list(double *dst, int ld, int n, const double *__restrict__ At, const int *__restrict__ list,
                              const double *__restrict__ wgt, int cnt, double sign, const Smem &S) {
  const int tid = threadIdx.x;
  for (int off = 0; off < cnt; off += PW) {
    const int w = (cnt - off < PW) ? cnt - off : PW;
    const int wpad = (w + 3) & ~3;   // the DMMA product loop walks 4 columns at a time: zero the pad columns
    // gather: thread = row, eight columns (list entries) requested before any is stored -- the one-element-at-a-time form chained
    // list[c] -> A'[:, list[c]] loads behind an integer division per element
    for (int i = tid; i < n; i += NT) {
      for (int c0 = 0; c0 < wpad; c0 += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          const int c = c0 + u;
          v[u] = (c < w) ? wgt[off + c] * At[(size_t)i + (size_t)n * list[off + c]] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 8; u++) if (c0 + u < wpad) S.panel[(c0 + u) * LDP + i] = v[u];
      }
    }
    __syncthreads();
    cta_rank_update_lower(dst, ld, nullptr, 0, 0.0, 0.0, false, 0, n, S.panel, 0, w, sign);
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// rank-k update / downdate of the instance's factor inside the CTA -- the cholmod_updown replacement of the batch engine
// (ldlupdate_entering_constraints / ldldowndate_leaving_constraints, src/solver_interface.c:407-441; recurrence of
// Modify/t_cholmod_updown_numkr.c:289-376 restated for L L').  <= KU ranks per sweep, entering rows (+) before leaving (-).
//
// Per column j with pivot d = l_jj^2 and, per rank r, the carried alpha_r and the current w_r[j]:
//     d_{r+1} = d_r + s_r w_r[j]^2 / alpha_r          gamma_r = -s_r w_r[j] / (alpha_r d_{r+1})        alpha_r <- alpha_r d_{r+1} / d_r
// and for the rows below:  t = l_ij / l_jj;  { w_r[i] -= w_r[j] t;  t -= gamma_r w_r[i] } over r;  l_ij = t sqrt(d_k).
// The chain over the ranks of one column only ever ADDS (s_r w^2 / alpha_r are known up front), so lane r of warp 0 owns
// rank r, the d_r come out of one warp prefix sum and each lane pays a single division: the serial cost per column does
// not grow with the rank.  The sweep walks 16-column panels staged in shared memory: warp 0 runs the recurrence on the
// 16 x 16 diagonal block (its rows of W in registers) and leaves the column coefficients in shared memory, then every thread
// transforms one row below the block.
// ------------------------------------------------------------------------------------------------
constexpr int KU = 8, UW = 16, WIN = 32, PBL = 33;   // ranks per sweep, columns per block, rows of the chain warp's window, block row pitch
constexpr int NRING = 3;                              // window-block buffers: in use by the chain / being requested / being written back
constexpr int LVW = 8;                                // row owners: columns of their row's entries in flight (shared-memory staging ring)
static_assert((KU * LDP + NRING * UW * PBL + LVW * LDP) * sizeof(double) <= kUnionBytes,
              "W block + the ring of window-block buffers + the row owners' staging ring must fit the union region");

struct UdCoef {   // in S.vs: the coefficients of one 16-column block, written column by column by the chain warp
  double winv[UW], d0[UW], dfin[2][UW];   // old pivots (1 / l_jj, l_jj^2) of the whole block, written at its start; dfin by block parity: the previous block's new pivots are still needed while it is written back
  __align__(16) double wj[UW][KU];
  __align__(16) double gam[UW][KU];
  __align__(16) double wpiv[KU];   // the pivot row's W as it stands before its column (published by that row's lane one column earlier)
  __align__(16) double cbuf[KU];   // per-rank pivot increments s_r w_r^2 / alpha_r of the current column
};
static_assert(sizeof(UdCoef) <= sizeof(double) * VS_LEN, "update coefficients must fit the staged-vector buffer");

__device__ __forceinline__ void ud_mbar_init(unsigned long long *bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
// arrive by lane 0 only, as a PREDICATED instruction (a branch would split the chain warp, see the note in the column loop)
__device__ __forceinline__ void ud_mbar_arrive_lane0(unsigned long long *bar, int lane) {
  asm volatile("{\n.reg .pred p;\nsetp.eq.s32 p, %1, 0;\n@p mbarrier.arrive.shared::cta.b64 _, [%0];\n}\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(lane) : "memory");
}
__device__ __forceinline__ void ud_mbar_wait(unsigned long long *bar, unsigned parity) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
#ifdef QB_UD_TESTWAIT
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(a), "r"(parity) : "memory");
#else
  asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}\n" ::"r"(a), "r"(parity) : "memory");
#endif
}
__device__ __forceinline__ void bp_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bp_cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// chain warp: request the 32 x 16 window block at (k0, k0) -- lane = row k0 + lane, columns c <= lane -- into Pb[c * PBL + lane]
__device__ __forceinline__ void ud_block_prefetch(double *Pb, const double *L, int ld, int k0, int n, int lane) {
  const int row = k0 + lane;
#pragma unroll
  for (int c = 0; c < UW; c++) {
    if (row < n && k0 + c < n && c <= lane) bp_cp_async8(&Pb[c * PBL + lane], L + (size_t)row + (size_t)ld * (k0 + c));
    else Pb[c * PBL + lane] = 0.0;
  }
  bp_cp_async_commit();
}

// one sweep: L <- chol(L L' + sum_{r < kpos} w_r w_r' - sum_{kpos <= r < k} w_r w_r'),  w_r = wgt[r] * A'[:, list[r]]
//
// Round-2 shape (profiles/r02f_ncu_source_kbp_solve.txt: the old panel-staged sweep spent 18 % of the kernel with seven warps parked
// on the barrier behind warp 0's recurrence and another 12 % loading / storing panels):
//   * the CHAIN WARP (warp 0) walks the diagonal.  Its 32 lanes are the rows k0 .. k0 + 31 of the current 16-column block (the 16 block
//     rows and 16 look-ahead rows); their L entries sit in a small shared buffer that was requested one block ahead (cp.async), their
//     rows of W in registers.  Per column it forms the (alpha, gamma) coefficients, publishes them and arrives on that column's mbarrier.
//   * every other thread owns ONE row of L for the whole sweep (thread t <-> row t), keeps that row of W in registers and its 16
//     entries of the current block in registers (loaded straight from global memory before the block starts, written straight back):
//     it waits on the column's mbarrier, applies the column, and so trails the chain warp by one column.  The rows-below pass, the
//     panel staging and two of the four barriers per block are gone; what remains per block is the chain itself.
// The arithmetic per row is unchanged (t = l / l_jj; { w_r -= w_r[j] t; t -= gamma_r w_r }; l' = t sqrt(d_k)).
// bars: UW mbarriers (count 1) initialised once per kernel; every block completes exactly one phase of each; phase = blocks so far.
// block-end rendezvous of the two roles (named barrier 1; __syncthreads() is barrier 0 and is used by the code around the sweep)
__device__ __forceinline__ void ud_block_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

// write the finished window block (rows k0 .. k0 + 31, columns k0 .. k0 + 15; unit-scaled entries in Pn, new pivots squared in dfin)
// back to L: l' = t sqrt(d_k), l_jj' = sqrt(d_k), rdiag = 1 / l_jj'.  Called by `nthr` threads with t = 0 .. nthr - 1.
__device__ __forceinline__ void ud_flush_block(double *L, int ld, int n, double *rdiag_g, int k0, const double *Pn, const double *dfin, int t, int nthr) {
  const int w = (n - k0 < UW) ? n - k0 : UW;
  for (int e = t; e < WIN * UW; e += nthr) {
    const int l = e & (WIN - 1), c = e / WIN;
    if (c < w && k0 + l < n && l >= c) {
      double ln, lninv;
      chol32::sqrt_and_rcp(dfin[c], ln, lninv);   // branch-free (no slow-path call); every user of the new pivot forms it this way
      L[(size_t)(k0 + l) + (size_t)ld * (k0 + c)] = (l == c) ? ln : Pn[c * PBL + l] * ln;
      if (l == c) rdiag_g[k0 + c] = lninv;
    }
  }
}

// role 1 of the sweep: the chain warp (warp 0).  Nothing but the recurrence: its window blocks are requested and written back by
// the row owners (a first version did both here and spent 40 % of every block outside the column loop).
__device__ __forceinline__ void ud_chain_role(int n, int k, int kpos, double *Wm, double *Pb0, UdCoef &cf, int *info, unsigned long long *bars,
                                              long long *pf) {
  const int lane = threadIdx.x & 31;
  long long tq = clock64();
#define PC(k) do { if (pf) { const long long t_ = clock64(); if (lane == 31) atomicAdd(reinterpret_cast<unsigned long long *>(pf + (k)), (unsigned long long)(t_ - tq)); tq = t_; } } while (0)
  double ial = 1.0;                          // lane r < k: 1 / alpha_r
  const double sg = (lane < kpos) ? 1.0 : -1.0;
  bool bad = false;
  int slot = 0, blk = 0;
  for (int k0 = 0; k0 < n; k0 += UW, slot = (slot + 1 == NRING) ? 0 : slot + 1, blk++) {
    const int w = (n - k0 < UW) ? n - k0 : UW;
    double *Pn = Pb0 + slot * (UW * PBL);
    double *dfin_out = cf.dfin[blk & 1];
    const bool rowvalid = k0 + lane < n;
    double wl[KU];
#pragma unroll
    for (int r = 0; r < KU; r++) wl[r] = rowvalid ? Wm[r * LDP + k0 + lane] : 0.0;
    // off the chain: the old pivots are not touched before their own column, so 1 / l_jj and l_jj^2 of all 16 columns are
    // formed up front (lane = column); the new pivots sqrt(d) and the column scaling are left to the write-back
    const double ldiag = (lane < w) ? Pn[lane * PBL + lane] : 1.0;
    double winv_l;   // 1 / l_jj: reciprocal seed + two Newton steps (the IEEE division subroutine has a slow-path call as well)
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(winv_l) : "d"(ldiag));
    winv_l = fma(winv_l, fma(-ldiag, winv_l, 1.0), winv_l);
    winv_l = fma(winv_l, fma(-ldiag, winv_l, 1.0), winv_l);
    const double d0_l = ldiag * ldiag;
    __syncwarp();
    PC(23);
    // Column step.  One dependent chain: pivot row's W (shared memory, left there by that row's lane at the end of the previous
    // step) -> per-rank increments c_r -> their prefix sums d_r (ONE exchange through shared memory and a depth-3 add tree; the first
    // version scanned with three dependent shuffle rounds at ~54 clocks each) -> reciprocal -> gamma_r, alpha_r -> publish ->
    // row update of the window.  There is no shuffle in the loop (the old pivots of the block sit in shared memory), so a lane-dependent
    // branch around a store cannot push later steps onto the slow collective path.
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < KU; r++) cf.wpiv[r] = wl[r];
    }
    if (lane < UW) { cf.winv[lane] = winv_l; cf.d0[lane] = d0_l; }
    double sgial = sg * ial;
    __syncwarp();
    for (int j = 0; j < w; j++) {
      const double winv = cf.winv[j], d0 = cf.d0[j];
      const double wj = (lane < k) ? cf.wpiv[lane & (KU - 1)] : 0.0;
      const double c = wj * wj * sgial;            // = sg wj^2 ial, same rounding as ((sg wj) wj) ial
      if (lane < KU) cf.cbuf[lane] = c;
      __syncwarp();
      double incl;
      {
        const double2 *cb = reinterpret_cast<const double2 *>(cf.cbuf);
        const double2 c01 = cb[0], c23 = cb[1], c45 = cb[2], c67 = cb[3];
        const int rr = lane & (KU - 1);
        const double m0 = c01.x, m1 = (rr >= 1) ? c01.y : 0.0, m2 = (rr >= 2) ? c23.x : 0.0, m3 = (rr >= 3) ? c23.y : 0.0;
        const double m4 = (rr >= 4) ? c45.x : 0.0, m5 = (rr >= 5) ? c45.y : 0.0, m6 = (rr >= 6) ? c67.x : 0.0, m7 = (rr >= 7) ? c67.y : 0.0;
        incl = ((m0 + m1) + (m2 + m3)) + ((m4 + m5) + (m6 + m7));
      }
      const double dnext = d0 + incl, dprev = d0 + (incl - c);
      if (lane < k && !(dnext > 0.0)) bad = true;
      double q;   // ial / dnext: reciprocal seed + two Newton steps instead of the IEEE division subroutine
      asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(dnext));
      q = fma(q, fma(-dnext, q, 1.0), q);
      q = fma(q, fma(-dnext, q, 1.0), q);
      q *= ial;
      const double gam = -sg * wj * q;
      ial = dprev * q;
      sgial = sg * ial;
      // publish column j: lanes < KU the rank coefficients, lane k - 1 (which holds d_k) the new pivot squared
      if (lane < KU) { cf.wj[j][lane] = (lane < k) ? wj : 0.0; cf.gam[j][lane] = (lane < k) ? gam : 0.0; }
      if (lane == k - 1) dfin_out[j] = dnext;
      __syncwarp();
      ud_mbar_arrive_lane0(bars + j, lane);   // column j is published: the row owners may apply it
      {
        double t = (rowvalid ? Pn[j * PBL + lane] : 0.0) * winv;
        const double2 *wj2 = reinterpret_cast<const double2 *>(cf.wj[j]), *gm2 = reinterpret_cast<const double2 *>(cf.gam[j]);
#pragma unroll
        for (int r2 = 0; r2 < KU / 2; r2++) {
          const double2 a = wj2[r2], g = gm2[r2];
          wl[2 * r2] = fma(-a.x, t, wl[2 * r2]);
          t = fma(-g.x, wl[2 * r2], t);
          wl[2 * r2 + 1] = fma(-a.y, t, wl[2 * r2 + 1]);
          t = fma(-g.y, wl[2 * r2 + 1], t);
        }
        if (lane > j && rowvalid) Pn[j * PBL + lane] = t;   // unit-scaled; the write-back multiplies by the new pivot
        if (lane == j + 1) {   // the next pivot row: leave its W for the next step
#pragma unroll
          for (int r = 0; r < KU; r++) cf.wpiv[r] = wl[r];
        }
      }
      __syncwarp();
    }
    PC(25);
    for (int j = w; j < UW; j++) ud_mbar_arrive_lane0(bars + j, lane);   // ragged last block: every barrier completes one phase per block
    if (lane >= UW && rowvalid) {   // the look-ahead rows are block rows of the next window: hand their W over
#pragma unroll
      for (int r = 0; r < KU; r++) Wm[r * LDP + k0 + lane] = wl[r];
    }
    __syncwarp();
    PC(26);
    ud_block_barrier();
    PC(27);
  }
  if (__any_sync(0xffffffffu, bad) && lane == 0) *info = 1;
#undef PC
}

// role 2 of the sweep: a row owner (threads >= WIN; thread t <-> row t).  Besides its own row it helps the chain warp: the previous
// window block is written back by all row owners, the next one is requested by warp 1.
__device__ __forceinline__ void ud_row_role(double *L, int ld, int n, double *rdiag_g, double *Wm, double *Pb0, double *Ls, const UdCoef &cf,
                                            unsigned long long *bars, unsigned phase) {
  const int row = threadIdx.x, lane = threadIdx.x & 31;
  const bool requester = (threadIdx.x >> 5) == 1;
  double wv[KU];
#pragma unroll
  for (int r = 0; r < KU; r++) wv[r] = (row < n) ? Wm[r * LDP + row] : 0.0;
  int slot = 0, blk = 0;
  // This row's entries of the block travel through a private staging ring in shared memory: Ls[c % LVW][row] is written by this
  // thread's own cp.async and read by this thread only, so no barrier is involved and no registers are held (16 entries in
  // registers spill at 64 registers per thread, and a spilled entry costs an L2 round trip per column).  Entry c + LVW (of this or
  // the next block) is requested as soon as entry c has been consumed, i.e. LVW columns ahead of its use.
  if (row >= WIN && row < n) {
#pragma unroll
    for (int c = 0; c < LVW; c++) {
      if (c < n) bp_cp_async8(&Ls[c * LDP + row], L + (size_t)row + (size_t)ld * c);
      bp_cp_async_commit();
    }
  }
  for (int k0 = 0; k0 < n; k0 += UW, slot = (slot + 1 == NRING) ? 0 : slot + 1, blk++) {
    const int w = (n - k0 < UW) ? n - k0 : UW;
    const unsigned par = phase & 1u;
    phase++;
    const int slot_next = (slot + 1 == NRING) ? 0 : slot + 1, slot_prev = (slot == 0) ? NRING - 1 : slot - 1;
    const bool below = row >= k0 + WIN && row < n, below_next = row >= k0 + UW + WIN && row < n;
    double *Lr = L + (size_t)row + (size_t)ld * k0;
    const double *dfin = cf.dfin[blk & 1];
    if (requester && k0 + UW < n) ud_block_prefetch(Pb0 + slot_next * (UW * PBL), L, ld, k0 + UW, n, lane);
    if (k0 > 0) ud_flush_block(L, ld, n, rdiag_g, k0 - UW, Pb0 + slot_prev * (UW * PBL), cf.dfin[(blk - 1) & 1], threadIdx.x - WIN, NT - WIN);
    if (below) {   // below the window: apply the block column by column as it is published
#pragma unroll
      for (int c = 0; c < UW; c++) {
        if (c < w) {
          bp_cp_async_wait_group<LVW - 1>();   // this thread's request for column c has landed (later groups may still be in flight)
          ud_mbar_wait(bars + c, par);
          double t = Ls[(c % LVW) * LDP + row] * cf.winv[c];
          const double2 *wj2 = reinterpret_cast<const double2 *>(cf.wj[c]), *gm2 = reinterpret_cast<const double2 *>(cf.gam[c]);
#pragma unroll
          for (int r2 = 0; r2 < KU / 2; r2++) {
            const double2 a = wj2[r2], g = gm2[r2];
            wv[2 * r2] = fma(-a.x, t, wv[2 * r2]);
            t = fma(-g.x, wv[2 * r2], t);
            wv[2 * r2 + 1] = fma(-a.y, t, wv[2 * r2 + 1]);
            t = fma(-g.y, wv[2 * r2 + 1], t);
          }
          double ln, lninv;
          chol32::sqrt_and_rcp(dfin[c], ln, lninv);
          Lr[(size_t)ld * c] = t * ln;
        }
        // refill the slot just consumed: column c + LVW of this block, or of the next block when this row stays below the window
        const int cn = c + LVW;
        if ((cn < UW || below_next) && k0 + cn < n) bp_cp_async8(&Ls[(c % LVW) * LDP + row], Lr + (size_t)ld * cn);
        bp_cp_async_commit();
      }
      if (row < k0 + WIN + UW) {   // this row enters the chain warp's window with the next block: hand its W over
#pragma unroll
        for (int r = 0; r < KU; r++) Wm[r * LDP + row] = wv[r];
      }
    }
    if (requester) bp_cp_async_wait_group<0>();
    ud_block_barrier();
  }
}

__device__ __noinline__ void cta_updown_sweep(double *L, int ld, int n, double *rdiag_g, const double *__restrict__ At,
                                              const int *__restrict__ list, const double *__restrict__ wgt, int k, int kpos, const Smem &S, int *info,
                                              unsigned long long *bars, unsigned &phase, long long *pf) {
  const int tid = threadIdx.x;
  double *Wm = S.panel, *Pb0 = S.panel + KU * LDP;
  UdCoef &cf = *reinterpret_cast<UdCoef *>(S.vs);
  long long tq = clock64();
#define PQ(k) do { if (pf) { const long long t_ = clock64(); if (tid == NT - 1) atomicAdd(reinterpret_cast<unsigned long long *>(pf + (k)), (unsigned long long)(t_ - tq)); tq = t_; } } while (0)
  for (int idx = tid; idx < KU * n; idx += NT) {   // gather the rows (zero columns beyond k)
    const int r = idx / n, i = idx - r * n;
    Wm[r * LDP + i] = (r < k) ? wgt[r] * At[(size_t)i + (size_t)n * list[r]] : 0.0;
  }
  if (tid < 32) { ud_block_prefetch(Pb0, L, ld, 0, n, tid); bp_cp_async_wait_group<0>(); }
  __syncthreads();
  PQ(22);
  if (tid < 32) ud_chain_role(n, k, kpos, Wm, Pb0, cf, info, bars, pf);
  else ud_row_role(L, ld, n, rdiag_g, Wm, Pb0, Pb0 + NRING * UW * PBL, cf, bars, phase);
  const int nblk = (n + UW - 1) / UW;
  phase += (unsigned)nblk;
  ud_flush_block(L, ld, n, rdiag_g, (nblk - 1) * UW, Pb0 + ((nblk - 1) % NRING) * (UW * PBL), cf.dfin[(nblk - 1) & 1], tid, NT);   // last window block
  __syncthreads();
  PQ(24);
#undef PQ
}

// ordered lists of the entering (candidate active, not in the factor) and leaving rows, entering first, weights sqrt(sigma):
// list_pos[0 .. ne + nl), w_pos likewise (set_entering_leaving_constraints, newton.c:134-149).  Returns counts through ctl.npos / nneg.
__device__ __noinline__ void p_updown_lists(const Args &P, int b) {
  __shared__ int warp_cnt0[NW], warp_cnt1[NW];
  __shared__ int base0, base1;
  const int m = P.m, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t om = (size_t)b * m;
  if (tid == 0) { base0 = 0; base1 = 0; }
  __syncthreads();
  for (int start = 0; start < m; start += NT) {
    const int i = start + tid;
    bool p0 = false, p1 = false;
    if (i < m) {
      const int a = P.active_cand[om + i], o = P.active_old[om + i];
      p0 = a && !o; p1 = !a && o;
    }
    const unsigned b0 = __ballot_sync(0xffffffffu, p0), b1 = __ballot_sync(0xffffffffu, p1);
    if (lane == 0) { warp_cnt0[warp] = __popc(b0); warp_cnt1[warp] = __popc(b1); }
    __syncthreads();
    int off0 = base0, off1 = base1;
    for (int w = 0; w < warp; w++) { off0 += warp_cnt0[w]; off1 += warp_cnt1[w]; }
    const unsigned lt = (1u << lane) - 1u;
    if (p0) P.list_pos[om + off0 + __popc(b0 & lt)] = i;
    if (p1) P.list_neg[om + off1 + __popc(b1 & lt)] = i;
    __syncthreads();
    if (tid == 0) { int t0 = 0, t1 = 0; for (int w = 0; w < NW; w++) { t0 += warp_cnt0[w]; t1 += warp_cnt1[w]; } base0 += t0; base1 += t1; }
    __syncthreads();
  }
  const int ne = base0, nl = base1;
  __syncthreads();
  for (int q = tid; q < nl; q += NT) P.list_pos[om + ne + q] = P.list_neg[om + q];   // append the leaving rows
  __syncthreads();
  for (int q = tid; q < ne + nl; q += NT) P.w_pos[om + q] = P.sqrt_sigma[om + P.list_pos[om + q]];
  if (tid == 0) { BCtl &c = P.ctl[b]; c.npos = ne; c.nneg = nl; }
}

// ------------------------------------------------------------------------------------------------
// active-set commit + ordered H-difference lists (batch.cu kb_lists, one CTA)
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ void p_lists(const Args &P, int b, int refac) {
  __shared__ int warp_cnt0[NW], warp_cnt1[NW];
  __shared__ int base0, base1, redo;
  const int m = P.m, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t om = (size_t)b * m;
  for (int i = tid; i < m; i += NT) P.active[om + i] = P.active_cand[om + i];
  if (!refac) return;
  int scratch_mode = P.ctl[b].scratch;
  const int nb_active = P.ctl[b].nb_active;
  for (int attempt = 0; attempt < 2; attempt++) {
    __syncthreads();
    if (tid == 0) { base0 = 0; base1 = 0; redo = 0; }
    __syncthreads();
    for (int start = 0; start < m; start += NT) {
      const int i = start + tid;
      bool p0 = false, p1 = false;
      double s0 = 0.0, s1 = 0.0;
      if (i < m) {
        const int a = P.active[om + i];
        const double wn = a ? P.sigma[om + i] : 0.0;
        const double wh = (scratch_mode || !P.activeH[om + i]) ? 0.0 : P.sigmaH[om + i];
        const double dw = wn - wh;
        if (dw > 0.0) { p0 = true; s0 = (wh == 0.0) ? P.sqrt_sigma[om + i] : sqrt(dw); }
        else if (dw < 0.0) { p1 = true; s1 = sqrt(-dw); }
      }
      const unsigned b0 = __ballot_sync(0xffffffffu, p0), b1 = __ballot_sync(0xffffffffu, p1);
      if (lane == 0) { warp_cnt0[warp] = __popc(b0); warp_cnt1[warp] = __popc(b1); }
      __syncthreads();
      int off0 = base0, off1 = base1;
      for (int w = 0; w < warp; w++) { off0 += warp_cnt0[w]; off1 += warp_cnt1[w]; }
      const unsigned lt = (1u << lane) - 1u;
      if (p0) { const int pos = off0 + __popc(b0 & lt); P.list_pos[om + pos] = i; P.w_pos[om + pos] = s0; }
      if (p1) { const int pos = off1 + __popc(b1 & lt); P.list_neg[om + pos] = i; P.w_neg[om + pos] = s1; }
      __syncthreads();
      if (tid == 0) { int t0 = 0, t1 = 0; for (int w = 0; w < NW; w++) { t0 += warp_cnt0[w]; t1 += warp_cnt1[w]; } base0 += t0; base1 += t1; }
      __syncthreads();
    }
    if (tid == 0 && !scratch_mode && base0 + base1 > nb_active) redo = 1;   // cheaper to rebuild from Q
    __syncthreads();
    if (!redo) break;
    scratch_mode = 1;
  }
  for (int i = tid; i < m; i += NT) { P.activeH[om + i] = P.active[om + i]; P.sigmaH[om + i] = P.sigma[om + i]; }
  if (tid == 0) {
    BCtl &c = P.ctl[b];
    c.npos = base0; c.nneg = base1; c.H_valid = 1; c.scratch = scratch_mode;
  }
}

// ordered list of the active rows with weights sqrt(sigma) (boost_gamma's A_J' Sigma_J A_J)
__device__ __noinline__ void p_active_list(const Args &P, int b) {
  __shared__ int warp_cnt[NW];
  __shared__ int base;
  const int m = P.m, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t om = (size_t)b * m;
  if (tid == 0) base = 0;
  __syncthreads();
  for (int start = 0; start < m; start += NT) {
    const int i = start + tid;
    const bool p = (i < m) && P.active[om + i];
    const unsigned bl = __ballot_sync(0xffffffffu, p);
    if (lane == 0) warp_cnt[warp] = __popc(bl);
    __syncthreads();
    int off = base;
    for (int w = 0; w < warp; w++) off += warp_cnt[w];
    if (p) { const int pos = off + __popc(bl & ((1u << lane) - 1u)); P.list_pos[om + pos] = i; P.w_pos[om + pos] = P.sqrt_sigma[om + i]; }
    __syncthreads();
    if (tid == 0) { int t = 0; for (int w = 0; w < NW; w++) t += warp_cnt[w]; base += t; }
    __syncthreads();
  }
  if (tid == 0) P.ctl[b].npos = base;
}

// boost_gamma path of the outer update (qpalm.c:613-627, iteration.c:159-211).  Returns with gamma / reset_newton set.
__device__ __noinline__ void p_boost(const Args &P, int b, double *scratch, const Smem &S) {
  __shared__ int need_gersh;
  const int n = P.n, m = P.m, tid = threadIdx.x;
  const BSet &st = P.st;
  const size_t om = (size_t)b * m, on = (size_t)b * n;
  double na = 0, ne = 0, nl = 0;
  for (int i = tid; i < m; i += NT) {
    const double t = P.y[om + i] / P.sigma[om + i];
    const double a = P.Ax[om + i] + t;
    P.Axys[om + i] = a;
    const int act = (a <= P.bmin[om + i]) || (a >= P.bmax[om + i]);
    P.active[om + i] = act;
    na += act; ne += act && !P.active_old[om + i]; nl += !act && P.active_old[om + i];
  }
  na = block_red<RED_SUM>(na, scratch); ne = block_red<RED_SUM>(ne, scratch); nl = block_red<RED_SUM>(nl, scratch);
  if (tid == 0) {
    BCtl &c = P.ctl[b];
    c.nb_active = (int)na; c.nb_enter = (int)ne; c.nb_leave = (int)nl; c.boost = 0;
    need_gersh = 0;
    if (ne == 0 && nl == 0) {
      c.boost = 1;
      if (na == 0) { c.gamma = 1e12; c.reset_newton = 1; }   // iteration.c:198-200
      else { c.scratch = 1; c.npos = 0; c.nneg = 0; need_gersh = 1; }
    } else if (c.gamma < st.gamma_max) { c.gamma = fmin(c.gamma * st.gamma_upd, st.gamma_max); c.reset_newton = 1; }
  }
  __syncthreads();
  if (need_gersh) {
    p_active_list(P, b);
    __syncthreads();
    double *Lb = P.L + (size_t)b * P.sLL;
    const int ld = P.ld;
    for (int idx = tid; idx < n * n; idx += NT) { const int j = idx / n, i = idx - j * n; if (i >= j) Lb[(size_t)i + (size_t)ld * j] = 0.0; }
    __syncthreads();
    cta_syrk_list(Lb, ld, n, P.At, P.list_pos + om, P.w_pos + om, P.ctl[b].npos, 1.0, S);
    __syncthreads();
    double ub = -1.0e300;
    for (int i = tid; i < n; i += NT) {   // Gershgorin: |row i| of the symmetric matrix
      double acc = 0.0;
      for (int j = 0; j <= i; j++) acc += fabs(Lb[(size_t)i + (size_t)j * ld]);
      for (int j = i + 1; j < n; j++) acc += fabs(Lb[(size_t)j + (size_t)i * ld]);
      ub = fmax(ub, acc);
    }
    ub = block_red<RED_MAX>(ub, scratch);
    if (tid == 0) {
      BCtl &c = P.ctl[b];
      c.gamma = fmax(st.gamma_max, 1e14 / ub);
      c.gamma_maxed = 1;
      c.reset_newton = 1;
    }
    __syncthreads();
  }
  // Qx / Qd shifts for a changed gamma (iteration.c:205-209); the Qd shift belongs to boost_gamma only
  const double g = P.ctl[b].gamma, gp = P.ctl[b].gamma_prev;
  if (g != gp) {
    const double tau = BSC(b, S_TAU);
    const int boosted = P.ctl[b].boost;
    for (int j = tid; j < n; j += NT) {
      P.Qx[on + j] = P.Qx[on + j] + (1.0 / g - 1.0 / gp) * P.x[on + j];
      if (boosted) P.Qd[on + j] = P.Qd[on + j] + (tau / g - tau / gp) * P.d[on + j];
    }
    __syncthreads();
    if (tid == 0) P.ctl[b].reset_newton = 1;
  }
}

// ------------------------------------------------------------------------------------------------
// line search (linesearch.c:14-120)
// ------------------------------------------------------------------------------------------------
__device__ __noinline__ void p_ls_build(const Args &P, int b, double *part) {
  const int n = P.n, m = P.m;
  const size_t on = (size_t)b * n, om = (size_t)b * m, o2 = (size_t)b * 2 * m;
  const double inv_gamma = 1 / P.ctl[b].gamma;
  double eta = 0, beta = 0;
  for (int j = threadIdx.x; j < n; j += NT) {
    double qd = P.Qd[on + j];
    const double dj = P.d[on + j];
    if (P.st.proximal) { qd = qd + inv_gamma * dj; P.Qd[on + j] = qd; }
    eta += dj * qd; beta += dj * P.df[on + j];
  }
  double a_part = 0, b_part = 0, n_l = 0;
  for (int i = threadIdx.x; i < m; i += NT) {
    const double ss = P.sqrt_sigma[om + i], sg = P.sigma[om + i], ax = P.Ax[om + i], yi = P.y[om + i];
    const double t = ss * P.Ad[om + i];
    double dl[2], al[2];
    dl[1] = t; dl[0] = t * -1;
    double u = ax - P.bmin[om + i]; u = sg * u; u = yi + u; al[0] = u / ss;
    u = P.bmax[om + i] - ax; u = sg * u; u = u - yi; al[1] = u / ss;
#pragma unroll
    for (int hh = 0; hh < 2; hh++) {
      const int idx = i + hh * m;
      const double s = al[hh] / dl[hh];
      const bool inL = s > 0, inP = dl[hh] > 0;
      P.keys[o2 + idx] = inL ? (unsigned long long)__double_as_longlong(s) : ~0ull;
      P.vals[o2 + idx] = (unsigned int)idx;
      const double d2 = dl[hh] * dl[hh], dalp = dl[hh] * al[hh];
      P.ls_da[o2 + idx] = inP ? d2 : -d2;
      P.ls_db[o2 + idx] = inP ? -dalp : dalp;
      if ((int)inL + (int)inP == 1) { a_part += d2; b_part += dalp; }
      n_l += inL;
    }
  }
  const double vals[5] = {eta, beta, a_part, b_part, n_l};
  const int ops[5] = {RED_SUM, RED_SUM, RED_SUM, RED_SUM, RED_SUM};
  const int slots[5] = {S_ETA, S_BETA, S_LS_A, S_LS_B, S_NL};
  block_red_multi<5>(P, b, vals, ops, slots, part);
}

// stable LSD radix sort of N <= SORT_MAX (key, val) pairs in shared memory; sorted pairs are written back to global
__device__ __forceinline__ void p_sort(int N, unsigned long long *kg, unsigned int *vg, const Smem &S) {
  // ping-pong between ONE shared-memory buffer and the instance's global (kg, vg) arrays (L2-resident, 23 KB): a second
  // shared buffer would cost the fourth CTA per SM
#if QB_BP_SORT_GLOBAL
  unsigned long long *k0 = reinterpret_cast<unsigned long long *>(S.u);
  unsigned int *v0 = reinterpret_cast<unsigned int *>(k0 + SORT_MAX);
  unsigned int *warp_cnt = v0 + SORT_MAX;            // [NW][256]
  unsigned long long *kalt = kg;
  unsigned int *valt = vg;
#else
  unsigned long long *k0 = reinterpret_cast<unsigned long long *>(S.u), *kalt = k0 + SORT_MAX;
  unsigned int *v0 = reinterpret_cast<unsigned int *>(kalt + SORT_MAX), *valt = v0 + SORT_MAX;
  unsigned int *warp_cnt = valt + SORT_MAX;          // [NW][256]
#endif
  unsigned int *running = warp_cnt + NW * 256;       // [256]
  unsigned int *hist = running + 256;                // [256]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < N; i += NT) { k0[i] = kg[i]; v0[i] = vg[i]; }
  __syncthreads();
  unsigned long long *kin = k0, *kout = kalt;
  unsigned int *vin = v0, *vout = valt;
  for (int pass = 0; pass < 8; pass++) {
    const int shift = pass * 8;
    for (int d = tid; d < 256; d += NT) hist[d] = 0;
    __syncthreads();
    for (int i = tid; i < N; i += NT) atomicAdd(&hist[(unsigned)(kin[i] >> shift) & 255u], 1u);
    __syncthreads();
    if (hist[(unsigned)(kin[0] >> shift) & 255u] == (unsigned)N) { __syncthreads(); continue; }   // all keys share this digit
    if (warp == 0) {
      unsigned int loc[8], sum = 0;
#pragma unroll
      for (int t = 0; t < 8; t++) { loc[t] = hist[lane * 8 + t]; sum += loc[t]; }
      unsigned int inc = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const unsigned int u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += u; }
      unsigned int run = inc - sum;
#pragma unroll
      for (int t = 0; t < 8; t++) { running[lane * 8 + t] = run; run += loc[t]; }
    }
    __syncthreads();
    for (int start = 0; start < N; start += NT) {
      for (int i = tid; i < NW * 256; i += NT) warp_cnt[i] = 0;
      __syncthreads();
      const int e = start + tid;
      const bool valid = e < N;
      unsigned long long k = 0; unsigned int v = 0; unsigned int dg = 0x10000u + lane;
      if (valid) { k = kin[e]; v = vin[e]; dg = (unsigned)(k >> shift) & 255u; }
      const unsigned peers = __match_any_sync(0xffffffffu, dg);
      const int rank = __popc(peers & ((1u << lane) - 1u));
      if (valid && rank == 0) warp_cnt[warp * 256 + dg] = __popc(peers);
      __syncthreads();
      if (valid) {
        unsigned int off = running[dg];
        for (int w = 0; w < warp; w++) off += warp_cnt[w * 256 + dg];
        kout[off + rank] = k; vout[off + rank] = v;
      }
      __syncthreads();
      for (int d = tid; d < 256; d += NT) {
        unsigned int t = 0;
#pragma unroll
        for (int w = 0; w < NW; w++) t += warp_cnt[w * 256 + d];
        running[d] += t;
      }
      __syncthreads();
    }
    unsigned long long *tk = kin; kin = kout; kout = tk;
    unsigned int *tv = vin; vin = vout; vout = tv;
  }
  __syncthreads();
  if (kin != kg) for (int i = tid; i < N; i += NT) { kg[i] = kin[i]; vg[i] = vin[i]; }
}

__device__ __noinline__ void p_ls_select(const Args &P, int b) {
  __shared__ double wa[NW], wb[NW];
  __shared__ double carry_a, carry_b;
  __shared__ int found;
  const size_t o2 = (size_t)b * 2 * P.m;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nL = (int)BSC(b, S_NL);
  if (tid == 0) { carry_a = BSC(b, S_ETA) + BSC(b, S_LS_A); carry_b = BSC(b, S_BETA) - BSC(b, S_LS_B); found = 0x7fffffff; }
  __syncthreads();
  for (int start = 0; start < nL; start += NT) {
    const int i = start + tid;
    double ta = 0.0, tb = 0.0, s = 0.0;
    if (i < nL) { const unsigned int idx = P.vals[o2 + i]; ta = P.ls_da[o2 + idx]; tb = P.ls_db[o2 + idx]; s = __longlong_as_double((long long)P.keys[o2 + i]); }
    double ia = ta, ib = tb;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double ua = __shfl_up_sync(0xffffffffu, ia, o), ub = __shfl_up_sync(0xffffffffu, ib, o);
      if (lane >= o) { ia += ua; ib += ub; }
    }
    if (lane == 31) { wa[warp] = ia; wb[warp] = ib; }
    __syncthreads();
    double oa = 0.0, ob = 0.0;
    for (int w = 0; w < warp; w++) { oa += wa[w]; ob += wb[w]; }
    double ea = __shfl_up_sync(0xffffffffu, ia, 1), eb = __shfl_up_sync(0xffffffffu, ib, 1);
    if (lane == 0) { ea = 0.0; eb = 0.0; }
    const double a_i = carry_a + (oa + ea), b_i = carry_b + (ob + eb);
    if (i < nL && (a_i * s + b_i > 0)) atomicMin(&found, i);
    __syncthreads();
    if (found != 0x7fffffff) {
      if (i == found) BSC(b, S_TAU) = -b_i / a_i;
      return;
    }
    if (tid == NT - 1) { carry_a += oa + ia; carry_b += ob + ib; }
    __syncthreads();
  }
  if (tid == 0) BSC(b, S_TAU) = -carry_b / carry_a;
}

__device__ void p_update_iterate(const Args &P, int b) {
  const int n = P.n, m = P.m;
  const double tau = BSC(b, S_TAU);
  const size_t on = (size_t)b * n, om = (size_t)b * m;
  for (int i = threadIdx.x; i < n; i += NT) {
    const double xi = P.x[on + i];
    P.x_prev[on + i] = xi; P.x[on + i] = xi + tau * P.d[on + i];
    const double qd = P.Qd[on + i] * tau;
    P.Qd[on + i] = qd; P.Qx[on + i] = P.Qx[on + i] + qd;
  }
  for (int i = threadIdx.x; i < m; i += NT) {
    const double ad = P.Ad[om + i] * tau;
    P.Ad[om + i] = ad; P.Ax[om + i] = P.Ax[om + i] + ad;
  }
}

// store_solution (termination.c:242-252) + compute_objective (iteration.c:231-270)
__device__ __noinline__ void p_store(const Args &P, int b, double *scratch) {
  const int n = P.n, m = P.m;
  const BSet &st = P.st;
  const size_t on = (size_t)b * n, om = (size_t)b * m;
  const double cinv = P.ctl[b].cinv, inv_gamma = 1 / P.ctl[b].gamma;
  double obj = 0;
  for (int i = threadIdx.x; i < n; i += NT) {
    const double xi = P.x[on + i];
    P.x_out[on + i] = st.scaling ? xi * P.D[i] : xi;
    if (st.proximal) obj += (0.5 * (P.Qx[on + i] - inv_gamma * xi) + P.q[on + i]) * xi;
    else obj += (0.5 * P.Qx[on + i] + P.q[on + i]) * xi;
  }
  for (int i = threadIdx.x; i < m; i += NT) {
    double v = P.yh[om + i];
    if (st.scaling) { v *= cinv; v = v * P.E[i]; }
    P.y_out[om + i] = v;
  }
  obj = block_red<RED_SUM>(obj, scratch);
  if (threadIdx.x == 0) { if (st.scaling) obj *= cinv; P.ctl[b].objective = obj + st.data_c; }
}

// ------------------------------------------------------------------------------------------------
// the persistent kernel
// ------------------------------------------------------------------------------------------------
#define PH(k) do { if (P.prof) { const long long t_ = clock64(); if (tid == NT - 1) atomicAdd(reinterpret_cast<unsigned long long *>(P.prof + (size_t)b * 32 + (k)), (unsigned long long)(t_ - t_ph)); t_ph = t_; } } while (0)
#ifndef QB_BP_MINB
#define QB_BP_MINB (QB_BP_NT <= 192 ? 4 : 3)
#endif
__global__ void __launch_bounds__(NT, QB_BP_MINB) kbp_solve(const Args P) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double scratch[32];
  __shared__ int s_b;
  __shared__ Flags s_f;
  __shared__ int s_info;
  __shared__ __align__(8) unsigned long long s_udbar[UW];   // per-column mbarriers of the update sweep (chain warp -> row owners)
  unsigned ud_phase = 0;
  if (threadIdx.x == 0) {
    for (int j = 0; j < UW; j++) ud_mbar_init(s_udbar + j, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  Smem S;
  S.u = smem_raw;
  S.panel = reinterpret_cast<double *>(smem_raw);
  S.vs = reinterpret_cast<double *>(smem_raw + kUnionBytes);
  S.v = S.vs + VS_LEN;
  S.rd = S.v + NMAX + 16;
  const int tid = threadIdx.x, n = P.n, m = P.m, ld = P.ld;
  const BSet &st = P.st;
  for (;;) {
    __syncthreads();
    if (tid == 0) s_b = atomicAdd(P.queue, 1);
    __syncthreads();
    const int b = s_b;
    if (b >= P.nb) return;
    const size_t on = (size_t)b * n, om = (size_t)b * m;
    double *Hb = P.H + (size_t)b * P.sLL, *Lb = P.L + (size_t)b * P.sLL, *rdg = P.rdiag + (size_t)b * P.sR;
    long long t_ph = clock64();
    if (P.prof && tid == 0) for (int k = 0; k < 32; k++) P.prof[(size_t)b * 32 + k] = 0;
    p_init(P, b, scratch);
    __syncthreads();
    PH(0);
    for (;;) {
      // ---- residuals + termination scalars ----
      p_res_m(P, b, S.vs);
      __syncthreads();
      PH(1);
      p_gemv_rows(n, m, n, P.At, P.yh + om, P.Atyh + on, 1.0, S);
      __syncthreads();
      PH(2);
      p_res_n(P, b, S.vs);
      __syncthreads();
      if (tid == 0) p_control(P, b, s_f);
      __syncthreads();
      const Flags f = s_f;
      __syncthreads();   // everyone holds a private copy before thread 0 touches s_f again
      PH(3);
      if (f.done) break;
      // ---- outer updates ----
      if (f.sigma) { p_update_sigma(P, b, scratch); __syncthreads(); }
      if (f.outer) { p_outer(P, b, f.outer); __syncthreads(); }
      if (st.proximal && f.boost) { p_boost(P, b, scratch, S); __syncthreads(); }
      PH(4);
      // ---- inner step ----
      if (f.inner) {
        int refac = f.refac, factor = f.factor;
        if (f.updown) {   // rank update of the factor instead of a refactorisation (newton.c:103-108)
          p_updown_lists(P, b);
          if (tid == 0) s_info = 0;
          __syncthreads();
          const int ne = P.ctl[b].npos, nl = P.ctl[b].nneg;
          for (int off = 0; off < ne + nl; off += KU) {
            const int k = (ne + nl - off < KU) ? ne + nl - off : KU;
            const int kpos = (ne - off < 0) ? 0 : ((ne - off < k) ? ne - off : k);
            cta_updown_sweep(Lb, ld, n, rdg, P.At, P.list_pos + om + off, P.w_pos + om + off, k, kpos, S, &s_info,
                             s_udbar, ud_phase, P.prof ? P.prof + (size_t)b * 32 : nullptr);
          }
          __syncthreads();
          if (tid == 0) {
            BCtl &c = P.ctl[b];
            c.n_updown += (ne + nl + KU - 1) / KU; c.updown_ranks += ne + nl;
            if (s_info) { c.n_updown_fail++; c.scratch = !c.H_valid; c.n_refac++; c.refac_J += c.nb_active; }
          }
          if (s_info) { refac = 1; factor = 1; }   // a downdate lost definiteness: rebuild the factor from H
          __syncthreads();
          PH(14);
        }
        p_lists(P, b, refac);
        __syncthreads();
        PH(5);
        if (refac) {
          const BCtl c = P.ctl[b];
          if (c.scratch) {
            for (int i = tid; i < n; i += NT) {   // H <- c Q (lower): thread = row, eight columns in flight
              for (int j0 = 0; j0 <= i; j0 += 8) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; u++) v[u] = (j0 + u <= i) ? P.Qs[(size_t)i + (size_t)n * (j0 + u)] : 0.0;
#pragma unroll
                for (int u = 0; u < 8; u++) if (j0 + u <= i) Hb[(size_t)i + (size_t)ld * (j0 + u)] = v[u] * c.c;
              }
            }
            __syncthreads();
          }
          for (int pass = 0; pass < 2; pass++)   // + entering / grown sigma, then - leaving (one inlined copy of the SYRK)
            cta_syrk_list(Hb, ld, n, P.At, (pass ? P.list_neg : P.list_pos) + om, (pass ? P.w_neg : P.w_pos) + om,
                          pass ? c.nneg : c.npos, pass ? -1.0 : 1.0, S);
        }
        PH(6);
        for (int i = tid; i < n; i += NT) S.v[i] = P.dphi[on + i] * -1;
        __syncthreads();
        if (factor) {
          const double beta = P.ctl[b].beta;
          if (tid == 0) s_info = 0;
          cta_potrf(Lb, ld, f.fq ? P.Qs : Hb, f.fq ? n : ld, f.fq ? P.ctl[b].c : 1.0, beta, n, rdg, S, &s_info, true,
                    P.prof ? P.prof + (size_t)b * 32 : nullptr);
        }
        PH(7);
        cta_chol_solve(Lb, ld, n, rdg, S, factor != 0);
        for (int i = tid; i < n; i += NT) P.d[on + i] = S.v[i];
        for (int i = tid; i < m; i += NT) P.active_old[om + i] = P.active[om + i];
        __syncthreads();
        PH(8);
        // ---- line search + iterate update ----
        p_gemv_rows(n, n, n, P.Qs, P.d + on, P.Qd + on, P.ctl[b].c, S);
        __syncthreads();
        PH(13);
        p_gemv_rows(m, n, m, P.Am, P.d + on, P.Ad + om, 1.0, S);
        __syncthreads();
        PH(9);
        p_ls_build(P, b, S.vs);
        __syncthreads();
        PH(10);
        p_sort(2 * m, P.keys + (size_t)b * 2 * m, P.vals + (size_t)b * 2 * m, S);
        __syncthreads();
        PH(11);
        p_ls_select(P, b);
        __syncthreads();
        p_update_iterate(P, b);
        __syncthreads();
        PH(12);
      }
      if (tid == 0) {   // end of iteration (qpalm.c:484, 711-735)
        BCtl &c = P.ctl[b];
        c.iter++;
        s_f.done = 0;
        if (c.iter >= st.max_iter) { c.status = QPALM_MAX_ITER_REACHED; c.done = 1; s_f.done = 1; }
      }
      __syncthreads();
      if (s_f.done) break;
    }
    __syncthreads();
    p_store(P, b, scratch);
  }
}

}  // namespace QB_BP_NAMESPACE (bp / bp4)

bool QB_BP_SUPPORTED(int n, int m) {
  return n >= 1 && n <= QB_BP_NAMESPACE::NMAX && m >= 1 && 2 * m <= QB_BP_NAMESPACE::SORT_MAX && m <= QB_BP_NAMESPACE::VS_LEN;
}

int QB_BP_SOLVE(QPALMB200Batch *B, int nb) {
  using namespace QB_BP_NAMESPACE;
  Engine *e = B->shared;
  Args P;
  memset(&P, 0, sizeof(P));
  P.nb = nb; P.n = B->n; P.m = B->m; P.ld = B->ld; P.st = B->set;
  P.At = e->At; P.Am = B->Am; P.Qs = B->Qs; P.D = e->D; P.Dinv = e->Dinv; P.E = e->E; P.Einv = e->Einv;
  P.q_raw = B->q_raw; P.bmin_raw = B->bmin_raw; P.bmax_raw = B->bmax_raw; P.x_out = B->x_out; P.y_out = B->y_out;
  P.q = B->q; P.bmin = B->bmin; P.bmax = B->bmax; P.x = B->x; P.y = B->y; P.Ax = B->Ax; P.Qx = B->Qx; P.Aty = B->Aty;
  P.x_prev = B->x_prev; P.x0 = B->x0; P.sigma = B->sigma; P.sigma_inv = B->sigma_inv; P.sqrt_sigma = B->sqrt_sigma;
  P.Axys = B->Axys; P.z = B->z; P.pri_res = B->pri_res; P.pri_res_in = B->pri_res_in; P.yh = B->yh; P.Atyh = B->Atyh;
  P.df = B->df; P.dphi = B->dphi; P.d = B->d; P.Qd = B->Qd; P.Ad = B->Ad;
  P.active = B->active; P.active_old = B->active_old; P.active_cand = B->active_cand; P.activeH = B->activeH;
  P.list_pos = B->list_pos; P.list_neg = B->list_neg; P.sigmaH = B->sigmaH; P.w_pos = B->w_pos; P.w_neg = B->w_neg;
  P.H = B->H; P.L = B->L; P.rdiag = B->invdiag;
  P.sLL = (long long)B->ld * B->npad; P.sR = (long long)B->npad * kPanel;
  P.prof = B->prof;
  P.keys = B->keys; P.vals = B->vals; P.ls_da = B->ls_da; P.ls_db = B->ls_db; P.scal = B->scal; P.ctl = B->ctl; P.queue = B->queue;
  static int ctas_per_sm = 0, num_sms = 0;
  if (!ctas_per_sm) {
    QB_CUDA_TRY(cudaFuncSetAttribute(kbp_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
    QB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kbp_solve, NT, kSmemBytes));
    int dev = 0;
    QB_CUDA_TRY(cudaGetDevice(&dev));
    QB_CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev));
    if (ctas_per_sm < 1) ctas_per_sm = 1;
  }
  int grid = ctas_per_sm * num_sms;
  if (grid > nb) grid = nb;
  QB_CUDA_TRY(cudaMemsetAsync(B->queue, 0, sizeof(int), B->stream));
  QB_LAUNCH(kbp_solve, grid, NT, kSmemBytes, B->stream, P);
  QB_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace qb
