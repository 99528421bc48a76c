// sparse.cu -- device side of the supernodal multifrontal Cholesky (see sparse.cuh for the design).
//
// Factor layout: supernode s owns permuted columns [first, first + ns) and the sorted rows R_s (nr of them) below;
// its panel is column-major (nf = ns + nr) x ns with leading dimension nf; its update block U_s is nr x nr (ld nr),
// lower triangle used.  "Front" index space of s: 0..ns-1 = own columns, ns..nf-1 = R_s.
//
// Numeric factorization, level l (all supernodes whose children are done):
//   levels whose fronts all fit in shared memory: k_mf_small, ONE launch does extend-add + partial factorization +
//                 write-back with the whole front resident in shared memory.
//   larger fronts: mfc::k_mf_front, ONE launch per level with a thread-block CLUSTER per front (extend-add over all warps
//                 of the cluster, then per 32 columns: rows below the block, cluster barrier, trailing 64 x 64 tiles while
//                 the chain CTA factors the next diagonal block in registers, cluster barrier).
//   The per-block launches the cluster kernel replaced stay behind QPALM_B200_MF_PER_BLOCK=1 (and as the fallback when a
//   cluster launch is refused): k_mf_extend (U_s <- 0, then panel_s / U_s += U_c scattered through rel_c for every child c in
//   a fixed order, each CTA owning a range of target columns => no atomics, bit-reproducible), then per 32-column block
//   k_mf_diag (32 x 32 Cholesky in registers, one warp per front), k_mf_trsm (rows below, one thread per row), k_mf_syrk
//   (trailing update of the rest of the panel and of U_s, 64 x 64 tiles, 4 x 4 per thread).  Same operations in the same
//   order: bit-identical factors.
// Solves: one CTA per supernode per level, per-supernode update vectors (pull-based): k_mf_fwd / k_mf_bwd, and for the levels
// of large fronts k_mf_fwd_big / k_mf_bwd_big, which issue every load of the (constant) factor ahead of the step that needs
// it -- only the right-hand side carries a dependence from block to block.
// Rank-k update/downdate: k_ud_mark marks the etree paths, k_ud_sweep walks them (one CTA, columns in ascending order).
#include "sparse.cuh"
#include "sparse_host.h"
#include "chol32.cuh"

#include <string.h>
#include <vector>

namespace qb {

struct SpDev {
  int n, nsuper;
  const int *perm, *iperm, *sn_of_col, *first, *rows_off, *rowidx, *rel, *sn_parent, *child_ptr, *child_idx, *lvl_sn;
  const long long *panel_off, *upd_off;
  const int *sgn;   // nullptr: positive definite (L L').  Otherwise the a-priori pivot signs per PERMUTED column of a
                    // quasi-definite matrix (KKT path): the factorization is L S L', S = diag(sgn)
};
__device__ __forceinline__ double sgn_of(const SpDev &d, int pcol) { return d.sgn ? (double)d.sgn[pcol] : 1.0; }

struct SparseChol {
  SymHost h;
  SparseCholInfo info;
  SpDev d;
  std::vector<void *> allocs;
  double *upd = nullptr;      // update blocks (sum nr^2)
  double *uvec = nullptr;     // per-supernode update vectors of the solves (sum nr)
  double *v = nullptr;        // permuted rhs / solution (n)
  double *W = nullptr;        // rank-k panel, n x 8 permuted
  int *mark = nullptr;        // nsuper
  int small_nf = 0;           // fronts up to this size use the shared-memory kernel
  int small_smem_max = 0;
  bool no_cluster = false;    // set when a cluster launch of mfc::k_mf_front was refused on this device / partition
};

template <typename T>
static int up_vec(SparseChol *sc, const T **dst, const std::vector<T> &src) {
  void *p = nullptr;
  const size_t bytes = sizeof(T) * (src.size() ? src.size() : 1);
  QB_CUDA_TRY(cudaMalloc(&p, bytes));
  if (!src.empty()) QB_CUDA_TRY(cudaMemcpy(p, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice));
  sc->allocs.push_back(p);
  *dst = (const T *)p;
  return 0;
}
template <typename T>
static int dev_buf(SparseChol *sc, T **dst, size_t count) {
  void *p = nullptr;
  QB_CUDA_TRY(cudaMalloc(&p, sizeof(T) * (count ? count : 1)));
  QB_CUDA_TRY(cudaMemset(p, 0, sizeof(T) * (count ? count : 1)));
  sc->allocs.push_back(p);
  *dst = (T *)p;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// small-front kernel configuration
// ------------------------------------------------------------------------------------------------
constexpr int kSmallMaxNf = 160;   // (160 * 161 / 2 + pad) * 8 B of shared memory < 227 KB
__host__ __device__ inline int tri_ld(int nf) { return nf | 1; }   // odd leading dimension: conflict-free column walks

static int upload_symbolic(SparseChol *sc, int n);

int sparse_chol_analyze(SparseChol **out, int n, int m, const int *Acsc_p, const int *Acsc_i, const int *Acsr_p,
                        const int *Acsr_j, const long long *Qp, const long long *Qi, bool force, cudaStream_t stream) {
  *out = nullptr;
  SparseChol *sc = new SparseChol();
  const int r = symbolic_analyze(n, m, Acsc_p, Acsc_i, Acsr_p, Acsr_j, Qp, Qi, force, &sc->h);
  if (r == 1) { delete sc; return 0; }
  if (r != 0) { delete sc; return 3; }
  SymHost &h = sc->h;
  // the sparse path pays off only when the factor is much smaller than the dense triangle
  const double dense_tri = 0.5 * (double)n * ((double)n + 1.0);
  if (!force && ((double)h.nnzL > 0.30 * dense_tri || n < 512)) { delete sc; return 0; }
  size_t free_b = 0, total_b = 0;
  QB_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
  const double need = 8.0 * ((double)h.nnzL * 2 + (double)h.upd_total) + (1u << 28);
  if (need > (double)free_b) {
    fprintf(stderr, "[qpalm_b200] sparse Newton path needs %.1f GB (factor %.1f GB, update blocks %.1f GB) but only %.1f GB "
                    "of HBM is free\n", need / 1e9, 8.0 * h.nnzL / 1e9, 8.0 * h.upd_total / 1e9, free_b / 1e9);
    delete sc; return 2;
  }
  if (int rc = upload_symbolic(sc, n)) { sparse_chol_destroy(sc); return rc; }
  (void)stream;
  *out = sc;
  return 0;
}

// Symbolic analysis of a general symmetric QUASI-DEFINITE pattern (the KKT matrix [Q + I/gamma, A_J'; A_J, -inv(Sigma_J)] of
// src/solver_interface.c:119-170) given by its lower triangle (int64 CSC, N x N); columns >= n_pos carry negative pivots.
// Vanderbei: every symmetric permutation of a quasi-definite matrix has an L D L' factorization whose pivot signs are those
// of the diagonal blocks, so the signs are known before the numeric phase and no pivoting is needed.
int sparse_ldl_analyze(SparseChol **out, int N, int n_pos, const long long *Kp, const long long *Ki, cudaStream_t stream) {
  *out = nullptr;
  SparseChol *sc = new SparseChol();
  std::vector<int> zero_p((size_t)N + 1, 0), none(1, 0);
  const int r = symbolic_analyze(N, 0, zero_p.data(), none.data(), none.data(), none.data(), Kp, Ki, true, &sc->h);
  if (r != 0) { delete sc; return 3; }
  size_t free_b = 0, total_b = 0;
  QB_CUDA_TRY(cudaMemGetInfo(&free_b, &total_b));
  const double need = 8.0 * ((double)sc->h.nnzL + (double)sc->h.upd_total) + (1u << 28);
  if (need > (double)free_b) { fprintf(stderr, "[qpalm_b200] KKT factor needs %.1f GB of HBM\n", need / 1e9); delete sc; return 2; }
  if (int rc = upload_symbolic(sc, N)) { sparse_chol_destroy(sc); return rc; }
  std::vector<int> sg((size_t)N);
  for (int pj = 0; pj < N; pj++) sg[pj] = sc->h.perm[pj] >= n_pos ? -1 : 1;
  if (int rc = up_vec(sc, &sc->d.sgn, sg)) { sparse_chol_destroy(sc); return rc; }
  (void)stream;
  *out = sc;
  return 0;
}

static int upload_symbolic(SparseChol *sc, int n) {
  SymHost &h = sc->h;
  SparseCholInfo &I = sc->info;
  I.n = n; I.nsuper = h.nsuper; I.nlevels = h.nlevels; I.max_ns = h.max_ns; I.max_nf = h.max_nf;
  I.nnzL = h.nnzL; I.upd_entries = h.upd_total; I.nnzS = h.nnzS; I.flops = h.flops;
  SpDev &d = sc->d;
  d.n = n; d.nsuper = h.nsuper;
  int rc = 0;
  rc |= up_vec(sc, &d.perm, h.perm); rc |= up_vec(sc, &d.iperm, h.iperm); rc |= up_vec(sc, &d.sn_of_col, h.sn_of_col);
  rc |= up_vec(sc, &d.first, h.sn_first); rc |= up_vec(sc, &d.rows_off, h.rows_off); rc |= up_vec(sc, &d.rowidx, h.rowidx);
  rc |= up_vec(sc, &d.rel, h.rel); rc |= up_vec(sc, &d.sn_parent, h.sn_parent); rc |= up_vec(sc, &d.child_ptr, h.child_ptr);
  rc |= up_vec(sc, &d.child_idx, h.child_idx); rc |= up_vec(sc, &d.lvl_sn, h.lvl_sn);
  rc |= up_vec(sc, &d.panel_off, h.panel_off); rc |= up_vec(sc, &d.upd_off, h.upd_off);
  rc |= dev_buf(sc, &sc->upd, (size_t)h.upd_total);
  rc |= dev_buf(sc, &sc->uvec, h.rowidx.size());
  rc |= dev_buf(sc, &sc->v, (size_t)n);
  rc |= dev_buf(sc, &sc->W, (size_t)n * 8);
  rc |= dev_buf(sc, &sc->mark, (size_t)h.nsuper);
  d.sgn = nullptr;
  return rc;
}

void sparse_chol_destroy(SparseChol *sc) {
  if (!sc) return;
  for (void *p : sc->allocs) cudaFree(p);
  delete sc;
}
const SparseCholInfo *sparse_chol_info(const SparseChol *sc) { return &sc->info; }
size_t sparse_chol_factor_doubles(const SparseChol *sc) { return (size_t)sc->h.nnzL; }

// ------------------------------------------------------------------------------------------------
// position of permuted row pr in the front of supernode s (binary search in R_s beyond the own columns)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int front_pos(const SpDev &d, int f, int ns, int ro, int nr, int pr) {
  if (pr < f + ns) return pr - f;
  int lo = 0, hi = nr;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (d.rowidx[ro + mid] < pr) lo = mid + 1; else hi = mid; }
  return ns + lo;
}

// ================================================================================================
// assembly of H = Q + A_J' Sigma_J A_J + beta I into the panels: one warp per (permuted) column
// ================================================================================================
__global__ void __launch_bounds__(256)
k_sp_assemble(SpDev d, double *panels, int with_Q, const int *__restrict__ Qp, const int *__restrict__ Qi,
              const double *__restrict__ Qx, const int *__restrict__ Cp, const int *__restrict__ Ci,
              const double *__restrict__ Cx, const int *__restrict__ Rp, const int *__restrict__ Rj,
              const double *__restrict__ Rx, const int *__restrict__ active, const double *__restrict__ sigma, double beta) {
  const int lane = threadIdx.x & 31;
  const int pj = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pj >= d.n) return;
  const int j = d.perm[pj], s = d.sn_of_col[pj], f = d.first[s], ns = d.first[s + 1] - f, ro = d.rows_off[s],
            nr = d.rows_off[s + 1] - ro, nf = ns + nr;
  double *col = panels + d.panel_off[s] + (size_t)(pj - f) * nf;
  for (int i = lane; i < nf; i += 32) col[i] = 0.0;
  __syncwarp();
  if (with_Q) {
    for (int k = Qp[j] + lane; k < Qp[j + 1]; k += 32) {
      const int pc = d.iperm[Qi[k]];
      if (pc >= pj) col[front_pos(d, f, ns, ro, nr, pc)] = Qx[k];
    }
    __syncwarp();
  }
  if (active) {
    for (int k = Cp[j]; k < Cp[j + 1]; k++) {
      const int r = Ci[k];
      if (!active[r]) continue;
      const double v = sigma[r] * Cx[k];
      for (int t = Rp[r] + lane; t < Rp[r + 1]; t += 32) {
        const int pc = d.iperm[Rj[t]];
        if (pc >= pj) col[front_pos(d, f, ns, ro, nr, pc)] += v * Rx[t];
      }
      __syncwarp();
    }
  }
  if (lane == 0) col[pj - f] += beta;
}

int sparse_chol_assemble(SparseChol *sc, cudaStream_t s, double *panels, bool with_Q, const int *Qp, const int *Qi,
                         const double *Qx, const int *Acsc_p, const int *Acsc_i, const double *Acsc_x, const int *Acsr_p,
                         const int *Acsr_j, const double *Acsr_x, const int *active, const double *sigma, double beta) {
  QB_LAUNCH(k_sp_assemble, cdiv(sc->d.n, 8), 256, 0, s, sc->d, panels, with_Q ? 1 : 0, Qp, Qi, Qx, Acsc_p, Acsc_i, Acsc_x,
            Acsr_p, Acsr_j, Acsr_x, active, sigma, beta);
  QB_CUDA_TRY(cudaGetLastError());
  return 0;
}

// ================================================================================================
// numeric factorization, general path
// ================================================================================================
struct Front {
  int s, f, ns, ro, nr, nf;
  double *P, *U;
};
__device__ __forceinline__ Front load_front(const SpDev &d, double *panels, double *upd, int s) {
  Front F;
  F.s = s; F.f = d.first[s]; F.ns = d.first[s + 1] - F.f; F.ro = d.rows_off[s]; F.nr = d.rows_off[s + 1] - F.ro;
  F.nf = F.ns + F.nr; F.P = panels + d.panel_off[s]; F.U = upd + d.upd_off[s];
  return F;
}
__device__ __forceinline__ double *front_at(const Front &F, int i, int j) {   // i >= j in front index space
  return (j < F.ns) ? F.P + (size_t)i + (size_t)j * F.nf : F.U + (size_t)(i - F.ns) + (size_t)(j - F.ns) * F.nr;
}

constexpr int kExtT = 16;   // target columns per CTA in the extend-add
__global__ void __launch_bounds__(256) k_mf_extend(SpDev d, double *panels, double *upd, int lvl_begin) {
  const Front F = load_front(d, panels, upd, d.lvl_sn[lvl_begin + blockIdx.x]);
  const int c0 = blockIdx.y * kExtT, c1 = min(F.nf, c0 + kExtT);
  if (c0 >= F.nf) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int tc = max(c0, F.ns) + warp; tc < c1; tc += nw)
    for (int i = tc + lane; i < F.nf; i += 32) *front_at(F, i, tc) = 0.0;
  __syncthreads();
  for (int ch = d.child_ptr[F.s]; ch < d.child_ptr[F.s + 1]; ch++) {
    const int c = d.child_idx[ch];
    const int cro = d.rows_off[c], cnr = d.rows_off[c + 1] - cro;
    const int *rel = d.rel + cro;
    const double *Uc = upd + d.upd_off[c];
    int a, b;
    { int lo = 0, hi = cnr; while (lo < hi) { const int mid = (lo + hi) >> 1; if (rel[mid] < c0) lo = mid + 1; else hi = mid; } a = lo; }
    { int lo = a, hi = cnr; while (lo < hi) { const int mid = (lo + hi) >> 1; if (rel[mid] < c1) lo = mid + 1; else hi = mid; } b = lo; }
    for (int jc = a + warp; jc < b; jc += nw) {
      const int tc = rel[jc];
      for (int ic = jc + lane; ic < cnr; ic += 32) *front_at(F, rel[ic], tc) += Uc[(size_t)ic + (size_t)jc * cnr];
    }
    __syncthreads();
  }
}

// 32 x 32 (w x w) Cholesky of the diagonal block kb of every front of the level: one warp per front, lane = row.
// The row lives in registers with static indices (chol32.cuh: template-unrolled columns, branch-free sqrt/reciprocal);
// rows / columns >= w are an identity pad.
__global__ void __launch_bounds__(32) k_mf_diag(SpDev d, double *panels, int lvl_begin, int kb, int *info) {
  const int s = d.lvl_sn[lvl_begin + blockIdx.x];
  const int f = d.first[s], ns = d.first[s + 1] - f, nf = ns + d.rows_off[s + 1] - d.rows_off[s];
  const int k0 = kb * 32;
  if (k0 >= ns) return;
  const int w = min(32, ns - k0), lane = threadIdx.x;
  double *B = panels + d.panel_off[s] + (size_t)k0 + (size_t)k0 * nf;
  double row[32];
#pragma unroll
  for (int j = 0; j < 32; j++) row[j] = (lane < w && j <= lane) ? B[(size_t)lane + (size_t)j * nf] : ((j == lane) ? 1.0 : 0.0);
  double dl = 1.0, dinv = 1.0;
  int badcol = -1;
  if (d.sgn) {   // quasi-definite block: L S L' with the a-priori pivot signs
    const unsigned negmask = __ballot_sync(0xffffffffu, lane < w && d.sgn[f + k0 + lane] < 0);
    chol32::fstep_signed<0>(row, lane, dl, negmask, badcol);
  } else chol32::fstep<0>(row, lane, dl, dinv, badcol);
#pragma unroll
  for (int j = 0; j < 32; j++) if (lane < w && j < lane) B[(size_t)lane + (size_t)j * nf] = row[j];
  if (lane < w) B[(size_t)lane + (size_t)lane * nf] = dl;
  if (badcol >= 0 && badcol < w && lane == 0) atomicExch(info, 1 + f + k0);
}

// rows below diagonal block kb: X <- X * Lkk^{-T}, one thread per row
__global__ void __launch_bounds__(128) k_mf_trsm(SpDev d, double *panels, int lvl_begin, int kb) {
  const int s = d.lvl_sn[lvl_begin + blockIdx.x];
  const int f = d.first[s], ns = d.first[s + 1] - f, nf = ns + d.rows_off[s + 1] - d.rows_off[s];
  const int k0 = kb * 32;
  if (k0 >= ns) return;
  const int w = min(32, ns - k0);
  const int i0 = k0 + w + blockIdx.y * 128;
  if (i0 >= nf) return;
  __shared__ double Ls[32][33];
  double *P = panels + d.panel_off[s];
  for (int t = threadIdx.x; t < 32 * 32; t += 128) {
    const int r = t & 31, c = t >> 5;
    double v = (r < w && c <= r) ? P[(size_t)(k0 + r) + (size_t)(k0 + c) * nf] : ((r == c) ? 1.0 : 0.0);
    if (r == c) v = 1.0 / v;          // the diagonal is stored inverted: one division per column instead of one per element
    if (c < w) v *= sgn_of(d, f + k0 + c);   // L S L': X = A21 inv(L11') inv(S), i.e. column c of the triangle carries s_c
    Ls[r][c] = v;
  }
  __syncthreads();
  const int i = i0 + threadIdx.x;
  if (i >= nf) return;
  double x[32];
#pragma unroll
  for (int t = 0; t < 32; t++) x[t] = (t < w) ? P[(size_t)i + (size_t)(k0 + t) * nf] : 0.0;
#pragma unroll
  for (int t = 0; t < 32; t++) {
    double a = x[t];
#pragma unroll
    for (int u = 0; u < t; u++) a = fma(-x[u], Ls[t][u], a);
    x[t] = a * Ls[t][t];
  }
#pragma unroll
  for (int t = 0; t < 32; t++) if (t < w) P[(size_t)i + (size_t)(k0 + t) * nf] = x[t];
}

// trailing update after block kb: F(i, j) -= sum_t P(i, k0 + t) P(j, k0 + t), i >= j >= k0 + w; 64 x 64 tiles
__global__ void __launch_bounds__(256) k_mf_syrk(SpDev d, double *panels, double *upd, int lvl_begin, int kb) {
  if (blockIdx.z > blockIdx.y) return;   // tile (y, z): y = row tile, z = column tile, lower part only
  const Front F = load_front(d, panels, upd, d.lvl_sn[lvl_begin + blockIdx.x]);
  const int k0 = kb * 32;
  if (k0 >= F.ns) return;
  const int w = min(32, F.ns - k0), base = k0 + w;
  const int i0 = base + blockIdx.y * 64, j0 = base + blockIdx.z * 64;
  if (i0 >= F.nf || j0 >= F.nf) return;
  __shared__ double As[32][65], Bs[32][65];
  for (int t = threadIdx.x; t < 32 * 64; t += 256) {
    const int r = t & 63, c = t >> 6;
    As[c][r] = (c < w && i0 + r < F.nf) ? F.P[(size_t)(i0 + r) + (size_t)(k0 + c) * F.nf] : 0.0;
    Bs[c][r] = (c < w && j0 + r < F.nf) ? sgn_of(d, F.f + k0 + c) * F.P[(size_t)(j0 + r) + (size_t)(k0 + c) * F.nf] : 0.0;   // F -= P S P'
  }
  __syncthreads();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;   // rows tx + 16 a, cols ty + 16 b
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; a++)
#pragma unroll
    for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
#pragma unroll 8
  for (int t = 0; t < 32; t++) {
    double av[4], bv[4];
#pragma unroll
    for (int a = 0; a < 4; a++) { av[a] = As[t][tx + 16 * a]; bv[a] = Bs[t][ty + 16 * a]; }
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
      for (int b = 0; b < 4; b++) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
  }
#pragma unroll
  for (int b = 0; b < 4; b++) {
    const int j = j0 + ty + 16 * b;
    if (j >= F.nf) continue;
#pragma unroll
    for (int a = 0; a < 4; a++) {
      const int i = i0 + tx + 16 * a;
      if (i < F.nf && i >= j) { double *p = front_at(F, i, j); *p -= acc[a][b]; }
    }
  }
}

// ================================================================================================
// numeric factorization, small fronts: whole front in shared memory (packed columns, ld = nf | 1)
// ================================================================================================
__global__ void __launch_bounds__(256) k_mf_small(SpDev d, double *panels, double *upd, int lvl_begin, int *info) {
  extern __shared__ double Fs[];
  const Front F = load_front(d, panels, upd, d.lvl_sn[lvl_begin + blockIdx.x]);
  const int nf = F.nf, ns = F.ns, ld = tri_ld(nf);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  // load the panel, zero the update part (full square storage, lower part used)
  for (int j = warp; j < nf; j += nw)
    for (int i = j + lane; i < nf; i += 32) Fs[i + j * ld] = (j < ns) ? F.P[(size_t)i + (size_t)j * nf] : 0.0;
  __syncthreads();
  for (int ch = d.child_ptr[F.s]; ch < d.child_ptr[F.s + 1]; ch++) {
    const int c = d.child_idx[ch];
    const int cro = d.rows_off[c], cnr = d.rows_off[c + 1] - cro;
    const int *rel = d.rel + cro;
    const double *Uc = upd + d.upd_off[c];
    for (int jc = warp; jc < cnr; jc += nw) {
      const int tc = rel[jc];
      for (int ic = jc + lane; ic < cnr; ic += 32) Fs[rel[ic] + tc * ld] += Uc[(size_t)ic + (size_t)jc * cnr];
    }
    __syncthreads();
  }
  // right-looking partial Cholesky of the first ns columns
  __shared__ int bad;
  if (tid == 0) bad = 0;
  for (int k = 0; k < ns; k++) {
    __syncthreads();
    const double sk = sgn_of(d, F.f + k);      // +1 (L L') or the a-priori pivot sign (L S L')
    const double dkk = sk * Fs[k + k * ld];
    if (!(dkk > 0.0) && tid == 0) bad = 1 + F.f + k;
    const double l = sqrt(dkk), linv = sk / l;
    __syncthreads();
    for (int i = k + tid; i < nf; i += blockDim.x) Fs[i + k * ld] = (i == k) ? l : Fs[i + k * ld] * linv;
    __syncthreads();
    // trailing update: columns j > k, rows i >= j
    for (int j = k + 1 + warp; j < nf; j += nw) {
      const double ljk = sk * Fs[j + k * ld];
      if (ljk == 0.0) continue;
      for (int i = j + lane; i < nf; i += 32) Fs[i + j * ld] = fma(-Fs[i + k * ld], ljk, Fs[i + j * ld]);
    }
  }
  __syncthreads();
  for (int j = warp; j < nf; j += nw) {
    if (j < ns) { for (int i = j + lane; i < nf; i += 32) F.P[(size_t)i + (size_t)j * nf] = Fs[i + j * ld]; }
    else { for (int i = j + lane; i < nf; i += 32) F.U[(size_t)(i - ns) + (size_t)(j - ns) * F.nr] = Fs[i + j * ld]; }
  }
  if (tid == 0 && bad) atomicExch(info, bad);
}

// ================================================================================================
// numeric factorization, large fronts: ONE launch per level, one thread-block CLUSTER per front
// ================================================================================================
// The per-block launches above (k_mf_extend + per 32 columns k_mf_diag / k_mf_trsm / k_mf_syrk: 97 dependent block steps
// at ~46 us each on the n = 90 000 grid problem) become one launch per level: the CTAs of a cluster share one front, meet
// at hardware cluster barriers (release / acquire; data crosses CTAs through L2, read with ld.global.cg) and fronts of a
// level no longer wait for each other between steps.  CTA 0 of a multi-CTA cluster is the CHAIN CTA: while the others
// apply block kb's trailing update it brings the NEXT 32 x 32 diagonal block up to date and factors it in registers, so
// the one-warp factorization is off the other CTAs' critical path.  Every floating-point operation is the one the
// per-block kernels perform, in the same order: the two paths give bit-identical factors (tests/test_gpu_sparse.py).
namespace mfc {
constexpr int NT = 256;
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ unsigned cta_rank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cta_count() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_id() { unsigned r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }

struct Smem {
  double Ls[32][33];
  union {
    struct { double As[32][65], Bs[32][65]; } t;     // trailing-update tile operands
    struct { double Sl[32][33], Dn[32][33]; } c;     // chain CTA: slab of the next block's rows, next diagonal block
  };
};

// 32 x 32 in-register factorization of the block held in B (global, ld nf) or Dn (shared), written to B
__device__ __forceinline__ void factor_block(const SpDev &d, double *B, int nf, int w, int pcol0, const double (*Dn)[33], int *info) {
  const int lane = threadIdx.x;
  double row[32];
#pragma unroll
  for (int j = 0; j < 32; j++)
    row[j] = (lane < w && j <= lane) ? (Dn ? Dn[lane][j] : __ldcg(B + (size_t)lane + (size_t)j * nf)) : ((j == lane) ? 1.0 : 0.0);
  double dl = 1.0, dinv = 1.0;
  int badcol = -1;
  if (d.sgn) {
    const unsigned negmask = __ballot_sync(0xffffffffu, lane < w && d.sgn[pcol0 + lane] < 0);
    chol32::fstep_signed<0>(row, lane, dl, negmask, badcol);
  } else chol32::fstep<0>(row, lane, dl, dinv, badcol);
#pragma unroll
  for (int j = 0; j < 32; j++) if (lane < w && j < lane) B[(size_t)lane + (size_t)j * nf] = row[j];
  if (lane < w) B[(size_t)lane + (size_t)lane * nf] = dl;
  if (badcol >= 0 && badcol < w && lane == 0) atomicExch(info, 1 + pcol0);
}

// optional phase clocks (QPALM_B200_MF_CLOCKS=1; tools/prof_config.py prints them): thread 0 of CTA 0 and of CTA 1 of the first
// front of each level add the clocks spent per phase
__device__ unsigned long long g_mf_clk[16];
#define MF_CLK(slot, who)                                                                                              \
  do {                                                                                                                 \
    if (clk_on && tid == 0 && cr == (who)) { const long long now_ = clock64(); atomicAdd(&g_mf_clk[slot], (unsigned long long)(now_ - t_last)); t_last = now_; } \
  } while (0)

__global__ void __launch_bounds__(NT) k_mf_front(SpDev d, double *panels, double *upd, int lvl_begin, int *info, int clocks) {
  __shared__ Smem sm;
  const int cr = (int)cta_rank(), cn = (int)cta_count();
  const bool clk_on = clocks && cluster_id() == 0;
  long long t_last = clock64();
  const Front F = load_front(d, panels, upd, d.lvl_sn[lvl_begin + (int)cluster_id()]);
  const int nf = F.nf, ns = F.ns;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = NT / 32;

  // ---- extend-add.  Within one child every target entry receives exactly one source entry, so the child's lower triangle is
  //      cut into (32-row segment, source column) items dealt round-robin to ALL warps of the cluster, eight items in flight
  //      per warp; the children follow each other in their fixed order with a cluster barrier between them (no atomics,
  //      bit-reproducible).  (Column-per-warp with an owner CTA per target column left most warps idle on the small clusters:
  //      75 us per level against 33 us for the per-block kernel's full grid.)
  const int gw = cr * NW + warp, GW = cn * NW;
  for (int tc = ns + gw; tc < nf; tc += GW)
    for (int i = tc + lane; i < nf; i += 32) F.U[(size_t)(i - ns) + (size_t)(tc - ns) * F.nr] = 0.0;
  cluster_sync();
  for (int ch = d.child_ptr[F.s]; ch < d.child_ptr[F.s + 1]; ch++) {
    const int c = d.child_idx[ch];
    const int cro = d.rows_off[c], cnr = d.rows_off[c + 1] - cro;
    const int *rel = d.rel + cro;
    const double *Uc = upd + d.upd_off[c];
    const int nS = (cnr + 31) >> 5, nitem = 32 * (nS * (nS + 1) / 2);   // item -> (S, jc): rows 32 S + lane, column jc < 32 (S + 1)
    for (int it0 = gw; it0 < nitem; it0 += GW * 8) {
      double *pt[8];
      double sv[8], tv[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int it = it0 + GW * u, blk = it >> 5;
        int S = (int)((sqrtf(8.0f * (float)blk + 1.0f) - 1.0f) * 0.5f);
        while ((S + 1) * (S + 2) / 2 <= blk) S++;
        while (S * (S + 1) / 2 > blk) S--;
        const int jc = ((blk - S * (S + 1) / 2) << 5) + (it & 31), ic = (S << 5) + lane;
        const bool ok = it < nitem && jc < cnr && ic < cnr && ic >= jc;
        pt[u] = ok ? front_at(F, rel[ic], rel[jc]) : nullptr;
        sv[u] = ok ? Uc[(size_t)ic + (size_t)jc * cnr] : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 8; u++) tv[u] = pt[u] ? __ldcg(pt[u]) : 0.0;
#pragma unroll
      for (int u = 0; u < 8; u++) if (pt[u]) *pt[u] = tv[u] + sv[u];
    }
    cluster_sync();
  }
  MF_CLK(0, 0);

  // ---- block 0 of the panel
  if (cr == 0 && warp == 0) factor_block(d, F.P, nf, min(32, ns), F.f, nullptr, info);
  cluster_sync();
  MF_CLK(2, 0);
  if (clk_on && tid == 0 && cr == 1) t_last = clock64();

  const int nkb = (ns + 31) / 32;
  for (int kb = 0; kb < nkb; kb++) {
    const int k0 = kb * 32, w = min(32, ns - k0), base = k0 + w;
    if (base >= nf) break;                       // nothing below the last block (root front); uniform over the cluster
    double *P = F.P;
    // (a) the factored diagonal block, diagonal inverted, column signs folded in (as k_mf_trsm)
    for (int t = tid; t < 32 * 32; t += NT) {
      const int r = t & 31, c = t >> 5;
      double v = (r < w && c <= r) ? __ldcg(P + (size_t)(k0 + r) + (size_t)(k0 + c) * nf) : ((r == c) ? 1.0 : 0.0);
      if (r == c) v = 1.0 / v;
      if (c < w) v *= sgn_of(d, F.f + k0 + c);
      sm.Ls[r][c] = v;
    }
    __syncthreads();
    MF_CLK(3, 0);
    // (b) rows below: X <- X inv(L11') inv(S); 32-row chunks round-robin over the warps of the cluster, lane = row.
    // The first chunk -- the rows of the NEXT diagonal block -- is warp 0 of the chain CTA's: it keeps a copy in shared memory
    // (the slab), so the chain CTA starts the look-ahead from its own data without waiting for the other CTAs' rows.
    const bool lookahead = cn > 1 && base < ns;   // the chain CTA owns the next diagonal block
    const bool chain_ahead = lookahead && cr == 0;
    const int wn = min(32, ns - base);
    double dold[4];
    if (chain_ahead) {                             // old values of the next diagonal block (final since the previous barrier)
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int c = warp + 8 * q;
        dold[q] = (lane < wn && c <= lane) ? __ldcg(P + (size_t)(base + lane) + (size_t)(base + c) * nf) : 0.0;
      }
    }
    for (int i0 = base + (cr * NW + warp) * 32; i0 < nf; i0 += cn * NW * 32) {
      const int i = i0 + lane;
      const bool slab = chain_ahead && i0 == base;          // (cr == 0, warp == 0, first chunk)
      if (slab && i >= nf) {
#pragma unroll
        for (int t = 0; t < 32; t++) sm.c.Sl[t][lane] = 0.0;
      }
      if (i < nf) {
        double x[32];
#pragma unroll
        for (int t = 0; t < 32; t++) x[t] = (t < w) ? __ldcg(P + (size_t)i + (size_t)(k0 + t) * nf) : 0.0;
        // right-looking order: as soon as x[u] is final every later column takes its term -- 31 - u independent FMAs per step
        // instead of one 528-long dependent chain; each x[t] still accumulates u = 0, 1, ... in order (same bits as k_mf_trsm)
#pragma unroll
        for (int u = 0; u < 32; u++) {
          x[u] *= sm.Ls[u][u];
#pragma unroll
          for (int t = u + 1; t < 32; t++) x[t] = fma(-x[u], sm.Ls[t][u], x[t]);
        }
#pragma unroll
        for (int t = 0; t < 32; t++) if (t < w) P[(size_t)i + (size_t)(k0 + t) * nf] = x[t];
        if (slab) {                                          // Sl[c][r] = P(base + r, k0 + c), r < wn
#pragma unroll
          for (int t = 0; t < 32; t++) sm.c.Sl[t][lane] = (t < w && lane < wn) ? x[t] : 0.0;
        }
      }
    }
    MF_CLK(4, 0);
    if (chain_ahead) cluster_arrive(); else cluster_sync();          // the chain CTA completes this barrier after its look-ahead
    MF_CLK(5, 0);
    if (clk_on && tid == 0 && cr == 1) t_last = clock64();
    // (c) trailing update F(i, j) -= sum_t P(i, k0 + t) s_t P(j, k0 + t), i >= j >= base
    if (cn > 1 && cr == 0) {
      if (lookahead) {
        __syncthreads();                                     // the slab (warp 0) is visible to the CTA
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const int c = warp + 8 * q;
          double acc = 0.0;
#pragma unroll 8
          for (int t = 0; t < 32; t++) acc = fma(sm.c.Sl[t][lane], sgn_of(d, F.f + k0 + min(t, w - 1)) * sm.c.Sl[t][c], acc);
          sm.c.Dn[lane][c] = dold[q] - acc;
        }
        __syncthreads();
        if (warp == 0) factor_block(d, P + (size_t)base + (size_t)base * nf, nf, wn, F.f + base, sm.c.Dn, info);
        cluster_wait();
      }
    } else {
      const int T = (nf - base + 63) / 64, ntile = T * (T + 1) / 2;
      const int nworker = cn > 1 ? cn - 1 : 1, me = cn > 1 ? cr - 1 : 0;
      const int tx = tid & 15, ty = tid >> 4;
      for (int idx = me; idx < ntile; idx += nworker) {
        int by = (int)((sqrtf(8.0f * (float)idx + 1.0f) - 1.0f) * 0.5f);
        while ((by + 1) * (by + 2) / 2 <= idx) by++;
        while (by * (by + 1) / 2 > idx) by--;
        const int bz = idx - by * (by + 1) / 2;
        const int i0 = base + by * 64, j0 = base + bz * 64;
        __syncthreads();      // the previous tile's operands are no longer needed
        for (int t = tid; t < 32 * 64; t += NT) {
          const int r = t & 63, c = t >> 6;
          sm.t.As[c][r] = (c < w && i0 + r < nf) ? __ldcg(P + (size_t)(i0 + r) + (size_t)(k0 + c) * nf) : 0.0;
          sm.t.Bs[c][r] = (c < w && j0 + r < nf) ? sgn_of(d, F.f + k0 + c) * __ldcg(P + (size_t)(j0 + r) + (size_t)(k0 + c) * nf) : 0.0;
        }
        // old values of the tile requested before the products.  One base pointer per tile column (entry (i, j) = colp[b] + i, in
        // the panel for j < ns, in the update block beyond): the per-entry address selection of front_at() -- two 64-bit
        // multiplies and a branch for each of the 32 accesses of a thread -- was a quarter of a tile's instructions.
        double *colp[4];
#pragma unroll
        for (int b = 0; b < 4; b++) {
          const int j = j0 + ty + 16 * b;
          colp[b] = (j < ns) ? F.P + (size_t)j * nf : F.U + (size_t)(j - ns) * F.nr - ns;
        }
        double old[4][4];
#pragma unroll
        for (int b = 0; b < 4; b++)
#pragma unroll
          for (int a = 0; a < 4; a++) {
            const int i = i0 + tx + 16 * a, j = j0 + ty + 16 * b;
            old[a][b] = (i < nf && j < nf && i >= j) ? __ldcg(colp[b] + i) : 0.0;
          }
        __syncthreads();
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
          for (int b = 0; b < 4; b++) acc[a][b] = 0.0;
#pragma unroll 8
        for (int t = 0; t < 32; t++) {
          double av[4], bv[4];
#pragma unroll
          for (int a = 0; a < 4; a++) { av[a] = sm.t.As[t][tx + 16 * a]; bv[a] = sm.t.Bs[t][ty + 16 * a]; }
#pragma unroll
          for (int a = 0; a < 4; a++)
#pragma unroll
            for (int b = 0; b < 4; b++) acc[a][b] = fma(av[a], bv[b], acc[a][b]);
        }
#pragma unroll
        for (int b = 0; b < 4; b++) {
          const int j = j0 + ty + 16 * b;
          if (j >= nf) continue;
#pragma unroll
          for (int a = 0; a < 4; a++) {
            const int i = i0 + tx + 16 * a;
            if (i >= nf || i < j) continue;
            if (lookahead && i < base + wn) continue;      // the chain CTA writes the next diagonal block
            colp[b][i] = old[a][b] - acc[a][b];
          }
        }
      }
      if (cn == 1 && base < ns) {               // single-CTA cluster: factor the next block after its update
        __syncthreads();
        if (warp == 0) factor_block(d, P + (size_t)base + (size_t)base * nf, nf, min(32, ns - base), F.f + base, nullptr, info);
      }
    }
    MF_CLK(6, 0); MF_CLK(7, 1);
    cluster_sync();
    MF_CLK(8, 0); MF_CLK(9, 1);
    if (clk_on && tid == 0 && cr == 0) atomicAdd(&g_mf_clk[10], 1ull);
  }
  if (clk_on && tid == 0 && cr == 0) atomicAdd(&g_mf_clk[11], 1ull);
}
}  // namespace mfc

static int launch_front_level(SparseChol *sc, cudaStream_t st, double *panels, int lvl_begin, int cnt, int mnf, int *info_dev) {
  // cluster size: a power of two <= 16, one wave of CTAs (one per SM), no more CTAs than chain + trailing tiles need
  const int T = cdiv(max(mnf - 32, 1), 64), ntile = T * (T + 1) / 2;
  int cn = 1;
  while (cn < 16 && cnt * (cn * 2) <= 148 && cn < ntile + 1) cn *= 2;
  static int max_cluster = -1;
  if (max_cluster < 0) {
    max_cluster = 8;
    if (cudaFuncSetAttribute(mfc::k_mf_front, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) == cudaSuccess) max_cluster = 16;
    else (void)cudaGetLastError();
  }
  int cap = max_cluster;
  if (const char *e = getenv("QPALM_B200_MF_CLUSTER")) cap = max(1, min(max_cluster, atoi(e)));   // tests: force small clusters
  while (cn > cap) cn /= 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(cnt * cn)); cfg.blockDim = dim3(mfc::NT); cfg.dynamicSmemBytes = 0; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cn; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  const bool prof = g_prof_on && prof_begin("k_mf_front", st);
  static const int clocks = getenv("QPALM_B200_MF_CLOCKS") ? 1 : 0;
  const cudaError_t err = cudaLaunchKernelEx(&cfg, mfc::k_mf_front, sc->d, panels, sc->upd, lvl_begin, info_dev, clocks);
  if (prof) prof_end(st);
  if (err != cudaSuccess) {   // e.g. a partition that cannot co-schedule the cluster: the caller takes the per-block launches
    (void)cudaGetLastError();
    return 1;
  }
  ++g_kernel_launches;
  return 0;
}

// debug: reads and clears the phase clocks of mfc::k_mf_front (12 counters)
extern "C" int qpalm_b200_mf_clocks(unsigned long long *out16) {
  QB_CUDA_TRY(cudaDeviceSynchronize());
  QB_CUDA_TRY(cudaMemcpyFromSymbol(out16, mfc::g_mf_clk, sizeof(unsigned long long) * 16));
  unsigned long long z[16] = {0};
  QB_CUDA_TRY(cudaMemcpyToSymbol(mfc::g_mf_clk, z, sizeof(z)));
  return 0;
}

int sparse_chol_factor(SparseChol *sc, cudaStream_t st, double *panels, int *info_dev) {
  const SymHost &h = sc->h;
  static bool attr_set = false;
  static bool per_block = false;      // QPALM_B200_MF_PER_BLOCK=1: the per-block launches (A/B and bit-identity tests)
  if (!attr_set) {
    QB_CUDA_TRY(cudaFuncSetAttribute(k_mf_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(double) * tri_ld(kSmallMaxNf) * kSmallMaxNf)));
    attr_set = true;
  }
  { const char *e = getenv("QPALM_B200_MF_PER_BLOCK"); per_block = e && atoi(e) != 0; }
  static bool dumped = false;
  const bool dump = !dumped && getenv("QPALM_B200_LEVELS") != nullptr;
  dumped = true;
  for (int l = 0; l < h.nlevels; l++) {
    const int b = h.lvl_ptr[l], cnt = h.lvl_ptr[l + 1] - b;
    const int mnf = h.lvl_max_nf[l], mns = h.lvl_max_ns[l];
    if (dump) fprintf(stderr, "[qpalm_b200 levels] level %d: fronts %d max nf %d max ns %d\n", l, cnt, mnf, mns);
    if (cnt <= 0) continue;
    if (mnf <= kSmallMaxNf) {
      const size_t smem = sizeof(double) * (size_t)tri_ld(mnf) * mnf;
      QB_LAUNCH(k_mf_small, cnt, mnf <= 32 ? 64 : (mnf <= 64 ? 128 : 256), smem, st, sc->d, panels, sc->upd, b, info_dev);
      continue;
    }
    if (!per_block && !sc->no_cluster) {
      const int rc = launch_front_level(sc, st, panels, b, cnt, mnf, info_dev);
      if (rc == 0) continue;
      if (rc < 0) return rc;
      sc->no_cluster = true;       // nothing of this level has run: fall through to the per-block launches, now and later
    }
    { dim3 g(cnt, cdiv(mnf, kExtT)); QB_LAUNCH(k_mf_extend, g, 256, 0, st, sc->d, panels, sc->upd, b); }
    for (int kb = 0; kb * 32 < mns; kb++) {
      QB_LAUNCH(k_mf_diag, cnt, 32, 0, st, sc->d, panels, b, kb, info_dev);
      const int rest = mnf - kb * 32 - 1;   // rows below the block (upper bound over the level)
      if (rest <= 0) continue;
      { dim3 g(cnt, cdiv(rest, 128)); QB_LAUNCH(k_mf_trsm, g, 128, 0, st, sc->d, panels, b, kb); }
      { const int T = cdiv(rest, 64); dim3 g(cnt, T, T); QB_LAUNCH(k_mf_syrk, g, 256, 0, st, sc->d, panels, sc->upd, b, kb); }
    }
  }
  QB_CUDA_TRY(cudaGetLastError());
  return 0;
}

// ================================================================================================
// solves
// ================================================================================================
__global__ void k_sp_permute_in(SpDev d, const double *__restrict__ rhs, double *v, double sgn) {
  const int pj = blockIdx.x * blockDim.x + threadIdx.x;
  if (pj < d.n) v[pj] = sgn * rhs[d.perm[pj]];
}
__global__ void k_sp_apply_sign(SpDev d, double *v) {
  const int pj = blockIdx.x * blockDim.x + threadIdx.x;
  if (pj < d.n) v[pj] *= (double)d.sgn[pj];
}
__global__ void k_sp_permute_out(SpDev d, const double *__restrict__ v, double *out) {
  const int pj = blockIdx.x * blockDim.x + threadIdx.x;
  if (pj < d.n) out[d.perm[pj]] = v[pj];
}

// forward: L y = b.  f (shared, nf) = [b_s ; 0] + children's update vectors; solve the ns x ns triangle in 32-column
// blocks; u_s = f[ns..) is handed to the parent.
// x / l from the correctly rounded reciprocal r = 1 / l: one multiply and two FMAs, no slow-path branch; the residual
// correction makes the quotient correctly rounded (Markstein), i.e. the value a division returns -- the ill-conditioned
// known-answer problems (tests/src/test_dua_inf_qp.c) flip iteration counts on a one-ulp difference here.
__device__ __forceinline__ double div_by_rcp(double x, double l, double r) {
  const double q = x * r;
  return fma(fma(-q, l, x), r, q);
}

__global__ void __launch_bounds__(256) k_mf_fwd(SpDev d, const double *__restrict__ panels, double *v, double *uvec, int lvl_begin) {
  extern __shared__ double fsh[];
  __shared__ double blk[32][33];
  __shared__ double dgl[32];
  const int s = d.lvl_sn[lvl_begin + blockIdx.x];
  const int f = d.first[s], ns = d.first[s + 1] - f, ro = d.rows_off[s], nr = d.rows_off[s + 1] - ro, nf = ns + nr;
  const double *P = panels + d.panel_off[s];
  const int tid = threadIdx.x;
  for (int i = tid; i < nf; i += blockDim.x) fsh[i] = (i < ns) ? v[f + i] : 0.0;
  __syncthreads();
  for (int ch = d.child_ptr[s]; ch < d.child_ptr[s + 1]; ch++) {
    const int c = d.child_idx[ch];
    const int cro = d.rows_off[c], cnr = d.rows_off[c + 1] - cro;
    for (int i = tid; i < cnr; i += blockDim.x) fsh[d.rel[cro + i]] += uvec[cro + i];
    __syncthreads();
  }
  for (int k0 = 0; k0 < ns; k0 += 32) {
    const int w = min(32, ns - k0);
    for (int t = tid; t < 32 * 32; t += blockDim.x) {
      const int r = t & 31, c = t >> 5;
      double e = (r < w && c <= r) ? P[(size_t)(k0 + r) + (size_t)(k0 + c) * nf] : ((r == c) ? 1.0 : 0.0);
      if (r == c) { dgl[r] = e; e = 1.0 / e; }     // inverted diagonal (one division per column, off the serial chain)
      blk[r][c] = e;
    }
    __syncthreads();
    if (tid < 32) {
      double x = (tid < w) ? fsh[k0 + tid] : 0.0;
#pragma unroll
      for (int j = 0; j < 32; j++) {
        const double xj = div_by_rcp(__shfl_sync(0xffffffffu, x, j), dgl[j], blk[j][j]);
        if (tid == j) x = xj; else if (tid > j) x = fma(-blk[tid][j], xj, x);
      }
      if (tid < w) fsh[k0 + tid] = x;
    }
    __syncthreads();
    for (int i = k0 + w + tid; i < nf; i += blockDim.x) {
      // 8 independent loads in flight per thread (a rolled loop pays one L2 round trip per column)
      const double *Pi = P + (size_t)i + (size_t)k0 * nf;
      double a0 = fsh[i], a1 = 0.0;
      int t = 0;
      for (; t + 8 <= w; t += 8) {
        double pv[8];
#pragma unroll
        for (int u = 0; u < 8; u++) pv[u] = Pi[(size_t)(t + u) * nf];
#pragma unroll
        for (int u = 0; u < 8; u += 2) { a0 = fma(-pv[u], fsh[k0 + t + u], a0); a1 = fma(-pv[u + 1], fsh[k0 + t + u + 1], a1); }
      }
      for (; t < w; t++) a0 = fma(-Pi[(size_t)t * nf], fsh[k0 + t], a0);
      fsh[i] = a0 + a1;
    }
    __syncthreads();
  }
  for (int i = tid; i < nf; i += blockDim.x) { if (i < ns) v[f + i] = fsh[i]; else uvec[ro + i - ns] = fsh[i]; }
}

// The same forward step for the levels of LARGE fronts (few CTAs, each streaming a panel of up to a few MB): only the
// right-hand side carries a dependence from block to block, the factor does not -- so every load of the factor is issued
// ahead of the step that needs it.  Per 32-column block: the values of this thread's first row below the block are requested
// BEFORE the one-warp substitution on the diagonal block and consumed after it, the next diagonal block travels through two
// registers per thread and is written to shared memory while the row update runs.  512 threads, so most steps have one row
// per thread.  (The generic kernel above paid, per block, one dependent L2 / HBM round trip for the diagonal block and two or
// three for the rows: 9-14 us per block at the top of the assembly tree against 4 us for the backward step.)  The arithmetic
// of every entry is that of the generic kernel, in the same order: identical bits.
constexpr int kFwdBigThreads = 512;
__global__ void __launch_bounds__(kFwdBigThreads) k_mf_fwd_big(SpDev d, const double *__restrict__ panels, double *v, double *uvec, int lvl_begin) {
  extern __shared__ double fsh[];
  __shared__ double blk[32][33];
  __shared__ double dgl[32];
  constexpr int NTB = kFwdBigThreads;
  const int s = d.lvl_sn[lvl_begin + blockIdx.x];
  const int f = d.first[s], ns = d.first[s + 1] - f, ro = d.rows_off[s], nr = d.rows_off[s + 1] - ro, nf = ns + nr;
  const double *P = panels + d.panel_off[s];
  const int tid = threadIdx.x;
  // element t (of 1024) of the 32 x 32 diagonal block at k0n, as the generic kernel loads it
  auto diag_elem = [&](int k0n, int t) -> double {
    const int r = t & 31, c = t >> 5, wn = min(32, ns - k0n);
    return (r < wn && c <= r) ? P[(size_t)(k0n + r) + (size_t)(k0n + c) * nf] : ((r == c) ? 1.0 : 0.0);
  };
  auto diag_store = [&](int t, double e) {
    const int r = t & 31, c = t >> 5;
    if (r == c) { dgl[r] = e; e = 1.0 / e; }
    blk[r][c] = e;
  };
  double nb0 = diag_elem(0, tid), nb1 = diag_elem(0, tid + NTB);      // in flight during the prologue
  for (int i = tid; i < nf; i += NTB) fsh[i] = (i < ns) ? v[f + i] : 0.0;
  __syncthreads();
  for (int ch = d.child_ptr[s]; ch < d.child_ptr[s + 1]; ch++) {
    const int c = d.child_idx[ch];
    const int cro = d.rows_off[c], cnr = d.rows_off[c + 1] - cro;
    for (int i = tid; i < cnr; i += NTB) fsh[d.rel[cro + i]] += uvec[cro + i];
    __syncthreads();
  }
  diag_store(tid, nb0); diag_store(tid + NTB, nb1);
  __syncthreads();
  for (int k0 = 0; k0 < ns; k0 += 32) {
    const int w = min(32, ns - k0), w8 = w & ~7;
    // requests that do not wait for the substitution below: this thread's first row under the block, the next diagonal block
    const int i1 = k0 + w + tid;
    double pv[32];
    {
      const double *Pi = P + (size_t)i1 + (size_t)k0 * nf;
#pragma unroll
      for (int u = 0; u < 32; u++) pv[u] = (i1 < nf && u < w) ? Pi[(size_t)u * nf] : 0.0;
    }
    const bool more = k0 + 32 < ns;
    if (more) { nb0 = diag_elem(k0 + 32, tid); nb1 = diag_elem(k0 + 32, tid + NTB); }
    if (tid < 32) {
      double x = (tid < w) ? fsh[k0 + tid] : 0.0;
#pragma unroll
      for (int j = 0; j < 32; j++) {
        const double xj = div_by_rcp(__shfl_sync(0xffffffffu, x, j), dgl[j], blk[j][j]);
        if (tid == j) x = xj; else if (tid > j) x = fma(-blk[tid][j], xj, x);
      }
      if (tid < w) fsh[k0 + tid] = x;
    }
    __syncthreads();
    if (i1 < nf) {
      double a0 = fsh[i1], a1 = 0.0;
#pragma unroll
      for (int u = 0; u < 32; u += 2) {
        if (u + 1 < w8) { a0 = fma(-pv[u], fsh[k0 + u], a0); a1 = fma(-pv[u + 1], fsh[k0 + u + 1], a1); }
      }
#pragma unroll
      for (int u = 0; u < 32; u++) if (u >= w8 && u < w) a0 = fma(-pv[u], fsh[k0 + u], a0);
      fsh[i1] = a0 + a1;
    }
    for (int i = i1 + NTB; i < nf; i += NTB) {      // fronts of more than 512 + 32 rows: the remaining rows
      const double *Pi = P + (size_t)i + (size_t)k0 * nf;
#pragma unroll
      for (int u = 0; u < 32; u++) pv[u] = (u < w) ? Pi[(size_t)u * nf] : 0.0;
      double a0 = fsh[i], a1 = 0.0;
#pragma unroll
      for (int u = 0; u < 32; u += 2) {
        if (u + 1 < w8) { a0 = fma(-pv[u], fsh[k0 + u], a0); a1 = fma(-pv[u + 1], fsh[k0 + u + 1], a1); }
      }
#pragma unroll
      for (int u = 0; u < 32; u++) if (u >= w8 && u < w) a0 = fma(-pv[u], fsh[k0 + u], a0);
      fsh[i] = a0 + a1;
    }
    if (more) { diag_store(tid, nb0); diag_store(tid + NTB, nb1); }     // blk / dgl are free: the substitution ended before the barrier
    __syncthreads();
  }
  for (int i = tid; i < nf; i += NTB) { if (i < ns) v[f + i] = fsh[i]; else uvec[ro + i - ns] = fsh[i]; }
}

// backward: L' x = y.  f = [y_s ; x(R_s)]; per 32-column block (descending): subtract the column dots with everything
// below the block (one warp per column), then the 32 x 32 transposed triangle.
__global__ void __launch_bounds__(256) k_mf_bwd(SpDev d, const double *__restrict__ panels, double *v, int lvl_begin) {
  extern __shared__ double fsh[];
  __shared__ double blk[32][33];
  __shared__ double dgl[32];
  const int s = d.lvl_sn[lvl_begin + blockIdx.x];
  const int f = d.first[s], ns = d.first[s + 1] - f, ro = d.rows_off[s], nr = d.rows_off[s + 1] - ro, nf = ns + nr;
  const double *P = panels + d.panel_off[s];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  for (int i = tid; i < nf; i += blockDim.x) fsh[i] = (i < ns) ? v[f + i] : v[d.rowidx[ro + i - ns]];
  __syncthreads();
  const int nb = (ns + 31) / 32;
  for (int kb = nb - 1; kb >= 0; kb--) {
    const int k0 = kb * 32, w = min(32, ns - k0);
    for (int t = tid; t < 32 * 32; t += blockDim.x) {
      const int r = t & 31, c = t >> 5;
      double e = (r < w && c <= r) ? P[(size_t)(k0 + r) + (size_t)(k0 + c) * nf] : ((r == c) ? 1.0 : 0.0);
      if (r == c) { dgl[r] = e; e = 1.0 / e; }
      blk[r][c] = e;
    }
    // A warp takes its (up to four with eight warps) columns TOGETHER: sixteen independent loads in flight per lane instead of four
    // (one column after the other paid nf / 128 dependent L2 round trips per column).  Per column the accumulation order is unchanged.
    for (int t0 = warp; t0 < w; t0 += 4 * nw) {
      const double *col[4];
      double a[4][4];
#pragma unroll
      for (int c = 0; c < 4; c++) {
        col[c] = P + (size_t)(k0 + min(t0 + c * nw, w - 1)) * nf;
#pragma unroll
        for (int q = 0; q < 4; q++) a[c][q] = 0.0;
      }
      int i = k0 + w + lane;
      for (; i + 96 < nf; i += 128) {
        double v[4][4];
#pragma unroll
        for (int c = 0; c < 4; c++)
#pragma unroll
          for (int q = 0; q < 4; q++) v[c][q] = col[c][i + 32 * q];
        const double f0 = fsh[i], f1 = fsh[i + 32], f2 = fsh[i + 64], f3 = fsh[i + 96];
#pragma unroll
        for (int c = 0; c < 4; c++) {
          a[c][0] = fma(v[c][0], f0, a[c][0]); a[c][1] = fma(v[c][1], f1, a[c][1]);
          a[c][2] = fma(v[c][2], f2, a[c][2]); a[c][3] = fma(v[c][3], f3, a[c][3]);
        }
      }
      for (; i < nf; i += 32) {
        const double f = fsh[i];
#pragma unroll
        for (int c = 0; c < 4; c++) a[c][0] = fma(col[c][i], f, a[c][0]);
      }
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const double sacc = warp_sum((a[c][0] + a[c][1]) + (a[c][2] + a[c][3]));
        if (lane == 0 && t0 + c * nw < w) fsh[k0 + t0 + c * nw] -= sacc;
      }
    }
    __syncthreads();
    if (tid < 32) {
      double x = (tid < w) ? fsh[k0 + tid] : 0.0;
#pragma unroll
      for (int j = 31; j >= 0; j--) {
        const double xj = div_by_rcp(__shfl_sync(0xffffffffu, x, j), dgl[j], blk[j][j]);
        if (tid == j) x = xj; else if (tid < j) x = fma(-blk[j][tid], xj, x);
      }
      if (tid < w) fsh[k0 + tid] = x;
    }
    __syncthreads();
  }
  for (int i = tid; i < ns; i += blockDim.x) v[f + i] = fsh[i];
}

// The backward step for the levels of LARGE fronts, with the factor loads issued ahead as in k_mf_fwd_big: while warp 0 runs the
// substitution of block kb, every warp already holds the requests for block kb - 1 -- its diagonal block (two registers per
// thread) and the first 256 rows of its two columns' dots (sixteen registers per lane).  512 threads = 16 warps, two columns
// per warp.  Per column the accumulation order is that of k_mf_bwd (four accumulators over 128-row strides, tail on the
// first): identical bits.
constexpr int kBwdBigThreads = 512;
__global__ void __launch_bounds__(kBwdBigThreads) k_mf_bwd_big(SpDev d, const double *__restrict__ panels, double *v, int lvl_begin) {
  extern __shared__ double fsh[];
  __shared__ double blk[32][33];
  __shared__ double dgl[32];
  constexpr int NTB = kBwdBigThreads, NWB = NTB / 32;
  const int s = d.lvl_sn[lvl_begin + blockIdx.x];
  const int f = d.first[s], ns = d.first[s + 1] - f, ro = d.rows_off[s], nr = d.rows_off[s + 1] - ro, nf = ns + nr;
  const double *P = panels + d.panel_off[s];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double nb0, nb1, pre[2][8];
  auto prefetch = [&](int kb) {
    const int k0 = kb * 32, w = min(32, ns - k0), i0 = k0 + w + lane;
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int t = tid + e * NTB, r = t & 31, c = t >> 5;
      const double x = (r < w && c <= r) ? P[(size_t)(k0 + r) + (size_t)(k0 + c) * nf] : ((r == c) ? 1.0 : 0.0);
      if (e == 0) nb0 = x; else nb1 = x;
    }
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const double *col = P + (size_t)(k0 + min(warp + c * NWB, w - 1)) * nf;
#pragma unroll
      for (int m = 0; m < 8; m++) pre[c][m] = (i0 + 32 * m < nf) ? col[i0 + 32 * m] : 0.0;
    }
  };
  const int nb = (ns + 31) / 32;
  prefetch(nb - 1);
  for (int i = tid; i < nf; i += NTB) fsh[i] = (i < ns) ? v[f + i] : v[d.rowidx[ro + i - ns]];
  __syncthreads();
  for (int kb = nb - 1; kb >= 0; kb--) {
    const int k0 = kb * 32, w = min(32, ns - k0);
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int t = tid + e * NTB, r = t & 31, c = t >> 5;
      double x = e == 0 ? nb0 : nb1;
      if (r == c) { dgl[r] = x; x = 1.0 / x; }
      blk[r][c] = x;
    }
    {
      const int i0 = k0 + w + lane;
      const int nfull = (nf - i0 - 96 > 0) ? (nf - i0 - 96 + 127) / 128 : 0;
      const double *col[2];
      double a[2][4];
#pragma unroll
      for (int c = 0; c < 2; c++) {
        col[c] = P + (size_t)(k0 + min(warp + c * NWB, w - 1)) * nf;
#pragma unroll
        for (int q = 0; q < 4; q++) a[c][q] = 0.0;
      }
#pragma unroll
      for (int it = 0; it < 2; it++) {
        if (it < nfull) {
          const int i = i0 + 128 * it;
          const double f0 = fsh[i], f1 = fsh[i + 32], f2 = fsh[i + 64], f3 = fsh[i + 96];
#pragma unroll
          for (int c = 0; c < 2; c++) {
            a[c][0] = fma(pre[c][4 * it], f0, a[c][0]); a[c][1] = fma(pre[c][4 * it + 1], f1, a[c][1]);
            a[c][2] = fma(pre[c][4 * it + 2], f2, a[c][2]); a[c][3] = fma(pre[c][4 * it + 3], f3, a[c][3]);
          }
        }
      }
      for (int it = 2; it < nfull; it++) {
        const int i = i0 + 128 * it;
        double vv[2][4];
#pragma unroll
        for (int c = 0; c < 2; c++)
#pragma unroll
          for (int q = 0; q < 4; q++) vv[c][q] = col[c][i + 32 * q];
        const double f0 = fsh[i], f1 = fsh[i + 32], f2 = fsh[i + 64], f3 = fsh[i + 96];
#pragma unroll
        for (int c = 0; c < 2; c++) {
          a[c][0] = fma(vv[c][0], f0, a[c][0]); a[c][1] = fma(vv[c][1], f1, a[c][1]);
          a[c][2] = fma(vv[c][2], f2, a[c][2]); a[c][3] = fma(vv[c][3], f3, a[c][3]);
        }
      }
      // tail (at most three rows per lane), on the first accumulator
      if (nfull == 0) {
#pragma unroll
        for (int m = 0; m < 3; m++) {
          const int i = i0 + 32 * m;
          if (i < nf) { const double fv = fsh[i]; a[0][0] = fma(pre[0][m], fv, a[0][0]); a[1][0] = fma(pre[1][m], fv, a[1][0]); }
        }
      } else if (nfull == 1) {
#pragma unroll
        for (int m = 0; m < 3; m++) {
          const int i = i0 + 128 + 32 * m;
          if (i < nf) { const double fv = fsh[i]; a[0][0] = fma(pre[0][4 + m], fv, a[0][0]); a[1][0] = fma(pre[1][4 + m], fv, a[1][0]); }
        }
      } else {
        for (int i = i0 + 128 * nfull; i < nf; i += 32) {
          const double fv = fsh[i];
          a[0][0] = fma(col[0][i], fv, a[0][0]); a[1][0] = fma(col[1][i], fv, a[1][0]);
        }
      }
#pragma unroll
      for (int c = 0; c < 2; c++) {
        const double sacc = warp_sum((a[c][0] + a[c][1]) + (a[c][2] + a[c][3]));
        if (lane == 0 && warp + c * NWB < w) fsh[k0 + warp + c * NWB] -= sacc;
      }
    }
    __syncthreads();
    if (kb > 0) prefetch(kb - 1);      // in flight during the substitution below
    if (tid < 32) {
      double x = (tid < w) ? fsh[k0 + tid] : 0.0;
#pragma unroll
      for (int j = 31; j >= 0; j--) {
        const double xj = div_by_rcp(__shfl_sync(0xffffffffu, x, j), dgl[j], blk[j][j]);
        if (tid == j) x = xj; else if (tid < j) x = fma(-blk[j][tid], xj, x);
      }
      if (tid < w) fsh[k0 + tid] = x;
    }
    __syncthreads();
  }
  for (int i = tid; i < ns; i += NTB) v[f + i] = fsh[i];
}

// QPALM_B200_MF_LEVEL_NAMES=1: the profiler (prof.cu) sees the solve launches under per-level names ("k_mf_fwd.L17")
static bool level_names() { static const bool on = getenv("QPALM_B200_MF_LEVEL_NAMES") != nullptr; return on; }
static const char *level_name(const char *base, int l) {
  static char names[2][64][24];
  static bool init = false;
  if (!init) {
    for (int k = 0; k < 2; k++) for (int i = 0; i < 64; i++) snprintf(names[k][i], sizeof(names[k][i]), "%s.L%02d", k == 0 ? "k_mf_fwd" : "k_mf_bwd", i);
    init = true;
  }
  return names[base[5] == 'f' ? 0 : 1][l < 63 ? l : 63];
}

int sparse_chol_solve(SparseChol *sc, cudaStream_t st, const double *panels, const double *rhs, double *out, bool negate) {
  const SymHost &h = sc->h;
  const int n = sc->d.n;
  static bool attr_set = false;
  if (!attr_set) {
    QB_CUDA_TRY(cudaFuncSetAttribute(k_mf_fwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    QB_CUDA_TRY(cudaFuncSetAttribute(k_mf_fwd_big, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    QB_CUDA_TRY(cudaFuncSetAttribute(k_mf_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    QB_CUDA_TRY(cudaFuncSetAttribute(k_mf_bwd_big, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  if (sizeof(double) * (size_t)h.max_nf > 200 * 1024) { fprintf(stderr, "[qpalm_b200] sparse solve: front of %d rows exceeds the shared-memory vector\n", h.max_nf); return 4; }
  const char *fg = getenv("QPALM_B200_MF_FWD_GENERIC");     // tests: the generic forward / backward kernels on every level (bit-identity check)
  const bool fwd_generic = fg && atoi(fg) != 0;
  QB_LAUNCH(k_sp_permute_in, cdiv(n, 256), 256, 0, st, sc->d, rhs, sc->v, negate ? -1.0 : 1.0);
  for (int l = 0; l < h.nlevels; l++) {
    const int b = h.lvl_ptr[l], cnt = h.lvl_ptr[l + 1] - b;
    if (cnt <= 0) continue;
    const int mnf = h.lvl_max_nf[l];
    const bool lp = g_prof_on && level_names() && prof_begin(level_name("k_mf_fwd", l), st);   // per-level timing (diagnostics)
    const int po = g_prof_on; if (lp) g_prof_on = 0;
    if (mnf > kSmallMaxNf && !fwd_generic) QB_LAUNCH(k_mf_fwd_big, cnt, kFwdBigThreads, sizeof(double) * (size_t)mnf, st, sc->d, panels, sc->v, sc->uvec, b);
    else QB_LAUNCH(k_mf_fwd, cnt, mnf <= 64 ? 64 : 256, sizeof(double) * (size_t)mnf, st, sc->d, panels, sc->v, sc->uvec, b);
    if (lp) { g_prof_on = po; prof_end(st); }
  }
  if (sc->d.sgn) QB_LAUNCH(k_sp_apply_sign, cdiv(n, 256), 256, 0, st, sc->d, sc->v);   // L S L' x = b: y <- inv(S) y = S y
  for (int l = h.nlevels - 1; l >= 0; l--) {
    const int b = h.lvl_ptr[l], cnt = h.lvl_ptr[l + 1] - b;
    if (cnt <= 0) continue;
    const int mnf = h.lvl_max_nf[l];
    const bool lp = g_prof_on && level_names() && prof_begin(level_name("k_mf_bwd", l), st);
    const int po = g_prof_on; if (lp) g_prof_on = 0;
    if (mnf > kSmallMaxNf && !fwd_generic) QB_LAUNCH(k_mf_bwd_big, cnt, kBwdBigThreads, sizeof(double) * (size_t)mnf, st, sc->d, panels, sc->v, b);
    else QB_LAUNCH(k_mf_bwd, cnt, mnf <= 64 ? 64 : 256, sizeof(double) * (size_t)mnf, st, sc->d, panels, sc->v, b);
    if (lp) { g_prof_on = po; prof_end(st); }
  }
  QB_LAUNCH(k_sp_permute_out, cdiv(n, 256), 256, 0, st, sc->d, sc->v, out);
  QB_CUDA_TRY(cudaGetLastError());
  return 0;
}

// ================================================================================================
// rank-k update / downdate (cholmod_updown replacement; recurrence of Modify/t_cholmod_updown_numkr.c:289-318 for LL')
// ================================================================================================
__global__ void k_ud_gather(SpDev d, const int *__restrict__ Rp, const int *__restrict__ Rj, const double *__restrict__ Rx,
                            const int *__restrict__ list, const double *__restrict__ scale, int scale_by_row, int cnt,
                            double *W, int *mark) {
  const int c = blockIdx.x;
  if (c >= cnt) return;
  const int row = list[c];
  const double sc = scale_by_row ? scale[row] : scale[c];
  for (int t = Rp[row] + threadIdx.x; t < Rp[row + 1]; t += blockDim.x) {
    const int pc = d.iperm[Rj[t]];
    W[(size_t)pc + (size_t)c * d.n] = Rx[t] * sc;
    int s = d.sn_of_col[pc];
    while (s >= 0 && atomicExch(&mark[s], 1) == 0) s = d.sn_parent[s];
  }
}

constexpr int kUdThreads = 512;
__global__ void __launch_bounds__(kUdThreads) k_ud_sweep(SpDev d, double *panels, double *W, int *mark, int k, int sign, int *info) {
  __shared__ double alpha[8], wj[8], gam[8];
  __shared__ double winv_s, lnew_s;
  __shared__ int skip_s, list[kUdThreads], nlist, wcount[kUdThreads / 32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = d.n;
  if (tid < 8) alpha[tid] = 1.0;
  const double sg = (double)sign;
  for (int base = 0; base < d.nsuper; base += kUdThreads) {
    // ordered compaction of the marked supernodes of this chunk
    const int sidx = base + tid;
    const int flag = (sidx < d.nsuper) ? mark[sidx] : 0;
    if (flag) mark[sidx] = 0;
    const unsigned bal = __ballot_sync(0xffffffffu, flag != 0);
    if (lane == 0) wcount[warp] = __popc(bal);
    __syncthreads();
    int off = 0, tot = 0;
    for (int wv = 0; wv < kUdThreads / 32; wv++) { if (wv < warp) off += wcount[wv]; tot += wcount[wv]; }
    if (flag) list[off + __popc(bal & ((1u << lane) - 1u))] = sidx;
    if (tid == 0) nlist = tot;
    __syncthreads();
    const int nl = nlist;
    for (int q = 0; q < nl; q++) {
      const int s = list[q];
      const int f = d.first[s], ns = d.first[s + 1] - f, ro = d.rows_off[s], nr = d.rows_off[s + 1] - ro, nf = ns + nr;
      double *P = panels + d.panel_off[s];
      for (int lc = 0; lc < ns; lc++) {
        const int j = f + lc;
        if (tid == 0) {
          bool any = false;
          for (int r = 0; r < k; r++) { wj[r] = W[(size_t)j + (size_t)r * n]; any |= (wj[r] != 0.0); }
          skip_s = any ? 0 : 1;
          if (any) {
            const double ljj = P[(size_t)lc + (size_t)lc * nf];
            const double winv = 1.0 / ljj;
            double dj = ljj * ljj;
            for (int r = 0; r < k; r++) {
              const double a = alpha[r] + sg * (wj[r] * wj[r]) / dj;
              dj *= a;
              gam[r] = -sg * wj[r] / dj;
              dj /= alpha[r];
              alpha[r] = a;
              W[(size_t)j + (size_t)r * n] = 0.0;
            }
            if (!(dj > 0.0)) atomicExch(info, 1 + j);
            const double lnew = sqrt(dj);
            winv_s = winv; lnew_s = lnew;
            P[(size_t)lc + (size_t)lc * nf] = lnew;
          }
        }
        __syncthreads();
        if (!skip_s) {
          const double winv = winv_s, lnew = lnew_s;
          for (int i = lc + 1 + tid; i < nf; i += kUdThreads) {
            const int gi = (i < ns) ? f + i : d.rowidx[ro + i - ns];
            double t = P[(size_t)i + (size_t)lc * nf] * winv;
            for (int r = 0; r < k; r++) {
              double wv = W[(size_t)gi + (size_t)r * n];
              wv -= wj[r] * t;
              t -= gam[r] * wv;
              W[(size_t)gi + (size_t)r * n] = wv;
            }
            P[(size_t)i + (size_t)lc * nf] = t * lnew;
          }
        }
        __syncthreads();
      }
    }
    __syncthreads();
  }
}

int sparse_chol_updown(SparseChol *sc, cudaStream_t st, double *panels, const int *Acsr_p, const int *Acsr_j,
                       const double *Acsr_x, const int *list, const double *scale, bool scale_by_row, int cnt, int sign,
                       int *info_dev) {
  if (cnt <= 0) return 0;
  if (cnt > 8) return 1;
  QB_CUDA_TRY(cudaMemsetAsync(sc->W, 0, sizeof(double) * (size_t)sc->d.n * cnt, st));
  QB_LAUNCH(k_ud_gather, cnt, 128, 0, st, sc->d, Acsr_p, Acsr_j, Acsr_x, list, scale, scale_by_row ? 1 : 0, cnt, sc->W, sc->mark);
  QB_LAUNCH(k_ud_sweep, 1, kUdThreads, 0, st, sc->d, panels, sc->W, sc->mark, cnt, sign, info_dev);
  QB_CUDA_TRY(cudaGetLastError());
  return 0;
}

// ================================================================================================
// Gershgorin row sums of the (unfactored) symmetric matrix held in the panels; helpers for the tests
// ================================================================================================
__global__ void __launch_bounds__(256) k_sp_abs_rowsums(SpDev d, const double *__restrict__ panels, double *out) {
  const int lane = threadIdx.x & 31;
  const int pj = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (pj >= d.n) return;
  const int s = d.sn_of_col[pj], f = d.first[s], ns = d.first[s + 1] - f, ro = d.rows_off[s], nf = ns + d.rows_off[s + 1] - ro;
  const int lc = pj - f;
  const double *col = panels + d.panel_off[s] + (size_t)lc * nf;
  double acc = 0.0;
  for (int i = lc + lane; i < nf; i += 32) {
    const double a = fabs(col[i]);
    acc += a;
    if (i > lc && a != 0.0) atomicAdd(&out[(i < ns) ? f + i : d.rowidx[ro + i - ns]], a);
  }
  acc = warp_sum(acc);
  if (lane == 0) atomicAdd(&out[pj], acc);
}
int sparse_chol_abs_rowsums(SparseChol *sc, cudaStream_t st, const double *panels, double *out_n) {
  QB_CUDA_TRY(cudaMemsetAsync(out_n, 0, sizeof(double) * (size_t)sc->d.n, st));
  QB_LAUNCH(k_sp_abs_rowsums, cdiv(sc->d.n, 8), 256, 0, st, sc->d, panels, out_n);
  QB_CUDA_TRY(cudaGetLastError());
  return 0;
}

int sparse_chol_download(SparseChol *sc, cudaStream_t st, const double *panels, double *L_host, long long *perm_host) {
  const SymHost &h = sc->h;
  const int n = h.n;
  std::vector<double> P((size_t)h.nnzL);
  QB_CUDA_TRY(cudaStreamSynchronize(st));
  QB_CUDA_TRY(cudaMemcpy(P.data(), panels, sizeof(double) * P.size(), cudaMemcpyDeviceToHost));
  memset(L_host, 0, sizeof(double) * (size_t)n * n);
  for (int s = 0; s < h.nsuper; s++) {
    const int f = h.sn_first[s], ns = h.sn_first[s + 1] - f, ro = h.rows_off[s], nr = h.rows_off[s + 1] - ro, nf = ns + nr;
    const double *p = P.data() + h.panel_off[s];
    for (int lc = 0; lc < ns; lc++)
      for (int i = lc; i < nf; i++) {
        const int gi = i < ns ? f + i : h.rowidx[ro + i - ns];
        L_host[(size_t)gi + (size_t)(f + lc) * n] = p[(size_t)i + (size_t)lc * nf];
      }
  }
  for (int i = 0; i < n; i++) perm_host[i] = h.perm[i];
  return 0;
}

}  // namespace qb
