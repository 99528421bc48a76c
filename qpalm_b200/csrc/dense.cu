// dense.cu -- FP64 dense kernels for the Newton system of QPALM (sm_100a).
//
// Replaces, for the Schur-complement path of the reference (src/solver_interface.c:319-519):
//   cholmod_aat + cholmod_add      -> dgemm_nt(lower_only) on gathered, sqrt(sigma)-scaled rows of A
//   cholmod_analyze/factorize_p    -> potrf_lower  (blocked right-looking, DMMA trailing updates)
//   cholmod_solve(CHOLMOD_LDLt)    -> chol_solve   (blocked forward/backward substitution)
//   cholmod_updown                 -> chol_updown  (rank-k recurrence of t_cholmod_updown_numkr.c:289-376
//                                                   restated for L L' and panelised for the GPU)
#include "dense.cuh"

namespace qb {

long long g_kernel_launches = 0;

// ================================================================================================
// DMMA NT GEMM:  C = beta*C + alpha * A * B'
// CTA tile 128 x 128 x 16, 8 warps (2 x 4), warp tile 64 x 32 = 8 x 4 m8n8k4 fragments,
// 3-stage cp.async pipeline, k-major shared tiles padded to 132 doubles (conflict-free 64-bit
// fragment loads: lane -> (k = lane & 3, row = lane >> 2) hits 16 distinct 8-byte banks per half warp).
// ================================================================================================
namespace gemm {
constexpr int BM = 128, BN = 128, BK = 16, STAGES = 3, SROW = BM + 4;
constexpr size_t kSmemBytes = sizeof(double) * 2 * STAGES * BK * SROW;

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gmem_src) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256, 1)
k_dgemm_nt(int K, const double *A, int lda, const double *B, int ldb,
           double *C, int ldc, double alpha, double beta, int lower_only,
           long long sA, long long sB, long long sC, const int *Kz, const int *maskz) {
  // blockIdx.z: batch instance (strides sA/sB/sC, optional per-instance K and activity mask)
  const int bm = blockIdx.x, bn = blockIdx.y;
  if (lower_only && bn > bm) return;
  if (maskz && !maskz[blockIdx.z]) return;
  if (Kz) K = Kz[blockIdx.z];
  if (K <= 0 && beta == 1.0) return;
  A += (size_t)blockIdx.z * sA; B += (size_t)blockIdx.z * sB; C += (size_t)blockIdx.z * sC;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  typedef double Tile[BK][SROW];
  Tile *As = reinterpret_cast<Tile *>(smem_raw);
  Tile *Bs = reinterpret_cast<Tile *>(smem_raw + sizeof(double) * STAGES * BK * SROW);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, wm = warp & 1, wn = warp >> 1;
  const double *Ag = A + (size_t)bm * BM;
  const double *Bg = B + (size_t)bn * BN;
  const int KT = K / BK;

  auto load_stage = [&](int stage, int kt) {
    const int k0 = kt * BK;
    const int mc = (tid & 63) * 2;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int kk = (tid >> 6) + 4 * i;
      cp_async16(&As[stage][kk][mc], Ag + mc + (size_t)(k0 + kk) * lda);
      cp_async16(&Bs[stage][kk][mc], Bg + mc + (size_t)(k0 + kk) * ldb);
    }
  };

  double acc[8][4][2];
#pragma unroll
  for (int i = 0; i < 8; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

#pragma unroll
  for (int s = 0; s < STAGES - 1; s++) {
    if (s < KT) load_stage(s, s);
    cp_async_commit();
  }
  const int fr = lane >> 2, fk = lane & 3;
  for (int kt = 0; kt < KT; kt++) {
    cp_async_wait<STAGES - 2>();
    __syncthreads();
    {
      const int nk = kt + STAGES - 1;
      if (nk < KT) load_stage(nk % STAGES, nk);
      cp_async_commit();
    }
    const int st = kt % STAGES;
    double a[2][8], b[2][4];
#pragma unroll
    for (int mb = 0; mb < 8; mb++) a[0][mb] = As[st][fk][wm * 64 + mb * 8 + fr];
#pragma unroll
    for (int nb = 0; nb < 4; nb++) b[0][nb] = Bs[st][fk][wn * 32 + nb * 8 + fr];
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
      const int cur = ks & 1, nxt = cur ^ 1;
      if (ks < 3) {
#pragma unroll
        for (int mb = 0; mb < 8; mb++) a[nxt][mb] = As[st][(ks + 1) * 4 + fk][wm * 64 + mb * 8 + fr];
#pragma unroll
        for (int nb = 0; nb < 4; nb++) b[nxt][nb] = Bs[st][(ks + 1) * 4 + fk][wn * 32 + nb * 8 + fr];
      }
#pragma unroll
      for (int mb = 0; mb < 8; mb++)
#pragma unroll
        for (int nb = 0; nb < 4; nb++) dmma884(acc[mb][nb][0], acc[mb][nb][1], a[cur][mb], b[cur][nb]);
    }
  }
  cp_async_wait<0>();
  __syncthreads();  // all global reads of this CTA have landed: C may alias A from here on

  const bool use_beta = (beta != 0.0);
#pragma unroll
  for (int mb = 0; mb < 8; mb++) {
    const int row = bm * BM + wm * 64 + mb * 8 + fr;
#pragma unroll
    for (int nb = 0; nb < 4; nb++) {
      const int col = bn * BN + wn * 32 + nb * 8 + 2 * fk;
      double *c0 = C + row + (size_t)col * ldc;
      double *c1 = c0 + ldc;
      double v0 = alpha * acc[mb][nb][0], v1 = alpha * acc[mb][nb][1];
      if (use_beta) { v0 += beta * (*c0); v1 += beta * (*c1); }
      *c0 = v0; *c1 = v1;
    }
  }
}
}  // namespace gemm

int dgemm_nt(cudaStream_t s, int M, int N, int K, const double *A, int lda, const double *B, int ldb,
             double *C, int ldc, double alpha, double beta, bool lower_only) {
  if (M <= 0 || N <= 0) return 0;
  if ((M % gemm::BM) || (N % gemm::BN) || (K % gemm::BK) || K <= 0) {
    fprintf(stderr, "[qpalm_b200] dgemm_nt: bad shape M=%d N=%d K=%d\n", M, N, K);
    return 1;
  }
  static bool attr_set = false;
  if (!attr_set) {
    QB_CUDA_TRY(cudaFuncSetAttribute(gemm::k_dgemm_nt, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)gemm::kSmemBytes));
    attr_set = true;
  }
  dim3 grid(M / gemm::BM, N / gemm::BN);
  QB_LAUNCH(gemm::k_dgemm_nt, grid, 256, gemm::kSmemBytes, s, K, A, lda, B, ldb, C, ldc, alpha, beta,
            lower_only ? 1 : 0, 0LL, 0LL, 0LL, (const int *)nullptr, (const int *)nullptr);
  QB_CUDA_TRY(cudaGetLastError());
  return 0;
}

int dgemm_nt_batched(cudaStream_t s, int nb, int M, int N, int K, const int *Kz, const double *A, int lda, long long sA,
                     const double *B, int ldb, long long sB, double *C, int ldc, long long sC, double alpha, double beta,
                     bool lower_only, const int *maskz) {
  if (M <= 0 || N <= 0 || nb <= 0) return 0;
  if ((M % gemm::BM) || (N % gemm::BN) || (K % gemm::BK)) return 1;
  static bool attr_set = false;
  if (!attr_set) {
    QB_CUDA_TRY(cudaFuncSetAttribute(gemm::k_dgemm_nt, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)gemm::kSmemBytes));
    attr_set = true;
  }
  dim3 grid(M / gemm::BM, N / gemm::BN, nb);
  QB_LAUNCH(gemm::k_dgemm_nt, grid, 256, gemm::kSmemBytes, s, K, A, lda, B, ldb, C, ldc, alpha, beta,
            lower_only ? 1 : 0, sA, sB, sC, Kz, maskz);
  QB_CUDA_TRY(cudaGetLastError());
  return 0;
}

// ================================================================================================
// 128 x 128 diagonal block: Cholesky factor + inverse of the factor, one CTA, all in shared memory.
// ================================================================================================
namespace diag {
// 128 x 128 diagonal block: Cholesky factor + inverse of the factor, one CTA, all in shared memory.
// Blocked in 32-column sub-panels: a 32 x 32 block is factorised by ONE WARP in registers (row per lane,
// pivots and multipliers exchanged with shuffles), the rows below are solved against it one row per
// thread (row in registers, factor entries broadcast from shared memory), then all 512 threads apply the
// rank-32 trailing update.  The inverse is formed block row by block row from the four 32 x 32 inverses.
constexpr int NB = 128, DS = 129, NT = 512, SB = 32, RSD = 97;
constexpr size_t kSmemBytes = sizeof(double) * (NB * DS + NB * (NB + 1) / 2 + SB * RSD + NB) + 16;

__device__ __forceinline__ int pidx(int i, int c) { return i * (i + 1) / 2 + c; }

__device__ void factor32_warp(double *As, double *rdiag, int base, int lane, int *info, int col0) {
  double a[SB];
#pragma unroll
  for (int c = 0; c < SB; c++) a[c] = (c <= lane) ? As[(base + lane) * DS + base + c] : 0.0;
  bool bad = false;
#pragma unroll
  for (int j = 0; j < SB; j++) {
    const double pjj = __shfl_sync(0xffffffffu, a[j], j);
    if (!(pjj > 0.0) && !bad) { bad = true; if (lane == 0 && info) atomicCAS(info, 0, col0 + base + j + 1); }
    const double ljj = sqrt(pjj), inv = 1.0 / ljj;
    if (lane == j) { a[j] = ljj; rdiag[base + j] = inv; }
    else if (lane > j) a[j] *= inv;
    // constant inner bounds + predicate (a j-dependent bound is unrolled before the outer loop and leaves a[] in
    // local memory); the dead half folds away after the outer unroll
#pragma unroll
    for (int c = 0; c < SB; c++) {
      if (c > j) {
        const double lcj = __shfl_sync(0xffffffffu, a[j], c);
        if (lane >= c) a[c] = fma(-a[j], lcj, a[c]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < SB; c++) if (c <= lane) As[(base + lane) * DS + base + c] = a[c];
}

// rows i > base+31: A(i, base..base+31) <- A(i, ..) * inv(L_bb)'  by forward substitution, one row per thread
__device__ void panel_solve(double *As, const double *rdiag, int base, int tid) {
  const int i = base + SB + tid;
  if (i >= NB) return;
  double v[SB];
#pragma unroll
  for (int c = 0; c < SB; c++) v[c] = As[i * DS + base + c];
#pragma unroll
  for (int c = 0; c < SB; c++) {
    double s = v[c];
#pragma unroll
    for (int t = 0; t < SB; t++) if (t < c) s = fma(-v[t], As[(base + c) * DS + base + t], s);
    v[c] = s * rdiag[base + c];
  }
#pragma unroll
  for (int c = 0; c < SB; c++) As[i * DS + base + c] = v[c];
}

// A(i,k) -= sum_t L(i, base+t) L(k, base+t) for base+32 <= k <= i < 128
__device__ void trailing_update(double *As, int base, int tid) {
  const int tx = tid & 31, ty = tid >> 5, start = base + SB;
  for (int i = start + ty; i < NB; i += NT / 32) {
    const double *Li = As + i * DS + base;
    for (int k = start + tx; k <= i; k += 32) {
      const double *Lk = As + k * DS + base;
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int t = 0; t < SB; t += 2) { s0 = fma(Li[t], Lk[t], s0); s1 = fma(Li[t + 1], Lk[t + 1], s1); }
      As[i * DS + k] -= (s0 + s1);
    }
  }
}

// X_bb = inv(L_bb) for the 32 x 32 diagonal block at `base`; lane = column of X
__device__ void inv32_warp(const double *As, double *Xs, const double *rdiag, int base, int lane) {
  double x[SB];
#pragma unroll
  for (int r = 0; r < SB; r++) {
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < SB; t++) if (t < r) s = fma(As[(base + r) * DS + base + t], x[t], s);   // x[t] == 0 for t < lane
    x[r] = (r < lane) ? 0.0 : ((r == lane) ? rdiag[base + r] : -s * rdiag[base + r]);
  }
#pragma unroll
  for (int r = 0; r < SB; r++) if (r >= lane) Xs[pidx(base + r, base + lane)] = x[r];
}

__device__ void inverse_from_factor(const double *As, double *Xs, double *Rs, const double *rdiag, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  if (warp < 4) inv32_warp(As, Xs, rdiag, warp * SB, lane);
  __syncthreads();
#pragma unroll
  for (int bi = 1; bi < 4; bi++) {
    const int width = SB * bi, r0 = SB * bi;
    // R = - L(block row bi, cols < r0) * X(rows < r0, cols < r0)
    for (int idx = tid; idx < SB * width; idx += NT) {
      const int r = idx / width, c = idx - r * width;
      const double *Lr = As + (r0 + r) * DS;
      double s = 0.0;
      for (int u = c; u < r0; u++) s = fma(Lr[u], Xs[pidx(u, c)], s);
      Rs[r * RSD + c] = -s;
    }
    __syncthreads();
    // X(block row bi, cols < r0) = X_bb * R
    for (int idx = tid; idx < SB * width; idx += NT) {
      const int r = idx / width, c = idx - r * width;
      double s = 0.0;
      for (int u = 0; u <= r; u++) s = fma(Xs[pidx(r0 + r, r0 + u)], Rs[u * RSD + c], s);
      Xs[pidx(r0 + r, c)] = s;
    }
    __syncthreads();
  }
}

// factor == 1: the block (lower) is factorised in place, then inverted.  factor == 0: the block already
// holds a Cholesky factor (after an update/downdate sweep) and only the inverse is formed.
// grid.x: diagonal block index (block_stride columns apart), grid.y: batch instance.
__global__ void __launch_bounds__(NT, 1)
k_diag_block(double *Lg, int ld, double *Xg, int *info, int col0, int factor, int block_stride,
             long long strideL, long long strideX, const int *mask) {
  if (mask && !mask[blockIdx.y]) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *As = reinterpret_cast<double *>(smem_raw);
  double *Xs = As + NB * DS;
  double *Rs = Xs + NB * (NB + 1) / 2;
  double *rdiag = Rs + SB * RSD;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  Lg += (size_t)blockIdx.y * strideL + (size_t)blockIdx.x * block_stride * (size_t)(ld + 1);
  Xg += (size_t)blockIdx.y * strideX + (size_t)blockIdx.x * NB * NB;
  if (info) info += blockIdx.y;
  col0 += blockIdx.x * block_stride;
  for (int idx = tid; idx < NB * NB; idx += NT) {
    const int i = idx & (NB - 1), k = idx >> 7;
    if (k <= i) As[i * DS + k] = Lg[i + (size_t)ld * k];
  }
  __syncthreads();
  if (factor) {
#pragma unroll 1
    for (int kb = 0; kb < 4; kb++) {
      const int base = kb * SB;
      if (warp == 0) factor32_warp(As, rdiag, base, lane, info, col0);
      __syncthreads();
      if (kb < 3) {
        panel_solve(As, rdiag, base, tid);
        __syncthreads();
        trailing_update(As, base, tid);
        __syncthreads();
      }
    }
    for (int idx = tid; idx < NB * NB; idx += NT) {
      const int i = idx & (NB - 1), k = idx >> 7;
      if (k <= i) Lg[i + (size_t)ld * k] = As[i * DS + k];
    }
  } else {
    if (tid < NB) rdiag[tid] = 1.0 / As[tid * DS + tid];
    __syncthreads();
  }
  inverse_from_factor(As, Xs, Rs, rdiag, tid);
  for (int idx = tid; idx < NB * NB; idx += NT) {
    const int i = idx & (NB - 1), c = idx >> 7;
    Xg[i + (size_t)NB * c] = (c <= i) ? Xs[pidx(i, c)] : 0.0;
  }
}
}  // namespace diag

static int diag_attr() {
  static bool set = false;
  if (!set) {
    QB_CUDA_TRY(cudaFuncSetAttribute(diag::k_diag_block, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)diag::kSmemBytes));
    set = true;
  }
  return 0;
}

// Two-level right-looking blocked Cholesky.  Inner panels are 128 wide (diagonal block factor+inverse in one
// CTA, panel solve as an in-place DMMA GEMM against the inverted block); their rank-128 updates are applied
// only inside the current 512-wide outer block.  Everything to the right of the outer block receives ONE
// rank-512 DMMA update per outer block, which amortises the C-tile read-modify-write of the trailing matrix
// four times better than rank-128 updates.
constexpr int kOuter = 512;

int potrf_lower_batched(cudaStream_t s, int nb, int npad, double *L, int ld, long long sL, double *invdiag, long long sX,
                        int *info_dev, const int *mask) {
  if (int e = diag_attr()) return e;
  for (int J0 = 0; J0 < npad; J0 += kOuter) {
    const int Jend = (J0 + kOuter < npad) ? J0 + kOuter : npad;
    for (int j0 = J0; j0 < Jend; j0 += kPanel) {
      double *Ljj = L + j0 + (size_t)j0 * ld;
      double *Xp = invdiag + (size_t)(j0 / kPanel) * kPanel * kPanel;
      QB_LAUNCH(diag::k_diag_block, dim3(1, nb), diag::NT, diag::kSmemBytes, s, Ljj, ld, Xp, info_dev, j0, 1, 0, sL, sX, mask);
      const int rem = npad - j0 - kPanel;
      if (rem <= 0) continue;
      double *L21 = Ljj + kPanel;
      // L21 <- A21 * inv(L11)'   (in place, one 128-wide tile column)
      if (int e = dgemm_nt_batched(s, nb, rem, kPanel, kPanel, nullptr, L21, ld, sL, Xp, kPanel, sX, L21, ld, sL, 1.0, 0.0, false, mask)) return e;
      // rank-128 update of the remaining columns of THIS outer block (all rows below)
      const int wcols = Jend - (j0 + kPanel);
      if (wcols > 0) {
        double *Cin = L + (j0 + kPanel) + (size_t)(j0 + kPanel) * ld;
        if (int e = dgemm_nt_batched(s, nb, rem, wcols, kPanel, nullptr, L21, ld, sL, L21, ld, sL, Cin, ld, sL, -1.0, 1.0, true, mask)) return e;
      }
    }
    // rank-(Jend-J0) update of everything to the right of the outer block
    const int rem = npad - Jend;
    if (rem > 0) {
      const double *P = L + Jend + (size_t)J0 * ld;
      double *A22 = L + Jend + (size_t)Jend * ld;
      if (int e = dgemm_nt_batched(s, nb, rem, rem, Jend - J0, nullptr, P, ld, sL, P, ld, sL, A22, ld, sL, -1.0, 1.0, true, mask)) return e;
    }
  }
  QB_CUDA_TRY(cudaGetLastError());
  return 0;
}

int potrf_lower(cudaStream_t s, int npad, double *L, int ld, double *invdiag, int *info_dev) {
  return potrf_lower_batched(s, 1, npad, L, ld, 0, invdiag, 0, info_dev, nullptr);
}

int trtri_diag_blocks(cudaStream_t s, int npad, const double *L, int ld, double *invdiag) {
  if (int e = diag_attr()) return e;
  const int nblk = npad / kPanel;
  QB_LAUNCH(diag::k_diag_block, nblk, diag::NT, diag::kSmemBytes, s, const_cast<double *>(L), ld, invdiag,
            (int *)nullptr, 0, 0, kPanel, 0LL, 0LL, (const int *)nullptr);
  QB_CUDA_TRY(cudaGetLastError());
  return 0;
}

// ================================================================================================
// blocked triangular solves.  One launch per 128-column block; the CTA that owns the next diagonal
// block also applies that block's inverse, so the following launch finds its x-block ready.
// ================================================================================================
namespace trsv {
constexpr int NB = 128, NT = 256;

// y(0..127) = X * v  (X lower, col-major 128x128) ; result written to out[], all NT threads call.
__device__ void apply_inv_lower(const double *__restrict__ X, const double *v_s, double *out, double *scratch) {
  const int r = threadIdx.x & (NB - 1), h = threadIdx.x >> 7;  // h in {0,1}: column halves
  double acc = 0.0;
  const int c0 = h * 64, c1 = c0 + 64;
#pragma unroll 8
  for (int c = c0; c < c1; c++) acc = fma(X[r + NB * c], v_s[c], acc);
  scratch[h * NB + r] = acc;
  __syncthreads();
  if (h == 0) out[r] = scratch[r] + scratch[NB + r];
}
// y(c) = sum_i X(i,c) v(i)  (X' * v)
__device__ void apply_inv_lower_t(const double *__restrict__ X, const double *v_s, double *out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int c = warp; c < NB; c += NT / 32) {
    double acc = 0.0;
#pragma unroll
    for (int q = 0; q < 4; q++) acc = fma(X[lane + 32 * q + NB * c], v_s[lane + 32 * q], acc);
    acc = warp_sum(acc);
    if (lane == 0) out[c] = acc;
  }
}

__global__ void __launch_bounds__(NT) k_first_fwd(const double *invdiag0, double *v, long long sX, long long sV, const int *mask) {
  if (mask && !mask[blockIdx.y]) return;
  invdiag0 += (size_t)blockIdx.y * sX; v += (size_t)blockIdx.y * sV;
  __shared__ double vs[NB], scratch[2 * NB];
  if (threadIdx.x < NB) vs[threadIdx.x] = v[threadIdx.x];
  __syncthreads();
  apply_inv_lower(invdiag0, vs, v, scratch);
}

// step b of the forward solve: rows of blocks > b get rhs -= L(rows, block b) * x_b; CTA 0 (block b+1)
// then forms x_{b+1} = inv(L_{b+1,b+1}) rhs_{b+1}.
__global__ void __launch_bounds__(NT) k_fwd_step(const double *__restrict__ L, int ld, const double *invdiag,
                                                 double *v, int b, long long sL, long long sX, long long sV, const int *mask) {
  if (mask && !mask[blockIdx.y]) return;
  L += (size_t)blockIdx.y * sL; invdiag += (size_t)blockIdx.y * sX; v += (size_t)blockIdx.y * sV;
  __shared__ double xs[NB], scratch[2 * NB], rs[NB];
  const int j0 = b * NB, r0 = (b + 1 + blockIdx.x) * NB;
  const int tid = threadIdx.x, r = tid & (NB - 1), h = tid >> 7;
  if (tid < NB) xs[tid] = v[j0 + tid];
  __syncthreads();
  const double *Lp = L + (size_t)(r0 + r) + (size_t)(j0 + h * 64) * ld;
  double acc = 0.0;
#pragma unroll 8
  for (int c = 0; c < 64; c++) acc = fma(Lp[(size_t)c * ld], xs[h * 64 + c], acc);
  scratch[h * NB + r] = acc;
  __syncthreads();
  if (h == 0) {
    const double nv = v[r0 + r] - (scratch[r] + scratch[NB + r]);
    if (blockIdx.x == 0) rs[r] = nv; else v[r0 + r] = nv;
  }
  if (blockIdx.x == 0) {
    __syncthreads();
    apply_inv_lower(invdiag + (size_t)(b + 1) * NB * NB, rs, v + r0, scratch);
  }
}

__global__ void __launch_bounds__(NT) k_first_bwd(const double *invdiag_last, double *v_last, long long sX, long long sV, const int *mask) {
  if (mask && !mask[blockIdx.y]) return;
  invdiag_last += (size_t)blockIdx.y * sX; v_last += (size_t)blockIdx.y * sV;
  __shared__ double vs[NB];
  if (threadIdx.x < NB) vs[threadIdx.x] = v_last[threadIdx.x];
  __syncthreads();
  apply_inv_lower_t(invdiag_last, vs, v_last);
}

// step b of the backward solve (b descending): columns of blocks < b get z -= L(block b rows, cols)' d_b;
// the CTA that owns block b-1 then forms d_{b-1} = inv(L_{b-1,b-1})' z_{b-1}.
__global__ void __launch_bounds__(NT) k_bwd_step(const double *__restrict__ L, int ld, const double *invdiag,
                                                 double *v, int b, long long sL, long long sX, long long sV, const int *mask) {
  if (mask && !mask[blockIdx.y]) return;
  L += (size_t)blockIdx.y * sL; invdiag += (size_t)blockIdx.y * sX; v += (size_t)blockIdx.y * sV;
  __shared__ double ds[NB], zs[NB];
  const int i0 = b * NB, c0 = blockIdx.x * NB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid < NB) ds[tid] = v[i0 + tid];
  __syncthreads();
  const bool next_diag = ((int)blockIdx.x == b - 1);
  for (int cc = warp * 16; cc < warp * 16 + 16; cc += 4) {
    double acc[4] = {0, 0, 0, 0};
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const double dv = ds[lane + 32 * q];
#pragma unroll
      for (int u = 0; u < 4; u++)
        acc[u] = fma(L[(size_t)(i0 + lane + 32 * q) + (size_t)(c0 + cc + u) * ld], dv, acc[u]);
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const double sacc = warp_sum(acc[u]);
      if (lane == 0) {
        const double nv = v[c0 + cc + u] - sacc;
        if (next_diag) zs[cc + u] = nv; else v[c0 + cc + u] = nv;
      }
    }
  }
  if (next_diag) {
    __syncthreads();
    apply_inv_lower_t(invdiag + (size_t)(b - 1) * NB * NB, zs, v + c0);
  }
}
}  // namespace trsv

int chol_solve_batched(cudaStream_t s, int nb, int npad, const double *L, int ld, long long sL, const double *invdiag,
                       long long sX, double *v, long long sV, const int *mask) {
  const int nblk = npad / kPanel;
  QB_LAUNCH(trsv::k_first_fwd, dim3(1, nb), trsv::NT, 0, s, invdiag, v, sX, sV, mask);
  for (int b = 0; b + 1 < nblk; b++)
    QB_LAUNCH(trsv::k_fwd_step, dim3(nblk - 1 - b, nb), trsv::NT, 0, s, L, ld, invdiag, v, b, sL, sX, sV, mask);
  QB_LAUNCH(trsv::k_first_bwd, dim3(1, nb), trsv::NT, 0, s, invdiag + (size_t)(nblk - 1) * kPanel * kPanel,
            v + (size_t)(nblk - 1) * kPanel, sX, sV, mask);
  for (int b = nblk - 1; b >= 1; b--)
    QB_LAUNCH(trsv::k_bwd_step, dim3(b, nb), trsv::NT, 0, s, L, ld, invdiag, v, b, sL, sX, sV, mask);
  QB_CUDA_TRY(cudaGetLastError());
  return 0;
}
int chol_solve(cudaStream_t s, int npad, const double *L, int ld, const double *invdiag, double *v) {
  return chol_solve_batched(s, 1, npad, L, ld, 0, invdiag, 0, v, 0, nullptr);
}

// ================================================================================================
// rank-k update / downdate, k <= 8 per sweep.
//
// CHOLMOD's kernel (Modify/t_cholmod_updown_numkr.c:289-376) walks the columns j of a unit-lower
// L D L' factor; per column and per rank vector r it forms (alpha_r, gamma_r) from w_r[j] and d_j and
// then, for every row i below j:   w_r[i] -= w_r[j] * L[i][j];  L[i][j] -= gamma_r * w_r[i].
// Here the factor is L L' (d_j = l_jj^2, unit column = column / l_jj) and the sweep is panelised in
// 32-column panels: the data-dependent scalar recurrence of a panel only needs the panel's 32 x 32
// diagonal block and its 32 rows of W ("phase 1", one warp, registers + shuffles); every row below is
// then transformed independently with the 32 x (2 + 2k) coefficients of the panel ("phase 2", one
// thread per row, coalesced column-major accesses).  Phase 1 of panel p+1 is executed by the phase-2
// CTA that owns those rows, so a sweep costs one launch per panel.
// ================================================================================================
namespace updown {
constexpr int PB = 32, KMAX = 8, CO = 2 + 2 * KMAX, NT = 128;

// one warp; lane = row of the diagonal block.  Reads/writes L block and W rows in global memory.
__device__ void phase1(double *Lg, int ld, double *Wg, int ldw, int k, int sign, int j0,
                       double *alpha_state, double *coef_out, int *info) {
  const int lane = threadIdx.x & 31;
  double lrow[PB], w[KMAX], alpha[KMAX];
#pragma unroll
  for (int j = 0; j < PB; j++) lrow[j] = (j <= lane) ? Lg[(size_t)(j0 + lane) + (size_t)(j0 + j) * ld] : 0.0;
#pragma unroll
  for (int r = 0; r < KMAX; r++) {
    w[r] = (r < k) ? Wg[(size_t)(j0 + lane) + (size_t)r * ldw] : 0.0;
    alpha[r] = (r < k) ? alpha_state[r] : 1.0;
  }
  const double sg = (double)sign;
  bool bad = false;
#pragma unroll
  for (int j = 0; j < PB; j++) {
    const double ljj = __shfl_sync(0xffffffffu, lrow[j], j);
    const double winv = 1.0 / ljj;
    double dj = ljj * ljj;
    double t = lrow[j] * winv;  // unit-column entry of this lane's row
#pragma unroll
    for (int r = 0; r < KMAX; r++) {
      if (r < k) {
        const double wj = __shfl_sync(0xffffffffu, w[r], j);
        const double a = alpha[r] + sg * (wj * wj) / dj;
        dj *= a;
        const double gam = -sg * wj / dj;
        dj /= alpha[r];
        alpha[r] = a;
        if (lane > j) { w[r] -= wj * t; t -= gam * w[r]; }
        if (lane == 0) { coef_out[j * CO + 2 + r] = wj; coef_out[j * CO + 2 + KMAX + r] = gam; }
      }
    }
    const double lnew = sqrt(dj);
    if (!(dj > 0.0)) bad = true;
    if (lane > j) lrow[j] = t * lnew;
    else if (lane == j) lrow[j] = lnew;
    if (lane == 0) { coef_out[j * CO + 0] = winv; coef_out[j * CO + 1] = lnew; }
  }
#pragma unroll
  for (int j = 0; j < PB; j++)
    if (j <= lane) Lg[(size_t)(j0 + lane) + (size_t)(j0 + j) * ld] = lrow[j];
#pragma unroll
  for (int r = 0; r < KMAX; r++)
    if (r < k) {
      Wg[(size_t)(j0 + lane) + (size_t)r * ldw] = 0.0;
      if (lane == 0) alpha_state[r] = alpha[r];
    }
  if (bad && lane == 0 && info) atomicExch(info, 1);
}

__global__ void __launch_bounds__(32) k_first(double *L, int ld, double *W, int ldw, int k, int sign,
                                               double *alpha_state, double *coef, int *info) {
  if (threadIdx.x < KMAX) alpha_state[threadIdx.x] = 1.0;
  __syncwarp();
  phase1(L, ld, W, ldw, k, sign, 0, alpha_state, coef, info);
}

// panel p (columns j0..j0+31): rows i >= j0+32.  coef_in: coefficients of panel p; CTA 0 afterwards
// runs phase 1 of panel p+1 (rows j0+32..j0+63 are the first 32 rows it owns) into coef_out.
__global__ void __launch_bounds__(NT) k_panel(double *L, int ld, double *W, int ldw, int k, int sign, int j0,
                                              int npad, double *alpha_state, const double *coef_in,
                                              double *coef_out, int *info) {
  __shared__ double cs[PB * CO];
  for (int i = threadIdx.x; i < PB * CO; i += NT) cs[i] = coef_in[i];
  __syncthreads();
  const int i = j0 + PB + blockIdx.x * NT + threadIdx.x;
  if (i < npad) {
    double w[KMAX];
#pragma unroll
    for (int r = 0; r < KMAX; r++) w[r] = (r < k) ? W[(size_t)i + (size_t)r * ldw] : 0.0;
    double *Lp = L + (size_t)i + (size_t)j0 * ld;
#pragma unroll 4
    for (int j = 0; j < PB; j++) {
      double t = Lp[(size_t)j * ld] * cs[j * CO + 0];
#pragma unroll
      for (int r = 0; r < KMAX; r++) {
        if (r < k) { w[r] -= cs[j * CO + 2 + r] * t; t -= cs[j * CO + 2 + KMAX + r] * w[r]; }
      }
      Lp[(size_t)j * ld] = t * cs[j * CO + 1];
    }
#pragma unroll
    for (int r = 0; r < KMAX; r++)
      if (r < k) W[(size_t)i + (size_t)r * ldw] = w[r];
  }
  if (blockIdx.x == 0 && j0 + PB < npad) {
    __syncthreads();  // this CTA's own global writes (rows j0+32..j0+63) are visible to its warp 0
    if (threadIdx.x < 32) phase1(L, ld, W, ldw, k, sign, j0 + PB, alpha_state, coef_out, info);
  }
}
}  // namespace updown

int chol_updown(cudaStream_t s, int npad, double *L, int ld, double *W, int ldw, int k, int sign,
                double *coef, int *info_dev) {
  using namespace updown;
  if (k <= 0) return 0;
  if (k > KMAX) return 1;
  double *alpha_state = coef + 2 * PB * CO;
  QB_LAUNCH(k_first, 1, 32, 0, s, L, ld, W, ldw, k, sign, alpha_state, coef, info_dev);
  const int npanels = npad / PB;
  for (int p = 0; p < npanels; p++) {
    const int j0 = p * PB;
    const int rows = npad - j0 - PB;
    if (rows <= 0) break;
    double *cin = coef + (size_t)(p & 1) * PB * CO, *cout = coef + (size_t)((p + 1) & 1) * PB * CO;
    QB_LAUNCH(k_panel, cdiv(rows, NT), NT, 0, s, L, ld, W, ldw, k, sign, j0, npad, alpha_state, cin, cout,
              info_dev);
  }
  QB_CUDA_TRY(cudaGetLastError());
  return 0;
}

// ================================================================================================
// small helpers
// ================================================================================================
__global__ void k_copy_lower_add_diag(int n, int npad, const double *__restrict__ H, double *__restrict__ L,
                                      int ld, double diag_add) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y;
  if (i >= npad || i < j) return;
  double v;
  if (i < n && j < n) v = H[(size_t)i + (size_t)j * ld] + ((i == j) ? diag_add : 0.0);
  else v = (i == j) ? 1.0 : 0.0;
  L[(size_t)i + (size_t)j * ld] = v;
}
int copy_lower_add_diag(cudaStream_t s, int n, int npad, const double *H, double *L, int ld, double diag_add) {
  dim3 grid(cdiv(npad, 256), npad);
  QB_LAUNCH(k_copy_lower_add_diag, grid, 256, 0, s, n, npad, H, L, ld, diag_add);
  QB_CUDA_TRY(cudaGetLastError());
  return 0;
}

// row i of the symmetric matrix: entries (i, j<=i) from row i of the lower triangle, (i, j>i) from column i
__global__ void k_sym_abs_rowsums(int n, const double *__restrict__ H, int ld, double *out) {
  __shared__ double scratch[32];
  const int i = blockIdx.x;
  double acc = 0.0;
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const double v = (j <= i) ? H[(size_t)i + (size_t)j * ld] : H[(size_t)j + (size_t)i * ld];
    acc += fabs(v);
  }
  acc = block_red<RED_SUM>(acc, scratch);
  if (threadIdx.x == 0) out[i] = acc;
}
int sym_abs_rowsums(cudaStream_t s, int n, const double *H, int ld, double *out) {
  QB_LAUNCH(k_sym_abs_rowsums, n, 256, 0, s, n, H, ld, out);
  QB_CUDA_TRY(cudaGetLastError());
  return 0;
}

}  // namespace qb
