// updown_flow.cu -- rank-k update / downdate of the dense Cholesky factor as ONE cooperative dataflow launch.
//
// Replaces cholmod_updown (Modify/cholmod_updown.c, kernel Modify/t_cholmod_updown_numkr.c:289-376) as called by
// ldlupdate_entering_constraints / ldldowndate_leaving_constraints / ldlupdate_sigma_changed
// (src/solver_interface.c:407-503) for the dense Newton system:   L L'  <-  L L' + W S W',   S = diag(+1 .. +1, -1 .. -1).
//
// CHOLMOD walks the columns one by one with a scalar (alpha, gamma) recurrence per rank, <= 8 ranks per sweep: n * k
// dependent steps.  Here the same factor is reached with a BLOCKED recurrence whose serial part does not grow with k.
// State: the k x k symmetric weight G (G = S at the start), so that the matrix still to be absorbed below panel p is
// W G W'.  For a 32-column panel with diagonal block L11, its rows of W  W1 (32 x k) and everything below (L21, W2):
//     B    = W1 G                       H11 = L11 L11' + B W1'          L11' = chol(H11)            (32 x 32, one warp)
//     V    = inv(L11) W1                Y   = inv(L11') B               Ca'  = inv(L11') L11
//     L21' = L21 Ca + W2 Y'             W2' = W2 - L21 V                G'   = G - Y' Y
// (block elimination of [L11 W1 G^1/2; L21 W2 G^1/2]: the projector onto the null space of the first block row gives the
// W2', G' pair.)  An update and a downdate differ only in the sign pattern of G, so entering AND leaving constraints go
// through one sweep.  Per panel the chain is: three 32 x 32 x k products, one in-register 32 x 32 Cholesky (chol32.cuh) and
// two triangular solves with 32 + k right-hand sides -- independent of n and nearly independent of k (k <= 64 per sweep).
//
// Dataflow: CTA c owns the 64-row blocks c, c + G, ...  It applies the published coefficients (Ca, Y, V) of every panel
// above its block to its rows (L tile prefetched into registers before it waits on the panel's flag), then runs the
// serial step for the two panels of its own diagonal block and publishes their coefficients (release / acquire flags in
// global memory, epoch-stamped so nothing is cleared between sweeps).  The cooperative launch guarantees co-residency.
#include "dense.cuh"
#include "chol32.cuh"
#include <map>
#include <mutex>

namespace qb {
namespace udflow {
constexpr int PB = 32, RB = 64, NT = 256, LS = PB + 1;

template <int KW>
struct Lay {   // shared-memory layout, in doubles
  static constexpr int WS = KW + 1;                       // row stride of the W block (odd: rows -> distinct banks)
  static constexpr int TW = KW / 8;                       // W columns per warp in the row transforms
  static constexpr int oWb = 0;                           // [RB][WS]   this block's rows of W (current state)
  static constexpr int oLt = oWb + RB * WS + ((RB * WS) & 1);   // [RB][LS]   L tile of the panel being applied (old values)
  static constexpr int oCoef = oLt + RB * LS + ((RB * LS) & 1); // Cc[PB][PB] | Yc[KW][PB] | Vc[PB][KW]  (same layout in the global ring)
  static constexpr int nCoef = PB * PB + 2 * PB * KW;
  static constexpr int oG = oCoef + nCoef;                // [KW][WS]
  static constexpr int oLs = oG + KW * WS;                // [PB][LS]   old diagonal block (zeros above the diagonal)
  static constexpr int oHs = oLs + PB * LS;               // [PB][LS]   H11, then the new diagonal block
  static constexpr int oBs = oHs + PB * LS;               // [PB][WS]   B = W1 G, later Y in natural layout
  static constexpr int oRd = oBs + PB * WS;               // rd[32] = 1 / diag(L11), rdn[32] = 1 / diag(L11')
  static constexpr int total = oRd + 2 * PB;
  static constexpr size_t bytes = sizeof(double) * (size_t)total;
};

__device__ __forceinline__ int ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_flag(const int *f, int epoch) {
  if (threadIdx.x == 0) while (ld_acquire(f) != epoch) {}
  __syncthreads();
}
__device__ __forceinline__ void publish(int *f, int epoch) {
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); st_release(f, epoch); }
}

// forward substitution  v <- inv(T) v  for one right-hand side per lane, T lower triangular in shared memory (row stride
// LS), rdiag = 1 / diag(T).  Right-looking, fully unrolled: 31 - R independent FMAs per step, broadcast operand loads.
template <int R>
__device__ __forceinline__ void trs(const double *T, const double *rdiag, double (&v)[PB]) {
  const double xr = v[R] * rdiag[R];
  v[R] = xr;
#pragma unroll
  for (int q = R + 1; q < PB; q++) v[q] = fma(-T[q * LS + R], xr, v[q]);
  if constexpr (R + 1 < PB) trs<R + 1>(T, rdiag, v);
}

// rows [rlo, RB) of the block:  L(:, panel) <- L Ca + W Y',  W <- W - L V.   Lt holds the old L tile, Cc / Yc / Vc the
// panel's coefficients.  lane -> rows (lane, lane + 32); warp w -> L columns 4w .. 4w+3 and W columns TW w .. TW w + TW - 1.
template <int KW>
__device__ __forceinline__ void phase2(double *sm, double *Lg, int ld, int rlo) {
  using S = Lay<KW>;
  constexpr int WS = S::WS, TW = S::TW;
  double *Wb = sm + S::oWb;
  const double *Lt = sm + S::oLt, *Cc = sm + S::oCoef, *Yc = Cc + PB * PB, *Vc = Yc + KW * PB;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int ra = lane, rb = lane + 32;
  double al[2][4], aw[2][TW];
#pragma unroll
  for (int j = 0; j < 4; j++) al[0][j] = al[1][j] = 0.0;
#pragma unroll
  for (int j = 0; j < TW; j++) aw[0][j] = aw[1][j] = 0.0;
  const int umax = 4 * w + 3;   // Ca is upper triangular: column c only takes u <= c
#pragma unroll
  for (int u = 0; u < PB; u++) {
    const double a0 = Lt[ra * LS + u], a1 = Lt[rb * LS + u];
    if (u <= umax) {
      const double2 c01 = *reinterpret_cast<const double2 *>(Cc + u * PB + 4 * w);
      const double2 c23 = *reinterpret_cast<const double2 *>(Cc + u * PB + 4 * w + 2);
      al[0][0] = fma(a0, c01.x, al[0][0]); al[0][1] = fma(a0, c01.y, al[0][1]);
      al[0][2] = fma(a0, c23.x, al[0][2]); al[0][3] = fma(a0, c23.y, al[0][3]);
      al[1][0] = fma(a1, c01.x, al[1][0]); al[1][1] = fma(a1, c01.y, al[1][1]);
      al[1][2] = fma(a1, c23.x, al[1][2]); al[1][3] = fma(a1, c23.y, al[1][3]);
    }
#pragma unroll
    for (int j = 0; j < TW; j += 2) {
      const double2 v2 = *reinterpret_cast<const double2 *>(Vc + u * KW + TW * w + j);
      aw[0][j] = fma(a0, v2.x, aw[0][j]); aw[0][j + 1] = fma(a0, v2.y, aw[0][j + 1]);
      aw[1][j] = fma(a1, v2.x, aw[1][j]); aw[1][j + 1] = fma(a1, v2.y, aw[1][j + 1]);
    }
  }
#pragma unroll 8
  for (int t = 0; t < KW; t++) {
    const double b0 = Wb[ra * WS + t], b1 = Wb[rb * WS + t];
    const double2 y01 = *reinterpret_cast<const double2 *>(Yc + t * PB + 4 * w);
    const double2 y23 = *reinterpret_cast<const double2 *>(Yc + t * PB + 4 * w + 2);
    al[0][0] = fma(b0, y01.x, al[0][0]); al[0][1] = fma(b0, y01.y, al[0][1]);
    al[0][2] = fma(b0, y23.x, al[0][2]); al[0][3] = fma(b0, y23.y, al[0][3]);
    al[1][0] = fma(b1, y01.x, al[1][0]); al[1][1] = fma(b1, y01.y, al[1][1]);
    al[1][2] = fma(b1, y23.x, al[1][2]); al[1][3] = fma(b1, y23.y, al[1][3]);
  }
  __syncthreads();   // every read of the old W block is done
  const bool wa = ra >= rlo;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    double *col = Lg + (size_t)(4 * w + j) * ld;
    if (wa) col[ra] = al[0][j];
    col[rb] = al[1][j];
  }
#pragma unroll
  for (int j = 0; j < TW; j++) {
    if (wa) Wb[ra * WS + TW * w + j] -= aw[0][j];
    Wb[rb * WS + TW * w + j] -= aw[1][j];
  }
  __syncthreads();
}

// serial step of one panel: rows pr .. pr+31 of the block are the panel's diagonal rows.  Leaves the coefficients in
// shared memory (and in the global ring slot), the new diagonal block in global memory and G' in shared memory.
template <int KW>
__device__ __forceinline__ void phase1(double *sm, double *Ldiag, int ld, int pr, double *coef_g, int *info) {
  using S = Lay<KW>;
  constexpr int WS = S::WS, TW = S::TW;
  double *Wb = sm + S::oWb, *Cc = sm + S::oCoef, *Yc = Cc + PB * PB, *Vc = Yc + KW * PB, *G = sm + S::oG;
  double *Ls = sm + S::oLs, *Hs = sm + S::oHs, *Bs = sm + S::oBs, *rd = sm + S::oRd, *rdn = rd + PB;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  for (int idx = tid; idx < PB * PB; idx += NT) {
    const int i = idx & 31, c = idx >> 5;
    Ls[i * LS + c] = (c <= i) ? Ldiag[i + (size_t)c * ld] : 0.0;
  }
  {   // B = W1 G   (lane = panel row, warp -> TW columns)
    double acc[TW];
#pragma unroll
    for (int j = 0; j < TW; j++) acc[j] = 0.0;
#pragma unroll 8
    for (int s = 0; s < KW; s++) {
      const double a = Wb[(pr + lane) * WS + s];
#pragma unroll
      for (int j = 0; j < TW; j++) acc[j] = fma(a, G[s * WS + TW * w + j], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < TW; j++) Bs[lane * WS + TW * w + j] = acc[j];
  }
  __syncthreads();
  if (tid < PB) rd[tid] = 1.0 / Ls[tid * LS + tid];
  {   // H11 = L11 L11' + B W1'   (lane = row, warp -> 4 columns)
    double acc[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int u = 0; u < PB; u++) {
      const double a = Ls[lane * LS + u];
#pragma unroll
      for (int j = 0; j < 4; j++) acc[j] = fma(a, Ls[(4 * w + j) * LS + u], acc[j]);
    }
#pragma unroll 8
    for (int t = 0; t < KW; t++) {
      const double b = Bs[lane * WS + t];
#pragma unroll
      for (int j = 0; j < 4; j++) acc[j] = fma(b, Wb[(pr + 4 * w + j) * WS + t], acc[j]);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) Hs[lane * LS + 4 * w + j] = acc[j];
  }
  __syncthreads();
  if (w == 0) {   // L11' = chol(H11), in registers (lane = row)
    double a[PB];
#pragma unroll
    for (int c = 0; c < PB; c++) a[c] = (c <= lane) ? Hs[lane * LS + c] : 0.0;
    double dl = 0.0, dinv = 0.0;
    int badcol = -1;
    chol32::fstep<0>(a, lane, dl, dinv, badcol);
    if (badcol >= 0 && lane == 0 && info) atomicExch(info, 1);
#pragma unroll
    for (int c = 0; c < PB; c++) if (c < lane) Hs[lane * LS + c] = a[c];
    Hs[lane * LS + lane] = dl;
    rdn[lane] = dinv;
  } else if (w <= KW / 32) {   // V = inv(L11) W1   (lane = column of W)
    const int t = (w - 1) * 32 + lane;
    double v[PB];
#pragma unroll
    for (int r = 0; r < PB; r++) v[r] = Wb[(pr + r) * WS + t];
    trs<0>(Ls, rd, v);
#pragma unroll
    for (int r = 0; r < PB; r++) Vc[r * KW + t] = v[r];
  }
  __syncthreads();
  if (w == 0) {   // Ca' = inv(L11') L11 : column `lane` of L11 as the right-hand side
    double v[PB];
#pragma unroll
    for (int r = 0; r < PB; r++) v[r] = Ls[r * LS + lane];
    trs<0>(Hs, rdn, v);
#pragma unroll
    for (int r = 0; r < PB; r++) Cc[lane * PB + r] = v[r];   // Cc[u][c] = Ca(u, c) = (inv(L11') L11)(c, u)
  } else if (w <= KW / 32) {   // Y = inv(L11') B
    const int t = (w - 1) * 32 + lane;
    double v[PB];
#pragma unroll
    for (int r = 0; r < PB; r++) v[r] = Bs[r * WS + t];
    trs<0>(Hs, rdn, v);
#pragma unroll
    for (int r = 0; r < PB; r++) { Yc[t * PB + r] = v[r]; Bs[r * WS + t] = v[r]; }   // Bs now holds Y (natural layout)
  }
  __syncthreads();
  {   // coefficients -> global ring, new diagonal block -> L
    const double2 *src = reinterpret_cast<const double2 *>(Cc);
    double2 *dst = reinterpret_cast<double2 *>(coef_g);
    for (int idx = tid; idx < S::nCoef / 2; idx += NT) dst[idx] = src[idx];
    for (int idx = tid; idx < PB * PB; idx += NT) {
      const int i = idx & 31, c = idx >> 5;
      if (c <= i) Ldiag[i + (size_t)c * ld] = Hs[i * LS + c];
    }
  }
}

// G <- G - Y' Y   (lane -> rows s = lane (+32), warp -> TW columns; Y in natural layout in Bs)
template <int KW>
__device__ __forceinline__ void update_G(double *sm) {
  using S = Lay<KW>;
  constexpr int WS = S::WS, TW = S::TW, NR = KW / 32;
  double *G = sm + S::oG;
  const double *Yn = sm + S::oBs;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double acc[NR][TW];
#pragma unroll
  for (int a = 0; a < NR; a++)
#pragma unroll
    for (int j = 0; j < TW; j++) acc[a][j] = 0.0;
#pragma unroll 8
  for (int r = 0; r < PB; r++) {
    double ys[NR], yt[TW];
#pragma unroll
    for (int a = 0; a < NR; a++) ys[a] = Yn[r * WS + lane + 32 * a];
#pragma unroll
    for (int j = 0; j < TW; j++) yt[j] = Yn[r * WS + TW * w + j];
#pragma unroll
    for (int a = 0; a < NR; a++)
#pragma unroll
      for (int j = 0; j < TW; j++) acc[a][j] = fma(ys[a], yt[j], acc[a][j]);
  }
#pragma unroll
  for (int a = 0; a < NR; a++)
#pragma unroll
    for (int j = 0; j < TW; j++) G[(lane + 32 * a) * WS + TW * w + j] -= acc[a][j];
}

template <int KW>
__global__ void __launch_bounds__(NT, 1)
k_updown_flow(double *L, int ld, const double *W, int ldw, int k, int kpos, int npad, double *coef_ring, double *g_ring,
              int *flag_c, int *flag_g, int epoch, int *info) {
  using S = Lay<KW>;
  constexpr int WS = S::WS;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sm = reinterpret_cast<double *>(smem_raw);
  double *Wb = sm + S::oWb, *Lt = sm + S::oLt, *Coef = sm + S::oCoef, *G = sm + S::oG;
  const int tid = threadIdx.x;
  const int nblk = npad / RB, npanels = npad / PB;
  for (int b = blockIdx.x; b < nblk; b += gridDim.x) {
    const int r0 = b * RB;
    __syncthreads();
    for (int idx = tid; idx < RB * KW; idx += NT) {
      const int r = idx & (RB - 1), t = idx >> 6;
      Wb[r * WS + t] = (t < k) ? __ldcg(W + (size_t)(r0 + r) + (size_t)t * ldw) : 0.0;
    }
    __syncthreads();
    // ---- panels above the block: apply the published coefficients ----
    for (int q = 0; q < 2 * b; q++) {
      double treg[RB * PB / NT];
      const double *Lq = L + (size_t)r0 + (size_t)q * PB * ld;
#pragma unroll
      for (int e = 0; e < RB * PB / NT; e++) {
        const int idx = tid + NT * e, r = idx & (RB - 1), u = idx >> 6;
        treg[e] = Lq[r + (size_t)u * ld];          // in flight while waiting for the panel
      }
      wait_flag(flag_c + q, epoch);
      {
        const double2 *src = reinterpret_cast<const double2 *>(coef_ring + (size_t)q * S::nCoef);
        double2 *dst = reinterpret_cast<double2 *>(Coef);
#pragma unroll 4
        for (int idx = tid; idx < S::nCoef / 2; idx += NT) dst[idx] = __ldcg(src + idx);
      }
#pragma unroll
      for (int e = 0; e < RB * PB / NT; e++) {
        const int idx = tid + NT * e, r = idx & (RB - 1), u = idx >> 6;
        Lt[r * LS + u] = treg[e];
      }
      __syncthreads();
      phase2<KW>(sm, L + (size_t)r0 + (size_t)q * PB * ld, ld, 0);
    }
    // ---- the block's own two panels ----
    for (int half = 0; half < 2; half++) {
      const int q = 2 * b + half;
      if (half == 0) {
        if (b == 0) {
          for (int idx = tid; idx < KW * KW; idx += NT) {
            const int s = idx / KW, t = idx - s * KW;
            G[s * WS + t] = (s == t) ? ((s < kpos || s >= k) ? 1.0 : -1.0) : 0.0;
          }
        } else {
          wait_flag(flag_g + b, epoch);
          const double *src = g_ring + (size_t)b * KW * KW;
          for (int idx = tid; idx < KW * KW; idx += NT) {
            const int s = idx / KW, t = idx - s * KW;
            G[s * WS + t] = __ldcg(src + idx);
          }
        }
        __syncthreads();
      }
      phase1<KW>(sm, L + (size_t)(r0 + PB * half) * (ld + 1), ld, PB * half, coef_ring + (size_t)q * S::nCoef, info);
      publish(flag_c + q, epoch);
      update_G<KW>(sm);
      __syncthreads();
      if (half == 0) {
        const double *Lq = L + (size_t)r0 + (size_t)q * PB * ld;
        for (int idx = tid; idx < PB * PB; idx += NT) {
          const int r = PB + (idx & 31), u = idx >> 5;
          Lt[r * LS + u] = Lq[r + (size_t)u * ld];
        }
        __syncthreads();
        phase2<KW>(sm, L + (size_t)r0 + (size_t)q * PB * ld, ld, PB);
      } else if (q + 1 < npanels) {
        double *dst = g_ring + (size_t)(b + 1) * KW * KW;
        for (int idx = tid; idx < KW * KW; idx += NT) {
          const int s = idx / KW, t = idx - s * KW;
          dst[idx] = G[s * WS + t];
        }
        publish(flag_g + b + 1, epoch);
      }
    }
  }
}

struct State { int *flags = nullptr; double *ring = nullptr; int cap_blk = 0; int epoch = 0; int max_grid[2] = {0, 0}; };
static std::mutex g_mu;
static std::map<cudaStream_t, State> g_state;

template <int KW>
static int launch(cudaStream_t s, State &st, int slot, int npad, double *L, int ld, const double *W, int ldw, int k, int kpos, int *info_dev) {
  using S = Lay<KW>;
  if (st.max_grid[slot] == 0) {
    int dev = 0, coop = 0, sms = 0, per_sm = 0;
    QB_CUDA_TRY(cudaGetDevice(&dev));
    QB_CUDA_TRY(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    QB_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    QB_CUDA_TRY(cudaFuncSetAttribute(k_updown_flow<KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::bytes));
    QB_CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_updown_flow<KW>, NT, S::bytes));
    st.max_grid[slot] = (coop && per_sm > 0) ? sms * per_sm : -1;
  }
  if (st.max_grid[slot] <= 0) return 1;
  const int nblk = npad / RB, npanels = npad / PB;
  const int grid = nblk < st.max_grid[slot] ? nblk : st.max_grid[slot];
  double *coef_ring = st.ring, *g_ring = st.ring + (size_t)npanels * Lay<64>::nCoef;
  int *flag_c = st.flags, *flag_g = st.flags + npanels;
  int epoch = st.epoch;
  void *args[] = {(void *)&L, (void *)&ld, (void *)&W, (void *)&ldw, (void *)&k, (void *)&kpos, (void *)&npad, (void *)&coef_ring,
                  (void *)&g_ring, (void *)&flag_c, (void *)&flag_g, (void *)&epoch, (void *)&info_dev};
  const bool prof = g_prof_on && prof_begin("udflow::k_updown_flow", s);
  const cudaError_t err = cudaLaunchCooperativeKernel((const void *)k_updown_flow<KW>, dim3(grid), dim3(NT), args, S::bytes, s);
  if (prof) prof_end(s);
  if (err != cudaSuccess) { (void)cudaGetLastError(); st.max_grid[slot] = -1; return 1; }
  ++g_kernel_launches;
  return 0;
}
}  // namespace udflow

int chol_updown_flow_max_rank() { return 64; }

// returns 0 when the sweep ran, 1 when the dataflow kernel is not available on this device / partition (caller uses the
// per-panel launches of chol_updown), < 0 on a CUDA error
int chol_updown_flow(cudaStream_t s, int npad, double *L, int ld, const double *W, int ldw, int k, int kpos, int *info_dev) {
  using namespace udflow;
  if (k <= 0) return 0;
  if (k > 64 || (npad % RB)) return 1;
  std::lock_guard<std::mutex> lk(g_mu);
  State &st = g_state[s];
  const int nblk = npad / RB, npanels = npad / PB;
  if (st.cap_blk < nblk) {
    if (st.flags) QB_CUDA_TRY(cudaFree(st.flags));
    if (st.ring) QB_CUDA_TRY(cudaFree(st.ring));
    st.cap_blk = nblk;
    QB_CUDA_TRY(cudaMalloc(&st.flags, sizeof(int) * (size_t)(npanels + nblk + 2)));
    QB_CUDA_TRY(cudaMemsetAsync(st.flags, 0, sizeof(int) * (size_t)(npanels + nblk + 2), s));
    QB_CUDA_TRY(cudaMalloc(&st.ring, sizeof(double) * ((size_t)npanels * Lay<64>::nCoef + (size_t)(nblk + 1) * 64 * 64)));
    st.epoch = 0;
  }
  st.epoch++;
  return k <= 32 ? launch<32>(s, st, 0, npad, L, ld, W, ldw, k, kpos, info_dev)
                 : launch<64>(s, st, 1, npad, L, ld, W, ldw, k, kpos, info_dev);
}

void chol_updown_flow_release(cudaStream_t s) {
  std::lock_guard<std::mutex> lk(udflow::g_mu);
  auto it = udflow::g_state.find(s);
  if (it == udflow::g_state.end()) return;
  if (it->second.flags) cudaFree(it->second.flags);
  if (it->second.ring) cudaFree(it->second.ring);
  udflow::g_state.erase(it);
}

}  // namespace qb
