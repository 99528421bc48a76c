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
// Dataflow, two roles in one cooperative launch:
//   * CTA 0 is the CHAIN: it runs the serial step of every panel, keeps G in shared memory and publishes the panel's
//     coefficients (Ca, Y, V) through a release / acquire flag.  What it needs from outside is the panel's rows of W with
//     all earlier panels applied.  The last D of those terms (D = 2 in the 32-column shape, 1 in the 64-column shape) it
//     applies itself (it keeps the last V blocks and the sub-diagonal L blocks, prefetched one panel ahead while they still
//     hold their old values); everything older comes as a 32 x k "strip" published by the consumer that owns those rows --
//     D panel-times of slack for the consumers.  Warps 0-6 compute; warp 7 does all the chain's global traffic
//     that is not on the critical path (coefficient stores + fence + flag, strip polling and loads two panels ahead) and
//     meets the compute warps through two named barriers.
//   * CTAs 1.. are CONSUMERS: CTA c owns the 64-row blocks c-1, c-1+(G-1), ...  It applies the published coefficients
//     of every panel above its block to its rows (L tile prefetched into registers before it waits on the flag) and
//     publishes the strips of its own two panels at the right moment.
// Flags are epoch-stamped, so nothing is cleared between sweeps.  The cooperative launch guarantees co-residency.
#include "dense.cuh"
#include "chol32.cuh"
#include <map>
#include <mutex>

namespace qb {
namespace udflow {
constexpr int PB = 32, RB = 64, NT = 256, NC = NT - 32, LS = PB + 1;   // chain CTA: NC compute threads + one I/O warp (warp 7)

template <int KW>
struct Lay {   // shared-memory layouts, in doubles (the two roles are exclusive per CTA, so their layouts overlap)
  static constexpr int D = (KW == 32) ? 2 : 1;            // panels the chain applies to a strip itself
  static constexpr bool HAS_W1T = (KW == 32);             // the 64-column shape has no room for the transposed panel rows
  static constexpr int WS = KW + 1;                       // row stride of W blocks (odd: rows -> distinct banks)
  static constexpr int GS = KW + 2;                       // even row stride: 128-bit loads of 4 consecutive columns
  static constexpr int TS = PB + 2;
  static constexpr int TW = KW / 8;                       // W columns per warp in the row transforms
  static constexpr int nCoef = PB * PB + 2 * PB * KW;     // Cc[PB][PB] | Yc[KW][PB] | Vc[PB][KW]  (layout of a global ring slot)
  static constexpr int ev(int x) { return x + (x & 1); }
  // consumer
  static constexpr int oWb = 0;                           // [RB][WS]   this block's rows of W (current state)
  static constexpr int oLt = ev(oWb + RB * WS);           // [RB][LS]   L tile of the panel being applied (old values)
  static constexpr int oCoef = ev(oLt + RB * LS);
  static constexpr int consumer_total = oCoef + nCoef;
  // chain
  static constexpr int cW1 = 0;                           // [PB][WS]   the panel's rows of W
  static constexpr int cBs = ev(cW1 + PB * WS);           // [PB][WS]   B = W1 G
  static constexpr int cCoef = ev(cBs + PB * WS);         // Cc | Yc    (V lives in the ring below)
  static constexpr int cVh = ev(cCoef + PB * PB + KW * PB);   // [D+1][PB][KW]   V of the last D + 1 panels (slot p % (D+1))
  static constexpr int cG = ev(cVh + (D + 1) * PB * KW);  // [KW][GS]
  static constexpr int cYn = ev(cG + KW * GS);            // [PB][GS]   Y in natural layout
  static constexpr int cLs = ev(cYn + PB * GS);           // [PB][LS]   old diagonal block (zeros above the diagonal)
  static constexpr int cLsT = ev(cLs + PB * LS);          // [PB][TS]   its transpose
  static constexpr int cHs = ev(cLsT + PB * TS);          // [PB][LS]   first half of H11, then the new diagonal block
  static constexpr int cRd = ev(cHs + PB * LS);           // rd[32] = 1 / diag(L11), rdn[32] = 1 / diag(L11')
  static constexpr int cW1T = ev(cRd + 2 * PB);           // [KW][TS]   transposed panel rows of W (HAS_W1T)
  static constexpr int cLsub = ev(cW1T + (HAS_W1T ? KW * TS : 0));   // [D+1][D*PB][LS]   sub-diagonal blocks (old values)
  static constexpr int cWst = ev(cLsub + (D + 1) * D * PB * LS);     // [2][PB][KW]       strips of the next two panels
  static constexpr int chain_total = cWst + 2 * PB * KW;
  static constexpr int total = chain_total > consumer_total ? chain_total : consumer_total;
  static constexpr size_t bytes = sizeof(double) * (size_t)total;
  static_assert(bytes <= 232448, "shared memory per CTA");
};

__device__ __forceinline__ int ld_acquire(const int *p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// named barriers of the chain CTA: 1 = its 224 compute threads; 2 / 3 = compute warps <-> the I/O warp (all 256 threads)
__device__ __forceinline__ void bar_compute() { asm volatile("bar.sync 1, 224;" ::: "memory"); }
__device__ __forceinline__ void bar_sync_io(int id) { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void bar_arrive_io(int id) { asm volatile("bar.arrive %0, 256;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void wait_flag(const int *f, int epoch) {
  if (threadIdx.x == 0) while (ld_acquire(f) != epoch) {}
  __syncthreads();
}
__device__ __forceinline__ void publish(int *f, int epoch) {
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); st_release(f, epoch); }
}

// register-tile product used by the chain's small GEMMs: lane -> rows ri = lane & 15 and ri + 16, 4 consecutive columns.
//   acc[a][j] += sgn * sum_k A_a[k * sa] * Bt[k * sb + j]      (Bt 16-byte aligned, sb even)
template <int K, bool NEG>
__device__ __forceinline__ void mac24(double (&acc)[2][4], const double *A0, const double *A1, int sa, const double *Bt, int sb) {
#pragma unroll 8
  for (int kk = 0; kk < K; kk++) {
    double a0 = A0[kk * sa], a1 = A1[kk * sa];
    if (NEG) { a0 = -a0; a1 = -a1; }
    const double2 b01 = *reinterpret_cast<const double2 *>(Bt + kk * sb);
    const double2 b23 = *reinterpret_cast<const double2 *>(Bt + kk * sb + 2);
    acc[0][0] = fma(a0, b01.x, acc[0][0]); acc[0][1] = fma(a0, b01.y, acc[0][1]);
    acc[0][2] = fma(a0, b23.x, acc[0][2]); acc[0][3] = fma(a0, b23.y, acc[0][3]);
    acc[1][0] = fma(a1, b01.x, acc[1][0]); acc[1][1] = fma(a1, b01.y, acc[1][1]);
    acc[1][2] = fma(a1, b23.x, acc[1][2]); acc[1][3] = fma(a1, b23.y, acc[1][3]);
  }
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

__device__ __forceinline__ double fast_rcp(double p) {   // 1 / p to about 1 ulp: 20-bit seed + two Newton steps
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(p));
#pragma unroll
  for (int it = 0; it < 2; it++) x = fma(x, fma(-p, x, 1.0), x);
  return x;
}

// In-register L D L' of a 32 x 32 block by one warp (lane = row): no square root on the column chain, and the shuffles
// of the unscaled column run under the reciprocal.  After the last step lane i holds the unit-lower row in a[c], c < i,
// and d_i in a[i].
template <int J>
__device__ __forceinline__ void ldl_step(double (&a)[PB], const int lane, int &badcol) {
  const double pjj = __shfl_sync(0xffffffffu, a[J], J);
  const double inv = fast_rcp(pjj);
  const double lij = a[J] * inv;
#pragma unroll
  for (int c = J + 1; c < PB; c++) {
    const double acj = __shfl_sync(0xffffffffu, a[J], c);   // d_J L(c, J)
    a[c] = fma(-lij, acj, a[c]);
  }
  if (!(pjj > 0.0) && badcol < 0) badcol = J;
  if (lane > J) a[J] = lij;
  if constexpr (J + 1 < PB) ldl_step<J + 1>(a, lane, badcol);
}

// ---------------------------------------------------------------------------------------------------------------
// the chain CTA
// ---------------------------------------------------------------------------------------------------------------
template <int KW>
__device__ void chain_io_warp(double *sm, const double *W, int ldw, int k, int npad, double *coef_ring, const double *strip_ring,
                              int *flag_c, const int *flag_s, int epoch) {
  using S = Lay<KW>;
  constexpr int D = S::D;
  const double *CY = sm + S::cCoef, *Vh = sm + S::cVh;
  double *Wst = sm + S::cWst;
  const int lane = threadIdx.x & 31, npanels = npad / PB;
  auto load_strip = [&](int p) {   // W rows of panel p with the panels < p - D applied (the original rows for the first D + 1 panels)
    double *dst = Wst + (size_t)(p & 1) * PB * KW;
    if (p <= D) {
      for (int idx = lane; idx < PB * KW; idx += 32) {
        const int r = idx & 31, t = idx >> 5;
        dst[r * KW + t] = (t < k) ? __ldcg(W + (size_t)(p * PB + r) + (size_t)t * ldw) : 0.0;
      }
    } else {
      if (lane == 0) while (ld_acquire(flag_s + p) != epoch) {}
      __syncwarp();
      const double2 *src = reinterpret_cast<const double2 *>(strip_ring + (size_t)p * PB * KW);
      double2 *d2 = reinterpret_cast<double2 *>(dst);
#pragma unroll 8
      for (int idx = lane; idx < PB * KW / 2; idx += 32) d2[idx] = __ldcg(src + idx);
    }
  };
  load_strip(0);
  if (npanels > 1) load_strip(1);
  bar_arrive_io(3);
  for (int p = 0; p < npanels; p++) {
    bar_sync_io(2);   // the panel's coefficients are complete in shared memory
    {
      double2 *dst = reinterpret_cast<double2 *>(coef_ring + (size_t)p * S::nCoef);
      const double2 *src = reinterpret_cast<const double2 *>(CY);
      constexpr int n1 = (PB * PB + KW * PB) / 2, n2 = PB * KW / 2;
#pragma unroll 8
      for (int idx = lane; idx < n1; idx += 32) dst[idx] = src[idx];
      const double2 *srcv = reinterpret_cast<const double2 *>(Vh + (size_t)(p % (D + 1)) * PB * KW);
#pragma unroll 8
      for (int idx = lane; idx < n2; idx += 32) dst[n1 + idx] = srcv[idx];
    }
    __threadfence();
    __syncwarp();
    if (lane == 0) st_release(flag_c + p, epoch);
    if (p + 2 < npanels) load_strip(p + 2);
    bar_arrive_io(3);   // coefficient buffers may be overwritten; strip p + 2 is in place
  }
}

template <int KW>
__device__ void chain_role(double *sm, double *L, int ld, int k, int kpos, int npad, int *info, long long *clk) {
  using S = Lay<KW>;
  constexpr int D = S::D, WS = S::WS, GS = S::GS, TS = S::TS;
  int nclk = 0;
  // instrumented runs (qpalm_b200_bench_updown_clocks): thread 0 stamps the stage boundaries of panels 100 and 101
#define UD_TICK(p) do { if (clk && tid == 0 && ((p) == 100 || (p) == 101) && nclk < 30) clk[nclk++] = clock64(); } while (0)
  double *W1 = sm + S::cW1, *W1T = sm + S::cW1T, *Bs = sm + S::cBs, *Cc = sm + S::cCoef, *Yc = Cc + PB * PB, *G = sm + S::cG;
  double *Ls = sm + S::cLs, *LsT = sm + S::cLsT, *Hs = sm + S::cHs, *Yn = sm + S::cYn, *rd = sm + S::cRd, *rdn = rd + PB;
  double *Vh = sm + S::cVh, *Lsub = sm + S::cLsub;
  const double *Wst = sm + S::cWst;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int ri = lane & 15, cg = lane >> 4;   // register-tile coordinates: warps 0-3 hold the tiles, columns 8 cw + 4 cg .. + 3
  const int npanels = npad / PB;
  // thread -> element maps of the prefetches
  //   diagonal block: idx = tid + NT e, e < 4:  i = idx & 31, c = idx >> 5
  //   sub-diagonal block (D * 32 rows x 32 columns): idx = tid + NT e:  r = idx % (32 D), u = idx / (32 D)
  constexpr int ND = (PB * PB + NC - 1) / NC, NS = (D * PB * PB + NC - 1) / NC;
  double pre_d[ND], pre_s[NS];
  auto issue_prefetch = [&](int p) {     // loads for panel p into registers (old values: nobody has touched these columns yet)
    const size_t j0 = (size_t)p * PB;
    const double *Ld = L + j0 + j0 * ld;
#pragma unroll
    for (int e = 0; e < ND; e++) {
      const int idx = tid + NC * e, i = idx & 31, c = idx >> 5;
      pre_d[e] = (idx < PB * PB && c <= i) ? Ld[i + (size_t)c * ld] : 0.0;
    }
#pragma unroll
    for (int e = 0; e < NS; e++) {
      const int idx = tid + NC * e, r = idx % (D * PB), u = idx / (D * PB);
      const size_t row = j0 + PB + r;
      pre_s[e] = (idx < D * PB * PB && row < (size_t)npad) ? L[row + (j0 + u) * ld] : 0.0;
    }
  };
  auto commit_prefetch = [&](int p) {    // registers -> shared memory (Ls, its transpose, sub-diagonal slot p % (D+1))
    double *dst = Lsub + (size_t)(p % (D + 1)) * (D * PB) * LS;
#pragma unroll
    for (int e = 0; e < ND; e++) {
      const int idx = tid + NC * e, i = idx & 31, c = idx >> 5;
      if (idx < PB * PB) { Ls[i * LS + c] = pre_d[e]; LsT[c * TS + i] = pre_d[e]; }
    }
#pragma unroll
    for (int e = 0; e < NS; e++) {
      const int idx = tid + NC * e, r = idx % (D * PB), u = idx / (D * PB);
      if (idx < D * PB * PB) dst[r * LS + u] = pre_s[e];
    }
  };

  for (int idx = tid; idx < KW * KW; idx += NC) {
    const int s = idx / KW, t = idx - s * KW;
    G[s * GS + t] = (s == t) ? ((s < kpos || s >= k) ? 1.0 : -1.0) : 0.0;
  }
  issue_prefetch(0);
  commit_prefetch(0);
  bar_sync_io(3);   // strips 0 and 1 are in shared memory

  for (int p = 0; p < npanels; p++) {
    UD_TICK(p);
    // ---- W1 = strip - sum_{d=1..D} L(panel p rows, panel p-d columns) V_{p-d} ----
    if (w < 4) for (int cw = w; cw < KW / 8; cw += 4) {
      const int c0 = 8 * cw + 4 * cg;
      const double *st = Wst + (size_t)(p & 1) * PB * KW;
      double acc[2][4];
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[a][j] = st[(ri + 16 * a) * KW + c0 + j];
#pragma unroll
      for (int d = 1; d <= D; d++) {
        if (p - d >= 0) {
          const double *Lb = Lsub + (size_t)((p - d) % (D + 1)) * (D * PB) * LS + ((d - 1) * PB + ri) * LS;
          mac24<PB, true>(acc, Lb, Lb + 16 * LS, 1, Vh + (size_t)((p - d) % (D + 1)) * PB * KW + c0, KW);
        }
      }
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int j = 0; j < 4; j++) {
          W1[(ri + 16 * a) * WS + c0 + j] = acc[a][j];
          if (S::HAS_W1T) W1T[(c0 + j) * TS + ri + 16 * a] = acc[a][j];
        }
    }
    if (tid < PB) rd[tid] = 1.0 / Ls[tid * LS + tid];
    bar_compute();
    UD_TICK(p);   // 1: strip combine
    if (p + 1 < npanels) issue_prefetch(p + 1);   // lands while this panel's serial step runs
    if (w < 4) for (int cw = w; cw < KW / 8; cw += 4) {   // B = W1 G
      const int c0 = 8 * cw + 4 * cg;
      double acc[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
      mac24<KW, false>(acc, W1 + ri * WS, W1 + (ri + 16) * WS, 1, G + c0, GS);
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int j = 0; j < 4; j++) Bs[(ri + 16 * a) * WS + c0 + j] = acc[a][j];
    }
    bar_compute();
    UD_TICK(p);   // 2: B
    if (w < 4) {   // H11 = L11 L11' + B W1'
      const int c0 = 8 * w + 4 * cg;
      double acc[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
      mac24<PB, false>(acc, Ls + ri * LS, Ls + (ri + 16) * LS, 1, LsT + c0, TS);
      if (S::HAS_W1T) mac24<KW, false>(acc, Bs + ri * WS, Bs + (ri + 16) * WS, 1, W1T + c0, TS);
      else {
#pragma unroll 8
        for (int t = 0; t < KW; t++) {
          const double a0 = Bs[ri * WS + t], a1 = Bs[(ri + 16) * WS + t];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const double b = W1[(c0 + j) * WS + t];
            acc[0][j] = fma(a0, b, acc[0][j]); acc[1][j] = fma(a1, b, acc[1][j]);
          }
        }
      }
#pragma unroll
      for (int a = 0; a < 2; a++)
#pragma unroll
        for (int j = 0; j < 4; j++) Hs[(ri + 16 * a) * LS + c0 + j] = acc[a][j];
    }
    bar_compute();
    UD_TICK(p);   // 3: H
    double *Vcur = Vh + (size_t)(p % (D + 1)) * PB * KW;
    if (w == 0) {   // L11' = chol(H11): L D L' in registers, then scaled by sqrt(D)
      double a[PB];
#pragma unroll
      for (int c = 0; c < PB; c++) a[c] = (c <= lane) ? Hs[lane * LS + c] : 0.0;
      int badcol = -1;
      ldl_step<0>(a, lane, badcol);
      if (badcol >= 0 && lane == 0 && info) atomicExch(info, 1);
      double dself = 0.0;
#pragma unroll
      for (int c = 0; c < PB; c++) dself = (c == lane) ? a[c] : dself;   // static register indices only
      const double sd = sqrt(dself);
      rdn[lane] = 1.0 / sd;
      double *Ld = L + (size_t)p * PB * (ld + 1);
#pragma unroll
      for (int c = 0; c < PB; c++) {
        const double sc = __shfl_sync(0xffffffffu, sd, c);
        const double v = (c < lane) ? a[c] * sc : sd;
        if (c <= lane) { Hs[lane * LS + c] = v; Ld[lane + (size_t)c * ld] = v; }
      }
    } else if (w <= KW / 32) {   // V = inv(L11) W1   (lane = column of W)
      const int t = (w - 1) * 32 + lane;
      double v[PB];
#pragma unroll
      for (int r = 0; r < PB; r++) v[r] = W1[r * WS + t];
      trs<0>(Ls, rd, v);
#pragma unroll
      for (int r = 0; r < PB; r++) Vcur[r * KW + t] = v[r];
    }
    if (p > 0) bar_sync_io(3);   // the I/O warp has stored the previous panel's coefficients: Cc / Yc may be overwritten
    bar_compute();
    UD_TICK(p);   // 4: chol | V solve
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
      for (int r = 0; r < PB; r++) { Yc[t * PB + r] = v[r]; Yn[r * GS + t] = v[r]; }
    }
    bar_compute();
    bar_arrive_io(2);   // hand the coefficients to the I/O warp (stores, fence, flag)
    UD_TICK(p);   // 5: Ca | Y solves
    // G <- G - Y' Y
#pragma unroll
    for (int half = 0; half < KW / 32; half++) {
      if (w < 4) for (int cw = w; cw < KW / 8; cw += 4) {
        const int c0 = 8 * cw + 4 * cg;
        double acc[2][4] = {{0.0, 0.0, 0.0, 0.0}, {0.0, 0.0, 0.0, 0.0}};
        mac24<PB, true>(acc, Yn + 32 * half + ri, Yn + 32 * half + ri + 16, GS, Yn + c0, GS);
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
          for (int j = 0; j < 4; j++) G[(32 * half + ri + 16 * a) * GS + c0 + j] += acc[a][j];
      }
    }
    UD_TICK(p);   // 6: G update
    if (p + 1 < npanels) commit_prefetch(p + 1);
    bar_compute();
    UD_TICK(p);   // 7: prefetch commit (waits for the loads) + barrier
  }
  bar_sync_io(3);   // pairs with the I/O warp's last arrival
#undef UD_TICK
}

// ---------------------------------------------------------------------------------------------------------------
// a consumer CTA
// ---------------------------------------------------------------------------------------------------------------
template <int KW>
__device__ void consumer_role(double *sm, double *L, int ld, const double *W, int ldw, int k, int npad, const double *coef_ring,
                              double *strip_ring, const int *flag_c, int *flag_s, int epoch, int first_block, int block_step) {
  using S = Lay<KW>;
  constexpr int WS = S::WS, D = S::D;
  double *Wb = sm + S::oWb, *Lt = sm + S::oLt, *Coef = sm + S::oCoef;
  const int tid = threadIdx.x;
  const int nblk = npad / RB;
  for (int b = first_block; b < nblk; b += block_step) {
    const int r0 = b * RB;
    __syncthreads();
    for (int idx = tid; idx < RB * KW; idx += NT) {
      const int r = idx & (RB - 1), t = idx >> 6;
      Wb[r * WS + t] = (t < k) ? __ldcg(W + (size_t)(r0 + r) + (size_t)t * ldw) : 0.0;
    }
    __syncthreads();
    for (int q = 0; q <= 2 * b; q++) {
      const int rlo = (q == 2 * b) ? PB : 0;   // the block's first panel only reaches its rows 32..63 (rows 0..31 are the panel itself)
      double treg[RB * PB / NT];
      double *Lq = L + (size_t)r0 + (size_t)q * PB * ld;
#pragma unroll
      for (int e = 0; e < RB * PB / NT; e++) {
        const int idx = tid + NT * e, r = idx & (RB - 1), u = idx >> 6;
        treg[e] = (r >= rlo) ? Lq[r + (size_t)u * ld] : 0.0;          // in flight while waiting for the panel
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
      phase2<KW>(sm, Lq, ld, rlo);
      // the chain applies the last D panels to a strip itself: publish the rows of panel ps once the panels < ps - D are in
      const int ps = q + D + 1;
      if (ps == 2 * b || ps == 2 * b + 1) {
        const double *src = Wb + (size_t)(ps - 2 * b) * PB * WS;
        double *dst = strip_ring + (size_t)ps * PB * KW;
        for (int idx = tid; idx < PB * KW; idx += NT) {
          const int r = idx / KW, t = idx - r * KW;
          dst[idx] = src[r * WS + t];
        }
        publish(flag_s + ps, epoch);
      }
    }
  }
}

template <int KW>
__global__ void __launch_bounds__(NT, 1)
k_updown_flow(double *L, int ld, const double *W, int ldw, int k, int kpos, int npad, double *coef_ring, double *strip_ring,
              int *flag_c, int *flag_s, int epoch, int *info, long long *clk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sm = reinterpret_cast<double *>(smem_raw);
  if (blockIdx.x == 0) {
    if (threadIdx.x >= NC) chain_io_warp<KW>(sm, W, ldw, k, npad, coef_ring, strip_ring, flag_c, flag_s, epoch);
    else chain_role<KW>(sm, L, ld, k, kpos, npad, info, clk);
  }
  else consumer_role<KW>(sm, L, ld, W, ldw, k, npad, coef_ring, strip_ring, flag_c, flag_s, epoch, (int)blockIdx.x - 1, (int)gridDim.x - 1);
}

struct State { int *flags = nullptr; double *ring = nullptr; int cap_blk = 0; int epoch = 0; int max_grid[2] = {0, 0}; };
static std::mutex g_mu;
static std::map<cudaStream_t, State> g_state;

template <int KW>
static int launch(cudaStream_t s, State &st, int slot, int npad, double *L, int ld, const double *W, int ldw, int k, int kpos, int *info_dev,
                  long long *clk) {
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
  if (st.max_grid[slot] < 2) return 1;
  const int grid = nblk + 1 < st.max_grid[slot] ? nblk + 1 : st.max_grid[slot];   // the chain CTA + consumers
  double *coef_ring = st.ring, *strip_ring = st.ring + (size_t)npanels * Lay<64>::nCoef;
  int *flag_c = st.flags, *flag_s = st.flags + npanels;
  int epoch = st.epoch;
  void *args[] = {(void *)&L, (void *)&ld, (void *)&W, (void *)&ldw, (void *)&k, (void *)&kpos, (void *)&npad, (void *)&coef_ring,
                  (void *)&strip_ring, (void *)&flag_c, (void *)&flag_s, (void *)&epoch, (void *)&info_dev, (void *)&clk};
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
int chol_updown_flow(cudaStream_t s, int npad, double *L, int ld, const double *W, int ldw, int k, int kpos, int *info_dev,
                     long long *clk_dev) {
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
    QB_CUDA_TRY(cudaMalloc(&st.flags, sizeof(int) * (size_t)(2 * npanels + 2)));
    QB_CUDA_TRY(cudaMemsetAsync(st.flags, 0, sizeof(int) * (size_t)(2 * npanels + 2), s));
    QB_CUDA_TRY(cudaMalloc(&st.ring, sizeof(double) * ((size_t)npanels * Lay<64>::nCoef + (size_t)npanels * PB * 64)));
    st.epoch = 0;
  }
  st.epoch++;
  return k <= 32 ? launch<32>(s, st, 0, npad, L, ld, W, ldw, k, kpos, info_dev, clk_dev)
                 : launch<64>(s, st, 1, npad, L, ld, W, ldw, k, kpos, info_dev, clk_dev);
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
