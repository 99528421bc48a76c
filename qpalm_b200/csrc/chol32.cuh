// chol32.cuh -- in-register Cholesky of a 32 x 32 block by ONE warp (lane = row), shared by the dense diagonal-block
// kernel (dense.cu) and the supernodal fronts (sparse.cu).
//
// The column loop is unrolled with templates so that every register index is static (a `#pragma unroll` over a
// column-dependent inner bound leaves the row in local memory: measured 233 us vs 67 us for the 128 x 128 block kernel).
// Per column the critical path is one shuffle, the branch-free sqrt/reciprocal below and one FMA.
#pragma once
#include "common.cuh"

namespace qb {
namespace chol32 {
constexpr int SB = 32;
constexpr unsigned FULL = 0xffffffffu;

// Branch-free sqrt(p) and 1/sqrt(p) for a normal positive p: hardware 1/sqrt seed (about 22 bits), two coupled Newton
// steps on (g, h) ~ (sqrt p, 1/(2 sqrt p)), Markstein's final correction for g (correctly rounded sqrt), then the
// reciprocal of g from the seed 2h with two FMA corrections.  Same values as sqrt() and 1.0 / sqrt() of the reference
// arithmetic for normal inputs, but without their slow-path calls, so the 32-column factorization stays one straight-line
// block the scheduler can interleave.  A non-positive p gives NaN (flagged by the caller).
__device__ __forceinline__ void sqrt_and_rcp(double p, double &g, double &x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(p));
  g = p * y;
  double h = 0.5 * y;
#pragma unroll
  for (int it = 0; it < 2; it++) {
    const double r = fma(-h, g, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
  }
  g = fma(fma(-g, g, p), h, g);
  x = h + h;
#pragma unroll
  for (int it = 0; it < 2; it++) x = fma(x, fma(-g, x, 1.0), x);
}

// After the call lane r holds L(r, c) in a[c] for c < r, its diagonal entry L(r, r) in dl and 1 / L(r, r) in dinv;
// badcol = first column with a non-positive pivot (or stays -1).  The upper triangle of a[] must come in as zeros.
template <int J>
__device__ __forceinline__ void fstep(double (&a)[SB], const int lane, double &dl, double &dinv, int &badcol) {
  const double pjj = __shfl_sync(FULL, a[J], J);
  double ljj, inv;
  sqrt_and_rcp(pjj, ljj, inv);
  const double lij = a[J] * inv;           // L(row, J)
#pragma unroll
  for (int c = J + 1; c < SB; c++) {
    const double lcj = __shfl_sync(FULL, a[J], c) * inv;   // L(c, J), the value lane c itself stores
    a[c] = fma(-lij, lcj, a[c]);
  }
  if (!(pjj > 0.0) && badcol < 0) badcol = J;
  a[J] = lij;
  if (lane == J) { dl = ljj; dinv = inv; }
  if constexpr (J + 1 < SB) fstep<J + 1>(a, lane, dl, dinv, badcol);
}

// Signed variant for quasi-definite blocks (the KKT factorization L S L', S = diag(+-1) known a priori: bit J of negmask set
// means pivot J is negative).  Column J: l_JJ = sqrt(s_J p), L(i,J) = s_J a_iJ / l_JJ, trailing a_ic -= s_J (a_iJ / l_JJ)(a_cJ / l_JJ).
template <int J>
__device__ __forceinline__ void fstep_signed(double (&a)[SB], const int lane, double &dl, const unsigned negmask, int &badcol) {
  const double sJ = ((negmask >> J) & 1u) ? -1.0 : 1.0;
  const double pjj = sJ * __shfl_sync(FULL, a[J], J);
  double ljj, inv;
  sqrt_and_rcp(pjj, ljj, inv);
  const double qi = a[J] * inv;
  const double mq = -sJ * qi;
#pragma unroll
  for (int c = J + 1; c < SB; c++) {
    const double qc = __shfl_sync(FULL, a[J], c) * inv;
    a[c] = fma(mq, qc, a[c]);
  }
  if (!(pjj > 0.0) && badcol < 0) badcol = J;
  a[J] = sJ * qi;
  if (lane == J) dl = ljj;
  if constexpr (J + 1 < SB) fstep_signed<J + 1>(a, lane, dl, negmask, badcol);
}

}  // namespace chol32
}  // namespace qb
