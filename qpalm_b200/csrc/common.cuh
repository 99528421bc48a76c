// common.cuh -- shared device/host helpers for the sm_100a kernels of libqpalm_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

namespace qb {

// ------------------------------------------------------------------------------------------------
// error handling: the library never falls back to the CPU; CUDA failures surface as error codes
// ------------------------------------------------------------------------------------------------
#define QB_CUDA_TRY(expr)                                                                          \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      fprintf(stderr, "[qpalm_b200] CUDA error %d (%s) at %s:%d: %s\n", (int)_e,                   \
              cudaGetErrorString(_e), __FILE__, __LINE__, #expr);                                  \
      return -(int)_e;                                                                             \
    }                                                                                              \
  } while (0)

#define QB_CUDA_CHECK_VOID(expr)                                                                   \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) {                                                                       \
      fprintf(stderr, "[qpalm_b200] CUDA error %d (%s) at %s:%d: %s\n", (int)_e,                   \
              cudaGetErrorString(_e), __FILE__, __LINE__, #expr);                                  \
    }                                                                                              \
  } while (0)

// every kernel launch of the library goes through this counter (bench.py reports `gpu_launches`).
// Selective per-kernel timing (prof.cu): when enabled for a name pattern, matching launches are bracketed by
// CUDA events on the launching stream; bench.py reads the per-kernel launch count and device time from it
// (the live `roofline.achieved` denominator).  Off by default: one integer test per launch.
extern long long g_kernel_launches;
extern int g_prof_on;
bool prof_begin(const char *name, cudaStream_t s);
void prof_end(cudaStream_t s);
#define QB_LAUNCH(kernel, grid, block, smem, stream, ...)                                          \
  do {                                                                                             \
    const bool _qb_p = qb::g_prof_on && qb::prof_begin(#kernel, (stream));                         \
    kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                    \
    ++qb::g_kernel_launches;                                                                       \
    if (_qb_p) qb::prof_end((stream));                                                             \
  } while (0)

constexpr int kWarp = 32;
constexpr int kPanel = 128;  // dense factor block size; every dense leading dimension is a multiple

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
static inline long long cdivll(long long a, long long b) { return (a + b - 1) / b; }
static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------------------------------------
// device reductions.  All floating-point sums are reduced in a fixed tree order so that a given
// launch configuration is bit-reproducible from run to run (the reference's re-solve test demands
// identical results, tests/src/test_basic_qp.c:298-305).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

enum RedOp { RED_SUM = 0, RED_MAX = 1, RED_MIN = 2 };

template <int OP>
__device__ __forceinline__ double red_combine(double a, double b) {
  if (OP == RED_SUM) return a + b;
  if (OP == RED_MAX) return fmax(a, b);
  return fmin(a, b);
}
template <int OP>
__device__ __forceinline__ double warp_red(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = red_combine<OP>(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide reduction; result valid in thread 0.  `scratch` holds >= 32 doubles of shared memory.
template <int OP>
__device__ __forceinline__ double block_red(double v, double *scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_red<OP>(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double ident = (OP == RED_SUM) ? 0.0 : (OP == RED_MAX ? -1.0e300 : 1.0e300);
    v = (lane < nw) ? scratch[lane] : ident;
    v = warp_red<OP>(v);
  }
  return v;
}

// streaming 128-bit read-only load (bypasses L1 allocation; for data touched once per kernel)
__device__ __forceinline__ double2 ld_stream2(const double *p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}

}  // namespace qb
