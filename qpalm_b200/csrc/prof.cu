// prof.cu -- selective per-kernel CUDA-event timing behind QB_LAUNCH (common.cuh).
// qpalm_b200_prof_enable("k_diag_block,k_dgemm_nt") times every launch whose kernel name contains one of the
// comma-separated patterns ("*" = all) with an event pair on the launching stream; qpalm_b200_prof_report
// synchronises, aggregates per kernel name and writes one JSON object.  Used by bench.py for the dominant
// kernel's average launch duration inside the timed region, and by tools/ for whole-step breakdowns.
#include "../../include/qpalm_b200.h"
#include "common.cuh"
#include <map>
#include <string>
#include <vector>
#include <string.h>

namespace qb {

int g_prof_on = 0;

namespace {
struct Rec { const char *name; cudaEvent_t e0, e1; };
std::vector<std::string> g_patterns;
std::vector<Rec> g_recs;
std::vector<cudaEvent_t> g_pool;
cudaEvent_t g_cur0 = nullptr, g_cur1 = nullptr;
const char *g_cur_name = nullptr;

cudaEvent_t get_event() {
  if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

bool prof_begin(const char *name, cudaStream_t s) {
  bool hit = false;
  for (const std::string &p : g_patterns)
    if (p == "*" || strstr(name, p.c_str())) { hit = true; break; }
  if (!hit) return false;
  g_cur0 = get_event(); g_cur1 = get_event(); g_cur_name = name;
  cudaEventRecord(g_cur0, s);
  return true;
}

void prof_end(cudaStream_t s) {
  cudaEventRecord(g_cur1, s);
  g_recs.push_back({g_cur_name, g_cur0, g_cur1});
}

}  // namespace qb

extern "C" int qpalm_b200_prof_enable(const char *patterns) {
  using namespace qb;
  g_patterns.clear();
  if (patterns && *patterns) {
    std::string all(patterns);
    size_t pos = 0;
    while (pos <= all.size()) {
      size_t c = all.find(',', pos);
      if (c == std::string::npos) c = all.size();
      if (c > pos) g_patterns.push_back(all.substr(pos, c - pos));
      pos = c + 1;
    }
  }
  g_prof_on = g_patterns.empty() ? 0 : 1;
  return 0;
}

extern "C" int qpalm_b200_prof_report(char *buf, size_t buflen) {
  using namespace qb;
  cudaDeviceSynchronize();
  std::map<std::string, std::pair<long long, double>> agg;
  for (const Rec &r : g_recs) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) { auto &a = agg[r.name]; a.first++; a.second += ms; }
    g_pool.push_back(r.e0); g_pool.push_back(r.e1);
  }
  g_recs.clear();
  std::string out = "{";
  bool first = true;
  for (auto &kv : agg) {
    char tmp[256];
    snprintf(tmp, sizeof tmp, "%s\"%s\": {\"launches\": %lld, \"ms\": %.6f}", first ? "" : ", ", kv.first.c_str(), kv.second.first, kv.second.second);
    out += tmp; first = false;
  }
  out += "}";
  if (!buf || buflen == 0) return (int)out.size() + 1;
  strncpy(buf, out.c_str(), buflen - 1);
  buf[buflen - 1] = 0;
  return 0;
}

// ---- latency micro-benchmarks (tools/microbench.py): SM clocks per dependent operation on one warp ----
namespace {
__global__ void k_lat(double *out, long long *clk, double x0) {
  double x = x0 + threadIdx.x * 1e-9, y = 1.000000001;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; i++) { x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); x = fma(x, y, 1e-9); }
  long long t1 = clock64();
  double s = x;
#pragma unroll 1
  for (int i = 0; i < 256; i++) { s = sqrt(s + 2.0); }
  long long t2 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; i++) { s = 1.0 / (s + 2.0); }
  long long t3 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; i++) { s = __shfl_sync(0xffffffffu, s, (i + 1) & 31) + 1.0; }
  long long t4 = clock64();
  double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
  long long t5 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; i++) {
    a0 = fma(a0, y, 1e-9); a1 = fma(a1, y, 1e-9); a2 = fma(a2, y, 1e-9); a3 = fma(a3, y, 1e-9);
    a4 = fma(a4, y, 1e-9); a5 = fma(a5, y, 1e-9); a6 = fma(a6, y, 1e-9); a7 = fma(a7, y, 1e-9);
  }
  long long t6 = clock64();
  out[threadIdx.x + blockIdx.x * blockDim.x] = s + a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    clk[0] = (t1 - t0) / 1024; clk[1] = (t2 - t1) / 256; clk[2] = (t3 - t2) / 256; clk[3] = (t4 - t3) / 256; clk[4] = (t6 - t5);
  }
}
}  // namespace

// out[0] dependent DFMA latency, [1] sqrt (+add) latency, [2] reciprocal (+add), [3] shfl.f64 (+add), [4] clocks for
// 256 x 8 independent DFMA per thread with `threads` threads per CTA on `ctas` CTAs (throughput probe)
extern "C" int qpalm_b200_microbench(int ctas, int threads, long long *out5) {
  double *d = nullptr; long long *c = nullptr;
  QB_CUDA_TRY(cudaMalloc(&d, sizeof(double) * (size_t)ctas * threads));
  QB_CUDA_TRY(cudaMalloc(&c, sizeof(long long) * 8));
  k_lat<<<ctas, threads>>>(d, c, 1.0);
  QB_CUDA_TRY(cudaDeviceSynchronize());
  k_lat<<<ctas, threads>>>(d, c, 1.0);
  QB_CUDA_TRY(cudaDeviceSynchronize());
  QB_CUDA_TRY(cudaMemcpy(out5, c, sizeof(long long) * 5, cudaMemcpyDeviceToHost));
  cudaFree(d); cudaFree(c);
  return 0;
}

// ---- micro-benchmark of the 16 x 16 in-warp Cholesky variants (SM clocks per block) ----
namespace {
constexpr int MB_LDP = 241, MB_SW = 16;
__device__ __forceinline__ void mb_diag_rolled(double *Pn, double *rd, int c0, int w) {
  const int lane = threadIdx.x & 31, row = c0 + lane;
  const int wend = (w - c0 < MB_SW) ? w - c0 : MB_SW;
  const bool live = lane < wend;
#pragma unroll 1
  for (int j = 0; j < wend; j++) {
    const double pjj = Pn[(c0 + j) * MB_LDP + c0 + j];
    __syncwarp();
    const double inv = rsqrt(pjj), ljj = pjj * inv;
    double l = 0.0;
    if (live && lane > j) { l = Pn[(c0 + j) * MB_LDP + row] * inv; Pn[(c0 + j) * MB_LDP + row] = l; }
    else if (lane == j) { Pn[(c0 + j) * MB_LDP + row] = ljj; rd[c0 + j] = inv; }
    __syncwarp();
    double tv[MB_SW], mv[MB_SW];
#pragma unroll
    for (int c = 1; c < MB_SW; c++) {
      const bool on = c > j && c < wend && lane >= c && live;
      tv[c] = on ? Pn[(c0 + c) * MB_LDP + row] : 0.0;
      mv[c] = on ? Pn[(c0 + j) * MB_LDP + c0 + c] : 0.0;
    }
#pragma unroll
    for (int c = 1; c < MB_SW; c++)
      if (c > j && c < wend && lane >= c && live) Pn[(c0 + c) * MB_LDP + row] = fma(-l, mv[c], tv[c]);
    __syncwarp();
  }
}
// column-oriented variant: lane = column c; each lane keeps its column's pending diagonal/sub-diagonal updates in
// registers?  (not possible without static indexing) -- instead: row-per-lane registers, j loop unrolled
__device__ __forceinline__ void mb_diag_regs(double *Pn, double *rd, int c0) {
  const int lane = threadIdx.x & 31, row = c0 + lane;
  double a[MB_SW];
#pragma unroll
  for (int c = 0; c < MB_SW; c++) a[c] = (lane < MB_SW) ? ((c <= lane) ? Pn[(c0 + c) * MB_LDP + row] : 0.0) : ((c == lane) ? 1.0 : 0.0);
#pragma unroll
  for (int j = 0; j < MB_SW; j++) {
    const double pjj = __shfl_sync(0xffffffffu, a[j], j);
    const double inv = rsqrt(pjj);
    if (lane == j) { a[j] = pjj * inv; rd[c0 + j] = inv; }
    else if (lane > j) a[j] *= inv;
#pragma unroll
    for (int c = 0; c < MB_SW; c++)
      if (c > j) { const double lcj = __shfl_sync(0xffffffffu, a[j], c); if (lane >= c) a[c] = fma(-a[j], lcj, a[c]); }
  }
#pragma unroll
  for (int c = 0; c < MB_SW; c++) if (lane < MB_SW && c <= lane) Pn[(c0 + c) * MB_LDP + row] = a[c];
}
__global__ void k_mb_diag(long long *clk, double *sink, int reps) {
  __shared__ double Pn[16 * MB_LDP];
  __shared__ double rd[64];
  const int lane = threadIdx.x;
  for (int v = 0; v < 2; v++) {
    long long total = 0;
    for (int r = 0; r < reps; r++) {
      for (int c = 0; c < 16; c++) Pn[c * MB_LDP + lane] = (c == lane) ? 40.0 + lane : 1.0 / (1 + c + lane);
      __syncwarp();
      const long long t0 = clock64();
      if (v == 0) mb_diag_rolled(Pn, rd, 0, 16); else mb_diag_regs(Pn, rd, 0);
      __syncwarp();
      total += clock64() - t0;
    }
    if (lane == 0) clk[v] = total / reps;
    sink[lane + 32 * v] = Pn[(lane & 15) * MB_LDP + lane] + rd[lane & 15];
  }
}
}  // namespace
extern "C" int qpalm_b200_microbench_diag(long long *out2) {
  long long *c = nullptr; double *d = nullptr;
  QB_CUDA_TRY(cudaMalloc(&c, 64)); QB_CUDA_TRY(cudaMalloc(&d, 8 * 64));
  k_mb_diag<<<1, 32>>>(c, d, 20);
  QB_CUDA_TRY(cudaDeviceSynchronize());
  QB_CUDA_TRY(cudaMemcpy(out2, c, 16, cudaMemcpyDeviceToHost));
  cudaFree(c); cudaFree(d);
  return 0;
}

// phase clocks of the 128 x 128 diagonal-block kernel (dense.cu); out32: clock64 samples at the phase boundaries
namespace qb { int diag_block_phase_clocks(long long *out32); }
extern "C" int qpalm_b200_microbench_diag_phases(long long *out32) { return qb::diag_block_phase_clocks(out32); }
