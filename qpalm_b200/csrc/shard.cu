// shard.cu -- row-sharding of one large dense QP over the GPUs of a box (BASELINE config 3).
//
// One process per GPU.  The constraint rows of A are split into contiguous blocks, one per rank; everything of length n or
// m (iterates, multipliers, sigma, bounds, line-search arrays) and Q / H / L stay replicated, so every rank executes the
// same kernels and the same host control flow on bit-identical data.  Only the three operations that touch A communicate:
//   A d        local rows            -> ncclAllGather of the m-vector            (every inner iteration)
//   A' yh      local partial (n)     -> ncclAllReduce(sum)                       (every iteration)
//   A_J' S A_J local SYRK partial    -> ncclAllReduce(sum) of the n x n record   (every refactorisation)
// plus max-allreduce / allgather of the Ruiz norms at setup.  NCCL is loaded lazily with dlopen (no link-time dependency:
// single-GPU users never need it); the unique id is created on rank 0 and distributed by the caller (torch.distributed,
// MPI, a file ...).
#include "../../include/qpalm_b200.h"
#include "common.cuh"
#include <dlfcn.h>
#include <string.h>

namespace qb {

namespace {
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { ncclSum = 0, ncclProd = 1, ncclMax = 2, ncclMin = 3 };
enum { ncclFloat64 = 8 };
typedef int (*fn_GetUniqueId)(ncclUniqueId *);
typedef int (*fn_CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
typedef int (*fn_CommDestroy)(ncclComm_t);
typedef int (*fn_AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*fn_AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
typedef const char *(*fn_GetErrorString)(int);

struct Nccl {
  void *lib = nullptr;
  fn_GetUniqueId GetUniqueId = nullptr; fn_CommInitRank CommInitRank = nullptr; fn_CommDestroy CommDestroy = nullptr;
  fn_AllReduce AllReduce = nullptr; fn_AllGather AllGather = nullptr; fn_GetErrorString GetErrorString = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
} g;

int load_nccl() {
  if (g.lib) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) { g.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (g.lib) break; }
  if (!g.lib) { fprintf(stderr, "[qpalm_b200] shard: cannot dlopen libnccl.so.2 (%s)\n", dlerror()); return 1; }
  g.GetUniqueId = (fn_GetUniqueId)dlsym(g.lib, "ncclGetUniqueId");
  g.CommInitRank = (fn_CommInitRank)dlsym(g.lib, "ncclCommInitRank");
  g.CommDestroy = (fn_CommDestroy)dlsym(g.lib, "ncclCommDestroy");
  g.AllReduce = (fn_AllReduce)dlsym(g.lib, "ncclAllReduce");
  g.AllGather = (fn_AllGather)dlsym(g.lib, "ncclAllGather");
  g.GetErrorString = (fn_GetErrorString)dlsym(g.lib, "ncclGetErrorString");
  if (!g.GetUniqueId || !g.CommInitRank || !g.AllReduce || !g.AllGather) { fprintf(stderr, "[qpalm_b200] shard: NCCL symbols missing\n"); return 1; }
  return 0;
}
int check(int rc, const char *what) {
  if (rc) fprintf(stderr, "[qpalm_b200] shard: %s failed: %s\n", what, g.GetErrorString ? g.GetErrorString(rc) : "?");
  return rc;
}
}  // namespace

int shard_world() { return g.comm ? g.world : 1; }
int shard_rank() { return g.comm ? g.rank : 0; }

int shard_allreduce(const double *send, double *recv, size_t count, bool max_op, cudaStream_t s) {
  if (!g.comm) return 0;
  return check(g.AllReduce(send, recv, count, ncclFloat64, max_op ? ncclMax : ncclSum, g.comm, s), "ncclAllReduce");
}
// in place: rank r's block sits at buf + r * count_per_rank
int shard_allgather(double *buf, size_t count_per_rank, cudaStream_t s) {
  if (!g.comm) return 0;
  return check(g.AllGather(buf + (size_t)g.rank * count_per_rank, buf, count_per_rank, ncclFloat64, g.comm, s), "ncclAllGather");
}

}  // namespace qb

extern "C" int qpalm_b200_shard_unique_id(char *out128) {
  using namespace qb;
  if (load_nccl()) return 1;
  ncclUniqueId id;
  if (check(g.GetUniqueId(&id), "ncclGetUniqueId")) return 2;
  memcpy(out128, id.internal, 128);
  return 0;
}

extern "C" int qpalm_b200_shard_init(int rank, int world, const char *id128) {
  using namespace qb;
  if (world <= 1) return 0;
  if (load_nccl()) return 1;
  if (g.comm) return 0;
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  if (check(g.CommInitRank(&g.comm, world, id, rank), "ncclCommInitRank")) { g.comm = nullptr; return 2; }
  g.rank = rank; g.world = world;
  return 0;
}

extern "C" void qpalm_b200_shard_finalize(void) {
  using namespace qb;
  if (g.comm && g.CommDestroy) g.CommDestroy(g.comm);
  g.comm = nullptr; g.rank = 0; g.world = 1;
}
