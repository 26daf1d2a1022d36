// Gradient exchange over NCCL (NVLink 5 / NVSwitch) behind the C ABI: cpt_nccl_unique_id / _init / _allreduce_sum_f32 / _destroy.
// The reference has no distributed code; this is the one exchange step of batch-sharded data parallelism, the
// SUM all-reduce of the flat gradient arena before the optimizer update (hook points: compyute/nn/modules/module.py:392-400,
// nn/optimizers.py:152,241).
//
// libnccl is bound at RUN time (dlopen of the libnccl.so.2 the process already has — the one torch ships — or the system's):
// the library itself has no link-time NCCL dependency and single-GPU users never load it.  Only the five entry points below
// are used; their prototypes are restated from nccl.h (NCCL 2.x ABI: ncclFloat32 = 7, ncclSum = 0, 128-byte unique id).
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"

namespace cpt {
namespace {

typedef struct ncclComm* ncclComm_t;
struct ncclUniqueId { char internal[128]; };
typedef int (*GetUniqueIdFn)(ncclUniqueId*);
typedef int (*CommInitRankFn)(ncclComm_t*, int, ncclUniqueId, int);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
typedef int (*CommDestroyFn)(ncclComm_t);
typedef const char* (*GetErrorStringFn)(int);

struct Nccl {
  void* handle = nullptr;
  GetUniqueIdFn get_unique_id = nullptr;
  CommInitRankFn comm_init_rank = nullptr;
  AllReduceFn all_reduce = nullptr;
  CommDestroyFn comm_destroy = nullptr;
  GetErrorStringFn error_string = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
} g;

int load() {
  if (g.handle) return CPT_OK;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  CPT_REQUIRE(h, CPT_ERR_UNSUPPORTED, "NCCL: libnccl.so.2 not found (%s)", dlerror());
  g.get_unique_id = reinterpret_cast<GetUniqueIdFn>(dlsym(h, "ncclGetUniqueId"));
  g.comm_init_rank = reinterpret_cast<CommInitRankFn>(dlsym(h, "ncclCommInitRank"));
  g.all_reduce = reinterpret_cast<AllReduceFn>(dlsym(h, "ncclAllReduce"));
  g.comm_destroy = reinterpret_cast<CommDestroyFn>(dlsym(h, "ncclCommDestroy"));
  g.error_string = reinterpret_cast<GetErrorStringFn>(dlsym(h, "ncclGetErrorString"));
  CPT_REQUIRE(g.get_unique_id && g.comm_init_rank && g.all_reduce && g.comm_destroy && g.error_string, CPT_ERR_UNSUPPORTED,
              "NCCL: libnccl lacks an expected entry point");
  g.handle = h;
  return CPT_OK;
}

int nccl_fail(int rc, const char* what) {
  set_error("NCCL: %s failed: %s", what, g.error_string ? g.error_string(rc) : "?");
  return CPT_ERR_CUDA;
}

}  // namespace
}  // namespace cpt

using namespace cpt;

extern "C" {

int cpt_nccl_unique_id(void* out128) {
  CPT_REQUIRE(out128, CPT_ERR_INVALID, "nccl_unique_id: null buffer");
  if (int e = load()) return e;
  ncclUniqueId id;
  if (int rc = g.get_unique_id(&id)) return nccl_fail(rc, "ncclGetUniqueId");
  memcpy(out128, &id, sizeof(id));
  return CPT_OK;
}

int cpt_nccl_init(int rank, int world, const void* unique_id128) {
  CPT_REQUIRE(unique_id128 && world >= 1 && rank >= 0 && rank < world, CPT_ERR_INVALID, "nccl_init: bad arguments");
  CPT_REQUIRE(!g.comm, CPT_ERR_INVALID, "nccl_init: communicator already initialised (call cpt_nccl_destroy first)");
  if (int e = load()) return e;
  ncclUniqueId id;
  memcpy(&id, unique_id128, sizeof(id));
  if (int rc = g.comm_init_rank(&g.comm, world, id, rank)) { g.comm = nullptr; return nccl_fail(rc, "ncclCommInitRank"); }
  g.rank = rank; g.world = world;
  return CPT_OK;
}

int cpt_nccl_world_size(void) { return g.comm ? g.world : 0; }

/* in place: ptr[i] <- sum over ranks of ptr[i]; asynchronous on `stream` like every other entry point */
int cpt_nccl_allreduce_sum_f32(float* ptr, int64_t count, void* stream) {
  CPT_REQUIRE(g.comm, CPT_ERR_INVALID, "nccl_allreduce_sum_f32: communicator not initialised");
  CPT_REQUIRE(ptr && count >= 0, CPT_ERR_INVALID, "nccl_allreduce_sum_f32: bad arguments");
  if (count == 0) return CPT_OK;
  if (int rc = g.all_reduce(ptr, ptr, (size_t)count, /*ncclFloat32*/ 7, /*ncclSum*/ 0, g.comm, as_stream(stream)))
    return nccl_fail(rc, "ncclAllReduce");
  return CPT_OK;  // NCCL's kernels, not ours: not counted in cpt_launch_count
}

int cpt_nccl_destroy(void) {
  if (!g.comm) return CPT_OK;
  const int rc = g.comm_destroy(g.comm);
  g.comm = nullptr;
  g.world = 1; g.rank = 0;
  return rc ? nccl_fail(rc, "ncclCommDestroy") : CPT_OK;
}

}  // extern "C"
