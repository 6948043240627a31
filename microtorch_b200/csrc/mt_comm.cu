// mt_comm.cu - data-parallel gradient exchange: NCCL all-reduce over NVLink 5 / NVSwitch.
//
// New component (the reference has no distributed code): one process
// per GPU, weight gradients summed across ranks on a dedicated communication stream so the
// exchange of layer L overlaps the backward GEMMs of layers < L.  NCCL is dlopen()ed so the
// library loads on a box without it; every entry point then fails with MT_ERR_NCCL.
#include <dlfcn.h>

#include "mt_common.cuh"

using namespace mt;

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclSuccess = 0 };
enum { ncclFloat32 = 7 };
enum { ncclSum = 0 };

struct Nccl {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};
Nccl& N() {
  static Nccl n;
  return n;
}

int need_lib() {
  if (!N().handle) {
    int rc = mt_comm_load(nullptr);
    if (rc) return rc;
  }
  return MT_OK;
}

}  // namespace

#define MT_NCCL(call)                                                                                   \
  do {                                                                                                  \
    ncclResult_t r__ = (call);                                                                          \
    if (r__ != ncclSuccess)                                                                             \
      return fail(MT_ERR_NCCL, "%s -> %s", #call, N().GetErrorString ? N().GetErrorString(r__) : "?"); \
  } while (0)

MT_API int mt_comm_load(const char* path) {
  Nccl& n = N();
  if (n.handle) return MT_OK;
  const char* lib = (path && *path) ? path : "libnccl.so.2";
  void* h = dlopen(lib, RTLD_NOW | RTLD_GLOBAL);
  if (!h) return fail(MT_ERR_NCCL, "mt_comm_load: dlopen(%s) failed: %s", lib, dlerror());
  n.GetUniqueId = (decltype(n.GetUniqueId))dlsym(h, "ncclGetUniqueId");
  n.CommInitRank = (decltype(n.CommInitRank))dlsym(h, "ncclCommInitRank");
  n.AllReduce = (decltype(n.AllReduce))dlsym(h, "ncclAllReduce");
  n.CommDestroy = (decltype(n.CommDestroy))dlsym(h, "ncclCommDestroy");
  n.GetErrorString = (decltype(n.GetErrorString))dlsym(h, "ncclGetErrorString");
  if (!n.GetUniqueId || !n.CommInitRank || !n.AllReduce || !n.CommDestroy) {
    dlclose(h);
    return fail(MT_ERR_NCCL, "mt_comm_load: %s lacks the NCCL entry points", lib);
  }
  n.handle = h;
  return MT_OK;
}

MT_API int mt_comm_unique_id(char* out128) {
  MT_CHECK_ARG(out128, "mt_comm_unique_id: null out pointer");
  int rc = need_lib();
  if (rc) return rc;
  ncclUniqueId id;
  MT_NCCL(N().GetUniqueId(&id));
  memcpy(out128, id.internal, 128);
  return MT_OK;
}

MT_API int mt_comm_init_rank(const char* id128, int rank, int world) {
  MT_REQUIRE_INIT();
  MT_CHECK_ARG(id128 && world >= 1 && rank >= 0 && rank < world, "mt_comm_init_rank: bad rank/world %d/%d", rank, world);
  int rc = need_lib();
  if (rc) return rc;
  if (N().comm) return fail(MT_ERR_INVALID, "mt_comm_init_rank: communicator already initialised");
  ncclUniqueId id;
  memcpy(id.internal, id128, 128);
  MT_NCCL(N().CommInitRank(&N().comm, world, id, rank));
  N().rank = rank;
  N().world = world;
  return MT_OK;
}

MT_API int mt_comm_world(int* rank, int* world) {
  if (rank) *rank = N().rank;
  if (world) *world = N().world;
  return MT_OK;
}

MT_API int mt_allreduce_sum_f32(float* buf, size_t n) {
  MT_REQUIRE_INIT();
  if (!N().comm) return fail(MT_ERR_NCCL, "mt_allreduce_sum_f32: call mt_comm_init_rank first");
  if (!n) return MT_OK;
  Global& G = g();
  // order the collective after everything already queued on the compute stream
  MT_CUDA(cudaEventRecord(G.ev_compute, G.stream));
  MT_CUDA(cudaStreamWaitEvent(G.comm_stream, G.ev_compute, 0));
  MT_NCCL(N().AllReduce(buf, buf, n, ncclFloat32, ncclSum, N().comm, G.comm_stream));
  G.launches++;
  return MT_OK;
}

MT_API int mt_comm_wait(void) {
  MT_REQUIRE_INIT();
  Global& G = g();
  MT_CUDA(cudaEventRecord(G.ev_comm, G.comm_stream));
  MT_CUDA(cudaStreamWaitEvent(G.stream, G.ev_comm, 0));
  return MT_OK;
}

MT_API int mt_comm_destroy(void) {
  Nccl& n = N();
  if (n.comm) {
    if (g().ready) cudaStreamSynchronize(g().comm_stream);
    n.CommDestroy(n.comm);
    n.comm = nullptr;
  }
  n.rank = 0;
  n.world = 1;
  return MT_OK;
}
