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
  ncclResult_t (*CommGetAsyncError)(ncclComm_t, ncclResult_t*) = nullptr;   // optional
  ncclResult_t (*CommAbort)(ncclComm_t) = nullptr;                          // optional
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
  n.CommGetAsyncError = (decltype(n.CommGetAsyncError))dlsym(h, "ncclCommGetAsyncError");
  n.CommAbort = (decltype(n.CommAbort))dlsym(h, "ncclCommAbort");
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

// Asynchronous failures (a peer died, a link went down) never show up in the return value of ncclAllReduce; they are
// polled here.  On an error the communicator is aborted - otherwise the queued collective, and every stream that
// waits for it, hangs forever - and the caller gets MT_ERR_NCCL.
static int poll_async_error(const char* where) {
  Nccl& n = N();
  if (!n.comm || !n.CommGetAsyncError) return MT_OK;
  ncclResult_t st = ncclSuccess;
  if (n.CommGetAsyncError(n.comm, &st) != ncclSuccess) return MT_OK;
  enum { ncclInProgress = 7 };
  if (st == ncclSuccess || st == ncclInProgress) return MT_OK;
  if (n.CommAbort) n.CommAbort(n.comm);
  n.comm = nullptr;
  return fail(MT_ERR_NCCL, "%s: asynchronous NCCL error: %s (communicator aborted)", where,
              n.GetErrorString ? n.GetErrorString(st) : "?");
}

MT_API int mt_comm_check(void) { return poll_async_error("mt_comm_check"); }

// the communication stream may start on `buf` once everything that writes it has run on the compute stream
static int order_comm_after_producer(const float* buf) {
  Global& G = g();
  if (buf && buf == G.dw_ready_ptr) {
    // the gradient mt_linear_bwd_f32 just produced: its dW GEMM (+ split-K fold) is marked by ev_dw_ready, so the
    // exchange does not wait for the dX GEMM queued behind it on the compute stream
    MT_CUDA(cudaStreamWaitEvent(G.comm_stream, G.ev_dw_ready, 0));
    G.dw_ready_ptr = nullptr;   // one-shot (the pool also forgets it when the block is freed)
  } else {
    MT_CUDA(cudaEventRecord(G.ev_compute, G.stream));
    MT_CUDA(cudaStreamWaitEvent(G.comm_stream, G.ev_compute, 0));
  }
  return MT_OK;
}

MT_API int mt_allreduce_sum_f32(float* buf, size_t n) {
  MT_REQUIRE_INIT();
  if (!N().comm) return fail(MT_ERR_NCCL, "mt_allreduce_sum_f32: call mt_comm_init_rank first");
  if (!n) return MT_OK;
  Global& G = g();
  int rc = order_comm_after_producer(buf);
  if (rc) return rc;
  if (G.trace_on) trace_mark("allreduce_begin", G.comm_stream, 1);
  pool_mark_foreign(buf, 2);
  MT_NCCL(N().AllReduce(buf, buf, n, ncclFloat32, ncclSum, N().comm, G.comm_stream));
  if (G.trace_on) trace_mark("allreduce_end", G.comm_stream, 1);
  G.launches++;
  return MT_OK;
}

namespace {
// p -= lr * g on the communication stream, right behind that gradient's all-reduce (same two roundings as
// sgd_multi_kernel, so the per-parameter and the fused update are bit-identical)
__global__ void __launch_bounds__(256) sgd_one_kernel(float* __restrict__ p, const float* __restrict__ g_, size_t n, float lr,
                                                      int vec) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t n4 = vec ? n / 4 : 0;
  for (size_t i = tid; i < n4; i += stride) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = mtd::ld_stream4(g_ + 4 * i);
    pv.x = __fsub_rn(pv.x, __fmul_rn(lr, gv.x));
    pv.y = __fsub_rn(pv.y, __fmul_rn(lr, gv.y));
    pv.z = __fsub_rn(pv.z, __fmul_rn(lr, gv.z));
    pv.w = __fsub_rn(pv.w, __fmul_rn(lr, gv.w));
    reinterpret_cast<float4*>(p)[i] = pv;
  }
  for (size_t j = n4 * 4 + tid; j < n; j += stride) p[j] = __fsub_rn(p[j], __fmul_rn(lr, g_[j]));
}
}  // namespace

// All-reduce `grad` over the ranks and apply  param -= lr * grad  on the communication stream as soon as the sum
// has arrived - while the compute stream is still busy with the backward GEMMs of the layers below.  The update is
// ordered after everything queued on the compute stream at the time of this call (the dX GEMM of the same layer
// reads `param`); the next reader of `param` on the compute stream must come after mt_comm_wait().
// Works without a communicator too (world 1): then it is just the deferred update.
MT_API int mt_allreduce_sgd_f32(float* param, float* grad, size_t n, float lr) {
  MT_REQUIRE_INIT();
  MT_CHECK_ARG(param && grad, "mt_allreduce_sgd_f32: null pointer");
  MT_CHECK_ARG((((uintptr_t)param | (uintptr_t)grad) & 3) == 0, "mt_allreduce_sgd_f32: pointers must be float aligned");
  if (!n) return MT_OK;
  Global& G = g();
  if (N().comm && N().world > 1) {
    int rc = order_comm_after_producer(grad);
    if (rc) return rc;
    if (G.trace_on) trace_mark("allreduce_begin", G.comm_stream, 1);
    MT_NCCL(N().AllReduce(grad, grad, n, ncclFloat32, ncclSum, N().comm, G.comm_stream));
    if (G.trace_on) trace_mark("allreduce_end", G.comm_stream, 1);
    G.launches++;
  }
  MT_CUDA(cudaEventRecord(G.ev_compute, G.stream));
  MT_CUDA(cudaStreamWaitEvent(G.comm_stream, G.ev_compute, 0));
  const int vec = (((uintptr_t)param | (uintptr_t)grad) & 15) == 0;
  pool_mark_foreign(param, 2);
  pool_mark_foreign(grad, 2);
  sgd_one_kernel<<<ew_grid((int64_t)(n / 4 + 1), 256 * 4), 256, 0, G.comm_stream>>>(param, grad, n, lr, vec);
  G.launches++;
  if (G.trace_on) trace_mark("sgd_one", G.comm_stream, 1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MT_ERR_CUDA, "sgd_one_kernel launch -> %s", cudaGetErrorString(e));
  return MT_OK;
}

MT_API int mt_comm_wait(void) {
  MT_REQUIRE_INIT();
  Global& G = g();
  int rc = poll_async_error("mt_comm_wait");
  if (rc) return rc;
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
