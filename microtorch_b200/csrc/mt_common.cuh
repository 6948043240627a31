// mt_common.cuh - shared state, error plumbing and device helpers of libmicrotorch_b200.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/microtorch_b200.h"

#define MT_API extern "C" __attribute__((visibility("default")))

namespace mt {

struct Global {
  bool ready = false;
  int device = -1;
  int sm_count = 0;
  int cc_major = 0, cc_minor = 0;
  size_t hbm_bytes = 0;
  size_t l2_bytes = 0;
  cudaStream_t stream = nullptr;       // compute stream
  cudaStream_t comm_stream = nullptr;  // collective stream
  cudaStream_t copy_stream = nullptr;  // host->device prefetch stream
  cudaEvent_t ev_copy = nullptr, ev_copy_gate = nullptr;
  cudaEvent_t ev_compute = nullptr, ev_comm = nullptr;
  uint64_t launches = 0;
  int gemm_engine_force = -1;
  int gemm_last_engine = -1;
  uint64_t gemm_mid_launches = 0;      // launches of the mid-size tcgen05 kernel (mt_gemm_tc_mid_launches)
  void* l2_flush_buf = nullptr;
  size_t l2_flush_bytes = 0;
  bool trace_on = false;               // mt_trace_*: one timed event per launch (a device-side timeline)
  // the dW of the last mt_linear_bwd_f32: an all-reduce of exactly this buffer may start as soon as dW is done,
  // without waiting for the dX GEMM queued behind it
  const void* dw_ready_ptr = nullptr;
  cudaEvent_t ev_dw_ready = nullptr;
};
Global& g();

// thread-local last-error message
void set_error(const char* fmt, ...);
int fail(int status, const char* fmt, ...);

#define MT_REQUIRE_INIT()                                                            \
  do {                                                                               \
    if (!mt::g().ready) return mt::fail(MT_ERR_NOT_INIT, "%s: mt_init() has not succeeded", __func__); \
  } while (0)

#define MT_CUDA(call)                                                                \
  do {                                                                               \
    cudaError_t e__ = (call);                                                        \
    if (e__ != cudaSuccess)                                                          \
      return mt::fail(MT_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call,     \
                      cudaGetErrorString(e__));                                      \
  } while (0)

#define MT_CHECK_ARG(cond, ...)                                                      \
  do {                                                                               \
    if (!(cond)) return mt::fail(MT_ERR_INVALID, __VA_ARGS__);                       \
  } while (0)

// timeline tracing (mt_trace_*): a timed event on `stream` labelled `name` (a string literal)
void trace_mark(const char* name, cudaStream_t stream, int stream_id);

// to be called right after a <<<>>> launch
#define MT_LAUNCHED()                                                                \
  do {                                                                               \
    mt::g().launches++;                                                              \
    if (mt::g().trace_on) mt::trace_mark(__func__, mt::g().stream, 0);               \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess)                                                          \
      return mt::fail(MT_ERR_CUDA, "%s:%d kernel launch -> %s", __FILE__, __LINE__,  \
                      cudaGetErrorString(e__));                                      \
  } while (0)

// pool (mt_pool.cu)
int pool_alloc(void** p, size_t bytes);
int pool_free(void* p);
void pool_mark_foreign(const void* p, int stream_bit);   // 1: copy stream, 2: communication stream

// Scratch block from the pool, returned on every exit path (pool_free is stream-ordered, so handing the block back
// right after the kernels that use it were queued is safe).
struct PoolScratch {
  void* ptr = nullptr;
  PoolScratch() = default;
  PoolScratch(const PoolScratch&) = delete;
  PoolScratch& operator=(const PoolScratch&) = delete;
  ~PoolScratch() { release(); }
  int alloc(size_t bytes) { release(); return pool_alloc(&ptr, bytes); }
  void release() { if (ptr) { pool_free(ptr); ptr = nullptr; } }
  float* f32() const { return static_cast<float*>(ptr); }
};

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// grid for a grid-stride elementwise kernel over `work` items with `per_block` items per block
inline int ew_grid(int64_t work, int64_t per_block) {
  int64_t blocks = ceil_div(work, per_block);
  int64_t cap = (int64_t)g().sm_count * 16;  // multiple of the SM count, 16 resident 256-thread CTAs/SM max
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

}  // namespace mt

// ------------------------------------------------------------------ device helpers
namespace mtd {

__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
// streaming 128-bit load that does not allocate in L1 (data touched once)
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum: warp shuffle, then a shared-memory tree over the warp leaders.
// Result valid in thread 0 (and broadcast to all when `bcast`).
template <int THREADS>
__device__ __forceinline__ float block_sum(float v, float* smem /* THREADS/32 floats */) {
  constexpr int W = THREADS / 32;
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < W ? smem[lane] : 0.f;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    if (lane == 0) smem[0] = t;
  }
  __syncthreads();
  float r = smem[0];
  __syncthreads();
  return r;
}

}  // namespace mtd
