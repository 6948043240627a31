// mt_p2p.cu - gradient exchange fused with the SGD update over NVLink peer memory (no NCCL in the data path).
//
// New component (the reference has no distributed code).  One process per GPU; every rank owns an
// ARENA in its HBM that all peers map through CUDA IPC (NVLink 5 / NVSwitch loads and stores):
//
//     [ flags: ready[slot][rank], done[slot][rank], tickets ] [ in: this rank's gradients ] [ out: the summed gradients ]
//
// A weight gradient that has just become final is exchanged and applied by three small kernels on the communication
// stream (the backward GEMMs of the layers below keep the compute stream busy meanwhile):
//
//   publish   grad -> own `in` region; fence; the last CTA raises ready[slot][me] = seq in EVERY peer's arena
//   exchange  waits until ready[slot][*] >= seq in its own arena, then for ITS 1/world slice of the tensor: loads the
//             slice from every rank's `in` (peer loads), adds them IN RANK ORDER - one fixed order on every rank, so the
//             replicas stay bit-identical and the result does not depend on a library's algorithm choice - and stores the
//             sum into every rank's `out` (peer stores): reduce-scatter and all-gather in one pass, 2 (W-1)/W x the
//             gradient bytes over NVLink per rank, the all-reduce minimum; fence; raises done[slot][me] = seq everywhere
//   apply     waits until done[slot][*] >= seq, then  grad <- sum  and  param <- param - lr * sum  (NumPy's two roundings,
//             bit-identical to sgd_multi_kernel) in one pass over the local `out` region
//
// seq grows by one per exchange of a slot, so flags are never reset and a peer that is a step behind can never be
// overtaken: a rank reaches `publish` of step s+1 only after its own `apply` of step s, which waited for every peer's
// `done` of step s.  All waits are bounded (a few seconds, then an error word in host-mapped memory that mt_comm_wait
// reports) - a dead peer cannot hang the GPU.
//
// mt_allreduce_sgd_f32 (mt_comm.cu, NCCL) stays as the fallback for tensors the arena cannot take.
#include "mt_common.cuh"

using namespace mt;

namespace {

constexpr int MAX_WORLD = 8;          // one NVSwitch box
constexpr int MAX_SLOTS = 256;        // distinct (parameter, size) pairs
constexpr int FLAG_WORDS = MAX_SLOTS * MAX_WORLD * 2 + MAX_SLOTS * 2;   // ready, done, two ticket counters per slot
constexpr size_t FLAG_BYTES = ((FLAG_WORDS * sizeof(unsigned) + 4095) / 4096) * 4096;
constexpr int THREADS = 512;
constexpr long long SPIN_NS = 8000000000ll;   // 8 s

struct Peers {
  float* in[MAX_WORLD];
  float* out[MAX_WORLD];
  unsigned* flags[MAX_WORLD];
};

struct Slot {
  unsigned long long key;   // caller's name for the tensor: the same on every rank (pointers are not)
  size_t n;
  size_t off;        // floats, into `in` and `out`
  unsigned seq;
};

struct P2P {
  bool active = false;
  int rank = 0, world = 1;
  void* arena = nullptr;              // own allocation
  size_t arena_bytes = 0, region_floats = 0, used_floats = 0;
  void* peer_base[MAX_WORLD] = {};
  Peers peers{};
  std::vector<Slot> slots;
  unsigned* err_host = nullptr;       // host-mapped error word (0 = fine)
  unsigned* err_dev = nullptr;
};
P2P& S() {
  static P2P s;
  return s;
}

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ long long now_ns() {
  long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ float4 ld_l2(const float* p) {      // L2 only: the line may have been written by a peer
  return __ldcg(reinterpret_cast<const float4*>(p));
}

__device__ __forceinline__ int ready_idx(int slot, int r) { return (slot * MAX_WORLD + r) * 2; }
__device__ __forceinline__ int done_idx(int slot, int r) { return (slot * MAX_WORLD + r) * 2 + 1; }
__device__ __forceinline__ int ticket_idx(int slot, int which) { return MAX_SLOTS * MAX_WORLD * 2 + slot * 2 + which; }

// thread 0 of the CTA waits until every rank's flag (ready or done) of `slot` in MY arena has reached seq
__device__ __forceinline__ void wait_all(const unsigned* my_flags, int slot, int world, unsigned seq, bool done,
                                         unsigned* err, unsigned code) {
  if (threadIdx.x == 0) {
    const long long t0 = now_ns();
    for (int r = 0; r < world; ++r) {
      const unsigned* f = my_flags + (done ? done_idx(slot, r) : ready_idx(slot, r));
      while ((int)(ld_acquire_sys(f) - seq) < 0) {
        if (now_ns() - t0 > SPIN_NS) {
          *err = code | ((unsigned)r << 8) | ((unsigned)slot << 16);
          __threadfence_system();
          break;
        }
        __nanosleep(200);
      }
    }
  }
  __syncthreads();
}

// all CTAs are done with their stores: the last one (ticket) raises this rank's flag in every peer's arena
__device__ __forceinline__ void signal_all(const Peers& peers, unsigned* my_flags, int slot, int which, int me, int world,
                                           unsigned seq, bool done) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned* ticket = my_flags + ticket_idx(slot, which);
    const unsigned arrived = atomicAdd(ticket, 1u);
    if (arrived == gridDim.x - 1) {
      *ticket = 0u;                                       // re-arm (the next launch on this slot is stream-ordered)
      __threadfence_system();
      for (int r = 0; r < world; ++r)
        st_release_sys(peers.flags[r] + (done ? done_idx(slot, me) : ready_idx(slot, me)), seq);
    }
  }
}

__global__ void __launch_bounds__(THREADS) p2p_publish_kernel(Peers peers, const float* __restrict__ grad, size_t off, size_t n4,
                                                              int slot, int me, int world, unsigned seq) {
  float* dst = peers.in[me] + off;
  const size_t stride = (size_t)gridDim.x * THREADS;
  for (size_t i = (size_t)blockIdx.x * THREADS + threadIdx.x; i < n4; i += stride)
    mtd::st4(dst + 4 * i, mtd::ld_stream4(grad + 4 * i));
  signal_all(peers, peers.flags[me], slot, 0, me, world, seq, false);
}

template <int W>
__global__ void __launch_bounds__(THREADS) p2p_exchange_kernel(Peers peers, size_t off, size_t n4, int slot, int me,
                                                               unsigned seq, unsigned* err) {
  wait_all(peers.flags[me], slot, W, seq, false, err, 1u);
  const size_t per = (n4 + W - 1) / W;                    // my slice, in float4 units
  const size_t lo = per * me, hi = lo + per < n4 ? lo + per : n4;
  const size_t stride = (size_t)gridDim.x * THREADS;
  for (size_t i = lo + (size_t)blockIdx.x * THREADS + threadIdx.x; i < hi; i += stride) {
    float4 v[W];
#pragma unroll
    for (int r = 0; r < W; ++r) v[r] = ld_l2(peers.in[r] + off + 4 * i);
    float4 s = v[0];
#pragma unroll
    for (int r = 1; r < W; ++r) {                          // rank order: the same sum, bit for bit, on every rank
      s.x = __fadd_rn(s.x, v[r].x);
      s.y = __fadd_rn(s.y, v[r].y);
      s.z = __fadd_rn(s.z, v[r].z);
      s.w = __fadd_rn(s.w, v[r].w);
    }
#pragma unroll
    for (int r = 0; r < W; ++r) mtd::st4(peers.out[(me + r) % W] + off + 4 * i, s);   // own copy first
  }
  signal_all(peers, peers.flags[me], slot, 1, me, W, seq, true);
}

__global__ void __launch_bounds__(THREADS) p2p_apply_kernel(Peers peers, float* __restrict__ param, float* __restrict__ grad,
                                                            size_t off, size_t n4, int slot, int me, int world,
                                                            unsigned seq, float lr, unsigned* err) {
  wait_all(peers.flags[me], slot, world, seq, true, err, 2u);
  const float* src = peers.out[me] + off;
  const size_t stride = (size_t)gridDim.x * THREADS;
  for (size_t i = (size_t)blockIdx.x * THREADS + threadIdx.x; i < n4; i += stride) {
    const float4 s = ld_l2(src + 4 * i);
    float4 p = *reinterpret_cast<const float4*>(param + 4 * i);
    p.x = __fsub_rn(p.x, __fmul_rn(lr, s.x));              // p + (-(lr * g)): two roundings, as NumPy does it
    p.y = __fsub_rn(p.y, __fmul_rn(lr, s.y));
    p.z = __fsub_rn(p.z, __fmul_rn(lr, s.z));
    p.w = __fsub_rn(p.w, __fmul_rn(lr, s.w));
    mtd::st4(param + 4 * i, p);
    mtd::st4(grad + 4 * i, s);                             // .grad holds the summed gradient, as after an all-reduce
  }
}

int order_after_producer(const float* buf) {               // same rule as mt_comm.cu
  Global& G = g();
  if (buf && buf == G.dw_ready_ptr) {
    MT_CUDA(cudaStreamWaitEvent(G.comm_stream, G.ev_dw_ready, 0));
    G.dw_ready_ptr = nullptr;
  } else {
    MT_CUDA(cudaEventRecord(G.ev_compute, G.stream));
    MT_CUDA(cudaStreamWaitEvent(G.comm_stream, G.ev_compute, 0));
  }
  return MT_OK;
}

}  // namespace

// Allocate this rank's arena and export its IPC handle (64 bytes) for the peers.
MT_API int mt_p2p_create(size_t arena_bytes, char* handle_out64) {
  MT_REQUIRE_INIT();
  MT_CHECK_ARG(handle_out64, "mt_p2p_create: null handle pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  P2P& s = S();
  if (s.arena) return fail(MT_ERR_INVALID, "mt_p2p_create: arena already exists");
  MT_CHECK_ARG(arena_bytes >= (64u << 20), "mt_p2p_create: arena must be at least 64 MiB");
  const size_t region = ((arena_bytes - FLAG_BYTES) / 2) & ~(size_t)4095;
  MT_CUDA(cudaMalloc(&s.arena, FLAG_BYTES + 2 * region));
  MT_CUDA(cudaMemset(s.arena, 0, FLAG_BYTES));
  MT_CUDA(cudaDeviceSynchronize());
  s.arena_bytes = FLAG_BYTES + 2 * region;
  s.region_floats = region / sizeof(float);
  s.used_floats = 0;
  cudaIpcMemHandle_t h;
  MT_CUDA(cudaIpcGetMemHandle(&h, s.arena));
  memcpy(handle_out64, &h, 64);
  if (!s.err_host) {
    MT_CUDA(cudaHostAlloc((void**)&s.err_host, sizeof(unsigned), cudaHostAllocMapped));
    *s.err_host = 0;
    MT_CUDA(cudaHostGetDevicePointer((void**)&s.err_dev, s.err_host, 0));
  }
  return MT_OK;
}

// Map every peer's arena.  `handles` = world x 64 bytes in rank order (this rank's own entry is ignored).
MT_API int mt_p2p_open(int rank, int world, const char* handles) {
  MT_REQUIRE_INIT();
  P2P& s = S();
  MT_CHECK_ARG(s.arena, "mt_p2p_open: call mt_p2p_create first");
  MT_CHECK_ARG(handles && world >= 2 && world <= MAX_WORLD && rank >= 0 && rank < world,
               "mt_p2p_open: bad rank/world %d/%d (peer memory serves 2..%d ranks of one box)", rank, world, MAX_WORLD);
  s.rank = rank;
  s.world = world;
  for (int r = 0; r < world; ++r) {
    void* base = s.arena;
    if (r != rank) {
      cudaIpcMemHandle_t h;
      memcpy(&h, handles + 64 * r, 64);
      MT_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    }
    s.peer_base[r] = base;
    char* b = static_cast<char*>(base);
    s.peers.flags[r] = reinterpret_cast<unsigned*>(b);
    s.peers.in[r] = reinterpret_cast<float*>(b + FLAG_BYTES);
    s.peers.out[r] = reinterpret_cast<float*>(b + FLAG_BYTES + s.region_floats * sizeof(float));
  }
  s.active = true;
  return MT_OK;
}

MT_API int mt_p2p_active(int* active) {
  if (active) *active = S().active ? 1 : 0;
  return MT_OK;
}

// Nonzero once a bounded wait inside an exchange kernel gave up (code | waited-for rank << 8 | slot << 16).
MT_API int mt_p2p_error(unsigned* code) {
  if (code) *code = S().err_host ? *(volatile unsigned*)S().err_host : 0u;
  return MT_OK;
}

// grad <- sum over ranks (rank order), param <- param - lr * sum, on the communication stream; see the file header.
// `key` names the tensor: every rank must pass the same key for the same parameter and call in the same order (slots of
// the arena are assigned on first use of a (key, n) pair; device pointers differ from rank to rank and cannot serve).
// MT_ERR_UNSUPPORTED (nothing queued) when the tensor cannot go through the arena - every rank takes the same decision,
// and the caller uses mt_allreduce_sgd_f32 instead.
MT_API int mt_p2p_allreduce_sgd_f32(float* param, float* grad, size_t n, float lr, unsigned long long key) {
  MT_REQUIRE_INIT();
  P2P& s = S();
  if (!s.active) return fail(MT_ERR_UNSUPPORTED, "mt_p2p_allreduce_sgd_f32: no peer arena");
  MT_CHECK_ARG(param && grad, "mt_p2p_allreduce_sgd_f32: null pointer");
  if (!n) return MT_OK;
  if (n % 4 || (((uintptr_t)param | (uintptr_t)grad) & 15))
    return fail(MT_ERR_UNSUPPORTED, "mt_p2p_allreduce_sgd_f32: needs 16-byte aligned tensors of 4k elements");
  int slot = -1;
  for (size_t i = 0; i < s.slots.size(); ++i)
    if (s.slots[i].key == key && s.slots[i].n == n) { slot = (int)i; break; }
  if (slot < 0) {
    const size_t need = (n + 1023) & ~(size_t)1023;
    if ((int)s.slots.size() >= MAX_SLOTS || s.used_floats + need > s.region_floats)
      return fail(MT_ERR_UNSUPPORTED, "mt_p2p_allreduce_sgd_f32: arena full (%zu of %zu floats used, %zu slots)",
                  s.used_floats, s.region_floats, s.slots.size());
    s.slots.push_back(Slot{key, n, s.used_floats, 0u});
    s.used_floats += need;
    slot = (int)s.slots.size() - 1;
  }
  Slot& sl = s.slots[slot];
  const unsigned seq = ++sl.seq;
  Global& G = g();
  int rc = order_after_producer(grad);
  if (rc) return rc;
  const size_t n4 = n / 4;
  const int copy_grid = (int)std::min<size_t>(64, (n4 + THREADS - 1) / THREADS);
  const int xchg_grid = (int)std::min<size_t>(32, ((n4 + s.world - 1) / s.world + THREADS - 1) / THREADS);
  pool_mark_foreign(param, 2);
  pool_mark_foreign(grad, 2);
  if (G.trace_on) trace_mark("p2p_begin", G.comm_stream, 1);
  p2p_publish_kernel<<<copy_grid, THREADS, 0, G.comm_stream>>>(s.peers, grad, sl.off, n4, slot, s.rank, s.world, seq);
#define MT_XCHG(WW) \
  p2p_exchange_kernel<WW><<<xchg_grid < 1 ? 1 : xchg_grid, THREADS, 0, G.comm_stream>>>(s.peers, sl.off, n4, slot, s.rank, seq, s.err_dev)
  switch (s.world) {
    case 2: MT_XCHG(2); break;
    case 3: MT_XCHG(3); break;
    case 4: MT_XCHG(4); break;
    case 5: MT_XCHG(5); break;
    case 6: MT_XCHG(6); break;
    case 7: MT_XCHG(7); break;
    default: MT_XCHG(8); break;
  }
#undef MT_XCHG
  if (G.trace_on) trace_mark("p2p_exchanged", G.comm_stream, 1);
  // the update overwrites `param`: everything queued on the compute stream so far (the dX GEMM of this layer reads it)
  // must have run first
  MT_CUDA(cudaEventRecord(G.ev_compute, G.stream));
  MT_CUDA(cudaStreamWaitEvent(G.comm_stream, G.ev_compute, 0));
  p2p_apply_kernel<<<ew_grid((int64_t)n4, THREADS * 2), THREADS, 0, G.comm_stream>>>(s.peers, param, grad, sl.off, n4, slot,
                                                                                      s.rank, s.world, seq, lr, s.err_dev);
  G.launches += 3;
  if (G.trace_on) trace_mark("p2p_applied", G.comm_stream, 1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(MT_ERR_CUDA, "peer-memory exchange launch -> %s", cudaGetErrorString(e));
  return MT_OK;
}

MT_API int mt_p2p_destroy(void) {
  P2P& s = S();
  if (!s.arena) return MT_OK;
  if (g().ready) cudaStreamSynchronize(g().comm_stream);
  for (int r = 0; r < s.world; ++r)
    if (r != s.rank && s.peer_base[r]) cudaIpcCloseMemHandle(s.peer_base[r]);
  cudaFree(s.arena);
  s = P2P{};
  return MT_OK;
}
