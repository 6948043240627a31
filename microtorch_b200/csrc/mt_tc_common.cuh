// mt_tc_common.cuh - device-side building blocks shared by the tcgen05 GEMM kernels (mt_gemm_tc.cu: the persistent
// CTA-pair kernel for large problems; mt_gemm_tc_mid.cu: the one-tile-per-CTA kernel for the mid-size regime):
// mbarrier / TMA / tcgen05 PTX wrappers, shared-memory and instruction descriptors, and the converter bodies that
// split an fp32 operand tile into its correction tiles on chip (see mt_gemm_tc.cu for the arithmetic).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>

#include "mt_common.cuh"

namespace tc {

constexpr int BM = 128;
constexpr int BK = 32;            // 32 fp32 = 128 B = one SWIZZLE_128B row
constexpr int UMMA_K = 8;         // kind::tf32: 32 B of K per instruction

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
// Bounded wait: a protocol bug must end in a trap (a clean CUDA error), never in a hung GPU.
// No printf / no calls here: a real call makes ptxas size the whole kernel for the smallest
// setmaxnreg budget (observed: 53 registers and the running sums spilled to local memory).
// The culprit is left in a host-mapped word instead (mt_gemm_tc_debug_word()).
static __device__ int* g_dbg_word = nullptr;   // one copy per translation unit (no -rdc): each sets its own
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int who) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {  // ~2 s
      if (g_dbg_word) {
        g_dbg_word[0] = 0x7100 | who;
        g_dbg_word[1] = blockIdx.x;
        g_dbg_word[2] = threadIdx.x;
        g_dbg_word[3] = parity;
        __threadfence_system();
      }
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// same with an L2 eviction-priority hint (createpolicy encodings: 0x12F0.. evict_first, 0x14F0.. evict_last)
__device__ __forceinline__ void tma_load_2d_hint(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1,
                                                 uint64_t policy) {
  if (policy == 0) { tma_load_2d(dst, map, bar, c0, c1); return; }
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// ---- cluster / CTA-pair helpers (cta_group::2)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (.release.cta), as CUTLASS's ClusterBarrier::arrive(cta_id) does: the cluster-scope release
  // compiles to MEMBAR.ALL.GPU in front of every arrive and showed up as a ~1.7 us converter<->MMA round trip
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok;
}
template <bool CLUSTER>
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, int who) {
  if (!CLUSTER) { mbar_wait(bar, parity, who); return; }
  if (mbar_try_wait_cluster(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      if (g_dbg_word) {
        g_dbg_word[0] = 0x7200 | who;
        g_dbg_word[1] = blockIdx.x;
        g_dbg_word[2] = threadIdx.x;
        g_dbg_word[3] = parity;
        __threadfence_system();
      }
      __trap();
    }
  }
}
template <int NCTA>
__device__ __forceinline__ void tmem_alloc_n(uint32_t dst_smem, uint32_t cols) {
  if (NCTA == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int NCTA>
__device__ __forceinline__ void tmem_dealloc_n(uint32_t taddr, uint32_t cols) {
  if (NCTA == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
  else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
template <int NCTA>
__device__ __forceinline__ void umma_tf32_n(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (NCTA == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
  }
}
// 16-bit inputs (bf16 here), fp32 accumulate: K = 16 per instruction
template <int NCTA>
__device__ __forceinline__ void umma_f16_n(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (NCTA == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
        : "memory");
  }
}
// tcgen05.commit: 1-CTA -> own barrier; CTA pair -> the barrier at the same offset in BOTH CTAs
template <int NCTA>
__device__ __forceinline__ void umma_commit_n(uint32_t bar) {
  if (NCTA == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  } else {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask)
                 : "memory");
  }
}
template <int R>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }

// ---------------------------------------------------------------- descriptors
// Shared-memory matrix descriptor (SM100 format, version 1).
//   K-major : SWIZZLE_128B (type 2).  Rows of 128 B (32 k), 8-row atoms 1024 B apart (SBO); LBO unused (1).
//             One UMMA_K slice = 32 B further along the row.
//   MN-major: 32-bit operands only exist as SWIZZLE_128B_BASE32B (type 1; TMA: SWIZZLE_128B_ATOM_32B).
//             A 32(mn) x 32(k) TMA box is 32 rows of 128 B; 4-row k-atoms 512 B apart (SBO), 32-element
//             MN atoms one box = 4096 B apart (LBO).  One UMMA_K slice (8 k) = 1024 B further.
__host__ __device__ constexpr uint64_t make_smem_desc_base(bool mn_major) {
  return mn_major ? ((uint64_t(4096 >> 4) << 16) | (uint64_t(512 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(1) << 61))
                  : ((uint64_t(1) << 16) | (uint64_t(1024 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(2) << 61));
}
__device__ __forceinline__ uint64_t smem_desc(uint64_t base, uint32_t addr) {
  return base | uint64_t((addr & 0x3FFFFu) >> 4);
}
// Instruction descriptor: D=f32, A=B=tf32, majors, N>>3 at [17,23), M>>4 at [24,29).
__host__ __device__ constexpr uint32_t make_idesc(int m, int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(a_mn) << 15) | (uint32_t(b_mn) << 16) |
         (uint32_t(n >> 3) << 17) | (uint32_t(m >> 4) << 24);
}

// "mix" mode (hi.hi on TF32, the two first-order corrections on BF16 at twice the rate): the correction operands are
// 16-bit K-major tiles of 32 k (64-byte rows) in SWIZZLE_64B (type 4): 8-row atoms 512 B apart (SBO), one UMMA_K slice
// (16 k) = 32 B further along the row.
// MN-major correction tiles: 32 (mn) x 32 (k) boxes of 2048 B (LBO: next 32 mn), 8-row k-atoms 512 B apart (SBO), one
// UMMA_K slice (16 k) = 1024 B further.
__host__ __device__ constexpr uint64_t make_smem_desc_base_16(bool mn_major) {
  return mn_major ? ((uint64_t(2048 >> 4) << 16) | (uint64_t(512 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(4) << 61))
                  : ((uint64_t(1) << 16) | (uint64_t(512 >> 4) << 32) | (uint64_t(1) << 46) | (uint64_t(4) << 61));
}
// D=f32, A=B=bf16
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n, bool a_mn, bool b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(a_mn) << 15) | (uint32_t(b_mn) << 16) | (uint32_t(n >> 3) << 17) |
         (uint32_t(m >> 4) << 24);
}

// hi/lo split of one 16-byte vector.  RN=true rounds hi to nearest tf32 (needs the explicit hi store);
// RN=false leaves hi = truncation, which is what the tensor core itself does with the raw fp32 bits.
template <bool RN>
__device__ __forceinline__ void split4(const float4 v, float4& hi, float4& lo) {
  const uint32_t add = RN ? 0x1000u : 0u;
  hi.x = __uint_as_float((__float_as_uint(v.x) + add) & 0xFFFFE000u);
  hi.y = __uint_as_float((__float_as_uint(v.y) + add) & 0xFFFFE000u);
  hi.z = __uint_as_float((__float_as_uint(v.z) + add) & 0xFFFFE000u);
  hi.w = __uint_as_float((__float_as_uint(v.w) + add) & 0xFFFFE000u);
  lo.x = __uint_as_float((__float_as_uint(v.x - hi.x) + 0x1000u) & 0xFFFFE000u);
  lo.y = __uint_as_float((__float_as_uint(v.y - hi.y) + 0x1000u) & 0xFFFFE000u);
  lo.z = __uint_as_float((__float_as_uint(v.z - hi.z) + 0x1000u) & 0xFFFFE000u);
  lo.w = __uint_as_float((__float_as_uint(v.w - hi.w) + 0x1000u) & 0xFFFFE000u);
}
// all-TF32 split with the CENTRED remainder (see centered_lo below): lo = tf32_rn((x - h (1 + c)) / (1 + c)).
// Here the other factor of a correction product is the TRUNCATED operand h = (x - l) / (1 + c), not x, so
//     A.B = S A_h.B_h + (1 + c)(A_l.B_h + A_h.B_l) + A_l.B_l,   S = (1 + c)^2
// and the remainders are pre-divided by (1 + c) (instead of S as in the mix mode) for the epilogue's single scale S.
__device__ __forceinline__ float centered_lo_tf32(float x);
__device__ __forceinline__ float4 split4_centered(const float4 v) {
  float4 lo;
  lo.x = __uint_as_float((__float_as_uint(centered_lo_tf32(v.x)) + 0x1000u) & 0xFFFFE000u);
  lo.y = __uint_as_float((__float_as_uint(centered_lo_tf32(v.y)) + 0x1000u) & 0xFFFFE000u);
  lo.z = __uint_as_float((__float_as_uint(centered_lo_tf32(v.z)) + 0x1000u) & 0xFFFFE000u);
  lo.w = __uint_as_float((__float_as_uint(centered_lo_tf32(v.w)) + 0x1000u) & 0xFFFFE000u);
  return lo;
}

// ---- converter bodies.  All loads of a batch are issued before any dependent math or store: written as a
// plain loop the compiler keeps load -> math -> store order per vector (the stores may alias the next load), which
// serialises ~150 cycles per float4 (measured: 2500 cycles per 32 KB k-block, the bottleneck of the whole kernel).
template <int VECS, bool WRITE_HI>
__device__ __forceinline__ void convert_tile(uint8_t* raw_b, uint8_t* lo_b, int ct) {
  constexpr int PER = VECS / 128;              // float4 per thread
  constexpr int U = PER < 8 ? PER : 8;
  float4* raw = reinterpret_cast<float4*>(raw_b);
  float4* lo = reinterpret_cast<float4*>(lo_b);
#pragma unroll
  for (int b = 0; b < PER; b += U) {
    float4 v[U], h[U], l[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = raw[ct + (b + u) * 128];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (WRITE_HI) split4<true>(v[u], h[u], l[u]);
      else l[u] = split4_centered(v[u]);          // hi = the tensor core's own truncation of the raw tile
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (WRITE_HI) raw[ct + (b + u) * 128] = h[u];
      lo[ct + (b + u) * 128] = l[u];
    }
  }
}
// mix mode: CENTERED remainder.  The tensor core TRUNCATES an fp32 operand to TF32, so with h = trunc_tf32(x) the plain
// remainder x - h lies in [0, 2^-10)|x| - one-sided, and the dropped lo.lo term has a non-zero mean (a systematic
// -2^-22 bias).  Writing x = h (1 + c) + l with c = 2^-11 centres it, |l| <= 2^-11 |h|, and since (1 + c) h = x - l:
//     A.B = (1 + c)^2 A_h.B_h  +  A_l.B  +  A.B_l  -  A_l.B_l                     (exact)
// The kernel accumulates  T = A_h.B_h + bf16(A_l / S).bf16(B) + bf16(A).bf16(B_l / S)  with S = (1 + c)^2 and the
// epilogue multiplies by S (one fp32 multiply per output): same three products, same operands, but the remainders that
// get rounded to BF16 are half as large - the rounding error of the two correction products, the error floor of the
// mix mode, halves - and the dropped A_l.B_l (2^-24 / 3) is unbiased.  h (1 + c) needs 22 bits and x - h (1 + c) 12:
// both exact in fp32 (measured before / after: tests/gemm_accuracy.py, profiles/r02_gemm_accuracy.jsonl).
constexpr float kMixCenter = 1.00048828125f;          // 1 + 2^-11
constexpr float kMixScale = 1.00097680091857910156f;  // S = (1 + 2^-11)^2 = 1 + 2^-10 + 2^-22, exact in fp32
constexpr float kMixInvScale = 0.99902415275573730469f;   // 1 / S to fp32
constexpr float kMixInvCenter = 0.99951195716857910156f;  // 1 / (1 + 2^-11) to fp32
__device__ __forceinline__ float centered_lo(float x) {
  return fmaf(-__uint_as_float(__float_as_uint(x) & 0xFFFFE000u), kMixCenter, x) * kMixInvScale;
}
__device__ __forceinline__ float centered_lo_tf32(float x) {
  return fmaf(-__uint_as_float(__float_as_uint(x) & 0xFFFFE000u), kMixCenter, x) * kMixInvCenter;
}
// mix mode, K-major source: the fp32 tile (SWIZZLE_128B: 16-byte chunk c of row r sits at chunk c ^ (r & 7)) is split
// into bf16(x) and bf16(x - trunc_tf32(x)) and written as two SWIZZLE_64B tiles (64-byte rows; 16-byte chunk q of row r
// sits at chunk q ^ ((r >> 1) & 3)).  A thread keeps its physical source chunk, so the logical column and the
// destination offset inside the row are per-thread constants; rows advance by 16 per step.
template <int VECS, int NTHR>
__device__ __forceinline__ void convert_tile_mix_k(const uint8_t* raw_b, uint8_t* hi16_b, uint8_t* lo16_b, int ct) {
  static_assert(NTHR % 64 == 0, "a step must advance whole groups of 8 rows so the swizzle terms stay per-thread constants");
  constexpr int PER = (VECS + NTHR - 1) / NTHR;                           // float4 per thread (last one may be out of range)
  constexpr int U = 3;
  const float4* raw = reinterpret_cast<const float4*>(raw_b);
  const int c = (ct & 7) ^ ((ct >> 3) & 7);                               // logical 16-byte chunk: k = 4c .. 4c+3
  const int dst0 = (ct >> 3) * 64 + ((((c >> 1) ^ ((ct >> 4) & 3))) << 4) + ((c & 1) << 3);
#pragma unroll
  for (int b = 0; b < PER; b += U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (b + u < PER && ct + (b + u) * NTHR < VECS) v[u] = raw[ct + (b + u) * NTHR];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (b + u < PER && ct + (b + u) * NTHR < VECS) {
        const float4 x = v[u];
        float4 lo;
        lo.x = centered_lo(x.x);
        lo.y = centered_lo(x.y);
        lo.z = centered_lo(x.z);
        lo.w = centered_lo(x.w);
        __nv_bfloat162 h0 = __floats2bfloat162_rn(x.x, x.y), h1 = __floats2bfloat162_rn(x.z, x.w);
        __nv_bfloat162 l0 = __floats2bfloat162_rn(lo.x, lo.y), l1 = __floats2bfloat162_rn(lo.z, lo.w);
        const int off = dst0 + (b + u) * (NTHR / 8) * 64;
        *reinterpret_cast<uint2*>(hi16_b + off) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
        *reinterpret_cast<uint2*>(lo16_b + off) = make_uint2(*reinterpret_cast<uint32_t*>(&l0), *reinterpret_cast<uint32_t*>(&l1));
      }
    }
  }
}
// mix mode, MN-major source (dW / dX operands): the raw tile is ROWS/32 TMA boxes of 32 (mn) x 32 (k) floats in
// SWIZZLE_128B_ATOM_32B = Swizzle<2,5,2>: row k of a box is 128 B = 32 mn, and the 32-byte chunk c of row k sits at chunk
// c ^ (k & 3).  The correction tiles stay MN-major: boxes of 32 (k) rows x 64 B (32 mn as bf16) in SWIZZLE_64B =
// Swizzle<2,4,3> (16-byte chunk q of row k sits at q ^ ((k >> 1) & 3)).  Four consecutive mn are 16 B of the source and
// 8 B of the destination, and the 32-byte source chunk index equals the 16-byte destination chunk index, so the
// conversion is element-wise on the byte image again: contiguous loads, contiguous 8-byte stores, no transpose.
template <int ROWS, int NTHR>
__device__ __forceinline__ void convert_tile_mix_mn(const uint8_t* raw_b, uint8_t* hi16_b, uint8_t* lo16_b, int ct) {
  constexpr int VECS = ROWS * 8;
  constexpr int PER = (VECS + NTHR - 1) / NTHR;          // float4 per thread (last one may be out of range)
  constexpr int U = 3;
  const float4* raw = reinterpret_cast<const float4*>(raw_b);
#pragma unroll
  for (int b = 0; b < PER; b += U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (b + u < PER && ct + (b + u) * NTHR < VECS) v[u] = raw[ct + (b + u) * NTHR];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = ct + (b + u) * NTHR;
      if (b + u < PER && i < VECS) {
        const int box = i >> 8, k = (i & 255) >> 3, p16 = i & 7;
        const int c32 = (p16 >> 1) ^ (k & 3);
        const int off = box * 2048 + k * 64 + ((c32 ^ ((k >> 1) & 3)) << 4) + ((p16 & 1) << 3);
        const float4 x = v[u];
        float4 lo;
        lo.x = centered_lo(x.x);
        lo.y = centered_lo(x.y);
        lo.z = centered_lo(x.z);
        lo.w = centered_lo(x.w);
        __nv_bfloat162 h0 = __floats2bfloat162_rn(x.x, x.y), h1 = __floats2bfloat162_rn(x.z, x.w);
        __nv_bfloat162 l0 = __floats2bfloat162_rn(lo.x, lo.y), l1 = __floats2bfloat162_rn(lo.z, lo.w);
        *reinterpret_cast<uint2*>(hi16_b + off) = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
        *reinterpret_cast<uint2*>(lo16_b + off) = make_uint2(*reinterpret_cast<uint32_t*>(&l0), *reinterpret_cast<uint32_t*>(&l1));
      }
    }
  }
}
// backward prologue: z = g * (1 - y^2) replaces the raw tile (the MMA truncates it to hi itself), lo(z) -> lo ring.
// `centred` = the centred remainder of split4_centered (the epilogue then scales by S); it must match what convert_tile
// does to the B operand of the same launch (write_hi = 0: centred, write_hi = 1: plain).
template <int VECS>
__device__ __forceinline__ void convert_tile_prologue(uint8_t* raw_b, const uint8_t* y_b, uint8_t* lo_b, int ct,
                                                      bool centred) {
  constexpr int PER = VECS / 128;
  constexpr int U = PER < 4 ? PER : 4;
  float4* raw = reinterpret_cast<float4*>(raw_b);
  const float4* ys = reinterpret_cast<const float4*>(y_b);
  float4* lo = reinterpret_cast<float4*>(lo_b);
#pragma unroll
  for (int b = 0; b < PER; b += U) {
    float4 g[U], y[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      g[u] = raw[ct + (b + u) * 128];
      y[u] = ys[ct + (b + u) * 128];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float4 z, hi;
      z.x = __fmul_rn(g[u].x, __fsub_rn(1.f, __fmul_rn(y[u].x, y[u].x)));
      z.y = __fmul_rn(g[u].y, __fsub_rn(1.f, __fmul_rn(y[u].y, y[u].y)));
      z.z = __fmul_rn(g[u].z, __fsub_rn(1.f, __fmul_rn(y[u].z, y[u].z)));
      z.w = __fmul_rn(g[u].w, __fsub_rn(1.f, __fmul_rn(y[u].w, y[u].w)));
      g[u] = z;
      if (centred) y[u] = split4_centered(z);   // y[u] now holds lo(z): same split as the B operand gets (convert_tile)
      else split4<false>(z, hi, y[u]);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      raw[ct + (b + u) * 128] = g[u];
      lo[ct + (b + u) * 128] = y[u];
    }
  }
}

// ---------------------------------------------------------------- host-side services of mt_gemm_tc.cu
// (shared with mt_gemm_tc_mid.cu)
// cached cuTensorMapEncodeTiled of a 2-D fp32 operand: dim0 = contiguous extent `inner`, dim1 = `outer` rows `ld` apart
int make_map(CUtensorMap* m, const float* base, long long inner, long long outer, long long ld, int box_inner,
             int box_outer, bool mn_major);
int* debug_word_device();                           // device pointer of the host-mapped barrier-timeout word
int prof_begin(double flops, double unit_flops);    // per-launch event timing for the bench roofline (mt_prof_*)
void prof_end(int token);
void precision_for(bool forward, int* mix, int* chunk_kb, int* dbg);

}  // namespace tc
