// mt_gemm_skinny.cu - HBM-bound kernels for Linear layers with a tiny output width (N <= 16).
//
// The 4096->10 head of BASELINE.json config #2 is three "GEMMs" of 5.4 GFLOP each that read or write
// 1 GiB: they are bandwidth problems, not tensor-core problems.  The generic
// 128x128 SIMT tile wastes 92 % of its work on them (0.48 / 1.47 / 1.05 ms measured); these kernels stream
// the big operand once with 128-bit coalesced accesses and keep the small one on chip.
//
//   skinny_fwd : Y[M,N]  = act(X[M,K] . W[N,K]^T + b)        W (N*K*4 B <= 200 KB) in shared memory
//   skinny_dx  : dX[M,K] = dZ[M,N] . W[N,K]  (* (1 - X^2))   W in shared memory, dZ rows in registers
//   skinny_dw  : dW[N,K] (+)= dZ[M,N]^T . X[M,K]             per-CTA row slabs, partials folded deterministically
//
// Algorithmic traffic: fwd 4*M*K + 4*M*N; dx 4*M*K (+4*M*K with the tanh' epilogue) + 4*M*N; dw 4*M*K + 4*M*N.
// Reference call sites: the same matmul / bkwd_a / bkwd_b of src/tensor/ops/overload.py:132-154 as the big GEMMs.
#include <algorithm>

#include "mt_common.cuh"
#include "mt_gemm.cuh"

using namespace mt;

namespace {

constexpr int THREADS = 512;
constexpr int RB = 4;                       // rows per warp iteration (register blocking over the smem-resident W)
int g_skinny_ring = 1;                      // dW: 1 = cp.async ring, 0 = register double buffer (mt_gemm_set_skinny_ring)

__device__ __forceinline__ float dot4_acc(const float4 a, const float4 b, float c) {
  return fmaf(a.w, b.w, fmaf(a.z, b.z, fmaf(a.y, b.y, fmaf(a.x, b.x, c))));
}

// ------------------------------------------------------------------------------------------------ forward
// W lives in shared memory CHUNK-major: the N rows' float4 of column-vector c sit together at Ws[c * NS + n] with an
// ODD stride NS = N | 1, so (a) one address per chunk and N LDS.128 with immediate offsets instead of an index
// computation per row, and (b) the 8 lanes of an LDS.128 phase hit 8 different 16-byte bank groups: conflict-free.
// N is a template parameter (no predicate, no dead accumulators).  Round 1's kernel ([n][K4] layout, N padded to a
// multiple of 4) issued 390 instructions per warp-chunk for 160 FMAs and was issue-bound at 63 % of HBM
// (profiles/r02_hbm_kernels_ncu_raw.csv: 205 M warp instructions, issue slots 63 % busy at 25 % occupancy).
template <int N>
__device__ __forceinline__ void fill_w_chunk_major(float4* Ws, const float* __restrict__ W, int K4, long long ldw) {
  constexpr int NS = N | 1;
  for (int i = threadIdx.x; i < N * K4; i += THREADS) {
    const int n = i / K4, c = i - n * K4;
    Ws[c * NS + n] = __ldg(reinterpret_cast<const float4*>(W + n * ldw) + c);
  }
}

template <int N>
__global__ void __launch_bounds__(THREADS) skinny_fwd_kernel(float* __restrict__ Y, const float* __restrict__ X,
                                                             const float* __restrict__ W, const float* __restrict__ bias,
                                                             long long M, long long K, long long ldx, long long ldw,
                                                             int act) {
  constexpr int NS = N | 1;
  extern __shared__ float4 Ws[];            // [K/4][NS]
  const int K4 = (int)(K >> 2);
  fill_w_chunk_major<N>(Ws, W, K4, ldw);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long warps = (long long)gridDim.x * (THREADS / 32);
  for (long long r0 = ((long long)blockIdx.x * (THREADS / 32) + warp) * RB; r0 < M; r0 += warps * RB) {
    float acc[RB][N];
    const float4* xr[RB];                   // rows past the end re-read row M-1 (never stored): no per-row predicates
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      xr[i] = reinterpret_cast<const float4*>(X + (r0 + i < M ? r0 + i : M - 1) * ldx);
#pragma unroll
      for (int n = 0; n < N; ++n) acc[i][n] = 0.f;
    }
    // software pipeline, two chunks deep: 2 x RB 128-bit loads per lane (64 KB per SM) are in flight while chunk c is
    // folded in (bandwidth x latency is ~44 KB per SM)
    // (N > 10: the third buffer does not fit under the 128-register cap of a 512-thread CTA - one chunk ahead)
    constexpr bool DEEP = N <= 10;
    constexpr int AHEAD = DEEP ? 64 : 32;
    float4 xa[RB], xb[RB];
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      xa[i] = lane < K4 ? mtd::ld_stream4(reinterpret_cast<const float*>(xr[i] + lane)) : make_float4(0.f, 0.f, 0.f, 0.f);
      if (DEEP)
        xb[i] = lane + 32 < K4 ? mtd::ld_stream4(reinterpret_cast<const float*>(xr[i] + lane + 32)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll(DEEP ? 3 : 1)
    for (int c = lane; c < K4; c += 32) {
      float4 xf[RB];
      const bool more = c + AHEAD < K4;
#pragma unroll
      for (int i = 0; i < RB; ++i)
        xf[i] = more ? mtd::ld_stream4(reinterpret_cast<const float*>(xr[i] + c + AHEAD)) : make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* wp = Ws + c * NS;
#pragma unroll
      for (int n = 0; n < N; ++n) {
        const float4 w = wp[n];
#pragma unroll
        for (int i = 0; i < RB; ++i) acc[i][n] = dot4_acc(xa[i], w, acc[i][n]);   // 4 FMAs, no separate add
      }
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        if (DEEP) { xa[i] = xb[i]; xb[i] = xf[i]; }
        else xa[i] = xf[i];
      }
    }
#pragma unroll
    for (int i = 0; i < RB; ++i)
#pragma unroll
      for (int n = 0; n < N; ++n) acc[i][n] = mtd::warp_sum(acc[i][n]);
    // lanes 0..N-1 each finish one output column of the RB rows
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      if (r0 + i < M) {
        float v = 0.f;
#pragma unroll
        for (int n = 0; n < N; ++n)
          if (lane == n) v = acc[i][n];
        if (lane < N) {
          if (bias) v += __ldg(bias + lane);
          if (act == MT_ACT_TANH) v = tanhf(v);
          Y[(r0 + i) * N + lane] = v;
        }
      }
    }
  }
}

// per-thread asynchronous prefetch queue: 16-byte cp.async (L2 -> shared, no registers held while in flight).
// Every thread reads back only the slots it filled itself, so no barrier is needed around the ring.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, int bytes /* 16, or 0 = zero-fill */) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }

constexpr int DW_STAGES = 4;                // ring depth of skinny_dw_ring_kernel (one X row slice per stage)

// ------------------------------------------------------------------------------------------------ dX
// same shared-memory plan as the forward kernel (chunk-major W, odd stride, N a template parameter)
template <int N>
__global__ void __launch_bounds__(THREADS) skinny_dx_kernel(float* __restrict__ dX, const float* __restrict__ dZ,
                                                            const float* __restrict__ W, const float* __restrict__ epi_tanh,
                                                            long long M, long long K, long long ldz, long long ldw) {
  constexpr int NS = N | 1;
  extern __shared__ float4 Ws[];            // [K/4][NS]
  const int K4 = (int)(K >> 2);
  fill_w_chunk_major<N>(Ws, W, K4, ldw);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long warps = (long long)gridDim.x * (THREADS / 32);
  for (long long r0 = ((long long)blockIdx.x * (THREADS / 32) + warp) * RB; r0 < M; r0 += warps * RB) {
    float dz[RB][N];
    long long roff[RB];                    // row offset into dX and (same shape) the saved activations
#pragma unroll
    for (int i = 0; i < RB; ++i) {
      const long long r = r0 + i < M ? r0 + i : M - 1;
      const float mine = lane < N ? __ldg(dZ + r * ldz + lane) : 0.f;
#pragma unroll
      for (int n = 0; n < N; ++n) dz[i][n] = __shfl_sync(0xffffffffu, mine, n);
      roff[i] = r * K;
    }
    float4 tc[RB];                         // saved activations of chunk c (tanh' epilogue), loaded one chunk ahead
#pragma unroll
    for (int i = 0; i < RB; ++i)
      tc[i] = (epi_tanh && lane < K4) ? mtd::ld_stream4(epi_tanh + roff[i] + 4 * lane) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int c = lane; c < K4; c += 32) {
      float4 tn[RB];
      const bool more = epi_tanh && (c + 32 < K4);
#pragma unroll
      for (int i = 0; i < RB; ++i) tn[i] = more ? mtd::ld_stream4(epi_tanh + roff[i] + 4 * (c + 32)) : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 o[RB];
#pragma unroll
      for (int i = 0; i < RB; ++i) o[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float4* wp = Ws + c * NS;
#pragma unroll
      for (int n = 0; n < N; ++n) {
        const float4 w = wp[n];
#pragma unroll
        for (int i = 0; i < RB; ++i) {
          o[i].x = fmaf(dz[i][n], w.x, o[i].x);
          o[i].y = fmaf(dz[i][n], w.y, o[i].y);
          o[i].z = fmaf(dz[i][n], w.z, o[i].z);
          o[i].w = fmaf(dz[i][n], w.w, o[i].w);
        }
      }
#pragma unroll
      for (int i = 0; i < RB; ++i) {
        if (r0 + i < M) {                  // warp-uniform
          float4 v = o[i];
          if (epi_tanh) {
            const float4 t = tc[i];
            v.x = __fmul_rn(v.x, __fsub_rn(1.f, __fmul_rn(t.x, t.x)));
            v.y = __fmul_rn(v.y, __fsub_rn(1.f, __fmul_rn(t.y, t.y)));
            v.z = __fmul_rn(v.z, __fsub_rn(1.f, __fmul_rn(t.z, t.z)));
            v.w = __fmul_rn(v.w, __fsub_rn(1.f, __fmul_rn(t.w, t.w)));
          }
          mtd::st4(dX + roff[i] + 4 * c, v);
        }
      }
#pragma unroll
      for (int i = 0; i < RB; ++i) tc[i] = tn[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------ dW
// CTA b sums rows [b*slab, (b+1)*slab) of dZ^T.X into partial[b][N][K]; thread t owns float4 columns t, t+512, ...
constexpr int DZ_ROWS = 64;                 // dZ rows staged in shared memory per step

template <int NP, int CG>
__global__ void __launch_bounds__(THREADS) skinny_dw_kernel(float* __restrict__ partial, const float* __restrict__ dZ,
                                                            const float* __restrict__ X, long long M, int N, long long K,
                                                            long long ldz, long long ldx, long long slab) {
  __shared__ float dzs[DZ_ROWS][NP];
  const int K4 = (int)(K >> 2);
  const long long m_begin = (long long)blockIdx.x * slab;
  const long long m_end = m_begin + slab < M ? m_begin + slab : M;
  float4 acc[CG][NP];
#pragma unroll
  for (int g = 0; g < CG; ++g)
#pragma unroll
    for (int n = 0; n < NP; ++n) acc[g][n] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long mb = m_begin; mb < m_end; mb += DZ_ROWS) {
    const int rows = (int)(m_end - mb < DZ_ROWS ? m_end - mb : DZ_ROWS);
    __syncthreads();
    for (int i = threadIdx.x; i < DZ_ROWS * NP; i += THREADS) {
      const int r = i / NP, n = i - r * NP;
      dzs[r][n] = (r < rows && n < N) ? __ldg(dZ + (mb + r) * ldz + n) : 0.f;
    }
    __syncthreads();
    float4 xc[CG];                         // row r of this thread's columns, loaded one row ahead of its FMAs
#pragma unroll
    for (int g = 0; g < CG; ++g) {
      const int c = threadIdx.x + g * THREADS;
      xc[g] = (c < K4) ? mtd::ld_stream4(X + mb * ldx + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll 2
    for (int r = 0; r < rows; ++r) {
      float4 xn[CG];
#pragma unroll
      for (int g = 0; g < CG; ++g) {
        const int c = threadIdx.x + g * THREADS;
        xn[g] = (c < K4 && r + 1 < rows) ? mtd::ld_stream4(X + (mb + r + 1) * ldx + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int n = 0; n < NP; ++n) {
        const float d = dzs[r][n];          // broadcast read
#pragma unroll
        for (int g = 0; g < CG; ++g) {
          acc[g][n].x = fmaf(d, xc[g].x, acc[g][n].x);
          acc[g][n].y = fmaf(d, xc[g].y, acc[g][n].y);
          acc[g][n].z = fmaf(d, xc[g].z, acc[g][n].z);
          acc[g][n].w = fmaf(d, xc[g].w, acc[g][n].w);
        }
      }
#pragma unroll
      for (int g = 0; g < CG; ++g) xc[g] = xn[g];
    }
  }
  float* out = partial + (long long)blockIdx.x * N * K;
#pragma unroll
  for (int g = 0; g < CG; ++g) {
    const int c = threadIdx.x + g * THREADS;
    if (c < K4) {
#pragma unroll
      for (int n = 0; n < NP; ++n)
        if (n < N) mtd::st4(out + (long long)n * K + 4 * c, acc[g][n]);
    }
  }
}

// dW with the X rows staged through a cp.async ring (DW_STAGES rows of this CTA's column slice, private slots per
// thread): DW_STAGES-1 rows stay in flight while one row is folded into the accumulators.
template <int NP, int CG>
__global__ void __launch_bounds__(THREADS) skinny_dw_ring_kernel(float* __restrict__ partial, const float* __restrict__ dZ,
                                                                 const float* __restrict__ X, long long M, int N, long long K,
                                                                 long long ldz, long long ldx, long long slab) {
  __shared__ float dzs[DZ_ROWS][NP];
  extern __shared__ float4 ringbuf[];       // [stage][CG][THREADS]
  const int K4 = (int)(K >> 2);
  const long long m_begin = (long long)blockIdx.x * slab;
  const long long m_end = m_begin + slab < M ? m_begin + slab : M;
  const long long total = m_end - m_begin;
  float4 acc[CG][NP];
#pragma unroll
  for (int g = 0; g < CG; ++g)
#pragma unroll
    for (int n = 0; n < NP; ++n) acc[g][n] = make_float4(0.f, 0.f, 0.f, 0.f);
  auto issue = [&](long long q, int stage) {
#pragma unroll
    for (int g = 0; g < CG; ++g) {
      const int c = threadIdx.x + g * THREADS;
      const bool ok = c < K4;
      cp_async16(ringbuf + (stage * CG + g) * THREADS + threadIdx.x, ok ? X + (m_begin + q) * ldx + 4 * c : X, ok ? 16 : 0);
    }
  };
#pragma unroll
  for (int s = 0; s < DW_STAGES; ++s) {
    if (s < total) issue(s, s);
    cp_async_commit();
  }
  long long q = 0;
  for (long long mb = m_begin; mb < m_end; mb += DZ_ROWS) {
    const int rows = (int)(m_end - mb < DZ_ROWS ? m_end - mb : DZ_ROWS);
    __syncthreads();
    for (int i = threadIdx.x; i < DZ_ROWS * NP; i += THREADS) {
      const int r = i / NP, n = i - r * NP;
      dzs[r][n] = (r < rows && n < N) ? __ldg(dZ + (mb + r) * ldz + n) : 0.f;
    }
    __syncthreads();
    for (int r = 0; r < rows; ++r, ++q) {
      const int stage = (int)(q % DW_STAGES);
      cp_async_wait<DW_STAGES - 1>();
      float4 xc[CG];
#pragma unroll
      for (int g = 0; g < CG; ++g) xc[g] = ringbuf[(stage * CG + g) * THREADS + threadIdx.x];
      if (q + DW_STAGES < total) issue(q + DW_STAGES, stage);
      cp_async_commit();
#pragma unroll
      for (int n = 0; n < NP; ++n) {
        const float d = dzs[r][n];          // broadcast read
#pragma unroll
        for (int g = 0; g < CG; ++g) {
          acc[g][n].x = fmaf(d, xc[g].x, acc[g][n].x);
          acc[g][n].y = fmaf(d, xc[g].y, acc[g][n].y);
          acc[g][n].z = fmaf(d, xc[g].z, acc[g][n].z);
          acc[g][n].w = fmaf(d, xc[g].w, acc[g][n].w);
        }
      }
    }
  }
  float* out = partial + (long long)blockIdx.x * N * K;
#pragma unroll
  for (int g = 0; g < CG; ++g) {
    const int c = threadIdx.x + g * THREADS;
    if (c < K4) {
#pragma unroll
      for (int n = 0; n < NP; ++n)
        if (n < N) mtd::st4(out + (long long)n * K + 4 * c, acc[g][n]);
    }
  }
}

inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

constexpr size_t SMEM_OPTIN = 232448;      // 227 KB per CTA

inline size_t skinny_w_bytes(long long n, long long k) { return (size_t)(n | 1) * (size_t)k * 4; }

template <int N>
int run_fwd(const GemmArgs& a, long long ldx, long long ldw) {
  const size_t smem = skinny_w_bytes(N, a.K);
  const int grid = (int)std::min<long long>(g().sm_count, ceil_div(a.M, (THREADS / 32) * RB));
  auto kern = skinny_fwd_kernel<N>;
  static bool set = false;
  if (!set) { MT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_OPTIN)); set = true; }
  kern<<<grid, THREADS, smem, g().stream>>>(a.C, a.A, a.B, a.bias, a.M, a.K, ldx, ldw, a.act);
  MT_LAUNCHED();
  return MT_OK;
}

template <int N>
int run_dx(const GemmArgs& a, long long ldz, long long ldw) {
  // GEMM view: M x N_out(=a.N, the wide K of the layer) with reduction a.K (= the layer's tiny N)
  const size_t smem = skinny_w_bytes(N, a.N);
  auto kern = skinny_dx_kernel<N>;
  static bool set = false;
  if (!set) { MT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_OPTIN)); set = true; }
  const int grid = (int)std::min<long long>(g().sm_count, ceil_div(a.M, (THREADS / 32) * RB));
  kern<<<grid, THREADS, smem, g().stream>>>(a.C, a.A, a.B, a.epi_tanh, a.M, a.N, ldz, ldw);
  MT_LAUNCHED();
  return MT_OK;
}

// exact-N dispatch (1..16)
#define MT_BY_N(n, CALL)                                                                                               \
  switch (n) {                                                                                                         \
    case 1: return CALL(1); case 2: return CALL(2); case 3: return CALL(3); case 4: return CALL(4);                    \
    case 5: return CALL(5); case 6: return CALL(6); case 7: return CALL(7); case 8: return CALL(8);                    \
    case 9: return CALL(9); case 10: return CALL(10); case 11: return CALL(11); case 12: return CALL(12);              \
    case 13: return CALL(13); case 14: return CALL(14); case 15: return CALL(15); default: return CALL(16);            \
  }

template <int NP>
int run_dw(const GemmArgs& a, long long ldz, long long ldx) {
  // GEMM view: rows M (= tiny layer N), columns a.N (= layer K), reduction a.K (= batch rows)
  const long long rows = a.K, Kc = a.N;
  const int K4 = (int)(Kc / 4);
  const int cg = (int)ceil_div(K4, THREADS);
  long long ctas = std::min<long long>((long long)g().sm_count, ceil_div(rows, 256));
  const long long slab = ceil_div(ceil_div(rows, ctas), DZ_ROWS) * DZ_ROWS;
  ctas = ceil_div(rows, slab);
  PoolScratch scratch;
  int rc = scratch.alloc(sizeof(float) * (size_t)ctas * a.M * Kc);
  if (rc) return rc;
  float* partial = scratch.f32();
  if (g_skinny_ring) {
#define MT_DWR(CGV)                                                                                                    \
  do {                                                                                                                 \
    auto kern = skinny_dw_ring_kernel<NP, CGV>;                                                                        \
    const size_t ring = (size_t)DW_STAGES * CGV * THREADS * sizeof(float4);                                           \
    static bool set = false;                                                                                           \
    if (!set) { MT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring)); set = true; } \
    kern<<<(int)ctas, THREADS, ring, g().stream>>>(partial, a.A, a.B, rows, (int)a.M, Kc, ldz, ldx, slab);             \
  } while (0)
    switch (cg) {
      case 1: MT_DWR(1); break;
      case 2: MT_DWR(2); break;
      case 3: MT_DWR(3); break;
      default: MT_DWR(4); break;
    }
#undef MT_DWR
  } else {
#define MT_DW(CGV) skinny_dw_kernel<NP, CGV><<<(int)ctas, THREADS, 0, g().stream>>>(partial, a.A, a.B, rows, (int)a.M, Kc, ldz, ldx, slab)
    switch (cg) {
      case 1: MT_DW(1); break;
      case 2: MT_DW(2); break;
      case 3: MT_DW(3); break;
      default: MT_DW(4); break;
    }
#undef MT_DW
  }
  MT_LAUNCHED();
  return splitk_finalize(a.C, partial, (int)ctas, a.M, Kc, nullptr, MT_ACT_NONE, a.accumulate, nullptr);
}

template <typename F4, typename F8, typename F12, typename F16>
int by_np(long long n, F4 f4, F8 f8, F12 f12, F16 f16) {
  if (n <= 4) return f4();
  if (n <= 8) return f8();
  if (n <= 12) return f12();
  return f16();
}

}  // namespace

namespace mt {

int gemm_skinny_try(const GemmArgs& a) {
  if (a.tanh_src) return MT_ERR_UNSUPPORTED;
  // ---- forward shape: tiny N, A row-major [M,K], B(k,n) = W[n*ldw + k]
  if (a.N <= 16 && a.N >= 1 && a.M >= 256 && a.K >= 128 && a.a_sk == 1 && a.b_sk == 1 && !a.accumulate && !a.epi_tanh) {
    const long long ldx = a.a_sm, ldw = a.b_sn;
    if (a.K % 4 == 0 && ldx % 4 == 0 && ldw % 4 == 0 && al16(a.A) && al16(a.B) && skinny_w_bytes(a.N, a.K) <= SMEM_OPTIN) {
#define MT_FWD(NN) run_fwd<NN>(a, ldx, ldw)
      MT_BY_N(a.N, MT_FWD)
#undef MT_FWD
    }
  }
  // ---- dX shape: tiny reduction, A row-major [M,Kr], B(k,n) = W[k*ldw + n], wide output
  if (a.K <= 16 && a.K >= 1 && a.M >= 256 && a.N >= 128 && a.a_sk == 1 && a.b_sn == 1 && !a.accumulate && !a.bias &&
      a.act == MT_ACT_NONE) {
    const long long ldz = a.a_sm, ldw = a.b_sk;
    if (a.N % 4 == 0 && ldw % 4 == 0 && al16(a.B) && al16(a.C) && (!a.epi_tanh || al16(a.epi_tanh)) &&
        skinny_w_bytes(a.K, a.N) <= SMEM_OPTIN) {
#define MT_DX(NN) run_dx<NN>(a, ldz, ldw)
      MT_BY_N(a.K, MT_DX)
#undef MT_DX
    }
  }
  // ---- dW shape: tiny M, A(m,k) = dZ[k*ldz + m], B(k,n) = X[k*ldx + n], long reduction
  if (a.M <= 16 && a.M >= 1 && a.N >= 128 && a.K >= 1024 && a.a_sm == 1 && a.b_sn == 1 && !a.bias && a.act == MT_ACT_NONE &&
      !a.epi_tanh) {
    const long long ldz = a.a_sk, ldx = a.b_sk;
    // accumulators per thread = CG x NP float4: beyond 24 the kernel spills (ptxas -v), leave those to the generic tile
    const long long cg_need = ceil_div(a.N / 4, THREADS), np_need = ceil_div(a.M, 4) * 4;
    if (a.N % 4 == 0 && ldx % 4 == 0 && al16(a.B) && al16(a.C) && cg_need <= 4 && cg_need * np_need <= 24)
      return by_np(a.M, [&] { return run_dw<4>(a, ldz, ldx); }, [&] { return run_dw<8>(a, ldz, ldx); },
                   [&] { return run_dw<12>(a, ldz, ldx); }, [&] { return run_dw<16>(a, ldz, ldx); });
  }
  return MT_ERR_UNSUPPORTED;
}

}  // namespace mt

// tuning knob (tools/, tests): dW kernel with the cp.async ring (1, default) or the register double buffer (0).
// (The same ring was tried for the forward kernel: 0.43 ms against 0.29 ms for the register double buffer, dropped.)
extern "C" __attribute__((visibility("default"))) int mt_gemm_set_skinny_ring(int v) {
  g_skinny_ring = v ? 1 : 0;
  return MT_OK;
}
