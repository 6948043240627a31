// mt_gemm_simt.cu - generic fp32 CUDA-core GEMM: any strides, any sizes, batched, split-K.
//
// Role: (1) the bare `@` of the reference for arbitrary operands (src/tensor/ops/overload.py:
// 132,142,153), (2) the skinny / unaligned Linear shapes the tcgen05 path does not take (the
// 4096->10 head, the 2-wide spiral-demo layers), (3) the on-device cross-check of the tcgen05
// kernels in the GPU tests.  Exact fp32 FMA accumulation.
//
//   C[b, m, n] (+)= act( sum_k A[b*a_sb + m*a_sm + k*a_sk] * B[b*b_sb + k*b_sk + n*b_sn] + bias[n] )
//
// Tiling: BM x BN x 16 per CTA, 256 threads, TM x TN register micro-tile, register
// prefetch of the next k-slab while the current one is consumed from shared memory.
// Skinny problems use a 256x16 / 16x256 tile; problems that cannot fill 148 SMs split K
// across CTAs into a partial buffer that mt_reduce_sum_f32 folds (deterministic).
#include <algorithm>

#include "mt_common.cuh"
#include "mt_gemm.cuh"

using namespace mt;

namespace {

constexpr int BK = 16;
constexpr int THREADS = 256;

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(THREADS) sgemm_kernel(const GemmArgs p) {
  static_assert((BM / TM) * (BN / TN) == THREADS, "thread tiling must cover the CTA tile");
  constexpr int A_PER = BM * BK / THREADS, B_PER = BN * BK / THREADS;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const long long m0 = (long long)blockIdx.y * BM, n0 = (long long)blockIdx.x * BN;
  long long batch = 0, split = 0;
  if (p.splits > 1) split = blockIdx.z; else batch = blockIdx.z;
  const long long kchunk = p.kchunk;
  const long long kbeg = split * kchunk;
  const long long kend = kbeg + kchunk < p.K ? kbeg + kchunk : p.K;

  const float* __restrict__ A = p.A + batch * p.a_sb;
  const float* __restrict__ B = p.B + batch * p.b_sb;
  const bool a_kfast = (p.a_sk == 1), b_nfast = (p.b_sn == 1);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float ra[A_PER], rb[B_PER];
  auto fetch = [&](long long k0) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      const int e = tid + i * THREADS;
      const int kk = a_kfast ? e % BK : e / BM, mm = a_kfast ? e / BK : e % BM;
      const long long gm = m0 + mm, gk = k0 + kk;
      ra[i] = (gm < p.M && gk < kend) ? __ldg(A + gm * p.a_sm + gk * p.a_sk) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      const int e = tid + i * THREADS;
      const int kk = b_nfast ? e / BN : e % BK, nn = b_nfast ? e % BN : e / BK;
      const long long gn = n0 + nn, gk = k0 + kk;
      rb[i] = (gn < p.N && gk < kend) ? __ldg(B + gk * p.b_sk + gn * p.b_sn) : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      const int e = tid + i * THREADS;
      const int kk = a_kfast ? e % BK : e / BM, mm = a_kfast ? e / BK : e % BM;
      As[kk][mm] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      const int e = tid + i * THREADS;
      const int kk = b_nfast ? e / BN : e % BK, nn = b_nfast ? e % BN : e / BK;
      Bs[kk][nn] = rb[i];
    }
  };

  fetch(kbeg);
  for (long long k0 = kbeg; k0 < kend; k0 += BK) {
    stash();
    __syncthreads();
    if (k0 + BK < kend) fetch(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs[kk][tx * TN + j];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  float* __restrict__ C = p.C + (p.splits > 1 ? split * p.M * p.N : batch * p.M * p.N);
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const long long gm = m0 + ty * TM + i;
    if (gm >= p.M) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const long long gn = n0 + tx * TN + j;
      if (gn >= p.N) continue;
      float v = acc[i][j];
      if (p.splits == 1) {
        if (p.bias) v += __ldg(p.bias + gn);
        if (p.act == MT_ACT_TANH) v = tanhf(v);
        if (p.epi_tanh) { const float t = __ldg(p.epi_tanh + gm * p.N + gn); v = __fmul_rn(v, __fsub_rn(1.f, __fmul_rn(t, t))); }
        if (p.accumulate) v += C[gm * p.N + gn];
      }
      C[gm * p.N + gn] = v;
    }
  }
}

// epilogue of a split-K run: C = act(sum + bias) (+ C)
__global__ void __launch_bounds__(256) splitk_epilogue_kernel(float* __restrict__ C, const float* __restrict__ sum,
                                                              const float* __restrict__ bias, long long M, long long N,
                                                              int act, int accumulate, const float* __restrict__ epi_tanh) {
  const long long total = M * N;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    float v = sum[i];
    if (bias) v += __ldg(bias + i % N);
    if (act == MT_ACT_TANH) v = tanhf(v);
    if (epi_tanh) { const float t = __ldg(epi_tanh + i); v = __fmul_rn(v, __fsub_rn(1.f, __fmul_rn(t, t))); }
    if (accumulate) v += C[i];
    C[i] = v;
  }
}

// few splits (the tcgen05 split-K: S <= 8): fold the stacked partials and run the epilogue in ONE pass, split order
// 0..S-1 (deterministic).  Traffic 4*M*N*(S+1) instead of 4*M*N*(S+3) for reduce + epilogue.
__global__ void __launch_bounds__(256) splitk_fold_kernel(float* __restrict__ C, const float* __restrict__ partial, int splits,
                                                          const float* __restrict__ bias, long long M, long long N, int act,
                                                          int accumulate, const float* __restrict__ epi_tanh) {
  const long long total4 = (M * N) >> 2;             // N % 4 == 0 and 16-byte aligned bases (checked by the caller)
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long plane = M * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += stride) {
    float4 v = mtd::ld_stream4(partial + 4 * i);
    for (int sp = 1; sp < splits; ++sp) {
      const float4 u = mtd::ld_stream4(partial + sp * plane + 4 * i);
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    if (bias) {
      const float4 b = __ldg(reinterpret_cast<const float4*>(bias + (4 * i) % N));
      v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
    }
    if (act == MT_ACT_TANH) { v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w); }
    if (epi_tanh) {
      const float4 t = mtd::ld_stream4(epi_tanh + 4 * i);
      v.x = __fmul_rn(v.x, __fsub_rn(1.f, __fmul_rn(t.x, t.x)));
      v.y = __fmul_rn(v.y, __fsub_rn(1.f, __fmul_rn(t.y, t.y)));
      v.z = __fmul_rn(v.z, __fsub_rn(1.f, __fmul_rn(t.z, t.z)));
      v.w = __fmul_rn(v.w, __fsub_rn(1.f, __fmul_rn(t.w, t.w)));
    }
    if (accumulate) {
      const float4 c = *reinterpret_cast<const float4*>(C + 4 * i);
      v.x += c.x; v.y += c.y; v.z += c.z; v.w += c.w;
    }
    mtd::st4(C + 4 * i, v);
  }
}

template <int BM, int BN, int TM, int TN>
int launch(const GemmArgs& a, long long batch) {
  dim3 grid((unsigned)ceil_div(a.N, BN), (unsigned)ceil_div(a.M, BM), (unsigned)(a.splits > 1 ? a.splits : batch));
  sgemm_kernel<BM, BN, TM, TN><<<grid, THREADS, 0, g().stream>>>(a);
  MT_LAUNCHED();
  return MT_OK;
}

}  // namespace

namespace mt {

static int g_skinny = 1;   // test hook: 0 routes tiny-N shapes through the generic tile as well
static int g_small_tiles = 1;   // test hook: 0 keeps the 128x128 tile for small problems (mt_gemm_set_small_tiles)

int splitk_finalize(float* C, const float* partial, int splits, long long M, long long N, const float* bias, int act,
                    int accumulate, const float* epi_tanh) {
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  if (splits <= 16 && N % 4 == 0 && al16(C) && al16(partial) && (!bias || al16(bias)) && (!epi_tanh || al16(epi_tanh))) {
    splitk_fold_kernel<<<ew_grid((M * N) >> 2, 256), 256, 0, g().stream>>>(C, partial, splits, bias, M, N, act, accumulate,
                                                                          epi_tanh);
    MT_LAUNCHED();
    return MT_OK;
  }
  PoolScratch sum;
  int rc = sum.alloc(sizeof(float) * (size_t)M * N);
  if (rc) return rc;
  rc = mt_reduce_sum_f32(sum.f32(), partial, 1, splits, M * N, nullptr, 0, 0, 0);
  if (rc) return rc;
  splitk_epilogue_kernel<<<ew_grid(M * N, 1024), 256, 0, g().stream>>>(C, sum.f32(), bias, M, N, act, accumulate, epi_tanh);
  MT_LAUNCHED();
  return MT_OK;
}

int gemm_simt(GemmArgs a, long long batch) {
  if (a.M == 0 || a.N == 0 || batch == 0) return MT_OK;
  MT_CHECK_ARG(batch <= 65535, "matmul: batch %lld exceeds 65535", batch);
  if (batch == 1 && g_skinny) {
    int rc = gemm_skinny_try(a);
    if (rc != MT_ERR_UNSUPPORTED) return rc;
  }
  // tile choice
  int shape = 0;  // 0: 128x128, 1: 256x16 (skinny N), 2: 16x256 (skinny M), 3: 64x64, 4: 32x32
  const long long sms = g().sm_count;
  if (a.N <= 16 && a.M >= 64) shape = 1;
  else if (a.M <= 16 && a.N >= 64) shape = 2;
  else if (g_small_tiles) {
    // small problems (the 200-row layers of the spiral demo): a 128x128 tile leaves one or two CTAs on 148 SMs, each
    // thread grinding through 64 mostly-padding outputs (30-60 us per launch measured); smaller tiles spread the same
    // work over more SMs and skip the padding
    if (ceil_div(a.M, 128) * ceil_div(a.N, 128) * batch < sms)
      shape = (ceil_div(a.M, 64) * ceil_div(a.N, 64) * batch >= sms) ? 3 : 4;
  }
  const long long bm = shape == 0 ? 128 : (shape == 1 ? 256 : (shape == 2 ? 16 : (shape == 3 ? 64 : 32)));
  const long long bn = shape == 0 ? 128 : (shape == 1 ? 16 : (shape == 2 ? 256 : (shape == 3 ? 64 : 32)));
  const long long tiles = ceil_div(a.M, bm) * ceil_div(a.N, bn) * batch;
  long long splits = 1;
  if (batch == 1 && tiles < 2 * sms && a.K >= 1024) {
    splits = std::min<long long>(ceil_div(2 * sms, tiles), a.K / 256);
    if (splits > 65535) splits = 65535;
    if (splits < 1) splits = 1;
  }
  long long kchunk = ceil_div(std::max<long long>(a.K, 1), splits);
  kchunk = ceil_div(kchunk, BK) * BK;
  splits = std::max<long long>(1, ceil_div(a.K, kchunk));
  a.splits = (int)splits;
  a.kchunk = kchunk;
  float* final_c = a.C;
  PoolScratch partial;
  if (splits > 1) {
    int rc = partial.alloc(sizeof(float) * splits * a.M * a.N);
    if (rc) return rc;
    a.C = partial.f32();
  }
  int rc;
  if (shape == 0) rc = launch<128, 128, 8, 8>(a, batch);
  else if (shape == 1) rc = launch<256, 16, 4, 4>(a, batch);
  else if (shape == 2) rc = launch<16, 256, 4, 4>(a, batch);
  else if (shape == 3) rc = launch<64, 64, 4, 4>(a, batch);
  else rc = launch<32, 32, 2, 2>(a, batch);
  if (rc) return rc;
  if (splits > 1) {
    rc = splitk_finalize(final_c, partial.f32(), (int)splits, a.M, a.N, a.bias, a.act, a.accumulate, a.epi_tanh);
    if (rc) return rc;
  }
  return MT_OK;
}

}  // namespace mt

MT_API int mt_matmul_f32(float* C, const float* A, const float* B, int64_t batch, int64_t M, int64_t N, int64_t K,
                         int64_t a_sb, int64_t a_sm, int64_t a_sk, int64_t b_sb, int64_t b_sk, int64_t b_sn) {
  MT_REQUIRE_INIT();
  MT_CHECK_ARG(C && A && B, "mt_matmul_f32: null pointer");
  MT_CHECK_ARG(batch >= 0 && M >= 0 && N >= 0 && K >= 0, "mt_matmul_f32: negative extent");
  GemmArgs a{};
  a.C = C; a.A = A; a.B = B; a.M = M; a.N = N; a.K = K;
  a.a_sb = a_sb; a.a_sm = a_sm; a.a_sk = a_sk; a.b_sb = b_sb; a.b_sk = b_sk; a.b_sn = b_sn;
  a.bias = nullptr; a.act = MT_ACT_NONE; a.accumulate = 0;
  // 2-D operands with TMA-compatible layouts go to the tensor cores
  if (batch == 1 && g().gemm_engine_force != 0) {
    int rc = gemm_tc_try(a);
    if (rc == MT_OK) { g().gemm_last_engine = 1; return MT_OK; }
    if (rc != MT_ERR_UNSUPPORTED) return rc;
    if (g().gemm_engine_force >= 1) return rc;
  }
  g().gemm_last_engine = 0;
  return gemm_simt(a, batch);
}

extern "C" __attribute__((visibility("default"))) int mt_gemm_set_small_tiles(int v) {
  mt::g_small_tiles = v ? 1 : 0;
  return MT_OK;
}

extern "C" __attribute__((visibility("default"))) int mt_gemm_set_skinny(int v) {
  mt::g_skinny = v ? 1 : 0;
  return MT_OK;
}

MT_API int mt_gemm_last_engine(int* engine) {
  MT_CHECK_ARG(engine, "mt_gemm_last_engine: null out pointer");
  *engine = g().gemm_last_engine;
  return MT_OK;
}

MT_API int mt_gemm_set_engine(int engine) {
  MT_CHECK_ARG(engine >= -1 && engine <= 2, "mt_gemm_set_engine: engine must be -1, 0, 1 or 2");
  g().gemm_engine_force = engine;
  return MT_OK;
}
