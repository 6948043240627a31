// mt_rng.cu - counter-based device random numbers for Tensor.randn / Tensor.uniform on device="cuda".
//
// Reference: src/tensor/tensor.py:402-434 - `_tensor(device).random.randn(*dims)` / `.random.uniform(low, high, dims)`,
// i.e. cupy.random on the CUDA path.  CuPy's generator cannot be reproduced bit for bit (and NumPy's Mersenne stream is
// a different one again), so parity here is by distribution; what this generator guarantees instead is a SPECIFIED
// stream: Philox4x32-10 (Salmon et al., SC'11) keyed by a 64-bit seed, block b of the stream = Philox(counter = (b, 0),
// key = seed), four 32-bit words per block.  tests/test_rng_gpu.py restates it in NumPy and checks the uniform stream
// bit for bit and the normal stream to 1e-6.  Stateless: the caller (microtorch_b200.cuda.random) owns seed and offset.
//
//   uniform : u = ((x >> 8) + 0.5) * 2^-24 in (0, 1), out = low + (high - low) * u (two fp32 roundings)
//   normal  : Box-Muller on word pairs (x0, x1), (x2, x3): r = sqrt(-2 log u0), z = r cos(2 pi u1), r sin(2 pi u1)
// HBM-bound (4 B written per element), 128-bit stores, grid-stride.
#include "mt_common.cuh"

using namespace mt;

namespace {

constexpr int THREADS = 256;

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ float u01(unsigned x) { return ((float)(x >> 8) + 0.5f) * 5.9604644775390625e-08f; }

template <bool NORMAL>
__global__ void __launch_bounds__(THREADS) rng_kernel(float* __restrict__ out, size_t n, unsigned long long seed,
                                                      unsigned long long offset, float low, float span, int vec) {
  const size_t blocks4 = (n + 3) / 4;
  const size_t stride = (size_t)gridDim.x * THREADS;
  const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
  for (size_t b = (size_t)blockIdx.x * THREADS + threadIdx.x; b < blocks4; b += stride) {
    const unsigned long long ctr = offset + b;
    const uint4 x = philox4x32_10(make_uint4((unsigned)ctr, (unsigned)(ctr >> 32), 0u, 0u), key);
    float4 v;
    if (NORMAL) {
      const float r0 = sqrtf(-2.f * logf(u01(x.x))), r1 = sqrtf(-2.f * logf(u01(x.z)));
      float s0, c0, s1, c1;
      sincospif(2.f * u01(x.y), &s0, &c0);
      sincospif(2.f * u01(x.w), &s1, &c1);
      v = make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
    } else {
      v.x = __fadd_rn(low, __fmul_rn(span, u01(x.x)));
      v.y = __fadd_rn(low, __fmul_rn(span, u01(x.y)));
      v.z = __fadd_rn(low, __fmul_rn(span, u01(x.z)));
      v.w = __fadd_rn(low, __fmul_rn(span, u01(x.w)));
    }
    const size_t i = 4 * b;
    if (vec && i + 3 < n) {
      mtd::st4(out + i, v);
    } else {
      const float e[4] = {v.x, v.y, v.z, v.w};
      for (int j = 0; j < 4 && i + j < n; ++j) out[i + j] = e[j];
    }
  }
}

}  // namespace

// out[i] = low + (high - low) * u_i, u_i the i-th uniform of the Philox stream (seed, starting at block `offset`).
// Consumes ceil(n / 4) blocks.
MT_API int mt_rng_uniform_f32(float* out, size_t n, unsigned long long seed, unsigned long long offset, float low, float high) {
  MT_REQUIRE_INIT();
  MT_CHECK_ARG(out || !n, "mt_rng_uniform_f32: null output");
  if (!n) return MT_OK;
  const int vec = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  rng_kernel<false><<<ew_grid((int64_t)((n + 3) / 4), THREADS), THREADS, 0, g().stream>>>(out, n, seed, offset, low, high - low, vec);
  MT_LAUNCHED();
  return MT_OK;
}

// out[i] = the i-th standard normal of the stream (Box-Muller on consecutive word pairs).  Consumes ceil(n / 4) blocks.
MT_API int mt_rng_normal_f32(float* out, size_t n, unsigned long long seed, unsigned long long offset) {
  MT_REQUIRE_INIT();
  MT_CHECK_ARG(out || !n, "mt_rng_normal_f32: null output");
  if (!n) return MT_OK;
  const int vec = (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  rng_kernel<true><<<ew_grid((int64_t)((n + 3) / 4), THREADS), THREADS, 0, g().stream>>>(out, n, seed, offset, 0.f, 1.f, vec);
  MT_LAUNCHED();
  return MT_OK;
}
