// mt_elementwise.cu - broadcasting elementwise forward / fused backward kernels (fp32).
//
// HBM-bound: 12 B/element for a binary op on equal shapes (8 B + broadcast operand
// otherwise), 8 B/element for a unary op.  Every fast path moves 128-bit vectors with
// coalesced warps and four independent vectors in flight per thread; grids are a multiple
// of the SM count and loop with a grid stride.
//
// Reference call sites replaced (paths relative to the reference checkout):
//   a + b, a * b, -a                 src/tensor/ops/overload.py:52,70,94
//   grad * b.data                    src/tensor/ops/overload.py:109
//   log/exp/tanh, x**p               src/tensor/ops/math.py:8,27,47,70
//   g / x, g * y, g*(p*x**(p-1)), g*(1-y**2)   src/tensor/ops/math.py:12,31,51,74 (one kernel each
//                                    instead of 1..3 kernels + temporaries)
//   np.add/multiply/true_divide(out=)   src/tensor/tensor.py:549,570,587,615
// Arithmetic uses the explicit round-to-nearest intrinsics so that nvcc does not contract
// a*b+c into an FMA: NumPy rounds after every operation and several results are compared
// bit-for-bit with it.
#include "mt_common.cuh"

using namespace mt;

namespace {

// NumPy's scalar-exponent fast paths: x**2 = x*x, x**-1 = 1/x, x**0.5 = sqrt,
// x**1 = x, x**0 = 1; everything else goes through powf.
__device__ __forceinline__ float np_pow(float x, float p) {
  if (p == 2.f) return __fmul_rn(x, x);
  if (p == 1.f) return x;
  if (p == 0.f) return 1.f;
  if (p == -1.f) return __fdiv_rn(1.f, x);
  if (p == 0.5f) return __fsqrt_rn(x);
  return powf(x, p);
}

template <int OP>
__device__ __forceinline__ float bin_op(float a, float b, float s) {
  if (OP == MT_BIN_ADD) return __fadd_rn(a, b);
  if (OP == MT_BIN_MUL) return __fmul_rn(a, b);
  if (OP == MT_BIN_SUB) return __fsub_rn(a, b);
  if (OP == MT_BIN_DIV) return __fdiv_rn(a, b);
  if (OP == MT_BIN_TANH_BWD) return __fmul_rn(a, __fsub_rn(1.f, __fmul_rn(b, b)));
  if (OP == MT_BIN_POW_BWD) return __fmul_rn(a, __fmul_rn(s, np_pow(b, s - 1.f)));
  if (OP == MT_BIN_EXP_BWD) return __fmul_rn(a, b);
  if (OP == MT_BIN_LT) return a < b ? 1.f : 0.f;
  if (OP == MT_BIN_GT) return a > b ? 1.f : 0.f;
  if (OP == MT_BIN_LE) return a <= b ? 1.f : 0.f;
  if (OP == MT_BIN_GE) return a >= b ? 1.f : 0.f;
  if (OP == MT_BIN_EQ) return a == b ? 1.f : 0.f;
  if (OP == MT_BIN_NE) return a != b ? 1.f : 0.f;
  if (OP == MT_BIN_MAX) return (a != a || b != b) ? __int_as_float(0x7fc00000) : fmaxf(a, b);   // NumPy propagates NaN
  if (OP == MT_BIN_MIN) return (a != a || b != b) ? __int_as_float(0x7fc00000) : fminf(a, b);
  return 0.f;
}

template <int OP>
__device__ __forceinline__ float un_op(float x, float s) {
  if (OP == MT_UN_NEG) return -x;
  if (OP == MT_UN_LOG) return logf(x);
  if (OP == MT_UN_TANH) return tanhf(x);
  if (OP == MT_UN_EXP) return expf(x);
  if (OP == MT_UN_POW) return np_pow(x, s);
  if (OP == MT_UN_SCALE) return __fmul_rn(x, s);
  if (OP == MT_UN_ADDS) return __fadd_rn(x, s);
  if (OP == MT_UN_ABS) return fabsf(x);
  return 0.f;
}

constexpr int THREADS = 256;
constexpr int UNROLL = 4;  // independent 128-bit vectors in flight per thread

// ---- flat: out[i] = op(a[i*SA], b[i*SB]), SA/SB in {0,1}; vector body + scalar tail
template <int OP, int SA, int SB>
__global__ void __launch_bounds__(THREADS) bin_flat_kernel(float* __restrict__ out, const float* __restrict__ a,
                                                           const float* __restrict__ b, size_t n, float s) {
  const size_t n4 = n >> 2;
  const size_t tid = (size_t)blockIdx.x * THREADS + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * THREADS;
  const float a0 = SA ? 0.f : __ldg(a), b0 = SB ? 0.f : __ldg(b);
  size_t i = tid;
  for (; i + (UNROLL - 1) * stride < n4; i += UNROLL * stride) {
    float4 va[UNROLL], vb[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      if (SA) va[u] = mtd::ld_stream4(a + 4 * (i + u * stride));
      if (SB) vb[u] = mtd::ld_stream4(b + 4 * (i + u * stride));
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      float4 r;
      r.x = bin_op<OP>(SA ? va[u].x : a0, SB ? vb[u].x : b0, s);
      r.y = bin_op<OP>(SA ? va[u].y : a0, SB ? vb[u].y : b0, s);
      r.z = bin_op<OP>(SA ? va[u].z : a0, SB ? vb[u].z : b0, s);
      r.w = bin_op<OP>(SA ? va[u].w : a0, SB ? vb[u].w : b0, s);
      mtd::st4(out + 4 * (i + u * stride), r);
    }
  }
  for (; i < n4; i += stride) {
    float4 va = SA ? mtd::ld_stream4(a + 4 * i) : make_float4(a0, a0, a0, a0);
    float4 vb = SB ? mtd::ld_stream4(b + 4 * i) : make_float4(b0, b0, b0, b0);
    float4 r;
    r.x = bin_op<OP>(va.x, vb.x, s);
    r.y = bin_op<OP>(va.y, vb.y, s);
    r.z = bin_op<OP>(va.z, vb.z, s);
    r.w = bin_op<OP>(va.w, vb.w, s);
    mtd::st4(out + 4 * i, r);
  }
  for (size_t j = (n4 << 2) + tid; j < n; j += stride)
    out[j] = bin_op<OP>(SA ? a[j] : a0, SB ? b[j] : b0, s);
}

// scalar flavour for mis-aligned views
template <int OP, int SA, int SB>
__global__ void __launch_bounds__(THREADS) bin_flat_scalar_kernel(float* __restrict__ out, const float* __restrict__ a,
                                                                  const float* __restrict__ b, size_t n, float s) {
  const size_t stride = (size_t)gridDim.x * THREADS;
  for (size_t j = (size_t)blockIdx.x * THREADS + threadIdx.x; j < n; j += stride)
    out[j] = bin_op<OP>(a[SA ? j : 0], b[SB ? j : 0], s);
}

// ---- 2-D: out[r,c] = op(a[r*ar + c*AC], b[r*br + c*BC]); AC/BC in {0,1}; C % 4 == 0,
// row strides % 4 == 0 (alignment), one float4 of a row per thread iteration.
template <int OP, int AC, int BC>
__global__ void __launch_bounds__(THREADS) bin_2d_kernel(float* __restrict__ out, const float* __restrict__ a,
                                                         const float* __restrict__ b, long long R, long long C4,
                                                         long long ar, long long br, float s) {
  const long long total = R * C4;
  const long long stride = (long long)gridDim.x * THREADS;
  for (long long i = (long long)blockIdx.x * THREADS + threadIdx.x; i < total; i += stride) {
    const long long r = i / C4, c = (i - r * C4) << 2;
    float4 va, vb;
    if (AC) va = mtd::ld_stream4(a + r * ar + c);
    else { float t = __ldg(a + r * ar); va = make_float4(t, t, t, t); }
    if (BC) vb = __ldg(reinterpret_cast<const float4*>(b + r * br + c));
    else { float t = __ldg(b + r * br); vb = make_float4(t, t, t, t); }
    float4 o;
    o.x = bin_op<OP>(va.x, vb.x, s);
    o.y = bin_op<OP>(va.y, vb.y, s);
    o.z = bin_op<OP>(va.z, vb.z, s);
    o.w = bin_op<OP>(va.w, vb.w, s);
    mtd::st4(out + (i << 2), o);
  }
}

// ---- generic N-D fallback (any strides, any alignment)
struct NdDesc {
  int ndim;
  long long shape[MT_MAX_DIMS];
  long long sa[MT_MAX_DIMS];
  long long sb[MT_MAX_DIMS];
};

template <int OP>
__global__ void __launch_bounds__(THREADS) bin_nd_kernel(float* __restrict__ out, const float* __restrict__ a,
                                                         const float* __restrict__ b, const __grid_constant__ NdDesc d,
                                                         long long total, float s) {
  const long long stride = (long long)gridDim.x * THREADS;
  for (long long i = (long long)blockIdx.x * THREADS + threadIdx.x; i < total; i += stride) {
    long long rem = i, oa = 0, ob = 0;
#pragma unroll 1
    for (int k = d.ndim - 1; k >= 0; --k) {
      long long q = rem / d.shape[k], x = rem - q * d.shape[k];
      oa += x * d.sa[k];
      ob += x * d.sb[k];
      rem = q;
    }
    out[i] = bin_op<OP>(a[oa], b[ob], s);
  }
}

// in-place: dst[idx . sd] = op(dst[idx . sd], src[idx . ss])
template <int OP>
__global__ void __launch_bounds__(THREADS) inplace_nd_kernel(float* dst, const float* __restrict__ src,
                                                             const __grid_constant__ NdDesc d, long long total) {
  const long long stride = (long long)gridDim.x * THREADS;
  for (long long i = (long long)blockIdx.x * THREADS + threadIdx.x; i < total; i += stride) {
    long long rem = i, od = 0, os = 0;
#pragma unroll 1
    for (int k = d.ndim - 1; k >= 0; --k) {
      long long q = rem / d.shape[k], x = rem - q * d.shape[k];
      od += x * d.sa[k];
      os += x * d.sb[k];
      rem = q;
    }
    dst[od] = bin_op<OP>(dst[od], src[os], 0.f);
  }
}

template <int OP>
__global__ void __launch_bounds__(THREADS) unary_kernel(float* __restrict__ out, const float* __restrict__ x, size_t n,
                                                        float s, int vec_ok) {
  const size_t tid = (size_t)blockIdx.x * THREADS + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * THREADS;
  const size_t n4 = vec_ok ? (n >> 2) : 0;
  size_t i = tid;
  for (; i + (UNROLL - 1) * stride < n4; i += UNROLL * stride) {
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = mtd::ld_stream4(x + 4 * (i + u * stride));
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      float4 r;
      r.x = un_op<OP>(v[u].x, s);
      r.y = un_op<OP>(v[u].y, s);
      r.z = un_op<OP>(v[u].z, s);
      r.w = un_op<OP>(v[u].w, s);
      mtd::st4(out + 4 * (i + u * stride), r);
    }
  }
  for (; i < n4; i += stride) {
    float4 v = mtd::ld_stream4(x + 4 * i), r;
    r.x = un_op<OP>(v.x, s);
    r.y = un_op<OP>(v.y, s);
    r.z = un_op<OP>(v.z, s);
    r.w = un_op<OP>(v.w, s);
    mtd::st4(out + 4 * i, r);
  }
  for (size_t j = (n4 << 2) + tid; j < n; j += stride) out[j] = un_op<OP>(x[j], s);
}

// ---- where(cond, a, b): dense fast path (all three operands contiguous) and a strided general path
__global__ void __launch_bounds__(THREADS) where_flat_kernel(float* __restrict__ out, const float* __restrict__ c,
                                                             const float* __restrict__ a, const float* __restrict__ b,
                                                             size_t n, int vec_ok) {
  const size_t tid = (size_t)blockIdx.x * THREADS + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * THREADS;
  const size_t n4 = vec_ok ? (n >> 2) : 0;
  for (size_t i = tid; i < n4; i += stride) {
    const float4 vc = mtd::ld_stream4(c + 4 * i), va = mtd::ld_stream4(a + 4 * i), vb = mtd::ld_stream4(b + 4 * i);
    float4 r;
    r.x = vc.x != 0.f ? va.x : vb.x;
    r.y = vc.y != 0.f ? va.y : vb.y;
    r.z = vc.z != 0.f ? va.z : vb.z;
    r.w = vc.w != 0.f ? va.w : vb.w;
    mtd::st4(out + 4 * i, r);
  }
  for (size_t j = (n4 << 2) + tid; j < n; j += stride) out[j] = c[j] != 0.f ? a[j] : b[j];
}

struct Nd3Desc {
  int ndim;
  long long shape[MT_MAX_DIMS];
  long long sc[MT_MAX_DIMS], sa[MT_MAX_DIMS], sb[MT_MAX_DIMS];
};
__global__ void __launch_bounds__(THREADS) where_nd_kernel(float* __restrict__ out, const float* __restrict__ c,
                                                           const float* __restrict__ a, const float* __restrict__ b,
                                                           const __grid_constant__ Nd3Desc d, long long total) {
  const long long stride = (long long)gridDim.x * THREADS;
  for (long long i = (long long)blockIdx.x * THREADS + threadIdx.x; i < total; i += stride) {
    long long rem = i, oc = 0, oa = 0, ob = 0;
#pragma unroll 1
    for (int k = d.ndim - 1; k >= 0; --k) {
      long long q = rem / d.shape[k], x = rem - q * d.shape[k];
      oc += x * d.sc[k];
      oa += x * d.sa[k];
      ob += x * d.sb[k];
      rem = q;
    }
    out[i] = c[oc] != 0.f ? a[oa] : b[ob];
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// drop extent-1 dims and merge neighbours that are contiguous for BOTH operands
int collapse(int ndim, const int64_t* shape, const int64_t* sa, const int64_t* sb, long long* oshape, long long* osa,
             long long* osb) {
  int n = 0;
  for (int k = 0; k < ndim; ++k) {
    if (shape[k] == 1) continue;
    if (n > 0 && osa[n - 1] == sa[k] * shape[k] && osb[n - 1] == sb[k] * shape[k]) {
      oshape[n - 1] *= shape[k];
      osa[n - 1] = sa[k];
      osb[n - 1] = sb[k];
    } else {
      oshape[n] = shape[k];
      osa[n] = sa[k];
      osb[n] = sb[k];
      ++n;
    }
  }
  return n;
}

template <int OP>
int launch_binary(float* out, const float* a, const float* b, int ndim, const int64_t* shape, const int64_t* sa,
                  const int64_t* sb, float s) {
  long long cs[MT_MAX_DIMS], ca[MT_MAX_DIMS], cb[MT_MAX_DIMS];
  long long total = 1;
  for (int k = 0; k < ndim; ++k) total *= shape[k];
  if (total == 0) return MT_OK;
  int n = collapse(ndim, shape, sa, sb, cs, ca, cb);
  cudaStream_t st = g().stream;
  if (n == 0) {  // single element
    n = 1;
    cs[0] = 1;
    ca[0] = 0;
    cb[0] = 0;
  }
  const bool al = aligned16(out) && aligned16(a) && aligned16(b);
  if (n == 1 && (ca[0] == 0 || ca[0] == 1) && (cb[0] == 0 || cb[0] == 1)) {
    const size_t N = (size_t)cs[0];
    const int grid = ew_grid((int64_t)(N / 4 + 1), THREADS * UNROLL);
    const int key = (ca[0] ? 2 : 0) | (cb[0] ? 1 : 0);
    if (al) {
      switch (key) {
        case 3: bin_flat_kernel<OP, 1, 1><<<grid, THREADS, 0, st>>>(out, a, b, N, s); break;
        case 2: bin_flat_kernel<OP, 1, 0><<<grid, THREADS, 0, st>>>(out, a, b, N, s); break;
        case 1: bin_flat_kernel<OP, 0, 1><<<grid, THREADS, 0, st>>>(out, a, b, N, s); break;
        default: bin_flat_kernel<OP, 0, 0><<<grid, THREADS, 0, st>>>(out, a, b, N, s); break;
      }
    } else {
      const int g1 = ew_grid((int64_t)N, THREADS * UNROLL);
      switch (key) {
        case 3: bin_flat_scalar_kernel<OP, 1, 1><<<g1, THREADS, 0, st>>>(out, a, b, N, s); break;
        case 2: bin_flat_scalar_kernel<OP, 1, 0><<<g1, THREADS, 0, st>>>(out, a, b, N, s); break;
        case 1: bin_flat_scalar_kernel<OP, 0, 1><<<g1, THREADS, 0, st>>>(out, a, b, N, s); break;
        default: bin_flat_scalar_kernel<OP, 0, 0><<<g1, THREADS, 0, st>>>(out, a, b, N, s); break;
      }
    }
    MT_LAUNCHED();
    return MT_OK;
  }
  // row pitch must keep float4 loads aligned only for an operand that is read along the row (inner stride 1);
  // a column operand (inner stride 0: one scalar per row, e.g. shape (R,1)) may have any pitch
  if (n == 2 && al && (ca[1] == 0 || ca[1] == 1) && (cb[1] == 0 || cb[1] == 1) && cs[1] % 4 == 0 &&
      (ca[1] == 0 || ca[0] % 4 == 0) && (cb[1] == 0 || cb[0] % 4 == 0) && ca[0] >= 0 && cb[0] >= 0) {
    const long long R = cs[0], C4 = cs[1] / 4;
    const int grid = ew_grid(R * C4, THREADS * UNROLL);
    const int key = (ca[1] ? 2 : 0) | (cb[1] ? 1 : 0);
    switch (key) {
      case 3: bin_2d_kernel<OP, 1, 1><<<grid, THREADS, 0, st>>>(out, a, b, R, C4, ca[0], cb[0], s); break;
      case 2: bin_2d_kernel<OP, 1, 0><<<grid, THREADS, 0, st>>>(out, a, b, R, C4, ca[0], cb[0], s); break;
      case 1: bin_2d_kernel<OP, 0, 1><<<grid, THREADS, 0, st>>>(out, a, b, R, C4, ca[0], cb[0], s); break;
      default: bin_2d_kernel<OP, 0, 0><<<grid, THREADS, 0, st>>>(out, a, b, R, C4, ca[0], cb[0], s); break;
    }
    MT_LAUNCHED();
    return MT_OK;
  }
  NdDesc d;
  d.ndim = n;
  for (int k = 0; k < n; ++k) {
    d.shape[k] = cs[k];
    d.sa[k] = ca[k];
    d.sb[k] = cb[k];
  }
  bin_nd_kernel<OP><<<ew_grid(total, THREADS * UNROLL), THREADS, 0, st>>>(out, a, b, d, total, s);
  MT_LAUNCHED();
  return MT_OK;
}

}  // namespace

MT_API int mt_ew_binary_f32(int op, float* out, const float* a, const float* b, int ndim, const int64_t* shape,
                            const int64_t* a_strides, const int64_t* b_strides, float scalar) {
  MT_REQUIRE_INIT();
  MT_CHECK_ARG(ndim >= 0 && ndim <= MT_MAX_DIMS, "mt_ew_binary_f32: ndim %d out of range", ndim);
  MT_CHECK_ARG(out && a && b, "mt_ew_binary_f32: null pointer");
  switch (op) {
    case MT_BIN_ADD: return launch_binary<MT_BIN_ADD>(out, a, b, ndim, shape, a_strides, b_strides, scalar);
    case MT_BIN_MUL: return launch_binary<MT_BIN_MUL>(out, a, b, ndim, shape, a_strides, b_strides, scalar);
    case MT_BIN_SUB: return launch_binary<MT_BIN_SUB>(out, a, b, ndim, shape, a_strides, b_strides, scalar);
    case MT_BIN_DIV: return launch_binary<MT_BIN_DIV>(out, a, b, ndim, shape, a_strides, b_strides, scalar);
    case MT_BIN_TANH_BWD: return launch_binary<MT_BIN_TANH_BWD>(out, a, b, ndim, shape, a_strides, b_strides, scalar);
    case MT_BIN_POW_BWD: return launch_binary<MT_BIN_POW_BWD>(out, a, b, ndim, shape, a_strides, b_strides, scalar);
    case MT_BIN_EXP_BWD: return launch_binary<MT_BIN_EXP_BWD>(out, a, b, ndim, shape, a_strides, b_strides, scalar);
    case MT_BIN_LT: return launch_binary<MT_BIN_LT>(out, a, b, ndim, shape, a_strides, b_strides, scalar);
    case MT_BIN_GT: return launch_binary<MT_BIN_GT>(out, a, b, ndim, shape, a_strides, b_strides, scalar);
    case MT_BIN_LE: return launch_binary<MT_BIN_LE>(out, a, b, ndim, shape, a_strides, b_strides, scalar);
    case MT_BIN_GE: return launch_binary<MT_BIN_GE>(out, a, b, ndim, shape, a_strides, b_strides, scalar);
    case MT_BIN_EQ: return launch_binary<MT_BIN_EQ>(out, a, b, ndim, shape, a_strides, b_strides, scalar);
    case MT_BIN_NE: return launch_binary<MT_BIN_NE>(out, a, b, ndim, shape, a_strides, b_strides, scalar);
    case MT_BIN_MAX: return launch_binary<MT_BIN_MAX>(out, a, b, ndim, shape, a_strides, b_strides, scalar);
    case MT_BIN_MIN: return launch_binary<MT_BIN_MIN>(out, a, b, ndim, shape, a_strides, b_strides, scalar);
  }
  return fail(MT_ERR_INVALID, "mt_ew_binary_f32: unknown op %d", op);
}

MT_API int mt_ew_inplace_f32(int op, float* dst, const int64_t* dst_strides, const float* src,
                             const int64_t* src_strides, int ndim, const int64_t* shape) {
  MT_REQUIRE_INIT();
  MT_CHECK_ARG(ndim >= 0 && ndim <= MT_MAX_DIMS, "mt_ew_inplace_f32: ndim %d out of range", ndim);
  MT_CHECK_ARG(op == MT_BIN_ADD || op == MT_BIN_MUL || op == MT_BIN_DIV || op == MT_BIN_SUB,
               "mt_ew_inplace_f32: op %d is not an in-place op", op);
  long long total = 1;
  bool dense = true;
  long long expect = 1;
  for (int k = ndim - 1; k >= 0; --k) {
    total *= shape[k];
    if (shape[k] != 1 && dst_strides[k] != expect) dense = false;
    expect *= shape[k];
  }
  if (total == 0) return MT_OK;
  if (dense) return mt_ew_binary_f32(op, dst, dst, src, ndim, shape, dst_strides, src_strides, 0.f);
  NdDesc d;
  d.ndim = ndim;
  for (int k = 0; k < ndim; ++k) {
    d.shape[k] = shape[k];
    d.sa[k] = dst_strides[k];
    d.sb[k] = src_strides[k];
  }
  const int grid = ew_grid(total, THREADS * UNROLL);
  switch (op) {
    case MT_BIN_ADD: inplace_nd_kernel<MT_BIN_ADD><<<grid, THREADS, 0, g().stream>>>(dst, src, d, total); break;
    case MT_BIN_MUL: inplace_nd_kernel<MT_BIN_MUL><<<grid, THREADS, 0, g().stream>>>(dst, src, d, total); break;
    case MT_BIN_SUB: inplace_nd_kernel<MT_BIN_SUB><<<grid, THREADS, 0, g().stream>>>(dst, src, d, total); break;
    default: inplace_nd_kernel<MT_BIN_DIV><<<grid, THREADS, 0, g().stream>>>(dst, src, d, total); break;
  }
  MT_LAUNCHED();
  return MT_OK;
}

MT_API int mt_ew_where_f32(float* out, const float* cond, const float* a, const float* b, int ndim, const int64_t* shape,
                           const int64_t* c_strides, const int64_t* a_strides, const int64_t* b_strides) {
  MT_REQUIRE_INIT();
  MT_CHECK_ARG(ndim >= 0 && ndim <= MT_MAX_DIMS, "mt_ew_where_f32: ndim %d out of range", ndim);
  MT_CHECK_ARG(out && cond && a && b, "mt_ew_where_f32: null pointer");
  long long total = 1, expect = 1;
  bool dense = true;
  for (int k = ndim - 1; k >= 0; --k) {
    total *= shape[k];
    if (shape[k] != 1 && (c_strides[k] != expect || a_strides[k] != expect || b_strides[k] != expect)) dense = false;
    expect *= shape[k];
  }
  if (total == 0) return MT_OK;
  if (dense) {
    const int vec = aligned16(out) && aligned16(cond) && aligned16(a) && aligned16(b);
    where_flat_kernel<<<ew_grid(total / 4 + 1, THREADS * UNROLL), THREADS, 0, g().stream>>>(out, cond, a, b, (size_t)total, vec);
  } else {
    Nd3Desc d;
    d.ndim = ndim;
    for (int k = 0; k < ndim; ++k) {
      d.shape[k] = shape[k];
      d.sc[k] = c_strides[k];
      d.sa[k] = a_strides[k];
      d.sb[k] = b_strides[k];
    }
    where_nd_kernel<<<ew_grid(total, THREADS * UNROLL), THREADS, 0, g().stream>>>(out, cond, a, b, d, total);
  }
  MT_LAUNCHED();
  return MT_OK;
}

MT_API int mt_ew_unary_f32(int op, float* out, const float* x, size_t n, float scalar) {
  MT_REQUIRE_INIT();
  if (n == 0) return MT_OK;
  MT_CHECK_ARG(out && x, "mt_ew_unary_f32: null pointer");
  const int vec = aligned16(out) && aligned16(x);
  const int grid = ew_grid((int64_t)(n / 4 + 1), THREADS * UNROLL);
  cudaStream_t st = g().stream;
  switch (op) {
    case MT_UN_NEG: unary_kernel<MT_UN_NEG><<<grid, THREADS, 0, st>>>(out, x, n, scalar, vec); break;
    case MT_UN_LOG: unary_kernel<MT_UN_LOG><<<grid, THREADS, 0, st>>>(out, x, n, scalar, vec); break;
    case MT_UN_TANH: unary_kernel<MT_UN_TANH><<<grid, THREADS, 0, st>>>(out, x, n, scalar, vec); break;
    case MT_UN_EXP: unary_kernel<MT_UN_EXP><<<grid, THREADS, 0, st>>>(out, x, n, scalar, vec); break;
    case MT_UN_POW: unary_kernel<MT_UN_POW><<<grid, THREADS, 0, st>>>(out, x, n, scalar, vec); break;
    case MT_UN_SCALE: unary_kernel<MT_UN_SCALE><<<grid, THREADS, 0, st>>>(out, x, n, scalar, vec); break;
    case MT_UN_ADDS: unary_kernel<MT_UN_ADDS><<<grid, THREADS, 0, st>>>(out, x, n, scalar, vec); break;
    case MT_UN_ABS: unary_kernel<MT_UN_ABS><<<grid, THREADS, 0, st>>>(out, x, n, scalar, vec); break;
    default: return fail(MT_ERR_INVALID, "mt_ew_unary_f32: unknown op %d", op);
  }
  MT_LAUNCHED();
  return MT_OK;
}
