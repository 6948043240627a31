// mt_loss_sgd.cu - fused loss value + dL/dpred kernels and the fused multi-tensor SGD update.
//
// Loss (reference: src/nn/loss.py:63-100 + Reduce.mean, src/tensor/ops/reduce.py:40-43):
//   one pass reads pred and target, writes dpred and one partial per CTA; the last CTA to
//   finish (threadfence + ticket) sums the partials in a fixed order.  12 B/element instead
//   of the 6 (MSE) / 7 (CE) / 13 (BCE) forward launches plus as many backward visits of the
//   reference graph.  HBM-bound.
// SGD (reference: src/nn/optimizer.py:37-39 -> src/tensor/tensor.py:564-571):
//   p -= lr*g for every parameter tensor in ONE launch (table of pointers passed by value),
//   optionally zeroing g in the same pass (Module.zero_grad, src/nn/module.py:83-89):
//   12 B/element (16 with the zeroing).  HBM-bound.
#include "mt_common.cuh"
#include "mt_loss.cuh"

using namespace mt;

namespace {

constexpr int THREADS = 256;

using mtd::loss_elem;

template <int KIND, bool LEAN>
__device__ __forceinline__ void loss_vec4(const float4 p, const float4 t, float scale, float4& g, float& a0, float& a1,
                                          float& a2, float& a3) {
  float v;
  loss_elem<KIND, LEAN>(p.x, t.x, scale, v, g.x); a0 += v;
  loss_elem<KIND, LEAN>(p.y, t.y, scale, v, g.y); a1 += v;
  loss_elem<KIND, LEAN>(p.z, t.z, scale, v, g.z); a2 += v;
  loss_elem<KIND, LEAN>(p.w, t.w, scale, v, g.w); a3 += v;
}
__device__ __forceinline__ bool lean4(const float4 p) {
  return mtd::bce_lean_ok(p.x) && mtd::bce_lean_ok(p.y) && mtd::bce_lean_ok(p.z) && mtd::bce_lean_ok(p.w);
}

template <int KIND, bool WRITE_GRAD>
__global__ void __launch_bounds__(THREADS) loss_kernel(float* __restrict__ loss_out, float* __restrict__ dpred,
                                                       const float* __restrict__ pred, const float* __restrict__ target,
                                                       size_t n, float scale, float* __restrict__ partials,
                                                       unsigned int* __restrict__ ticket, int vec) {
  __shared__ float smem[THREADS / 32];
  __shared__ bool is_last;
  const size_t tid = (size_t)blockIdx.x * THREADS + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * THREADS;
  const size_t n4 = vec ? (n >> 2) : 0;   // a mis-aligned view (an odd-offset slice) takes the scalar loop for everything
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
  size_t i = tid;
  for (; i + stride < n4; i += 2 * stride) {  // two independent vector pairs in flight
    float4 p0 = mtd::ld_stream4(pred + 4 * i), t0 = mtd::ld_stream4(target + 4 * i);
    float4 p1 = mtd::ld_stream4(pred + 4 * (i + stride)), t1 = mtd::ld_stream4(target + 4 * (i + stride));
    float4 g0, g1;
    if (KIND == MT_LOSS_BCE && lean4(p0) && lean4(p1)) {
      loss_vec4<KIND, true>(p0, t0, scale, g0, acc0, acc1, acc2, acc3);
      loss_vec4<KIND, true>(p1, t1, scale, g1, acc0, acc1, acc2, acc3);
    } else {
      loss_vec4<KIND, false>(p0, t0, scale, g0, acc0, acc1, acc2, acc3);
      loss_vec4<KIND, false>(p1, t1, scale, g1, acc0, acc1, acc2, acc3);
    }
    if (WRITE_GRAD) {
      mtd::st4(dpred + 4 * i, g0);
      mtd::st4(dpred + 4 * (i + stride), g1);
    }
  }
  for (; i < n4; i += stride) {
    float4 p0 = mtd::ld_stream4(pred + 4 * i), t0 = mtd::ld_stream4(target + 4 * i), g0;
    loss_vec4<KIND, false>(p0, t0, scale, g0, acc0, acc1, acc2, acc3);
    if (WRITE_GRAD) mtd::st4(dpred + 4 * i, g0);
  }
  for (size_t j = (n4 << 2) + tid; j < n; j += stride) {
    float v, gj;
    loss_elem<KIND>(pred[j], target[j], scale, v, gj);
    acc0 += v;
    if (WRITE_GRAD) dpred[j] = gj;
  }
  const float block_total = mtd::block_sum<THREADS>((acc0 + acc1) + (acc2 + acc3), smem);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = block_total;
    __threadfence();
    const unsigned int done = atomicAdd(ticket, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {  // fixed-order final sum over the per-CTA partials
    __threadfence();
    float s = 0.f;
    for (unsigned int k = threadIdx.x; k < gridDim.x; k += THREADS) s += __ldcg(partials + k);
    s = mtd::block_sum<THREADS>(s, smem);
    if (threadIdx.x == 0) {
      loss_out[0] = __fmul_rn(s, scale);   // sum * float32(1/N)   (mean = sum * count**-1)
      *ticket = 0u;                        // re-arm for the next launch
    }
  }
}

struct LossScratch {
  float* partials = nullptr;
  unsigned int* ticket = nullptr;
  int cap = 0;
};
LossScratch& scratch() {
  static LossScratch s;
  return s;
}

// ------------------------------------------------------------------ multi-tensor SGD
struct SgdTable {
  enum { CAP = 64 };
  float* p[CAP];
  float* g[CAP];
  unsigned long long start4[CAP + 1];  // prefix of float4 counts
  unsigned long long n[CAP];           // element counts
  unsigned char vec[CAP];              // both pointers 16 B aligned -> 128-bit accesses
};

template <bool ZERO>
__global__ void __launch_bounds__(THREADS) sgd_multi_kernel(const __grid_constant__ SgdTable t, int count, float lr) {
  const unsigned long long total4 = t.start4[count];
  const unsigned long long stride = (unsigned long long)gridDim.x * THREADS;
  int cur = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * THREADS + threadIdx.x; i < total4; i += stride) {
    while (i >= t.start4[cur + 1]) ++cur;
    const unsigned long long k = i - t.start4[cur];
    float* __restrict__ pp = t.p[cur];
    float* __restrict__ gg = t.g[cur];
    const unsigned long long n = t.n[cur];
    if (4 * k + 3 < n && t.vec[cur]) {
      float4 pv = *reinterpret_cast<float4*>(pp + 4 * k);
      float4 gv = *reinterpret_cast<const float4*>(gg + 4 * k);
      // p + (-(lr*g)): two roundings, exactly as NumPy does it (no FMA contraction)
      pv.x = __fsub_rn(pv.x, __fmul_rn(lr, gv.x));
      pv.y = __fsub_rn(pv.y, __fmul_rn(lr, gv.y));
      pv.z = __fsub_rn(pv.z, __fmul_rn(lr, gv.z));
      pv.w = __fsub_rn(pv.w, __fmul_rn(lr, gv.w));
      *reinterpret_cast<float4*>(pp + 4 * k) = pv;
      if (ZERO) *reinterpret_cast<float4*>(gg + 4 * k) = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      const unsigned long long je = (4 * k + 4 < n) ? 4 * k + 4 : n;   // ragged tail or a mis-aligned view
      for (unsigned long long j = 4 * k; j < je; ++j) {
        pp[j] = __fsub_rn(pp[j], __fmul_rn(lr, gg[j]));
        if (ZERO) gg[j] = 0.f;
      }
    }
  }
}

}  // namespace

MT_API int mt_loss_fwd_bwd_f32(int kind, float* loss_out, float* dpred, const float* pred, const float* target,
                               size_t n, float scale) {
  MT_REQUIRE_INIT();
  MT_CHECK_ARG(loss_out && pred && target, "mt_loss_fwd_bwd_f32: null pointer");
  MT_CHECK_ARG(n > 0, "mt_loss_fwd_bwd_f32: empty input");
  MT_CHECK_ARG(((uintptr_t)pred & 3) == 0 && ((uintptr_t)target & 3) == 0 && ((uintptr_t)dpred & 3) == 0,
               "mt_loss_fwd_bwd_f32: operands must be float aligned");
  // basic-slice views (X[i:i+b]) start anywhere: 128-bit accesses only when all three operands allow them
  const int vec = (((uintptr_t)pred | (uintptr_t)target | (uintptr_t)dpred) & 15) == 0;
  LossScratch& S = scratch();
  const int grid = ew_grid((int64_t)(n / 4 + 1), THREADS * 2);
  if (S.cap < grid) {
    if (S.partials) {
      pool_free(S.partials);
      pool_free(S.ticket);
    }
    const int cap = g().sm_count * 16;
    int rc = pool_alloc((void**)&S.partials, sizeof(float) * cap);
    if (rc) return rc;
    rc = pool_alloc((void**)&S.ticket, 512);
    if (rc) return rc;
    MT_CUDA(cudaMemsetAsync(S.ticket, 0, 512, g().stream));
    S.cap = cap;
  }
  cudaStream_t st = g().stream;
#define MT_LOSS(K)                                                                                              \
  do {                                                                                                          \
    if (dpred) loss_kernel<K, true><<<grid, THREADS, 0, st>>>(loss_out, dpred, pred, target, n, scale, S.partials, S.ticket, vec); \
    else loss_kernel<K, false><<<grid, THREADS, 0, st>>>(loss_out, dpred, pred, target, n, scale, S.partials, S.ticket, vec); \
  } while (0)
  switch (kind) {
    case MT_LOSS_MSE: MT_LOSS(MT_LOSS_MSE); break;
    case MT_LOSS_CE_REF: MT_LOSS(MT_LOSS_CE_REF); break;
    case MT_LOSS_BCE: MT_LOSS(MT_LOSS_BCE); break;
    default: return fail(MT_ERR_INVALID, "mt_loss_fwd_bwd_f32: unknown loss kind %d", kind);
  }
#undef MT_LOSS
  MT_LAUNCHED();
  return MT_OK;
}

MT_API int mt_sgd_multi_f32(float* const* params, float* const* grads, const size_t* sizes, int count, float lr,
                            int zero_grads) {
  MT_REQUIRE_INIT();
  MT_CHECK_ARG(count >= 0, "mt_sgd_multi_f32: negative count");
  int done = 0;
  while (done < count) {
    SgdTable t;
    int k = 0;
    unsigned long long acc = 0;
    while (done < count && k < SgdTable::CAP) {
      if (sizes[done]) {
        MT_CHECK_ARG(params[done] && grads[done], "mt_sgd_multi_f32: null tensor %d", done);
        MT_CHECK_ARG(((uintptr_t)params[done] & 3) == 0 && ((uintptr_t)grads[done] & 3) == 0,
                     "mt_sgd_multi_f32: tensor %d not float aligned", done);
        t.vec[k] = (((uintptr_t)params[done] | (uintptr_t)grads[done]) & 15) == 0;
        t.p[k] = params[done];
        t.g[k] = grads[done];
        t.n[k] = sizes[done];
        t.start4[k] = acc;
        acc += (sizes[done] + 3) / 4;
        ++k;
      }
      ++done;
    }
    t.start4[k] = acc;
    if (!k) continue;
    const int grid = ew_grid((int64_t)acc, THREADS * 4);
    if (zero_grads) sgd_multi_kernel<true><<<grid, THREADS, 0, g().stream>>>(t, k, lr);
    else sgd_multi_kernel<false><<<grid, THREADS, 0, g().stream>>>(t, k, lr);
    MT_LAUNCHED();
  }
  return MT_OK;
}
