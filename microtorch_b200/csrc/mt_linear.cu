// mt_linear.cu - Linear forward / backward entry points on top of the two GEMM engines.
//
// Forward  (src/nn/layer/linear.py:48-62 + src/nn/activation.py:11-12):
//     Y = act(X . W^T + b)          one GEMM, bias add and tanh in its epilogue
// Backward (src/tensor/ops/overload.py:137-154, src/tensor/ops/math.py:74):
//     dZ = dY * (1 - Y^2)           fused into the operand prologue of both GEMMs (tcgen05 engine)
//     dW (+)= dZ^T . X              K = batch rows: the "un-broadcast" reduction of the 3-D case is
//                                   just this long-K GEMM over the flattened (B*S) rows
//     dX  = dZ . W
// No bias gradient exists in the reference (bias is added in place outside the tape, Q1).
#include "mt_common.cuh"
#include "mt_gemm.cuh"

using namespace mt;

namespace {

// run on tcgen05 when the engine takes the shape, otherwise on the SIMT kernel (needs dZ explicit)
int run(GemmArgs a, int* engine_min) {
  if (g().gemm_engine_force != 0) {
    int rc = gemm_tc_try(a);
    if (rc == MT_OK) return MT_OK;
    if (rc != MT_ERR_UNSUPPORTED) return rc;
    if (g().gemm_engine_force >= 1)
      return fail(MT_ERR_UNSUPPORTED, "tcgen05 engine forced but shape M=%lld N=%lld K=%lld is not supported by it",
                  a.M, a.N, a.K);
  }
  *engine_min = 0;
  if (a.tanh_src) return fail(MT_ERR_INVALID, "internal: SIMT engine needs an explicit dZ");
  return gemm_simt(a, 1);
}

}  // namespace

MT_API int mt_linear_fwd_f32(float* Y, const float* X, const float* W, const float* bias, int64_t M, int64_t N,
                             int64_t K, int act) {
  MT_REQUIRE_INIT();
  MT_CHECK_ARG(Y && X && W, "mt_linear_fwd_f32: null pointer");
  MT_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "mt_linear_fwd_f32: negative extent");
  MT_CHECK_ARG(act == MT_ACT_NONE || act == MT_ACT_TANH, "mt_linear_fwd_f32: unknown activation %d", act);
  if (M == 0 || N == 0) return MT_OK;
  GemmArgs a{};
  a.C = Y; a.A = X; a.B = W; a.M = M; a.N = N; a.K = K;
  a.a_sm = K; a.a_sk = 1;     // X row-major [M,K]
  a.b_sk = 1; a.b_sn = K;     // B(k,n) = W[n,k]
  a.bias = bias; a.act = act;
  a.forward = 1;
  int engine = 1;
  int rc = run(a, &engine);
  g().gemm_last_engine = engine;
  return rc;
}

// Above this many dY elements the tanh' prologue is NOT fused into the two GEMMs: the Y tile costs a third of every
// raw stage (3 instead of 5 TMA stages in flight) and measured 10.2 vs 7.9 ms per 65536x4096x4096 GEMM, while one
// explicit dZ pass is 0.55 ms (profiles/r01_gemm_final_matrix.jsonl).  Small problems keep the fused form (one launch less).
static long long g_prologue_fuse_max = 1ll << 22;
extern "C" __attribute__((visibility("default"))) int mt_linear_set_prologue_fuse_max(long long elems) {
  g_prologue_fuse_max = elems;
  return MT_OK;
}

MT_API int mt_linear_bwd_f32(float* dX, float* dW, const float* dY, const float* Yact, const float* X, const float* W,
                             int64_t M, int64_t N, int64_t K, int flags) {
  const int accumulate_dW = flags & MT_LINEAR_ACCUMULATE_DW;
  const bool dx_preact = (flags & MT_LINEAR_DX_TIMES_TANH_GRAD_OF_X) != 0;
  MT_REQUIRE_INIT();
  MT_CHECK_ARG(dW && dY && X && W, "mt_linear_bwd_f32: null pointer");
  MT_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "mt_linear_bwd_f32: negative extent");
  if (N == 0 || K == 0) return MT_OK;
  int engine = 1;
  const float* dZ = dY;          // explicit dZ, materialised only if an engine needs it
  PoolScratch dz_scratch;
  float* dZ_tmp = nullptr;
  auto need_explicit = [&]() -> int {
    if (!Yact || dZ_tmp) return MT_OK;
    int rc = dz_scratch.alloc(sizeof(float) * (size_t)M * N);
    if (rc) return rc;
    dZ_tmp = dz_scratch.f32();
    const int64_t shape[1] = {M * N}, st[1] = {1};
    rc = mt_ew_binary_f32(MT_BIN_TANH_BWD, dZ_tmp, dY, Yact, 1, shape, st, st, 0.f);
    dZ = dZ_tmp;
    return rc;
  };
  if (Yact && (long long)M * N > g_prologue_fuse_max) {   // large: one explicit dZ pass, then plain GEMMs
    int rc0 = need_explicit();
    if (rc0) return rc0;
  }
  auto attempt = [&](GemmArgs a) -> int {
    // first with the fused prologue on tcgen05, then (if refused) SIMT on an explicit dZ
    if (g().gemm_engine_force != 0) {
      if (Yact && !dZ_tmp && gemm_tc_mid_wants(a)) {   // the mid-size kernel has no prologue flavour: one small dZ pass
        int rc0 = need_explicit();
        if (rc0) return rc0;
      }
      GemmArgs t = a;
      t.A = dZ_tmp ? dZ_tmp : dY;
      t.tanh_src = dZ_tmp ? nullptr : Yact;
      int rc = gemm_tc_try(t);
      if (rc == MT_OK) return MT_OK;
      if (rc != MT_ERR_UNSUPPORTED) return rc;
      if (g().gemm_engine_force >= 1)
        return fail(MT_ERR_UNSUPPORTED, "tcgen05 engine forced but backward shape M=%lld N=%lld K=%lld is not supported",
                    a.M, a.N, a.K);
    }
    engine = 0;
    int rc = need_explicit();
    if (rc) return rc;
    a.A = dZ;
    a.tanh_src = nullptr;
    return gemm_simt(a, 1);
  };
  int rc = MT_OK;
  if (M == 0) {
    if (!accumulate_dW) rc = mt_memset_zero(dW, sizeof(float) * (size_t)N * K);
    return rc;
  }
  {  // dW[N,K] (+)= dZ^T . X      GEMM  M'=N, N'=K, K'=M
    GemmArgs a{};
    a.C = dW; a.B = X; a.M = N; a.N = K; a.K = M;
    a.a_sm = 1; a.a_sk = N;     // A(m'=n, k'=row) = dY[row, n]
    a.b_sk = K; a.b_sn = 1;     // B(k'=row, n'=k) = X[row, k]
    a.accumulate = accumulate_dW;
    rc = attempt(a);
    if (rc == MT_OK) {   // dW is complete here: a gradient exchange of it need not wait for the dX GEMM below
      Global& G = g();
      if (cudaEventRecord(G.ev_dw_ready, G.stream) == cudaSuccess) G.dw_ready_ptr = dW;
      else { cudaGetLastError(); G.dw_ready_ptr = nullptr; }
    }
  }
  if (rc == MT_OK && dX) {  // dX[M,K] = dZ . W    GEMM  M'=M, N'=K, K'=N
    GemmArgs a{};
    a.C = dX; a.B = W; a.M = M; a.N = K; a.K = N;
    a.a_sm = N; a.a_sk = 1;     // A(m, k'=n) = dY[m, n]
    a.b_sk = K; a.b_sn = 1;     // B(k'=n, n'=k) = W[n, k]
    if (dx_preact) a.epi_tanh = X;   // X is the tanh output of the producing layer: emit ITS dZ directly
    rc = attempt(a);
  }
  g().gemm_last_engine = engine;
  return rc;
}
