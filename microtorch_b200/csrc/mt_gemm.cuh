// mt_gemm.cuh - shared argument block of the two GEMM engines (tcgen05 and SIMT).
#pragma once
#include "mt_common.cuh"

namespace mt {

// C[m,n] (+)= act( sum_k A'(m,k) * B(k,n) + bias[n] ),  C row-major [M,N] (ld = N)
//   A(m,k)  = A[m*a_sm + k*a_sk]            (+ b*a_sb for the batched SIMT path)
//   B(k,n)  = B[k*b_sk + n*b_sn]
//   A'(m,k) = A(m,k) * (1 - T(m,k)^2) when `tanh_src` != NULL, T addressed like A
//             (tanh' fused into the operand prologue; tcgen05 engine only)
struct GemmArgs {
  float* C;
  const float* A;
  const float* B;
  long long M, N, K;
  long long a_sb, a_sm, a_sk, b_sb, b_sk, b_sn;
  const float* bias;      // [N] or NULL
  const float* tanh_src;  // same layout as A, or NULL
  const float* epi_tanh;  // [M,N] like C, or NULL: C = result * (1 - epi_tanh^2)  (emit the producer's pre-activation
                          // gradient straight from the dX epilogue: tanh' fused into the PRODUCER instead of a prologue)
  int act;                // mt_act
  int accumulate;         // C += result
  int forward;            // a Linear forward product (its precision can be set apart: mt_gemm_tc_set_fwd)
  // SIMT split-K bookkeeping (filled by gemm_simt)
  int splits;
  long long kchunk;
};

int gemm_simt(GemmArgs a, long long batch);
// C = act(sum_s partial[s] + bias) (+ C): deterministic fold of split-K partials stacked as [splits, M, N]
int splitk_finalize(float* C, const float* partial, int splits, long long M, long long N, const float* bias, int act,
                    int accumulate, const float* epi_tanh);
// HBM-bound kernels for tiny output widths (N <= 16); MT_ERR_UNSUPPORTED when the shape is not one of theirs
int gemm_skinny_try(const GemmArgs& a);
// MT_OK if the tcgen05 engine ran it, MT_ERR_UNSUPPORTED if the layout/shape is not one it
// takes (caller falls back to SIMT), anything else is a real error.
int gemm_tc_try(const GemmArgs& a);
// the mid-size tcgen05 kernel (mt_gemm_tc_mid.cu: one 128 x 128 tile x K slice per CTA, slices folded through DSMEM):
// `wants` = the dispatcher's size window (or engine 2 forced), `try` = run it (MT_ERR_UNSUPPORTED: layout not taken)
bool gemm_tc_mid_wants(const GemmArgs& a);
int gemm_tc_mid_try(const GemmArgs& a);

}  // namespace mt
