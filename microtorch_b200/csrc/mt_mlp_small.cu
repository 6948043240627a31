// mt_mlp_small.cu - the whole training loop of a small Linear(+Tanh) stack in ONE kernel launch.
//
// Reference path: the user loop of demo.ipynb:5178-5189 -
//     for k in range(steps): model.zero_grad(); pred = model(X); loss = loss_fn(pred, T); loss.backward(); opt.step()
// with full-batch X / T that never change (src/nn/sequential.py:37-48, src/nn/layer/linear.py:48-62,
// src/nn/loss.py:35-100, src/nn/optimizer.py:37-39).  At the notebook's sizes (200 samples, 2->64->32->16->1) one step
// is ~1.6 MFLOP: every launch-per-op or even graph-per-step scheme is pure launch/dependency latency (SURVEY.md
// section 8f item 1).  Here a thread-block CLUSTER (up to 8 CTAs on 8 SMs) keeps weights, activations and gradients
// in shared memory and iterates the steps on chip; HBM sees the inputs once, one loss value per step, and the final
// weights / last gradients.  The batch rows are split over the CTAs (data parallel inside the cluster), every CTA
// holds a full replica of the weights, and the only exchange per step is the sum of the per-CTA partial dW (and
// partial loss) read from the peers' shared memory (DSMEM) after ONE cluster barrier.
//
// Per step (phases separated by __syncthreads, rows = this CTA's share of the batch):
//   forward   A_i = tanh(A_{i-1} W_i^T + b_i)                         i = 1..L
//   loss      partial value (block sum) and G_L = dL/dZ_L (pred map a*y+b and tanh' folded in)
//   backward  G_{i-1} = (G_i W_i) * (1 - A_{i-1}^2)  i = L..2, then every layer's partial dW_i = G_i^T A_{i-1}
//             -> scratch in one barrier-free phase
//   exchange  barrier.cluster; reduce-scatter (unit u summed by rank u % C in rank order) ; barrier.cluster;
//             all-gather + W_i = W_i - lr * dW_i (two roundings, like NumPy) - every replica applies the same bits.
// Biases receive no gradient in the reference and stay constant.
//
// The three matrix products of a layer run through one register-tiled routine (block_gemm): TMxTN outputs per
// work item, the reduction split over R lanes of a warp (shuffle fold) so that tiny layers still fill the CTA.
// All shared-memory pitches are odd, which keeps every access pattern of the three products conflict-free or
// broadcast.  Deterministic: fixed work decomposition, no atomics.
#include <algorithm>

#include "mt_common.cuh"
#include "mt_loss.cuh"

using namespace mt;

namespace {

constexpr int THREADS = 1024;
constexpr int MAXL = MT_MLP_MAX_LAYERS;
constexpr int SMEM_MAX = 227 * 1024 - 256;   // dynamic part: the CTA limit minus this kernel's static shared memory

struct Phase {            // one matrix product, planned on the host so the kernel does no integer division:
  int tile;               //   tile shape code (0: 4x4, 1: 2x2, 2: 1x1)
  int log2R;              //   lanes per work item = 1 << log2R (they split the reduction)
  int MT, NT;             //   tiles along m and n
  unsigned nt_magic;      //   ceil(2^32 / NT): item / NT == umulhi(item, nt_magic) while item * NT < 2^32 (NT > 1)
  int passes;             //   ceil(MT * NT / (threads >> log2R))
};

struct Params {
  int L, B, steps, loss_kind, has_map, timing;
  int rows_per;               // batch rows per CTA of the cluster
  float pa, pb, lr, scale;
  int d[MAXL + 1], act[MAXL];
  int pA[MAXL + 1], pW[MAXL];
  int offA[MAXL + 1], offG[MAXL + 1], offW[MAXL], offB[MAXL], offT;   // offG[i]: dL/dZ_i (i >= 1), same pitch as A_i
  int offS, offP, offSW[MAXL], offSL; // scratch S (partial dW of every layer, then the partial loss) and P (summed slices)
  Phase fwd[MAXL], dw[MAXL], dx[MAXL];
  float* W[MAXL];
  const float* bias[MAXL];
  float* gradW[MAXL];
  const float* X;
  const float* T;
  float* losses;
};

// phase clocks of rank 0 (debug / tuning: mt_mlp_debug_clocks): fwd, loss, dX chain, dW, cluster barrier, update
__device__ unsigned long long g_mlp_clk[8];
#define MLP_TICK(slot)                                            \
  do {                                                            \
    if (p.timing && rank == 0 && tid == 0) {                      \
      const long long now = clock64();                            \
      g_mlp_clk[slot] += (unsigned long long)(now - t_prev);      \
      t_prev = now;                                               \
    }                                                             \
  } while (0)

__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned cluster_size() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// read one float of CTA `rank`'s shared memory at the address that `local` has in this CTA.  The asm is volatile
// (stays behind the cluster barrier, itself a volatile asm) but carries no memory clobber, so several of these
// loads can be in flight at once.
__device__ __forceinline__ float ld_peer(const float* local, unsigned rank) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(local);
  unsigned ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra));
  return v;
}
// sum over the cluster's ranks, rank order, all loads issued before the first add
__device__ __forceinline__ float sum_peers(const float* local, unsigned csize) {
  float v[8];
#pragma unroll
  for (unsigned c = 0; c < 8; ++c) v[c] = c < csize ? ld_peer(local, c) : 0.f;
  float t = v[0];
#pragma unroll
  for (unsigned c = 1; c < 8; ++c) t += v[c];
  return t;
}
__device__ __forceinline__ float4 ld_peer4(const float* local, unsigned rank) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(local);
  unsigned ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(ra));
  return v;
}
// the same for four consecutive floats (16-byte aligned): one remote request moves 16 bytes per lane - the
// SM-to-SM network is request-bound, so this is 4x cheaper than four scalar reads
__device__ __forceinline__ float4 sum_peers4(const float* local, unsigned csize) {
  const unsigned a = (unsigned)__cvta_generic_to_shared(local);
  float4 v[8];
#pragma unroll
  for (unsigned c = 0; c < 8; ++c) {
    v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < csize) {
      unsigned ra;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(c));
      asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];"
                   : "=f"(v[c].x), "=f"(v[c].y), "=f"(v[c].z), "=f"(v[c].w)
                   : "r"(ra));
    }
  }
  float4 t = v[0];
#pragma unroll
  for (unsigned c = 1; c < 8; ++c) {
    t.x += v[c].x; t.y += v[c].y; t.z += v[c].z; t.w += v[c].w;
  }
  return t;
}

__device__ __forceinline__ float block_sum_rt(float v, float* red) {      // blockDim.x threads; ends with a barrier
  v = mtd::warp_sum(v);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < nw; ++w) t += red[w];                                // fixed order, same value in every thread
  __syncthreads();
  return t;
}

// C[m,n] = sum_k A[m*a_sm + k*a_sk] * B[k*b_sk + n*b_sn], delivered to epi(m, n, value).
// Work item = TMxTN outputs {tm + i*MT} x {tn + j*NT} (strided so neighbouring lanes touch neighbouring n);
// R lanes share an item and split k.
template <int TM, int TN, class Epi>
__device__ __forceinline__ void block_gemm(const Phase ph, int M, int N, int K, const float* __restrict__ A, int a_sm,
                                           int a_sk, const float* __restrict__ Bm, int b_sk, int b_sn, Epi epi) {
  const int MT = ph.MT, NT = ph.NT;
  const int items = MT * NT;
  const int R = 1 << ph.log2R;
  const int r = threadIdx.x & (R - 1);
  const int per_pass = (int)blockDim.x >> ph.log2R;
  // every lane of a warp runs the same number of passes (items are padded), so the shuffles below are convergent
  for (int pass = 0; pass < ph.passes; ++pass) {
    const int item = pass * per_pass + ((int)threadIdx.x >> ph.log2R);
    const bool live = item < items && M > 0;       // M can be smaller than planned (last ranks of a ragged batch) or 0
    const int tm = live ? (NT == 1 ? item : (int)__umulhi((unsigned)item, ph.nt_magic)) : 0;
    const int tn = live ? item - tm * NT : 0;
    int am[TM], bn[TN];
#pragma unroll
    for (int i = 0; i < TM; ++i) am[i] = min(tm + i * MT, M - 1) * a_sm;
#pragma unroll
    for (int j = 0; j < TN; ++j) bn[j] = min(tn + j * NT, N - 1) * b_sn;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;
    if (live) {
      const float* pa = A + r * a_sk;
      const float* pb = Bm + r * b_sk;
      const int sa = a_sk << ph.log2R, sb = b_sk << ph.log2R;
#pragma unroll 8
      for (int k = r; k < K; k += R) {
        float a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM; ++i) a[i] = pa[am[i]];
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = pb[bn[j]];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        pa += sa;
        pb += sb;
      }
    }
    for (int o = R >> 1; o > 0; o >>= 1) {
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] += __shfl_xor_sync(0xffffffffu, acc[i][j], o);
    }
    if (live && r == 0) {
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          const int m = tm + i * MT, n = tn + j * NT;
          if (m < M && n < N) epi(m, n, acc[i][j]);
        }
    }
  }
}

template <class Epi>
__device__ __forceinline__ void run_gemm(Phase ph, int M, int N, int K, const float* A, int a_sm, int a_sk, const float* Bm,
                                         int b_sk, int b_sn, Epi epi) {
  switch (ph.tile) {
    case 0: block_gemm<4, 4>(ph, M, N, K, A, a_sm, a_sk, Bm, b_sk, b_sn, epi); break;
    case 1: block_gemm<2, 2>(ph, M, N, K, A, a_sm, a_sk, Bm, b_sk, b_sn, epi); break;
    default: block_gemm<1, 1>(ph, M, N, K, A, a_sm, a_sk, Bm, b_sk, b_sn, epi); break;
  }
}

template <int KIND>
__global__ void __launch_bounds__(THREADS, 1) mlp_fit_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float red[THREADS / 32];
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int L = p.L;
  const unsigned rank = cluster_rank(), csize = cluster_size();
  const int r_begin = (int)rank * p.rows_per;
  const int B = max(0, min(p.rows_per, p.B - r_begin));       // rows of this CTA
  // ---- stage weights and biases (full replica), this CTA's rows of X and T
  for (int i = 0; i < L; ++i) {
    const int dout = p.d[i + 1], din = p.d[i];
    for (int e = tid; e < dout * din; e += nthr) {
      const int n = e / din, k = e - n * din;
      sm[p.offW[i] + n * p.pW[i] + k] = p.W[i][e];
    }
    for (int e = tid; e < dout; e += nthr) sm[p.offB[i] + e] = p.bias[i] ? p.bias[i][e] : 0.f;
  }
  for (int e = tid; e < B * p.d[0]; e += nthr) {
    const int b = e / p.d[0], k = e - b * p.d[0];
    sm[p.offA[0] + b * p.pA[0] + k] = p.X[(size_t)r_begin * p.d[0] + e];
  }
  const int nout = B * p.d[L];
  for (int e = tid; e < nout; e += nthr) sm[p.offT + e] = p.T[(size_t)r_begin * p.d[L] + e];
  __syncthreads();

  long long t_prev = clock64();
  for (int step = 0; step < p.steps; ++step) {
    float* S = sm + p.offS;
    // ---- forward
    for (int i = 0; i < L; ++i) {
      const float* Ain = sm + p.offA[i];
      float* Aout = sm + p.offA[i + 1];
      const float* bias = sm + p.offB[i];
      const int po = p.pA[i + 1], act = p.act[i];
      run_gemm(p.fwd[i], B, p.d[i + 1], p.d[i], Ain, p.pA[i], 1, sm + p.offW[i], 1, p.pW[i],
               [=](int b, int n, float v) {
                 v = __fadd_rn(v, bias[n]);
                 Aout[b * po + n] = act ? tanhf(v) : v;
               });
      __syncthreads();
    }
    MLP_TICK(0);
    // ---- partial loss value and G_L = dL/dZ_L
    {
      const float* AL = sm + p.offA[L];
      float* GL = sm + p.offG[L];
      const int dl = p.d[L], pl = p.pA[L], act = p.act[L - 1];
      float part = 0.f;
      for (int e = tid; e < nout; e += nthr) {
        const int b = e / dl, n = e - b * dl;
        const float a = AL[b * pl + n];
        const float pred = p.has_map ? __fadd_rn(__fmul_rn(a, p.pa), p.pb) : a;
        float val, grad;
        mtd::loss_elem<KIND>(pred, sm[p.offT + e], p.scale, val, grad);
        part += val;
        if (p.has_map) grad = __fmul_rn(grad, p.pa);
        if (act) grad = __fmul_rn(grad, __fsub_rn(1.f, __fmul_rn(a, a)));
        GL[b * pl + n] = grad;
      }
      const float total = block_sum_rt(part, red);                 // ends with a barrier
      if (tid == 0) S[p.offSL] = total;
    }
    MLP_TICK(1);
    // ---- backward chain: G_{i-1}[b,k] = (sum_n G_i[b,n] * W_i[n,k]) * tanh'(A_{i-1}[b,k])
    for (int i = L - 1; i >= 1; --i) {
      const float* G = sm + p.offG[i + 1];
      const float* Ain = sm + p.offA[i];
      float* Gin = sm + p.offG[i];
      const int dout = p.d[i + 1], din = p.d[i], pg = p.pA[i + 1], pa = p.pA[i], pw = p.pW[i];
      const int act = p.act[i - 1];
      run_gemm(p.dx[i], B, din, dout, G, pg, 1, sm + p.offW[i], pw, 1, [=](int b, int k, float v) {
        if (act) {
          const float a = Ain[b * pa + k];
          v = __fmul_rn(v, __fsub_rn(1.f, __fmul_rn(a, a)));
        }
        Gin[b * pa + k] = v;
      });
      __syncthreads();
    }
    MLP_TICK(2);
    // ---- every layer's partial dW[n,k] = sum_b G_i[b,n] * A_{i-1}[b,k] in one phase (independent work, no barrier
    //      between layers; zeros when this CTA has no rows)
    for (int i = 0; i < L; ++i) {
      float* Sw = S + p.offSW[i];
      const int pw = p.pW[i];
      run_gemm(p.dw[i], p.d[i + 1], p.d[i], B, sm + p.offG[i + 1], 1, p.pA[i + 1], sm + p.offA[i], p.pA[i], 1,
               [=](int n, int k, float v) { Sw[n * pw + k] = v; });
    }
    MLP_TICK(3);
    // ---- exchange, reduce-scatter then all-gather through the peers' shared memory.  The weight regions are
    //      contiguous and the scratch mirrors their layout element for element (pad elements hold don't-care values),
    //      so the unit of exchange is a float4 of the flat scratch.  Unit u is owned by rank u % C: the owner adds the C
    //      partials in rank order (one value for the whole cluster -> bit-identical replicas) and publishes the sum in
    //      its P buffer; after a second barrier every CTA reads each unit once from its owner and applies
    //      W = W - lr * dW (two roundings, like NumPy).  Each SM serves 2/C of the scratch per step instead of all of it.
    cluster_sync();
    MLP_TICK(4);
    const bool last = (step == p.steps - 1);
    float* P = sm + p.offP;
    const int total4 = p.offSL >> 2;
    const int lgc = __ffs((int)csize) - 1;                 // cluster sizes are powers of two
    // the owner keeps only its own slice of the sums: unit u = rank + C*j lives at P[j]
    for (int u = (int)rank + (int)csize * tid; u < total4; u += (int)csize * nthr)
      *reinterpret_cast<float4*>(P + 4 * (u >> lgc)) = sum_peers4(S + 4 * u, csize);
    float loss_total = 0.f;
    if (rank == 0 && tid == 0) loss_total = sum_peers(S + p.offSL, csize);
    cluster_sync();
    {
      float* Wall = sm + p.offW[0];
      const unsigned cmask = csize - 1;                      // cluster sizes are powers of two
      // four units per trip: four remote loads in flight per thread (a DSMEM load costs ~200+ cycles)
      for (int u0 = tid; u0 < total4; u0 += 4 * nthr) {
        float4 gs[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int u = u0 + q * nthr;
          gs[q] = u < total4 ? ld_peer4(P + 4 * (u >> lgc), (unsigned)u & cmask) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int u = u0 + q * nthr;
          if (u < total4) {
            float4 w = *reinterpret_cast<float4*>(Wall + 4 * u);
            w.x = __fsub_rn(w.x, __fmul_rn(p.lr, gs[q].x));
            w.y = __fsub_rn(w.y, __fmul_rn(p.lr, gs[q].y));
            w.z = __fsub_rn(w.z, __fmul_rn(p.lr, gs[q].z));
            w.w = __fsub_rn(w.w, __fmul_rn(p.lr, gs[q].w));
            *reinterpret_cast<float4*>(Wall + 4 * u) = w;
            // last step: keep the summed gradient for the write-out (S is idle after the second barrier)
            if (last && rank == 0) *reinterpret_cast<float4*>(S + 4 * u) = gs[q];
          }
        }
      }
    }
    if (rank == 0 && tid == 0) p.losses[step] = __fmul_rn(loss_total, p.scale);
    __syncthreads();
    MLP_TICK(5);
  }
  cluster_sync();            // no CTA may exit while a peer can still read its shared memory
  // ---- final weights (all replicas are identical) and the last step's summed gradients
  if (rank == 0) {
    const float* S = sm + p.offS;                        // where the last step parked the summed gradients
    for (int i = 0; i < L; ++i) {
      const int dout = p.d[i + 1], din = p.d[i];
      for (int e = tid; e < dout * din; e += nthr) {
        const int n = e / din, k = e - n * din;
        p.W[i][e] = sm[p.offW[i] + n * p.pW[i] + k];
        if (p.gradW[i]) p.gradW[i][e] = S[p.offSW[i] + n * p.pW[i] + k];
      }
    }
  }
}

inline int pow2_floor(int v) {
  int r = 1;
  while (r * 2 <= v) r *= 2;
  return r;
}

// Pick tile shape and reduction split for one product.  Measured on the notebook model: the phases are LATENCY bound
// (a cost model that minimised shared-memory load + shuffle instructions - larger tiles, fewer lanes - ran 1.5x slower),
// so the rule is: the largest tile that still yields >= 128 work items, then split the reduction over as many lanes as
// it takes to occupy the CTA.
int plan(Phase* out, int M, int N, int K, int nthreads) {
  Phase ph;
  M = std::max(M, 1);
  const int i44 = (int)(ceil_div(M, 4) * ceil_div(N, 4)), i22 = (int)(ceil_div(M, 2) * ceil_div(N, 2));
  int t;
  if (i44 >= 128) { ph.tile = 0; t = 4; }
  else if (i22 >= 128) { ph.tile = 1; t = 2; }
  else { ph.tile = 2; t = 1; }
  ph.MT = (int)ceil_div(M, t);
  ph.NT = (int)ceil_div(N, t);
  const long long items = (long long)ph.MT * ph.NT;
  if (items * ph.NT >= (1ll << 31)) return MT_ERR_UNSUPPORTED;
  int R = items >= nthreads ? 1 : pow2_floor((int)(nthreads / items));
  R = std::min(R, 32);
  R = std::min(R, pow2_floor(std::max(K, 1)));
  R = std::max(R, 1);
  ph.log2R = 0;
  while ((1 << ph.log2R) < R) ++ph.log2R;
  ph.nt_magic = ph.NT == 1 ? 0u : (unsigned)(((1ull << 32) + ph.NT - 1) / ph.NT);   // NT == 1: the kernel skips the divide
  ph.passes = (int)ceil_div(items, nthreads >> ph.log2R);
  *out = ph;
  return MT_OK;
}

inline int odd(int v) { return v | 1; }

int g_mlp_timing = 0;
int g_mlp_threads = 0;       // 0 = automatic (256 per CTA in a cluster, 1024 alone): mt_mlp_set_threads
int g_mlp_cluster = 0;       // 0 = automatic cluster size, else forced (1, 2, 4, 8): mt_mlp_set_cluster

}  // namespace

MT_API int mt_mlp_fit_f32(const mt_mlp_desc* desc, const float* X, const float* T, int64_t batch, int loss_kind,
                          int has_pred_map, float pred_scale, float pred_shift, float lr, float loss_scale, int steps,
                          float* losses) {
  MT_REQUIRE_INIT();
  MT_CHECK_ARG(desc && X && T && losses, "mt_mlp_fit_f32: null pointer");
  MT_CHECK_ARG(desc->n_layers >= 1 && desc->n_layers <= MAXL, "mt_mlp_fit_f32: n_layers must be in [1, %d]", MAXL);
  MT_CHECK_ARG(batch >= 1 && steps >= 0, "mt_mlp_fit_f32: batch >= 1 and steps >= 0 required");
  MT_CHECK_ARG(loss_kind == MT_LOSS_MSE || loss_kind == MT_LOSS_CE_REF || loss_kind == MT_LOSS_BCE,
               "mt_mlp_fit_f32: unknown loss kind %d", loss_kind);
  if (steps == 0) return MT_OK;
  if (batch > (1 << 20)) return fail(MT_ERR_UNSUPPORTED, "mt_mlp_fit_f32: batch does not fit on chip");
  Params p;
  memset(&p, 0, sizeof(p));
  p.L = desc->n_layers;
  p.B = (int)batch;
  p.steps = steps;
  p.loss_kind = loss_kind;
  p.has_map = has_pred_map ? 1 : 0;
  p.timing = g_mlp_timing;
  p.pa = pred_scale;
  p.pb = pred_shift;
  p.lr = lr;
  p.scale = loss_scale;
  p.X = X;
  p.T = T;
  p.losses = losses;
  // cluster size: split the batch over up to 8 CTAs while every CTA keeps at least 8 rows
  int csize = 1;
  while (csize < 8 && batch / (csize * 2) >= 8) csize *= 2;
  if (g_mlp_cluster > 0) csize = g_mlp_cluster;
  p.rows_per = (int)ceil_div(batch, csize);
  // measured on the notebook model: 256 threads per CTA beat 128 / 512 / 1024 inside a cluster (barrier and shuffle cost)
  const int nthreads = g_mlp_threads > 0 ? g_mlp_threads : (csize > 1 ? 256 : THREADS);
  size_t off = 0;
  for (int i = 0; i <= p.L; ++i) {
    MT_CHECK_ARG(desc->dims[i] >= 1 && desc->dims[i] <= 4096, "mt_mlp_fit_f32: dims[%d] = %d out of range", i, desc->dims[i]);
    p.d[i] = desc->dims[i];
    p.pA[i] = odd(p.d[i]);
  }
  for (int i = 0; i < p.L; ++i) {
    MT_CHECK_ARG(desc->weight[i], "mt_mlp_fit_f32: weight[%d] is null", i);
    p.W[i] = desc->weight[i];
    p.bias[i] = desc->bias[i];
    p.gradW[i] = desc->grad_weight[i];
    p.act[i] = desc->act[i] == MT_ACT_TANH ? 1 : 0;
    MT_CHECK_ARG(desc->act[i] == MT_ACT_NONE || desc->act[i] == MT_ACT_TANH, "mt_mlp_fit_f32: act[%d] unknown", i);
    p.pW[i] = odd(p.d[i]);
    const size_t wsz = ((size_t)p.d[i + 1] * p.pW[i] + 3) & ~(size_t)3;   // whole float4s, 16-byte aligned regions
    p.offW[i] = (int)off;             // weight regions are contiguous from offset 0; the scratch mirrors them
    p.offSW[i] = (int)off;
    off += wsz;
    if (plan(&p.fwd[i], p.rows_per, p.d[i + 1], p.d[i], nthreads) || plan(&p.dw[i], p.d[i + 1], p.d[i], p.rows_per, nthreads) ||
        plan(&p.dx[i], p.rows_per, p.d[i], p.d[i + 1], nthreads))
      return fail(MT_ERR_UNSUPPORTED, "mt_mlp_fit_f32: layer %d is too large for the on-chip trainer", i);
  }
  size_t sw = off;                   // scratch: all partial dW (same offsets as the weights), then the partial loss
  p.offSL = (int)sw;
  sw += 4;
  for (int i = 0; i < p.L; ++i) {
    p.offB[i] = (int)off;
    off += ((size_t)p.d[i + 1] + 3) & ~(size_t)3;
  }
  for (int i = 0; i <= p.L; ++i) {
    p.offA[i] = (int)off;
    off += (size_t)p.rows_per * p.pA[i];
  }
  for (int i = 1; i <= p.L; ++i) {
    p.offG[i] = (int)off;
    off += (size_t)p.rows_per * p.pA[i];
  }
  p.offT = (int)off;
  off += (size_t)p.rows_per * p.d[p.L];
  off = (off + 3) & ~(size_t)3;
  p.offS = (int)off;
  p.offP = (int)(off + sw);
  off += sw + (size_t)ceil_div((long long)(sw / 4), csize) * 4;      // S in full, P only this rank's 1/C of the units
  const size_t bytes = off * sizeof(float);
  if (bytes > (size_t)SMEM_MAX)
    return fail(MT_ERR_UNSUPPORTED, "mt_mlp_fit_f32: model + batch need %zu bytes of shared memory per CTA (limit %d)", bytes,
                SMEM_MAX);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(csize, 1, 1);
  cfg.blockDim = dim3(nthreads, 1, 1);
  cfg.dynamicSmemBytes = bytes;
  cfg.stream = g().stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#define MT_FIT(KIND)                                                                                              \
  do {                                                                                                            \
    auto kern = mlp_fit_kernel<KIND>;                                                                             \
    static bool set = false;                                                                                      \
    if (!set) { MT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX)); set = true; } \
    MT_CUDA(cudaLaunchKernelEx(&cfg, kern, p));                                                                   \
  } while (0)
  switch (loss_kind) {
    case MT_LOSS_MSE: MT_FIT(MT_LOSS_MSE); break;
    case MT_LOSS_CE_REF: MT_FIT(MT_LOSS_CE_REF); break;
    default: MT_FIT(MT_LOSS_BCE); break;
  }
#undef MT_FIT
  MT_LAUNCHED();
  return MT_OK;
}

// tuning / test knob: force the cluster size of mt_mlp_fit_f32 (0 = automatic)
extern "C" __attribute__((visibility("default"))) int mt_mlp_set_cluster(int csize) {
  if (!(csize == 0 || csize == 1 || csize == 2 || csize == 4 || csize == 8))
    return mt::fail(MT_ERR_INVALID, "mt_mlp_set_cluster: cluster size must be 0, 1, 2, 4 or 8");
  g_mlp_cluster = csize;
  return MT_OK;
}

// debug: enable (1) / disable (0) the phase clocks of rank 0 and read them back (8 x uint64 SM cycles, accumulated)
extern "C" __attribute__((visibility("default"))) int mt_mlp_debug_clocks(int enable, unsigned long long* out8) {
  g_mlp_timing = enable ? 1 : 0;
  if (out8) {
    MT_CUDA(cudaStreamSynchronize(g().stream));
    MT_CUDA(cudaMemcpyFromSymbol(out8, g_mlp_clk, sizeof(unsigned long long) * 8));
    unsigned long long zero[8] = {0};
    MT_CUDA(cudaMemcpyToSymbol(g_mlp_clk, zero, sizeof(zero)));
  }
  return MT_OK;
}

extern "C" __attribute__((visibility("default"))) int mt_mlp_set_threads(int nthreads) {
  if (!(nthreads == 0 || nthreads == 128 || nthreads == 256 || nthreads == 512 || nthreads == 1024))
    return mt::fail(MT_ERR_INVALID, "mt_mlp_set_threads: 0, 128, 256, 512 or 1024");
  g_mlp_threads = nthreads;
  return MT_OK;
}
