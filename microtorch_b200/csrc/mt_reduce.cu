// mt_reduce.cu - deterministic sum reductions over a [outer, red, inner] view (fp32).
//
// One primitive serves Reduce.sum (src/tensor/ops/reduce.py:11), the un-broadcast
// reductions of BaseOps.bkwd_broadcast (src/tensor/ops/base.py:28-54) and - with the fused
// `other` operand - the `grad * b.data` that precedes them in mul backward
// (src/tensor/ops/overload.py:109-112), so the product is never written to HBM.
//
// HBM-bound: 4 B/element read (8 B with `other`), output negligible.  Structure: 128-bit
// coalesced loads with four independent vectors in flight per thread, warp-shuffle
// reduction, shared-memory tree across warps, and - when one pass cannot fill 148 SMs - a
// split of the reduced axis across CTAs followed by a second tiny pass over the partials.
// No float atomics: the summation tree is fixed by the shape, so results are reproducible.
#include <algorithm>

#include "mt_common.cuh"

using namespace mt;

namespace {

constexpr int THREADS = 256;

// ------------------------------------------------------------ inner == 1: rows
// item = (row o, split s); the CTA sums x[o, r0:r1] (* w) and writes one partial.
// With splits > 1 the partials are folded by the LAST CTA to finish (threadfence + ticket, the loss kernel's idiom):
// one warp per row sums its `splits` partials lane-strided and then by the shuffle tree - a fixed order - so the
// second launch of round 1 (10 us of launch latency behind a 45 us pass at 2^26 elements) is gone.
template <bool HAS_W, bool VEC>
__global__ void __launch_bounds__(THREADS) reduce_rows_kernel(float* __restrict__ out, const float* __restrict__ x,
                                                              long long outer, long long red, long long splits,
                                                              long long chunk, const float* __restrict__ w,
                                                              long long so, long long sr, float* __restrict__ final_out,
                                                              unsigned int* __restrict__ ticket) {
  __shared__ float smem[THREADS / 32];
  __shared__ bool is_last;
  const long long items = outer * splits;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const long long o = item / splits, s = item - o * splits;
    const long long r0 = s * chunk;
    long long r1 = r0 + chunk;
    if (r1 > red) r1 = red;
    const float* px = x + o * red;
    const float* pw = HAS_W ? w + o * so : nullptr;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    if (VEC) {  // r0, red multiples of 4 and bases 16 B aligned (checked on the host)
      const long long v0 = r0 >> 2, v1 = r1 >> 2;
      long long i = v0 + threadIdx.x;
      for (; i + 3 * THREADS < v1; i += 4 * THREADS) {
        float4 a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          a[u] = mtd::ld_stream4(px + 4 * (i + u * THREADS));
          if (HAS_W) b[u] = mtd::ld_stream4(pw + 4 * (i + u * THREADS));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (HAS_W) {
            acc0 += a[u].x * b[u].x; acc1 += a[u].y * b[u].y; acc2 += a[u].z * b[u].z; acc3 += a[u].w * b[u].w;
          } else {
            acc0 += a[u].x; acc1 += a[u].y; acc2 += a[u].z; acc3 += a[u].w;
          }
        }
      }
      for (; i < v1; i += THREADS) {
        float4 a = mtd::ld_stream4(px + 4 * i);
        if (HAS_W) {
          float4 b = mtd::ld_stream4(pw + 4 * i);
          acc0 += a.x * b.x; acc1 += a.y * b.y; acc2 += a.z * b.z; acc3 += a.w * b.w;
        } else {
          acc0 += a.x; acc1 += a.y; acc2 += a.z; acc3 += a.w;
        }
      }
      for (long long j = (v1 << 2) + threadIdx.x; j < r1; j += THREADS)  // tail of the last split
        acc0 += HAS_W ? px[j] * pw[j] : px[j];
    } else {
      for (long long j = r0 + threadIdx.x; j < r1; j += THREADS) acc0 += HAS_W ? px[j] * pw[j * sr] : px[j];
    }
    float t = mtd::block_sum<THREADS>((acc0 + acc1) + (acc2 + acc3), smem);
    if (threadIdx.x == 0) out[item] = t;
  }
  if (ticket) {
    if (threadIdx.x == 0) {            // thread 0 wrote every partial of this CTA
      __threadfence();
      is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
      __threadfence();
      const int lane = threadIdx.x & 31;
      for (long long o = threadIdx.x >> 5; o < outer; o += THREADS / 32) {
        float acc = 0.f;
        for (long long k = lane; k < splits; k += 32) acc += __ldcg(out + o * splits + k);
        acc = mtd::warp_sum(acc);
        if (lane == 0) final_out[o] = acc;
      }
      if (threadIdx.x == 0) *ticket = 0u;   // re-arm for the next launch
    }
  }
}

// short rows: one warp per row
template <bool HAS_W>
__global__ void __launch_bounds__(THREADS) reduce_rows_warp_kernel(float* __restrict__ out, const float* __restrict__ x,
                                                                   long long outer, long long red,
                                                                   const float* __restrict__ w, long long so,
                                                                   long long sr) {
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (THREADS / 32);
  for (long long o = (long long)blockIdx.x * (THREADS / 32) + (threadIdx.x >> 5); o < outer; o += warps) {
    const float* px = x + o * red;
    float acc = 0.f;
    for (long long j = lane; j < red; j += 32) acc += HAS_W ? px[j] * w[o * so + j * sr] : px[j];
    acc = mtd::warp_sum(acc);
    if (lane == 0) out[o] = acc;
  }
}

// ------------------------------------------------------------ inner > 1: columns
// block = 256 threads = TX column-vectors x TY row lanes, TX = 2^tx_bits chosen on the host from `inner` (a (2^22, 8)
// gradient folded onto a (1, 8) operand has 2 column-vectors: TX = 2, TY = 128 - with the fixed 64 x 4 shape of round 1
// only 8 of 256 threads had work and the kernel ran at 17 % of HBM); grid = (column tiles, splits, outer tiles).
// The row lanes are folded through shared memory in a fixed binary tree; each thread stores its V sums as ONE vector,
// so the stores are conflict-free (the old [TY][TX*V+1] float layout was 4-way conflicted: 7.9 M conflicts per launch).
constexpr int CT = 256;

// one block's share of one (outer index, split): sums rows [r0, r1) of its TX column-vectors and stores the V sums of
// every column at `po`.  COHERENT: the rows were written by other CTAs of THIS launch (the fold below): L2 loads (ld.cg)
// instead of the non-coherent streaming path.
template <bool HAS_W, int V, bool COHERENT>
__device__ __forceinline__ void cols_block(float* __restrict__ po, const float* __restrict__ px, const float* __restrict__ pw,
                                           long long r0, long long r1, long long inner, long long sr, int tx_bits,
                                           bool col_ok, float* smem) {
  const int TYr = CT >> tx_bits;
  const int ty = threadIdx.x >> tx_bits;
  auto ld4 = [](const float* q) { return COHERENT ? __ldcg(reinterpret_cast<const float4*>(q)) : mtd::ld_stream4(q); };
  auto ld1 = [](const float* q) { return COHERENT ? __ldcg(q) : *q; };
  float acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = 0.f;
  if (col_ok) {
    long long r = r0 + ty;
    if (V == 4) {
      for (; r + 3 * TYr < r1; r += 4 * TYr) {
        float4 a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          a[u] = ld4(px + (r + u * TYr) * inner);
          if (HAS_W) b[u] = mtd::ld_stream4(pw + (r + u * TYr) * sr);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (HAS_W) {
            acc[0] += a[u].x * b[u].x; acc[1] += a[u].y * b[u].y; acc[2] += a[u].z * b[u].z; acc[3] += a[u].w * b[u].w;
          } else {
            acc[0] += a[u].x; acc[1] += a[u].y; acc[2] += a[u].z; acc[3] += a[u].w;
          }
        }
      }
      for (; r < r1; r += TYr) {
        float4 a = ld4(px + r * inner);
        if (HAS_W) {
          float4 b = mtd::ld_stream4(pw + r * sr);
          acc[0] += a.x * b.x; acc[1] += a.y * b.y; acc[2] += a.z * b.z; acc[3] += a.w * b.w;
        } else {
          acc[0] += a.x; acc[1] += a.y; acc[2] += a.z; acc[3] += a.w;
        }
      }
    } else {
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;      // four independent loads in flight
      for (; r + 3 * TYr < r1; r += 4 * TYr) {
        if (HAS_W) {
          a0 += ld1(px + r * inner) * pw[r * sr];
          a1 += ld1(px + (r + TYr) * inner) * pw[(r + TYr) * sr];
          a2 += ld1(px + (r + 2 * TYr) * inner) * pw[(r + 2 * TYr) * sr];
          a3 += ld1(px + (r + 3 * TYr) * inner) * pw[(r + 3 * TYr) * sr];
        } else {
          a0 += ld1(px + r * inner);
          a1 += ld1(px + (r + TYr) * inner);
          a2 += ld1(px + (r + 2 * TYr) * inner);
          a3 += ld1(px + (r + 3 * TYr) * inner);
        }
      }
      for (; r < r1; r += TYr) a0 += HAS_W ? ld1(px + r * inner) * pw[r * sr] : ld1(px + r * inner);
      acc[0] = (a0 + a1) + (a2 + a3);
    }
  }
  if (V == 4) reinterpret_cast<float4*>(smem)[threadIdx.x] = make_float4(acc[0], acc[1], acc[2], acc[3]);
  else smem[threadIdx.x] = acc[0];
  __syncthreads();
  for (int half = TYr >> 1; half > 0; half >>= 1) {      // fixed tree over the row lanes: deterministic
    if (ty < half) {
      if (V == 4) {
        float4 a = reinterpret_cast<float4*>(smem)[threadIdx.x];
        const float4 b = reinterpret_cast<float4*>(smem)[threadIdx.x + (half << tx_bits)];
        a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
        reinterpret_cast<float4*>(smem)[threadIdx.x] = a;
      } else {
        smem[threadIdx.x] += smem[threadIdx.x + (half << tx_bits)];
      }
    }
    __syncthreads();
  }
  if (ty == 0 && col_ok) {
    if (V == 4) *reinterpret_cast<float4*>(po) = reinterpret_cast<float4*>(smem)[threadIdx.x];
    else po[0] = smem[threadIdx.x];
  }
  __syncthreads();
}

// `tickets` != NULL (splits > 1): `out` holds the partials [outer, splits, inner]; the last of the `splits` CTAs of a
// (column tile, outer lane) to finish folds that tile's partials - the same block routine over the rows of the partial
// array - into `final_out`, so a split column reduction is ONE launch.
template <bool HAS_W, int V>
__global__ void __launch_bounds__(CT) reduce_cols_kernel(float* __restrict__ out, const float* __restrict__ x,
                                                         long long outer, long long red, long long inner,
                                                         long long splits, long long chunk,
                                                         const float* __restrict__ w, long long so, long long sr,
                                                         long long si, int tx_bits, float* __restrict__ final_out,
                                                         unsigned int* __restrict__ tickets) {
  __shared__ __align__(16) float smem[CT * V];
  __shared__ bool is_last;
  const int TXr = 1 << tx_bits;
  const int tx = threadIdx.x & (TXr - 1);
  const long long col = ((long long)blockIdx.x * TXr + tx) * V;
  const bool col_ok = col < inner;
  const long long s = blockIdx.y;
  const long long r0 = s * chunk;
  long long r1 = r0 + chunk;
  if (r1 > red) r1 = red;
  for (long long o = blockIdx.z; o < outer; o += gridDim.z)
    cols_block<HAS_W, V, false>(out + (o * splits + s) * inner + col, x + (o * red) * inner + col,
                                HAS_W ? w + o * so + col * si : nullptr, r0, r1, inner, sr, tx_bits, col_ok, smem);
  if (tickets) {
    unsigned int* ticket = tickets + (size_t)blockIdx.z * gridDim.x + blockIdx.x;
    __threadfence();                   // this CTA's partials (written by its ty == 0 threads) before the ticket
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(ticket, 1u) == gridDim.y - 1);
    __syncthreads();
    if (is_last) {
      __threadfence();
      for (long long o = blockIdx.z; o < outer; o += gridDim.z)
        cols_block<false, V, true>(final_out + o * inner + col, out + (o * splits) * inner + col, nullptr, 0, splits, inner,
                                   0, tx_bits, col_ok, smem);
      if (threadIdx.x == 0) *ticket = 0u;   // re-arm for the next launch
    }
  }
}

// few rows, many columns (the fold of stacked split-K partials, x.sum(axis=0) of a short-fat array): one thread per
// column-vector walks all `red` rows itself - fully coalesced, no shared memory, no idle row lanes.
template <bool HAS_W, int V>
__global__ void __launch_bounds__(CT) reduce_cols_flat_kernel(float* __restrict__ out, const float* __restrict__ x,
                                                              long long outer, long long red, long long inner,
                                                              const float* __restrict__ w, long long so, long long sr,
                                                              long long si) {
  const long long vcols = inner / V;                      // V == 4: inner % 4 == 0 (checked on the host)
  const long long total = outer * vcols;
  const long long stride = (long long)gridDim.x * CT;
  for (long long t = (long long)blockIdx.x * CT + threadIdx.x; t < total; t += stride) {
    const long long o = t / vcols, col = (t - o * vcols) * V;
    const float* px = x + (o * red) * inner + col;
    const float* pw = HAS_W ? w + o * so + col * si : nullptr;
    if (V == 4) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      long long r = 0;
      for (; r + 3 < red; r += 4) {
        float4 a[4], b[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          a[u] = mtd::ld_stream4(px + (r + u) * inner);
          if (HAS_W) b[u] = mtd::ld_stream4(pw + (r + u) * sr);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (HAS_W) { acc.x += a[u].x * b[u].x; acc.y += a[u].y * b[u].y; acc.z += a[u].z * b[u].z; acc.w += a[u].w * b[u].w; }
          else { acc.x += a[u].x; acc.y += a[u].y; acc.z += a[u].z; acc.w += a[u].w; }
        }
      }
      if (r < red) {
        // the last 1 - 3 rows (ALL the rows of a 2- or 3-way split-K fold): every load issued before the first addition -
        // written as a loop the additions serialised the round trips (0.72 of HBM on a (2, 2^25) fold: 16 bytes in
        // flight per thread)
        const int left = (int)(red - r);
        float4 a[3], b[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          if (u < left) {
            a[u] = mtd::ld_stream4(px + (r + u) * inner);
            if (HAS_W) b[u] = mtd::ld_stream4(pw + (r + u) * sr);
          }
        }
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          if (u < left) {
            if (HAS_W) { acc.x += a[u].x * b[u].x; acc.y += a[u].y * b[u].y; acc.z += a[u].z * b[u].z; acc.w += a[u].w * b[u].w; }
            else { acc.x += a[u].x; acc.y += a[u].y; acc.z += a[u].z; acc.w += a[u].w; }
          }
        }
      }
      *reinterpret_cast<float4*>(out + o * inner + col) = acc;
    } else {
      float acc = 0.f;
      for (long long r = 0; r < red; ++r) acc += HAS_W ? px[r * inner] * pw[r * sr] : px[r * inner];
      out[o * inner + col] = acc;
    }
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// tickets of the single-launch split reductions: zero when idle (the last CTA re-arms its own), one stream -> no sharing
constexpr long long TICKETS = 4096;
unsigned int* g_tickets = nullptr;
int g_single_launch = 1;               // 0: round 1's two-launch form (mt_reduce_set_single_launch, A/B measurements)
int ensure_tickets() {
  if (g_tickets) return MT_OK;
  MT_CUDA(cudaMalloc((void**)&g_tickets, sizeof(unsigned int) * TICKETS));
  MT_CUDA(cudaMemset(g_tickets, 0, sizeof(unsigned int) * TICKETS));
  return MT_OK;
}

// resident CTAs per SM of one reduce_rows_kernel instantiation (asked once)
template <bool HAS_W, bool VEC>
int rows_resident() {
  static int n = 0;
  if (!n) {
    int q = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, reduce_rows_kernel<HAS_W, VEC>, THREADS, 0) != cudaSuccess || q < 1) q = 4;
    n = std::min(q, 8);
  }
  return n;
}

int reduce_rows(float* out, const float* x, long long outer, long long red, const float* w, long long so, long long sr) {
  cudaStream_t st = g().stream;
  const int sms = g().sm_count;
  if (red <= 256 && outer >= 64) {
    const int grid = (int)std::min<long long>(ceil_div(outer, THREADS / 32), (long long)sms * 16);
    if (w) reduce_rows_warp_kernel<true><<<grid, THREADS, 0, st>>>(out, x, outer, red, w, so, sr);
    else reduce_rows_warp_kernel<false><<<grid, THREADS, 0, st>>>(out, x, outer, red, w, so, sr);
    MT_LAUNCHED();
    return MT_OK;
  }
  long long splits = 1;
  if (outer < 2LL * sms) {
    splits = std::min<long long>(ceil_div(4LL * sms, outer), ceil_div(red, 4096));
    if (splits < 1) splits = 1;
  }
  long long chunk = ceil_div(red, splits);
  chunk = (chunk + 3) & ~3LL;
  splits = ceil_div(red, chunk);
  const bool vec = aligned16(x) && red % 4 == 0 && (!w || (aligned16(w) && sr == 1 && so % 4 == 0));
  float* dst = out;
  PoolScratch scratch;
  float* partial = nullptr;
  if (splits > 1) {
    int rc = scratch.alloc(sizeof(float) * outer * splits);
    if (rc) return rc;
    partial = scratch.f32();
    dst = partial;
  }
  unsigned int* ticket = nullptr;
  if (splits > 1 && g_single_launch) {
    int rc = ensure_tickets();
    if (rc) return rc;
    ticket = g_tickets;
  }
  // ONE wave: the persistent grid is sized from the occupancy of the instantiation that runs.  A fixed 8 CTAs per SM was
  // right at 31 registers; the ticket fold took the kernels to 34 / 38, 6 CTAs fit, and the 1184-CTA grid ran as a wave and
  // a third - sum(axis=1) of 2^28 elements 192 us instead of 153 under ncu.  (Forcing 32 registers instead costs the
  // long-loop sum(all) case 5 %.)
#define MT_ROWS(HW, VC)                                                                                               \
  do {                                                                                                                \
    const int grid = (int)std::min<long long>(outer * splits, (long long)sms * rows_resident<HW, VC>());              \
    reduce_rows_kernel<HW, VC><<<grid, THREADS, 0, st>>>(dst, x, outer, red, splits, chunk, w, so, sr, out, ticket);  \
  } while (0)
  if (w) { if (vec) MT_ROWS(true, true); else MT_ROWS(true, false); }
  else { if (vec) MT_ROWS(false, true); else MT_ROWS(false, false); }
#undef MT_ROWS
  MT_LAUNCHED();
  if (splits > 1 && !ticket) {
    int rc = reduce_rows(out, partial, outer, splits, nullptr, 0, 0);
    return rc;
  }
  return MT_OK;
}

template <bool HAS_W, int V>
int cols_resident() {
  static int n = 0;
  if (!n) {
    int q = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, reduce_cols_kernel<HAS_W, V>, CT, 0) != cudaSuccess || q < 1) q = 4;
    n = std::min(q, 8);
  }
  return n;
}

int reduce_cols(float* out, const float* x, long long outer, long long red, long long inner, const float* w,
                long long so, long long sr, long long si, bool may_split = true) {
  cudaStream_t st = g().stream;
  const int sms = g().sm_count;
  const bool vec = inner % 4 == 0 && aligned16(x) && aligned16(out) &&
                   (!w || (aligned16(w) && si == 1 && sr % 4 == 0 && so % 4 == 0));
  const int V = vec ? 4 : 1;
  const long long vcols = ceil_div(inner, (long long)V);
  if (red <= 16 && outer * vcols >= (long long)sms * CT) {      // short-fat: one thread per column-vector
    const int grid = ew_grid(outer * vcols, CT);
#define MT_FLAT(HW, VV) reduce_cols_flat_kernel<HW, VV><<<grid, CT, 0, st>>>(out, x, outer, red, inner, w, so, sr, si)
    if (w) { if (vec) MT_FLAT(true, 4); else MT_FLAT(true, 1); }
    else { if (vec) MT_FLAT(false, 4); else MT_FLAT(false, 1); }
#undef MT_FLAT
    MT_LAUNCHED();
    return MT_OK;
  }
  int tx_bits = 0;
  while (tx_bits < 6 && (1LL << tx_bits) < vcols) ++tx_bits;    // TX = min(64, pow2 >= column-vectors)
  const long long oz = std::min<long long>(outer, 65535);
  // few column tiles and many rows (tall-thin: (131072, 512) is 2 tiles of 64 vectors): the reduced axis is split ~300
  // ways and ONE CTA per tile folds the partials with only 4 row lanes - ~19 serial L2 round trips while the machine
  // idles.  Narrower tiles (down to 16 vectors = 256 contiguous bytes per row, still whole lines) give more tiles, fewer
  // splits per tile and 4x the row lanes: the fold shrinks to 2 round trips.
  if (may_split && vec)
    while (tx_bits > 4 && ceil_div(vcols, 1LL << tx_bits) * oz * 16 <= sms && red >= 4096) --tx_bits;
  const int TXr = 1 << tx_bits, TYr = CT >> tx_bits;
  const long long tiles = ceil_div(vcols, (long long)TXr);
  long long splits = 1;
  // CTAs of the instantiation that will run that fit on an SM (4 at 64 registers, 6 at 40): one full wave of them
  const long long fill = (long long)sms * (w ? (vec ? cols_resident<true, 4>() : cols_resident<true, 1>())
                                             : (vec ? cols_resident<false, 4>() : cols_resident<false, 1>()));
  if (may_split && tiles * oz < fill) {
    // fill the machine (one wave) but leave every row lane at least ~8 rows of work
    splits = std::min<long long>(fill / (tiles * oz), ceil_div(red, (long long)TYr * 8));
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
  }
  const long long chunk = ceil_div(red, splits);
  splits = ceil_div(red, chunk);
  float* dst = out;
  PoolScratch scratch;
  float* partial = nullptr;
  if (splits > 1) {
    int rc = scratch.alloc(sizeof(float) * outer * splits * inner);
    if (rc) return rc;
    partial = scratch.f32();
    dst = partial;
  }
  dim3 grid((unsigned)tiles, (unsigned)splits, (unsigned)oz);
  unsigned int* tickets = nullptr;
  if (splits > 1 && g_single_launch && tiles * oz <= TICKETS - 1) {
    int rc = ensure_tickets();
    if (rc) return rc;
    tickets = g_tickets + 1;          // slot 0 belongs to reduce_rows
  }
#define MT_COLS(HW, VV)                                                                                                  \
  reduce_cols_kernel<HW, VV><<<grid, CT, 0, st>>>(dst, x, outer, red, inner, splits, chunk, w, so, sr, si, tx_bits, out, \
                                                  tickets)
  if (w) { if (vec) MT_COLS(true, 4); else MT_COLS(true, 1); }
  else { if (vec) MT_COLS(false, 4); else MT_COLS(false, 1); }
#undef MT_COLS
  MT_LAUNCHED();
  if (splits > 1 && !tickets) {
    // the fold of the partials is ONE more launch (it used to split again: three launches for a 2^26-element
    // sum(axis=0), 10 us of launch latency on a 43 us kernel)
    int rc = reduce_cols(out, partial, outer, splits, inner, nullptr, 0, 0, 0, false);
    return rc;
  }
  return MT_OK;
}

// ------------------------------------------------------------ max / min (Reduce.max / Reduce.min forward)
// Selection, so the result is bit-exact and order-independent; NaN wins like np.max / np.min.
template <bool IS_MAX>
__device__ __forceinline__ float pick(float a, float b) {
  return ((IS_MAX ? a > b : a < b) || a != a) ? a : b;
}

template <bool IS_MAX>
__device__ __forceinline__ float block_pick(float v, float* smem) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v = pick<IS_MAX>(v, __shfl_xor_sync(0xffffffffu, v, d));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  v = smem[0];
#pragma unroll
  for (int w = 1; w < THREADS / 32; ++w) v = pick<IS_MAX>(v, smem[w]);
  return v;
}

// inner == 1: item = (row, split); every thread starts from the first element of its item's range
template <bool IS_MAX>
__global__ void __launch_bounds__(THREADS) extreme_rows_kernel(float* __restrict__ out, const float* __restrict__ x,
                                                               long long outer, long long red, long long splits,
                                                               long long chunk) {
  __shared__ float smem[THREADS / 32];
  const long long items = outer * splits;
  for (long long item = blockIdx.x; item < items; item += gridDim.x) {
    const long long o = item / splits, s = item - o * splits;
    const long long r0 = s * chunk;
    long long r1 = r0 + chunk;
    if (r1 > red) r1 = red;
    const float* px = x + o * red;
    float acc = px[r0];
    const bool vec = ((reinterpret_cast<uintptr_t>(px + r0) & 15) == 0);
    long long j = r0;
    if (vec) {
      const long long nv = (r1 - r0) >> 2;
      for (long long i = threadIdx.x; i < nv; i += THREADS) {
        const float4 a = mtd::ld_stream4(px + r0 + 4 * i);
        acc = pick<IS_MAX>(acc, pick<IS_MAX>(pick<IS_MAX>(a.x, a.y), pick<IS_MAX>(a.z, a.w)));
      }
      j = r0 + (nv << 2);
    }
    for (long long k = j + threadIdx.x; k < r1; k += THREADS) acc = pick<IS_MAX>(acc, px[k]);
    const float t = block_pick<IS_MAX>(acc, smem);
    if (threadIdx.x == 0) out[item] = t;
  }
}

// inner > 1: one thread per (outer, split, column), coalesced across columns
template <bool IS_MAX>
__global__ void __launch_bounds__(THREADS) extreme_cols_kernel(float* __restrict__ out, const float* __restrict__ x,
                                                               long long outer, long long red, long long inner,
                                                               long long splits, long long chunk) {
  const long long total = outer * splits * inner;
  const long long stride = (long long)gridDim.x * THREADS;
  for (long long t = (long long)blockIdx.x * THREADS + threadIdx.x; t < total; t += stride) {
    const long long col = t % inner, os = t / inner;
    const long long s = os % splits, o = os / splits;
    const long long r0 = s * chunk;
    long long r1 = r0 + chunk;
    if (r1 > red) r1 = red;
    const float* px = x + (o * red) * inner + col;
    float acc = px[r0 * inner];
    for (long long r = r0 + 1; r < r1; ++r) acc = pick<IS_MAX>(acc, __ldg(px + r * inner));
    out[t] = acc;
  }
}

template <bool IS_MAX>
int reduce_extreme(float* out, const float* x, long long outer, long long red, long long inner) {
  cudaStream_t st = g().stream;
  const long long sms = g().sm_count;
  long long splits = 1;
  if (inner == 1) {
    if (outer < 2 * sms) splits = std::max<long long>(1, std::min<long long>(ceil_div(4 * sms, outer), ceil_div(red, 4096)));
  } else if (outer * inner < 64 * sms * THREADS / 32) {
    splits = std::max<long long>(1, std::min<long long>(ceil_div(64 * sms * THREADS / 32, outer * inner), ceil_div(red, 64)));
  }
  long long chunk = ceil_div(red, splits);
  chunk = (chunk + 3) & ~3LL;
  splits = ceil_div(red, chunk);
  float* dst = out;
  PoolScratch scratch;
  float* partial = nullptr;
  if (splits > 1) {
    int rc = scratch.alloc(sizeof(float) * outer * splits * inner);
    if (rc) return rc;
    partial = scratch.f32();
    dst = partial;
  }
  if (inner == 1) {
    const int grid = (int)std::min<long long>(outer * splits, sms * 8);
    extreme_rows_kernel<IS_MAX><<<grid, THREADS, 0, st>>>(dst, x, outer, red, splits, chunk);
  } else {
    const int grid = (int)std::min<long long>(ceil_div(outer * splits * inner, THREADS), sms * 16);
    extreme_cols_kernel<IS_MAX><<<grid, THREADS, 0, st>>>(dst, x, outer, red, inner, splits, chunk);
  }
  MT_LAUNCHED();
  if (splits > 1) {
    int rc = reduce_extreme<IS_MAX>(out, partial, outer, splits, inner);
    return rc;
  }
  return MT_OK;
}

}  // namespace

// diagnostic knob: 0 = split reductions as two launches (main pass + fold, round 1), 1 (default) = one launch whose last
// CTA folds the partials
extern "C" __attribute__((visibility("default"))) int mt_reduce_set_single_launch(int v) {
  g_single_launch = v ? 1 : 0;
  return MT_OK;
}

MT_API int mt_reduce_extreme_f32(int is_max, float* out, const float* x, int64_t outer, int64_t red, int64_t inner) {
  MT_REQUIRE_INIT();
  MT_CHECK_ARG(out && x, "mt_reduce_extreme_f32: null pointer");
  MT_CHECK_ARG(outer >= 0 && red >= 0 && inner >= 0, "mt_reduce_extreme_f32: negative extent");
  if (outer * inner == 0) return MT_OK;
  MT_CHECK_ARG(red > 0, "mt_reduce_extreme_f32: zero-size array to reduction operation which has no identity");
  return is_max ? reduce_extreme<true>(out, x, outer, red, inner) : reduce_extreme<false>(out, x, outer, red, inner);
}

MT_API int mt_reduce_sum_f32(float* out, const float* x, int64_t outer, int64_t red, int64_t inner, const float* other,
                             int64_t so, int64_t sr, int64_t si) {
  MT_REQUIRE_INIT();
  MT_CHECK_ARG(out && x, "mt_reduce_sum_f32: null pointer");
  MT_CHECK_ARG(outer >= 0 && red >= 0 && inner >= 0, "mt_reduce_sum_f32: negative extent");
  if (outer * inner == 0) return MT_OK;
  if (red == 0) return mt_memset_zero(out, sizeof(float) * outer * inner);
  if (inner == 1) return reduce_rows(out, x, outer, red, other, so, sr);
  return reduce_cols(out, x, outer, red, inner, other, so, sr, si);
}
