// mt_gemm_tc_mid.cu - the tcgen05 split-fp32 GEMM for the MID-SIZE regime (2^25 .. 2^32 multiply-adds).
//
// Same contract and the same arithmetic as mt_gemm_tc.cu (C (+)= act(A.B + bias) (* (1 - T^2)), fp32 in / out, error-
// compensated split on chip: TF32 hi.hi + two BF16 corrections, or 3xTF32, K accumulated in TMEM in chunks that are
// folded into IEEE fp32 register sums), replacing the same reference call sites (src/tensor/ops/overload.py:132,142,153
// through src/nn/layer/linear.py:48) - but organised for problems that cannot fill the persistent CTA-pair kernel:
// layers of 128 - 1024 features at a few thousand rows.  There the big kernel is latency-bound (one 256 x 256 tile per
// SM pair walks its K loop serially at ~0.85 us per k-block behind six converter warps, its split-K partials go
// through HBM and a second launch), and the SIMT tile is FMA-bound at ~20 TFLOP/s (profiles/r02_engine_crossover.jsonl).
//
// Shape of this kernel:
//   * one CTA per (output tile, K slice): grid (m-blocks, n-blocks, S), NOT persistent - small tiles and K slices put
//     every SM to work on a problem of a few hundred tiles' worth of k-blocks;
//   * the S K-slices of a tile form ONE THREAD-BLOCK CLUSTER, S = 1 .. 8 (any number up to the portable cluster size).
//     Each CTA leaves its fp32 partial tile in its own shared memory; after one cluster barrier CTA s sums rows
//     [128 s / S, 128 (s+1) / S) of all S partials through distributed shared memory (ld.shared::cluster, slice order
//     0 .. S-1: deterministic) and all 16 warps run the epilogue (scale, bias, tanh, tanh', accumulate) on them with
//     coalesced 512-byte row stores.  No partial ever touches HBM and there is no fold launch;
//   * fourteen converter warps instead of six: with one tile per CTA nobody needs a dedicated epilogue warpgroup
//     during the K loop, so warps 8-15 convert too and drain a finished TMEM chunk one k-block late (its MMAs have
//     retired by then, the conversion pipeline never stalls on the tensor core);
//   * two tile flavours (template NCTA): 128 x 128 with single-CTA MMAs (cta_group::1: no pair handshakes, no cluster
//     barrier in the prologue) and 256 x 256 CTA pairs (cta_group::2 inside a cluster of 2 S CTAs) with four times the
//     flops per k-block for 1.2x the shared-memory traffic; the host picks flavour and S from a measured cost model.
//
//   warp 0      TMA producer (one lane)                 full[r]
//   warp 1      MMA issuer (one lane; pair: leader)     empty[r], lofree[l], tfull[a]
//   warps 2-7   converters (56 registers)               conv[l]
//   warps 8-15  converters + chunk drain + staging of the partial tile (200 registers)   conv[l], tempty[a]
//   all warps   cluster fold + epilogue (128 registers)
//
// Roofline: the K loop is bound by the SHARED-MEMORY PORT, not the tensor pipe - per k-block of 32 a 128 x 128 tile moves
// 32 KB in by TMA, 32 + 32 KB through the converters and 64 KB of operand reads = 160 KB at 128 B/clk = 1250 cycles for
// 512 cycles of MMAs (measured 1254) - with launch / prologue / fold latency around it; DESIGN.md section 4.2 has the
// phase clocks, the ncu counters and the measured crossover against both neighbours.
#include <algorithm>

#include "mt_common.cuh"
#include "mt_gemm.cuh"
#include "mt_tc_common.cuh"

using namespace mt;

namespace tcm {

using namespace tc;

constexpr int BNL = 128;                      // B rows a CTA loads (a CTA pair holds half of a 256-wide tile each)
constexpr int NUM_THREADS = 512;
constexpr int CONV_WARPS = 14;
constexpr int CONV_THREADS = CONV_WARPS * 32;   // 448 = 7 * 64 (convert_tile_mix_k needs whole groups of 8 rows per step)
constexpr int A_TILE = BM * BK * 4;             // 16 KB
constexpr int B_TILE = BNL * BK * 4;            // 16 KB
constexpr int RAW_STAGE = A_TILE + B_TILE;      // TMA destination
constexpr int LO_STAGE = A_TILE + B_TILE;       // converter output: mix [A_hi16 | A_lo16 | B_hi16 | B_lo16], else [A_lo | B_lo]
constexpr int MAX_STAGES = 8;
constexpr int AUX = 1024 /*align slack*/ + 512 /*barriers*/;
constexpr int SMEM_MAX = 232448;
constexpr int REGS_LOW = 56, REGS_HIGH = 200, REGS_COMMON = 128;
constexpr int MAX_SPLITS = 8;                   // portable cluster size

struct Params {
  float* C;
  const float* bias;
  const float* epi_tanh;
  long long M, N, K;
  int act, accumulate;
  int k_blocks, kb_per_split;
  int chunk_kb;
  int raw_stages, lo_stages;
  int mix;
  int dbg;
  long long* stamps;   // diagnostics: clock64 at the phase boundaries of CTA (0,0,0), or NULL
};

__device__ __forceinline__ uint32_t cluster_nctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_ctaid_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctaid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_ctaid_z() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctaid.z;" : "=r"(r));
  return r;
}
// tcgen05.commit of a CTA pair that lives in a larger cluster: arrive on the barrier at this offset in the CTAs of `mask`
__device__ __forceinline__ void umma_commit_pair(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ float4 ld_cluster_f4(uint32_t cluster_addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(cluster_addr)
               : "memory");
  return v;
}

// all-TF32 split, any number of converter threads: lo tile = tf32(centred remainder), same byte image as the raw tile
template <int VECS, int NTHR>
__device__ __forceinline__ void convert_tile_tf32(const uint8_t* raw_b, uint8_t* lo_b, int ct) {
  constexpr int PER = (VECS + NTHR - 1) / NTHR;
  const float4* raw = reinterpret_cast<const float4*>(raw_b);
  float4* lo = reinterpret_cast<float4*>(lo_b);
  float4 v[PER];
#pragma unroll
  for (int u = 0; u < PER; ++u)
    if (ct + u * NTHR < VECS) v[u] = raw[ct + u * NTHR];
#pragma unroll
  for (int u = 0; u < PER; ++u)
    if (ct + u * NTHR < VECS) lo[ct + u * NTHR] = split4_centered(v[u]);
}

template <bool A_MN, bool B_MN>
__device__ __forceinline__ void convert_stage(uint8_t* sa, uint8_t* sb, uint8_t* sa_lo, uint8_t* sb_lo, int ct, int mix) {
  if (mix) {
    if (A_MN) convert_tile_mix_mn<BM, CONV_THREADS>(sa, sa_lo, sa_lo + A_TILE / 2, ct);
    else convert_tile_mix_k<A_TILE / 16, CONV_THREADS>(sa, sa_lo, sa_lo + A_TILE / 2, ct);
    if (B_MN) convert_tile_mix_mn<BNL, CONV_THREADS>(sb, sb_lo, sb_lo + B_TILE / 2, ct);
    else convert_tile_mix_k<B_TILE / 16, CONV_THREADS>(sb, sb_lo, sb_lo + B_TILE / 2, ct);
  } else {
    convert_tile_tf32<A_TILE / 16, CONV_THREADS>(sa, sa_lo, ct);
    convert_tile_tf32<B_TILE / 16, CONV_THREADS>(sb, sb_lo, ct);
  }
}

struct Bars {
  uint32_t base;
  __device__ __forceinline__ uint32_t full(int r) const { return base + 8u * r; }
  __device__ __forceinline__ uint32_t empty(int r) const { return base + 8u * (MAX_STAGES + r); }
  __device__ __forceinline__ uint32_t conv(int l) const { return base + 8u * (2 * MAX_STAGES + l); }
  __device__ __forceinline__ uint32_t lofree(int l) const { return base + 8u * (3 * MAX_STAGES + l); }
  __device__ __forceinline__ uint32_t tfull(int a) const { return base + 8u * (4 * MAX_STAGES + a); }
  __device__ __forceinline__ uint32_t tempty(int a) const { return base + 8u * (4 * MAX_STAGES + 2 + a); }
  __device__ __forceinline__ uint32_t slot() const { return base + 8u * (4 * MAX_STAGES + 4); }
};

// NCTA = 1: one CTA per 128 x 128 tile and K slice; cluster = the S slices of the tile.
// NCTA = 2: a CTA PAIR (tcgen05 cta_group::2) per 256 x 256 tile and K slice; cluster = 2 x S CTAs (x = CTA of the pair,
//   z = slice).  As in mt_gemm_tc.cu each CTA holds its own 128 rows of A and of the accumulator and half of the B tile,
//   the leader's MMA thread drives both tensor cores, converters and drainers report to the LEADER's barriers and the
//   commits are multicast to both CTAs.  Per flop a pair tile moves half the operand bytes through shared memory of a
//   128 x 128 tile - the K loop of the small tile is bound by the shared-memory port (~0.8 us per k-block, measured), not
//   by the tensor pipe - so problems with enough 256 x 256 tiles take this flavour.
template <bool A_MN, bool B_MN, int NCTA>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_mid_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, const Params p) {
  constexpr bool PAIR = NCTA == 2;
  constexpr int BN = BNL * NCTA;                 // accumulator columns = tile width
  constexpr int TMEM_COLS = 2 * BN;              // two chunk accumulators
  constexpr int HALF = BN / 2;                   // accumulator columns per drain thread
  constexpr int ROW4 = BN / 4;                   // float4 per row of the partial tile
  const int R = p.raw_stages, L = p.lo_stages;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t lo_base = smem_base + R * RAW_STAGE;
  const Bars bar{lo_base + uint32_t(L) * LO_STAGE};
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + R * RAW_STAGE + L * LO_STAGE + 8 * (4 * MAX_STAGES + 4));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cx = PAIR ? cluster_ctaid_x() : 0u;            // CTA of the pair (0 = leader)
  const uint32_t slice = cluster_ctaid_z();                     // this CTA's K slice
  const uint32_t S = cluster_nctarank() / NCTA;
  const uint32_t leader = slice * NCTA;                         // cluster rank of the pair's leader (x is the fastest dim)
  const int kb0 = int(slice) * p.kb_per_split;
  const int kb1 = min(kb0 + p.kb_per_split, p.k_blocks);
  const int nkb = kb1 > kb0 ? kb1 - kb0 : 0;    // an empty slice contributes a zero partial
  const int chunk_kb = p.chunk_kb;
  const int m0 = blockIdx.x * BM;                               // pair: blockIdx.x = 2 * (256-row block) + cx
  const int n0 = blockIdx.y * BN;                               // tile origin; this CTA loads B rows n0 + cx * 128 ...

  // TMA of k-block i of this CTA's slice into raw stage `stage`
  auto issue_loads = [&](int i, int stage) {
    const uint32_t sa = smem_base + stage * RAW_STAGE;
    const uint32_t sb = sa + A_TILE;
    const uint32_t fb = bar.full(stage);
    mbar_expect_tx(fb, RAW_STAGE);
    const int k0 = (kb0 + i) * BK;
    const int nb = n0 + int(cx) * BNL;
    if (A_MN) {
#pragma unroll
      for (int j = 0; j < BM / 32; ++j) tma_load_2d(sa + j * 4096, &map_a, fb, m0 + j * 32, k0);
    } else {
      tma_load_2d(sa, &map_a, fb, k0, m0);
    }
    if (B_MN) {
#pragma unroll
      for (int j = 0; j < BNL / 32; ++j) tma_load_2d(sb + j * 4096, &map_b, fb, nb + j * 32, k0);
    } else {
      tma_load_2d(sb, &map_b, fb, k0, nb);
    }
  };

  if (p.stamps && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 256) p.stamps[0] = clock64();
  const int pre = (p.dbg & 16) ? 0 : min(R, nkb);   // loads issued before the prologue barrier (dbg 16: none)
  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    for (int r = 0; r < R; ++r) {
      mbar_init(bar.full(r), 1);
      mbar_init(bar.empty(r), 1);
    }
    for (int l = 0; l < L; ++l) {
      mbar_init(bar.conv(l), CONV_WARPS * NCTA);
      mbar_init(bar.lofree(l), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar.tfull(a), 1);
      mbar_init(bar.tempty(a), 8 * NCTA);
    }
    fence_barrier_init();
    fence_proxy_async();
    // the first loads only touch this CTA's own barriers and shared memory: start them now, their latency
    // (~1.5 us) overlaps the TMEM allocation and the barrier below
    for (int i = 0; i < pre; ++i) issue_loads(i, i);
  }
  if (warp == 1) tmem_alloc_n<NCTA>(bar.slot(), TMEM_COLS);
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();   // pair: the peer's barriers must exist before remote arrives
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  const bool stamping = p.stamps && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 256;
  auto stamp = [&](int k) { if (stamping) p.stamps[k] = clock64(); };
  stamp(1);

  auto arrive_leader = [&](uint32_t b) {
    if (PAIR) mbar_arrive_cluster(mapa_shared(b, leader));
    else mbar_arrive(b);
  };
  // one k-block of conversion work for converter thread `ct`, and the hand-over to the MMA warp
  auto convert_block = [&](int stage, int lo, uint32_t phase, uint32_t lphase, int ct) {
    mbar_wait(bar.lofree(lo), lphase ^ 1, 0x15);
    mbar_wait(bar.full(stage), phase, 0x13);
    uint8_t* sa = smem_gen + stage * RAW_STAGE;
    uint8_t* sb = sa + A_TILE;
    uint8_t* sa_lo = smem_gen + R * RAW_STAGE + lo * LO_STAGE;
    uint8_t* sb_lo = sa_lo + A_TILE;
    if (!(p.dbg & 1)) convert_stage<A_MN, B_MN>(sa, sb, sa_lo, sb_lo, ct, p.mix);
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) arrive_leader(bar.conv(lo));
  };

  if (warp < 8) {
    reg_dec<REGS_LOW>();
    if (warp == 0) {
      // ===================== TMA producer =====================
      if (lane == 0) {
        int stage = pre == R ? 0 : pre;
        uint32_t phase = pre == R ? 1 : 0;
        for (int i = pre; i < nkb; ++i) {
          mbar_wait(bar.empty(stage), phase ^ 1, 0x10);
          issue_loads(i, stage);
          if (++stage == R) { stage = 0; phase ^= 1; }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer (pair: the leader CTA only) =====================
      if (lane == 0 && cx == 0) {
        constexpr uint32_t idesc = make_idesc(BM * NCTA, BN, A_MN, B_MN);
        constexpr uint64_t adesc_base = make_smem_desc_base(A_MN);
        constexpr uint64_t bdesc_base = make_smem_desc_base(B_MN);
        constexpr uint32_t a_kstep = A_MN ? 1024u : 32u;
        constexpr uint32_t b_kstep = B_MN ? 1024u : 32u;
        constexpr uint32_t idesc16 = make_idesc_bf16(BM * NCTA, BN, A_MN, B_MN);
        constexpr uint64_t da16 = make_smem_desc_base_16(A_MN), db16 = make_smem_desc_base_16(B_MN);
        constexpr uint32_t a16step = A_MN ? 1024u : 32u, b16step = B_MN ? 1024u : 32u;
        const uint16_t pair_mask = uint16_t(3u << leader);
        auto commit = [&](uint32_t b) {
          if (PAIR) umma_commit_pair(b, pair_mask);
          else umma_commit_n<1>(b);
        };
        int stage = 0, lo = 0;
        uint32_t lphase = 0;
        int c = 0;
        for (int i0 = 0; i0 < nkb; i0 += chunk_kb, ++c) {
          const int acc = c & 1;
          mbar_wait(bar.tempty(acc), ((c >> 1) & 1) ^ 1, 0x11);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * BN;
          const int i1 = min(i0 + chunk_kb, nkb);
          for (int i = i0; i < i1; ++i) {
            mbar_wait(bar.conv(lo), lphase, 0x12);
            tc_fence_after();
            const uint32_t sa = smem_base + stage * RAW_STAGE;
            const uint32_t sb = sa + A_TILE;
            const uint32_t sa_lo = lo_base + lo * LO_STAGE;
            const uint32_t sb_lo = sa_lo + A_TILE;
            if (p.mix) {
              const uint32_t a16hi = sa_lo, a16lo = sa_lo + A_TILE / 2;
              const uint32_t b16hi = sb_lo, b16lo = sb_lo + B_TILE / 2;
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                umma_f16_n<NCTA>(d_tmem, smem_desc(da16, a16lo + j * a16step), smem_desc(db16, b16hi + j * b16step), idesc16,
                                 (i > i0) || (j > 0));
                umma_f16_n<NCTA>(d_tmem, smem_desc(da16, a16hi + j * a16step), smem_desc(db16, b16lo + j * b16step), idesc16,
                                 1u);
              }
#pragma unroll
              for (int j = 0; j < BK / UMMA_K; ++j)
                umma_tf32_n<NCTA>(d_tmem, smem_desc(adesc_base, sa + j * a_kstep), smem_desc(bdesc_base, sb + j * b_kstep),
                                  idesc, 1u);
            } else {
#pragma unroll
              for (int j = 0; j < BK / UMMA_K; ++j) {
                const uint64_t a_hi = smem_desc(adesc_base, sa + j * a_kstep);
                const uint64_t a_lo = smem_desc(adesc_base, sa_lo + j * a_kstep);
                const uint64_t b_hi = smem_desc(bdesc_base, sb + j * b_kstep);
                const uint64_t b_lo = smem_desc(bdesc_base, sb_lo + j * b_kstep);
                umma_tf32_n<NCTA>(d_tmem, a_lo, b_hi, idesc, (i > i0) || (j > 0));   // small terms first
                umma_tf32_n<NCTA>(d_tmem, a_hi, b_lo, idesc, 1u);
                umma_tf32_n<NCTA>(d_tmem, a_hi, b_hi, idesc, 1u);
              }
            }
            commit(bar.empty(stage));
            commit(bar.lofree(lo));
            if (++stage == R) stage = 0;
            if (++lo == L) { lo = 0; lphase ^= 1; }
          }
          commit(bar.tfull(acc));
        }
      }
    } else {
      // ===================== converters, low-register flavour (warps 2..7) =====================
      const int ct = threadIdx.x - 64;
      int stage = 0, lo = 0;
      uint32_t phase = 0, lphase = 0;
      for (int i = 0; i < nkb; ++i) {
        convert_block(stage, lo, phase, lphase, ct);
        if (++stage == R) { stage = 0; phase ^= 1; }
        if (++lo == L) { lo = 0; lphase ^= 1; }
      }
    }
    // the fold + epilogue below is common code: every warp goes back to the launch allocation.  The registers come from
    // the high warps' hand-back, so wait for it (named barrier 1): a CTA with a short or empty K slice gets here before
    // the high warps have even taken their 200 - growing first would leave them blocked in setmaxnreg.inc for ever
    asm volatile("bar.sync 1, 512;" ::: "memory");
    reg_inc<REGS_COMMON>();
  } else {
    // ===================== converters + chunk drain + cluster fold + epilogue (warps 8..15) =====================
    reg_inc<REGS_HIGH>();
    const int ct = threadIdx.x - 64;
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may touch
    const int half = (warp - 8) >> 2;          // which half of the accumulator columns
    const uint32_t lane_base = uint32_t(quad * 32) << 16;
    float run[HALF];
    const int nchunks = (nkb + chunk_kb - 1) / chunk_kb;
    const int late = chunk_kb >= 2 ? 1 : 0;    // drain a chunk this many k-blocks after its last conversion
    int drained = 0;
    auto drain = [&](int c) {
      const int acc = c & 1;
      mbar_wait(bar.tfull(acc), (c >> 1) & 1, 0x14);
      tc_fence_after();
      const uint32_t taddr = tmem_base + lane_base + acc * BN + half * HALF;
#pragma unroll
      for (int q = 0; q < HALF / 32; ++q) {
        uint32_t r[32];
        tmem_ld32(taddr + q * 32, r);
        tmem_ld_wait();
        if (c == 0) {
#pragma unroll
          for (int j = 0; j < 32; ++j) run[q * 32 + j] = __uint_as_float(r[j]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) run[q * 32 + j] += __uint_as_float(r[j]);   // IEEE RN fold of the chunks
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_leader(bar.tempty(acc));
    };
    {
      int stage = 0, lo = 0;
      uint32_t phase = 0, lphase = 0;
      for (int i = 0; i < nkb; ++i) {
        convert_block(stage, lo, phase, lphase, ct);
        if (i == 0) stamp(2);
        if (++stage == R) { stage = 0; phase ^= 1; }
        if (++lo == L) { lo = 0; lphase ^= 1; }
        while (drained < nchunks && (drained + 1) * chunk_kb - 1 + late <= i) drain(drained++);
      }
      stamp(3);
      while (drained < nchunks) drain(drained++);
      stamp(4);
    }
    if (nkb == 0) {
#pragma unroll
      for (int j = 0; j < HALF; ++j) run[j] = 0.f;
    }
    // ---- partial tile -> own shared memory (the rings are dead: every MMA that reads this CTA's stages has retired,
    // the last tfull says so).  Row-major [128][BN] fp32, 16-byte chunk c of row r at chunk c ^ (r & 7): conflict-free
    // for the row-per-lane writes here and for the chunk-per-lane reads of the fold.
    {
      const int row = quad * 32 + lane;
      float4* st = reinterpret_cast<float4*>(smem_gen);
#pragma unroll
      for (int j = 0; j < HALF / 4; ++j) {
        const int c4 = half * (HALF / 4) + j;
        st[row * ROW4 + (c4 ^ (row & 7))] = make_float4(run[4 * j], run[4 * j + 1], run[4 * j + 2], run[4 * j + 3]);
      }
    }
    reg_dec<REGS_COMMON>();
    asm volatile("bar.arrive 1, 512;" ::: "memory");
  }

  // ===================== cluster fold + epilogue: all 16 warps =====================
  // CTA (pair member cx, slice z) owns rows [z * rows_per, (z + 1) * rows_per) of its 128-row band: it sums the S partials
  // of those rows in slice order (deterministic) through distributed shared memory and stores them with coalesced rows.
  // A thread keeps its column chunk (512 is a multiple of the row length), so the bias is loaded once, before the
  // barrier; the loop is unrolled four items deep with every load of a batch issued before the arithmetic - written as a
  // plain loop this phase cost 5.3 us of a 8.7 us launch (profiles/r02_mid_kernel.jsonl, phase clocks).
  const bool solo = (S * NCTA == 1);             // no peer: CTA-scope barrier, plain shared-memory loads
  const int tid = threadIdx.x;
  const int c4 = tid % ROW4;
  const long long gn = (long long)n0 + c4 * 4;
  const bool col_ok = gn < p.N;                  // N % 4 == 0 (checked on the host)
  float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p.bias && col_ok) bias4 = __ldg(reinterpret_cast<const float4*>(p.bias + gn));
  stamp(5);
  if (solo) __syncthreads(); else cluster_sync_all();
  stamp(6);
  {
    const int rows_per = (BM + int(S) - 1) / int(S);
    const int row_begin = min(int(slice) * rows_per, BM);
    const int row_end = min(row_begin + rows_per, BM);
    constexpr int RSTEP = NUM_THREADS / ROW4;    // rows between two items of a thread
    constexpr int U = 4;
    uint32_t peer[MAX_SPLITS];
#pragma unroll
    for (int q = 0; q < MAX_SPLITS; ++q) peer[q] = solo ? smem_base : mapa_shared(smem_base, (q < int(S) ? q : 0) * NCTA + cx);
    for (int r0 = row_begin + tid / ROW4; r0 < row_end; r0 += RSTEP * U) {
      float4 acc[U], ext[U], old[U];
      bool ok[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int row = r0 + u * RSTEP;
        const long long gr = (long long)m0 + row;
        ok[u] = row < row_end && gr < p.M && col_ok;
        if (ok[u] && p.epi_tanh) ext[u] = __ldg(reinterpret_cast<const float4*>(p.epi_tanh + gr * p.N + gn));
        if (ok[u] && p.accumulate) old[u] = *reinterpret_cast<const float4*>(p.C + gr * p.N + gn);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int row = r0 + u * RSTEP;
        if (row < row_end) {
          const uint32_t off = uint32_t(row * ROW4 + (c4 ^ (row & 7))) << 4;
          if (solo) {
            acc[u] = *reinterpret_cast<const float4*>(smem_gen + off);
          } else {
            acc[u] = ld_cluster_f4(peer[0] + off);
#pragma unroll
            for (int q = 1; q < MAX_SPLITS; ++q)
              if (q < int(S)) {
                const float4 v = ld_cluster_f4(peer[q] + off);
                acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w;
              }
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (!ok[u]) continue;
        const long long gr = (long long)m0 + r0 + u * RSTEP;
        float4 v = acc[u];
        v.x = fmaf(v.x, kMixScale, bias4.x); v.y = fmaf(v.y, kMixScale, bias4.y);   // centred split: see mt_tc_common.cuh
        v.z = fmaf(v.z, kMixScale, bias4.z); v.w = fmaf(v.w, kMixScale, bias4.w);
        if (p.act == MT_ACT_TANH) { v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w); }
        if (p.epi_tanh) {
          const float4 y = ext[u];
          v.x = __fmul_rn(v.x, __fsub_rn(1.f, __fmul_rn(y.x, y.x)));
          v.y = __fmul_rn(v.y, __fsub_rn(1.f, __fmul_rn(y.y, y.y)));
          v.z = __fmul_rn(v.z, __fsub_rn(1.f, __fmul_rn(y.z, y.z)));
          v.w = __fmul_rn(v.w, __fsub_rn(1.f, __fmul_rn(y.w, y.w)));
        }
        if (p.accumulate) { v.x += old[u].x; v.y += old[u].y; v.z += old[u].z; v.w += old[u].w; }
        *reinterpret_cast<float4*>(p.C + gr * p.N + gn) = v;
      }
    }
  }
  stamp(7);
  if (!solo) cluster_sync_all();   // nobody leaves while a peer may still read its partial
  stamp(8);

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_n<NCTA>(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------- host side
static int g_mid_on = 1;
static long long g_mid_min_macs = 1ll << 25;   // below: the SIMT tile (launch-bound either way, SIMT launches lighter)
static long long g_mid_max_macs = 1ll << 32;   // from here up the persistent CTA-pair kernel is as fast or faster (its tile epilogues overlap the next tile's K loop)
static int g_mid_splits = 0;                   // 0 = choose per problem, else force this many K slices (1, 2, 4, 8)
static int g_mid_ncta = 0;                     // 0 = choose per problem, 1 = 128 x 128 tiles, 2 = 256 x 256 pair tiles
static int g_mid_dbg = 0;
static long long* g_mid_stamps = nullptr;     // device buffer of 16 clock64 stamps (mt_gemm_tc_mid_set_stamps)

static int ensure_debug_word() {
  static bool set = false;
  if (set) return MT_OK;
  int* dev = debug_word_device();
  if (!dev) return fail(MT_ERR_CUDA, "gemm_mid: no host-mapped debug word");
  MT_CUDA(cudaMemcpyToSymbol(g_dbg_word, &dev, sizeof(dev)));
  set = true;
  return MT_OK;
}

struct Plan { int ncta, splits; double cost; };

// CTAs that can be resident at once when launched as clusters of `cluster` CTAs (clusters do not span GPCs, so big
// clusters strand a few SMs per GPC); measured once per (kernel flavour, cluster size) with the occupancy API
template <bool A_MN, bool B_MN, int NCTA>
static long long resident_ctas(int cluster) {
  static long long cache[MAX_SPLITS * 2 + 1] = {0};
  if (cache[cluster]) return cache[cluster];
  cudaFuncSetAttribute(gemm_mid_kernel<A_MN, B_MN, NCTA>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(NCTA, 1, (unsigned)(cluster / NCTA));
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_MAX;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCTA;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = cluster / NCTA;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, gemm_mid_kernel<A_MN, B_MN, NCTA>, &cfg) != cudaSuccess || n < 1) {
    cudaGetLastError();
    n = g().sm_count / cluster;
  }
  cache[cluster] = (long long)n * cluster;
  return cache[cluster];
}

// Tile flavour and K slices per tile.  Cost model in microseconds per launch, fitted to the sweep and the phase clocks in
// profiles/r02_mid_kernel.jsonl (back-to-back launches, every (flavour, S) forced on seven shapes; the fit is within one
// 2 us launch tick of 45 of the 48 single-shape measurements):
//     T = 3.5 + waves * (k-blocks per slice * t_kb + fixed)
//   128 x 128 tile:  t_kb = 0.68 / 0.76 us (BF16-corrected / 3xTF32) - the shared-memory port: 160 KB per k-block at
//                    128 B/clk = 1250 cycles, the K loop measures 1254 - fixed = 6.2 us (prologue 0.6, first TMA round
//                    trip + conversion 0.8, last MMAs + drain 1.1, staging 0.5, two cluster barriers 0.7 each, fold +
//                    epilogue 2.5: tanhf and ~20 B/clk of DSMEM);
//   pair tile:       t_kb = 0.82 / 1.0 us for four times the flops, fixed = 10.2 us (13 us unsplit: a CTA then runs the
//                    epilogue of 128 x 256 outputs alone).
// A wave holds as many CTAs as the occupancy API says are resident for that cluster size (clusters do not span GPCs:
// 148 CTAs in clusters of 1 - 2, 132 - 135 in clusters of 3 - 6, 120 in clusters of 8).
template <bool A_MN, bool B_MN>
static Plan choose_plan(long long M, long long N, int k_blocks, int mix) {
  Plan best{1, 1, 1e30};
  for (int ncta = 1; ncta <= 2; ++ncta) {
    if (g_mid_ncta && ncta != g_mid_ncta) continue;
    const long long tiles = ceil_div(M, BM * ncta) * ceil_div(N, BNL * ncta);
    const double t_kb = ncta == 1 ? (mix ? 0.68 : 0.76) : (mix ? 0.82 : 1.0);
    const int s_max = MAX_SPLITS / ncta;
    for (int s = 1; s <= s_max; ++s) {
      if (g_mid_splits && s != std::min(g_mid_splits, s_max)) continue;
      if (!g_mid_splits && s > 1 && (s > k_blocks || ceil_div(k_blocks, s) * (s - 1) >= k_blocks)) continue;   // no empty slices
      const long long cap = ncta == 1 ? resident_ctas<A_MN, B_MN, 1>(s) : resident_ctas<A_MN, B_MN, 2>(2 * s);
      const long long waves = ceil_div(tiles * s * ncta, cap);
      const double fixed = ncta == 1 ? 6.2 : (s == 1 ? 13.0 : 10.2);
      const double cost = 3.5 + (double)waves * ((double)ceil_div(k_blocks, s) * t_kb + fixed);
      if (cost < best.cost - 1e-9) best = Plan{ncta, s, cost};
    }
  }
  return best;
}

template <bool A_MN, bool B_MN, int NCTA>
static int launch_plan(const GemmArgs& a, long long lda, long long ldb, int S, int mix, int chunk, int dbg) {
  constexpr int BN = BNL * NCTA;
  CUtensorMap ma, mb;
  int rc;
  if (A_MN) rc = make_map(&ma, a.A, a.M, a.K, lda, 32, BK, true);
  else rc = make_map(&ma, a.A, a.K, a.M, lda, BK, BM, false);
  if (rc) return rc;
  if (B_MN) rc = make_map(&mb, a.B, a.N, a.K, ldb, 32, BK, true);
  else rc = make_map(&mb, a.B, a.K, a.N, ldb, BK, BNL, false);
  if (rc) return rc;
  Params p;
  p.C = a.C;
  p.bias = a.bias;
  p.epi_tanh = a.epi_tanh;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.act = a.act;
  p.accumulate = a.accumulate;
  p.k_blocks = (int)ceil_div(a.K, BK);
  p.mix = mix;
  p.dbg = (dbg & 1) | (g_mid_dbg & 16);
  p.stamps = g_mid_stamps;
  const long long m_blocks = ceil_div(a.M, BM * NCTA), n_blocks = ceil_div(a.N, BN);
  if (n_blocks > 65535 || m_blocks * NCTA > 0x7fffffffLL) return MT_ERR_UNSUPPORTED;
  p.kb_per_split = (int)ceil_div(p.k_blocks, S);
  p.chunk_kb = chunk > 0 ? chunk : p.kb_per_split;
  {  // ring depths: never deeper than the K slice, but the partial tile (128 x BN fp32) needs room afterwards
    int l = std::min(3, std::max(1, p.kb_per_split));
    int r = std::min((SMEM_MAX - AUX - l * LO_STAGE) / RAW_STAGE, std::max(2, p.kb_per_split));
    if (r > MAX_STAGES) r = MAX_STAGES;
    const int need = BM * BN * 4 / RAW_STAGE;    // stages' worth of staging room (2 or 4)
    if (r + l < need) r = need - l;
    p.raw_stages = r;
    p.lo_stages = l;
  }
  const int smem_bytes = p.raw_stages * RAW_STAGE + p.lo_stages * LO_STAGE + AUX;
  rc = ensure_debug_word();
  if (rc) return rc;
  auto kern = gemm_mid_kernel<A_MN, B_MN, NCTA>;
  static bool attr_set = false;
  if (!attr_set) {
    MT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX));
    attr_set = true;
  }
  const double flops = 2.0 * (double)a.M * (double)a.N * (double)a.K;
  const int token = prof_begin(flops, flops * (p.mix ? 4.0 : 6.0));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(m_blocks * NCTA), (unsigned)n_blocks, (unsigned)S);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = g().stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCTA;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = S;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, ma, mb, p);
  prof_end(token);
  if (le == cudaErrorLaunchOutOfResources || le == cudaErrorInvalidConfiguration || le == cudaErrorInvalidClusterSize) {
    // a device (or partition) that cannot host this cluster shape: leave the problem to the other GPU engines instead of
    // failing the call - unless this kernel was asked for by name (engine 2)
    cudaGetLastError();
    if (g().gemm_engine_force != 2) return MT_ERR_UNSUPPORTED;
  }
  if (le != cudaSuccess) return fail(MT_ERR_CUDA, "gemm_mid launch failed: %s", cudaGetErrorString(le));
  g().launches++;
  g().gemm_mid_launches++;
  if (g().trace_on) {
    char label[56];
    snprintf(label, sizeof(label), "tc_mid%d %lldx%lldx%lld %c%c s%d", NCTA, a.M, a.N, a.K, A_MN ? 'm' : 'k', B_MN ? 'm' : 'k', S);
    trace_mark(label, g().stream, 0);
  }
  return MT_OK;
}

template <bool A_MN, bool B_MN>
static int launch(const GemmArgs& a, long long lda, long long ldb) {
  int mix, chunk, dbg;
  precision_for(a.forward != 0, &mix, &chunk, &dbg);
  const Plan plan = choose_plan<A_MN, B_MN>(a.M, a.N, (int)ceil_div(a.K, BK), mix);
  if (plan.ncta == 2) return launch_plan<A_MN, B_MN, 2>(a, lda, ldb, plan.splits, mix, chunk, dbg);
  return launch_plan<A_MN, B_MN, 1>(a, lda, ldb, plan.splits, mix, chunk, dbg);
}

}  // namespace tcm

namespace mt {

// does the mid-size kernel take this problem (shape regime only; layout checks happen in gemm_tc_mid_try)?
bool gemm_tc_mid_wants(const GemmArgs& a) {
  if (!tcm::g_mid_on) return false;
  if (g().gemm_engine_force == 2) return true;
  if (g().gemm_engine_force >= 0) return false;
  if (a.M < 128 || a.N < 64 || a.K < 32) return false;
  const long long macs = a.M * a.N * a.K;
  return macs >= tcm::g_mid_min_macs && macs < tcm::g_mid_max_macs;
}

int gemm_tc_mid_try(const GemmArgs& a) {
  if (a.tanh_src) return MT_ERR_UNSUPPORTED;       // no prologue flavour: callers hand over an explicit dZ
  bool a_mn, b_mn;
  long long lda, ldb;
  if (a.a_sk == 1 && a.a_sm >= a.K) { a_mn = false; lda = a.a_sm; }
  else if (a.a_sm == 1 && a.a_sk >= a.M) { a_mn = true; lda = a.a_sk; }
  else return MT_ERR_UNSUPPORTED;
  if (a.b_sk == 1 && a.b_sn >= a.K) { b_mn = false; ldb = a.b_sn; }
  else if (a.b_sn == 1 && a.b_sk >= a.N) { b_mn = true; ldb = a.b_sk; }
  else return MT_ERR_UNSUPPORTED;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(a.A) || !al16(a.B) || !al16(a.C) || (a.bias && !al16(a.bias)) || (a.epi_tanh && !al16(a.epi_tanh)))
    return MT_ERR_UNSUPPORTED;
  if (lda % 4 || ldb % 4 || a.N % 4) return MT_ERR_UNSUPPORTED;
  if (a.M < 1 || a.N < 4 || a.K < 1) return MT_ERR_UNSUPPORTED;
  if (!a_mn && !b_mn) return tcm::launch<false, false>(a, lda, ldb);
  if (!a_mn && b_mn) return tcm::launch<false, true>(a, lda, ldb);
  if (a_mn && !b_mn) return tcm::launch<true, false>(a, lda, ldb);
  return tcm::launch<true, true>(a, lda, ldb);
}

}  // namespace mt

// diagnostic knobs (tools/engine_crossover.py): on/off, the multiply-add window the dispatcher gives this kernel
// (log2 bounds), and a forced number of K slices (0 = chosen per problem)
extern "C" __attribute__((visibility("default"))) int mt_gemm_tc_mid_launches(uint64_t* count) {
  if (count) *count = mt::g().gemm_mid_launches;
  return MT_OK;
}
extern "C" __attribute__((visibility("default"))) int mt_gemm_tc_mid_set(int on, int min_log2_macs, int max_log2_macs,
                                                                          int splits) {
  tcm::g_mid_on = on ? 1 : 0;
  if (min_log2_macs > 0) tcm::g_mid_min_macs = 1ll << min_log2_macs;
  if (max_log2_macs > 0) tcm::g_mid_max_macs = 1ll << max_log2_macs;
  tcm::g_mid_splits = splits < 0 ? 0 : splits;
  return MT_OK;
}
// diagnostics: CTAs resident at once for clusters of 1..8 CTAs of the forward (K-major x K-major) kernels: out[0..7] the
// 128 x 128 flavour, out[8..15] the pair flavour (even cluster sizes only, else 0)
extern "C" __attribute__((visibility("default"))) int mt_gemm_tc_mid_resident(long long* out16) {
  MT_REQUIRE_INIT();
  for (int c = 1; c <= 8; ++c) {
    out16[c - 1] = tcm::resident_ctas<false, false, 1>(c);
    out16[8 + c - 1] = (c % 2 == 0) ? tcm::resident_ctas<false, false, 2>(c) : 0;
  }
  return MT_OK;
}
// diagnostics: phase clocks of CTA (0,0,0) go to this device buffer of >= 16 int64 (NULL = off): 0 entry, 1 after the
// prologue barrier, 2 first k-block converted, 3 last k-block converted, 4 last chunk drained, 5 partial staged,
// 6 after the cluster barrier, 7 fold + epilogue done, 8 after the final cluster barrier
extern "C" __attribute__((visibility("default"))) int mt_gemm_tc_mid_set_stamps(long long* device_buffer) {
  tcm::g_mid_stamps = device_buffer;
  return MT_OK;
}
// tile flavour: 0 = chosen per problem, 1 = 128 x 128 single-CTA tiles, 2 = 256 x 256 CTA-pair tiles; dbg bit 4 (16) = no
// TMA loads before the prologue barrier
extern "C" __attribute__((visibility("default"))) int mt_gemm_tc_mid_set_tile(int ncta, int dbg) {
  tcm::g_mid_ncta = (ncta == 1 || ncta == 2) ? ncta : 0;
  tcm::g_mid_dbg = dbg;
  return MT_OK;
}
