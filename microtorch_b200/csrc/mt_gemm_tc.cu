// mt_gemm_tc.cu - fp32-accurate GEMM on the Blackwell tensor cores (tcgen05 + TMEM + TMA), split-fp32.
//
// Replaces cuBLAS SGEMM behind the reference's `a.data @ b.data`, `grad @ b.T`, `a.T @ grad`
// (src/tensor/ops/overload.py:132,142,153) as used by Linear (src/nn/layer/linear.py:48).
//
//   C[M,N] (+)= act( A'[M,K] . B[K,N] + bias[N] )          fp32 in, fp32 out
//
// Error-compensated split done ON CHIP: x = hi + lo with hi = trunc_tf32(x), lo = x - hi (exact in fp32);
//   A.B ~= A_hi.B_hi + A_lo.B_hi + A_hi.B_lo       (lo.lo ~ 2^-22 is dropped), fp32 accumulate in TMEM.
// The leading product runs on TF32 from the RAW fp32 tile: TMA lands it in shared memory and the tensor core truncates
// it to TF32 itself, so the raw tile IS hi.  The two first-order corrections only need ~8 bits of each factor (they are
// 2^-10 of the result), so they run as BF16 products at twice the TF32 rate: four "converter" warps write
// bf16(x) and bf16(x - hi) - half the bytes of the fp32 tile each - into a second ring.  Per k-slice of 32: 4 TF32 MMAs
// (K8) + 4 BF16 MMAs (K16) = 8 tensor slots instead of the 12 of a 3xTF32 split: the GEMMs of config #2 are 1.18-1.41x
// faster, at the same measured accuracy (1e-6 forward, 1.6-1.8e-6 backward; tools/mix_bringup.py, tools/mix_ab.py).
// mt_gemm_tc_set_mix(0) selects the all-TF32 split (lo rounded to tf32, same layout as the raw tile); the backward
// prologue variant (dZ = dY * (1 - Y^2) fused into the conversion, small problems only) always uses it.
// No split copies ever touch HBM.  For large problems the tanh' factor is applied in the PRODUCING GEMM's epilogue
// (epi_tanh) instead of a prologue.
//
// Persistent grid: one CTA PAIR (cluster of 2, tcgen05 cta_group::2) per two SMs loops over 256 x 256 output tiles
// (single CTAs, 128 x 128 tiles, when N <= 128).  Each CTA = 16 warps in 4 warpgroups:
//   warp 0      TMA producer      (one lane)            full[r]  <- bytes landed (own A rows, own half of B)
//   warps 2-7   converters        (192 threads; 4-7 in the all-TF32 mode)  conv[l]  <- correction tiles ready, arrives on the LEADER's barrier
//   warp 1      MMA issuer        (leader CTA, 1 lane)  empty[r], lofree[l], tfull[a] <- tcgen05.commit multicast
//   warps 8-15  accumulate/epilogue (256 threads)       tempty[a] <- accumulator drained (leader's barrier)
// Two shared-memory rings: a deep raw ring (TMA destination) and a shallow correction ring (converter output).
// The tensor core adds into its fp32 TMEM accumulator with truncation (measured: error grows
// linearly, ~2^-24 per tcgen05.mma, 7.5e-5 at K=4096), so K is cut into chunks of CHUNK_KB
// k-blocks: each chunk accumulates into one of two TMEM buffers (ping-pong) and the epilogue
// warps fold finished chunks into IEEE fp32 running sums held in registers (setmaxnreg gives
// them 200 registers, the data-movement warps 56).  Chunk c+1's MMAs overlap chunk c's drain
// and the previous tile's bias/tanh/store epilogue.  DESIGN.md section 4.1 has the measurements behind each choice.
// Operands may be K-major or MN-major (UMMA descriptor transpose bits), so forward (X.W^T), dX
// (dZ.W) and dW (dZ^T.X) all read the row-major activations/weights in place - no transposes; the correction tiles
// keep the major of their source (K-major: SWIZZLE_64B rows of 32 k; MN-major: SWIZZLE_64B boxes of 32 k x 32 mn).
//
// Roofline: tensor pipe.  One k-block (32) of a 256x256 pair tile is 4 TF32 + 4 BF16 tcgen05.mma of 128 cycles each;
// logical 2*M*N*K flops are reported against (bf16 dense peak)/4 = (TF32 peak)/2.
#include <algorithm>
#include <vector>

#include "mt_common.cuh"
#include "mt_gemm.cuh"
#include "mt_tc_common.cuh"

using namespace mt;

namespace tc {

constexpr int NUM_THREADS = 512;  // 4 warpgroups
constexpr int NUM_ACC = 2;
constexpr int REGS_MOVERS = 56;   // producer / MMA / converter warpgroups
constexpr int REGS_EPILOGUE = 200;

struct Params {
  float* C;
  const float* bias;
  const float* epi_tanh;   // [M,N]: C = result * (1 - epi_tanh^2), or NULL
  long long M, N, K;
  int act, accumulate;
  int m_blocks, n_blocks, k_blocks;
  int chunk_kb;       // k-blocks accumulated in TMEM before a flush to the fp32 register sums
  int write_hi;       // converters store hi explicitly (else the MMA's own truncation of the raw fp32 is hi)
  int splits, kb_per_split;    // split-K: work item = (tile, split); split s covers k-blocks [s*kbps, (s+1)*kbps)
                               // and writes its partial to C + s*M*N (folded by splitk_finalize)
  int raw_stages, lo_stages;   // ring depths (runtime: the split of the 227 KB is a tuning knob)
  int group_n;        // tile order: n-blocks per column group (see work_decode)
  unsigned long long hint_a, hint_b;   // L2 eviction-priority policies of the A / B tile loads (0 = none)
  int store_cs;       // epilogue stores C with st.global.cs (streaming: the output is not re-read by this kernel)
  int mix;            // 1: corrections as BF16 products from K-major [A_hi16 | A_lo16 | B_hi16 | B_lo16] tiles in the lo ring
  int dbg;            // timing experiments only (results are wrong): 1 skip the conversion, 8 signal before converting, 2 issue hi*hi only,
                      // 4 issue no MMA at all
};

// Shared-memory plan.  Two rings with different lifetimes:
//   raw ring : TMA destination [A | B | (Y tile for the tanh' prologue)], alive from the TMA issue until the MMAs
//              that read it retire - deep (up to 5 stages), because the TMA round trip (~1.2-2.4 us measured,
//              profiles/r01_gemm_pipeline_experiments.jsonl) is longer than one k-block of tensor work (~0.85 us);
//   lo ring  : converter output [A_lo | B_lo], alive from the conversion until the same MMAs retire - 2 stages,
//              enough for the conversion of k-block i+1 to overlap the MMAs of k-block i.
template <int BN, int NCTA, bool PRO>
struct Cfg {
  static constexpr int BN_LOCAL = BN / NCTA;            // B rows this CTA loads (CTA pair: half of N each)
  static constexpr int A_TILE = BM * BK * 4;
  static constexpr int B_TILE = BN_LOCAL * BK * 4;
  static constexpr int RAW_STAGE = A_TILE + B_TILE + (PRO ? A_TILE : 0);
  static constexpr int LO_STAGE = A_TILE + B_TILE;
  static constexpr int SMEM_MAX = 232448;               // 227 KB opt-in limit per CTA
  static constexpr int MAX_STAGES = 8;                  // per ring (barrier arrays are sized for this)
  static constexpr int AUX = 1024 /*align slack*/ + 512 /*barriers*/;
  static constexpr int TMEM_COLS = (NUM_ACC * BN <= 256) ? 256 : 512;
  static constexpr int EPI_COLS = BN / 2;               // accumulator columns per epilogue thread
  static int smem_bytes(int r, int l) { return r * RAW_STAGE + l * LO_STAGE + AUX; }
};


// NCTA = 1: one CTA per 128 x BN tile.
// NCTA = 2: a CTA PAIR (cluster of 2, tcgen05 cta_group::2) per 256 x BN tile.  Each CTA holds its own 128
//   rows of A and of the accumulator (its own TMEM) and HALF of the B tile; the leader's single MMA thread
//   drives both tensor cores, which share the two B halves - so per SM the operand traffic out of shared
//   memory drops by a third and a stage shrinks from 96 KB to 64 KB (3 stages instead of 2).
//   TMA and conversion stay CTA-local; the hand-offs that cross the pair are: converters -> leader's conv[s]
//   (remote mbarrier arrive), MMA -> both CTAs' empty[s] / tfull[a] (multicast commit), epilogues -> leader's tempty[a].
template <int BN, bool A_MN, bool B_MN, int NCTA, bool PRO>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_splitfp32_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                  const __grid_constant__ CUtensorMap map_t, const Params p) {
  using C = Cfg<BN, NCTA, PRO>;
  constexpr int MS = C::MAX_STAGES;
  const int R = p.raw_stages, L = p.lo_stages;
  constexpr bool PAIR = (NCTA == 2);
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  // layout: [raw ring][lo ring][barriers]
  const uint32_t lo_base = smem_base + R * C::RAW_STAGE;
  const uint32_t bar_base = lo_base + L * C::LO_STAGE;
  auto full_bar = [&](int r) { return bar_base + 8u * r; };                 // raw stage r: TMA bytes landed
  auto empty_bar = [&](int r) { return bar_base + 8u * (MS + r); };         // raw stage r: MMAs retired
  auto conv_bar = [&](int l) { return bar_base + 8u * (2 * MS + l); };      // lo stage l: k-block converted
  auto lofree_bar = [&](int l) { return bar_base + 8u * (3 * MS + l); };    // lo stage l: MMAs retired
  auto tfull_bar = [&](int a) { return bar_base + 8u * (4 * MS + a); };
  auto tempty_bar = [&](int a) { return bar_base + 8u * (4 * MS + NUM_ACC + a); };
  constexpr int kSlot = 8 * (4 * MS + 2 * NUM_ACC);
  const uint32_t tmem_slot = bar_base + kSlot;
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + R * C::RAW_STAGE + L * C::LO_STAGE + kSlot);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = PAIR ? cluster_ctarank() : 0u;
  const int unit = PAIR ? (blockIdx.x >> 1) : blockIdx.x;          // CTA (or CTA pair) index
  const int num_units = PAIR ? (gridDim.x >> 1) : gridDim.x;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&map_a);
    prefetch_tmap(&map_b);
    if (PRO) prefetch_tmap(&map_t);
    for (int r = 0; r < R; ++r) {
      mbar_init(full_bar(r), 1);
      mbar_init(empty_bar(r), 1);
    }
    for (int l = 0; l < L; ++l) {
      mbar_init(conv_bar(l), ((!PRO && p.mix) ? 6 : 4) * NCTA);   // converter warps of every CTA of the pair
      mbar_init(lofree_bar(l), 1);
    }
    for (int a = 0; a < NUM_ACC; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8 * NCTA);    // epilogue warps of every CTA of the pair
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_n<NCTA>(tmem_slot, C::TMEM_COLS);
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  const int num_tiles = p.m_blocks * p.n_blocks;
  const int num_work = num_tiles * p.splits;   // (tile of (128*NCTA) x BN, k-split) pairs
  const int chunk_kb = p.chunk_kb;
  // Work item w -> (k-split, m-block, n-block).  The units take items round-robin, so the items [i*U, (i+1)*U) run at the
  // same time ("a wave") and should share operand panels in L2:
  //   * the k-split is the SLOWEST index: a wave reads one K range, so all its tiles walk K in step and every A / B
  //     k-slice is fetched from HBM once and shared through L2 by the tiles of its row / column;
  //   * inside a split the n-blocks are cut into column groups of `group_n`; a group is walked m-block by m-block
  //     with its n-blocks fastest.  A wave is then a (U / group_n) x group_n rectangle of tiles - it touches
  //     U/group_n + group_n operand panels instead of U/n_blocks + n_blocks - and the group's B panels stay hot in L2
  //     from wave to wave while A streams through once per group.
  // group_n == n_blocks is the plain row-major order.
  const int group_items = p.m_blocks * p.group_n;
  auto work_split = [&](int w) { return w / num_tiles; };
  auto work_mn = [&](int w, int& mb, int& nb) {
    const int t = w % num_tiles;
    const int grp = t / group_items;
    const int r = t - grp * group_items;
    const int width = min(p.group_n, p.n_blocks - grp * p.group_n);
    mb = r / width;
    nb = grp * p.group_n + (r - mb * width);
  };
  auto work_kb0 = [&](int w) { return work_split(w) * p.kb_per_split; };
  auto work_kb1 = [&](int w) { return min((work_split(w) + 1) * p.kb_per_split, p.k_blocks); };

  if (warp < 8) {
    reg_dec<REGS_MOVERS>();
    if (warp == 0) {
      // ===================== TMA producer (every CTA loads its own A rows and its share of B) =====================
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t tx_bytes = C::RAW_STAGE;
        for (int w = unit; w < num_work; w += num_units) {
          int mb, nb;
          work_mn(w, mb, nb);
          const int m0 = mb * (BM * NCTA) + int(cta_rank) * BM;
          const int n0 = nb * BN + int(cta_rank) * C::BN_LOCAL;
          for (int kb = work_kb0(w), kbe = work_kb1(w); kb < kbe; ++kb) {
            mbar_wait(empty_bar(stage), phase ^ 1, 0);
            const uint32_t sa = smem_base + stage * C::RAW_STAGE;
            const uint32_t sb = sa + C::A_TILE;
            const uint32_t sa_lo = sb + C::B_TILE;        // prologue only: the Y tile rides in the raw stage
            const uint32_t fb = full_bar(stage);
            mbar_expect_tx(fb, tx_bytes);
            const int k0 = kb * BK;
            if (A_MN) {
#pragma unroll
              for (int j = 0; j < BM / 32; ++j) {   // 32(m) x 32(k) boxes, 4096 B each
                tma_load_2d_hint(sa + j * 4096, &map_a, fb, m0 + j * 32, k0, p.hint_a);
                if (PRO) tma_load_2d(sa_lo + j * 4096, &map_t, fb, m0 + j * 32, k0);
              }
            } else {
              tma_load_2d_hint(sa, &map_a, fb, k0, m0, p.hint_a);    // 32(k) x 128(m) box
              if (PRO) tma_load_2d(sa_lo, &map_t, fb, k0, m0);
            }
            if (B_MN) {
#pragma unroll
              for (int j = 0; j < C::BN_LOCAL / 32; ++j) tma_load_2d_hint(sb + j * 4096, &map_b, fb, n0 + j * 32, k0, p.hint_b);
            } else {
              tma_load_2d_hint(sb, &map_b, fb, k0, n0, p.hint_b);
            }
            if (++stage == R) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      // ===================== MMA issuer (leader CTA only) =====================
      if (lane == 0 && cta_rank == 0) {
        constexpr uint32_t idesc = make_idesc(BM * NCTA, BN, A_MN, B_MN);
        constexpr uint64_t adesc_base = make_smem_desc_base(A_MN);
        constexpr uint64_t bdesc_base = make_smem_desc_base(B_MN);
        constexpr uint32_t a_kstep = A_MN ? 1024u : 32u;   // bytes per UMMA_K slice
        constexpr uint32_t b_kstep = B_MN ? 1024u : 32u;
        int stage = 0, lo = 0;
        uint32_t lphase = 0;
        uint32_t cc = 0;                                   // chunk counter across tiles -> TMEM buffer + phase
        for (int w = unit; w < num_work; w += num_units) {
          const int kb_end = work_kb1(w);
          for (int kb0 = work_kb0(w); kb0 < kb_end; kb0 += chunk_kb, ++cc) {
            const int acc = cc & 1;
            mbar_wait_t<PAIR>(tempty_bar(acc), ((cc >> 1) & 1) ^ 1, 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + acc * BN;
            const int kb1 = min(kb0 + chunk_kb, kb_end);
            for (int kb = kb0; kb < kb1; ++kb) {
              mbar_wait_t<PAIR>(conv_bar(lo), lphase, 2);
              tc_fence_after();
              const uint32_t sa = smem_base + stage * C::RAW_STAGE;
              const uint32_t sb = sa + C::A_TILE;
              const uint32_t sa_lo = lo_base + lo * C::LO_STAGE;
              const uint32_t sb_lo = sa_lo + C::A_TILE;
              if (!PRO && p.mix) {
                // mix mode: two BF16 correction products (2 x K16 each), then hi.hi on TF32 (4 x K8)
                constexpr uint32_t idesc16 = make_idesc_bf16(BM * NCTA, BN, A_MN, B_MN);
                constexpr uint64_t da16 = make_smem_desc_base_16(A_MN), db16 = make_smem_desc_base_16(B_MN);
                constexpr uint32_t a16step = A_MN ? 1024u : 32u, b16step = B_MN ? 1024u : 32u;   // bytes per K16 slice
                const uint32_t a16hi = sa_lo, a16lo = sa_lo + C::A_TILE / 2;
                const uint32_t b16hi = sb_lo, b16lo = sb_lo + C::B_TILE / 2;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  umma_f16_n<NCTA>(d_tmem, smem_desc(da16, a16lo + j * a16step), smem_desc(db16, b16hi + j * b16step), idesc16,
                                   (kb > kb0) || (j > 0));
                  umma_f16_n<NCTA>(d_tmem, smem_desc(da16, a16hi + j * a16step), smem_desc(db16, b16lo + j * b16step), idesc16,
                                   1u);
                }
#pragma unroll
                for (int j = 0; j < BK / UMMA_K; ++j)
                  umma_tf32_n<NCTA>(d_tmem, smem_desc(adesc_base, sa + j * a_kstep), smem_desc(bdesc_base, sb + j * b_kstep),
                                    idesc, 1u);
              } else
#pragma unroll
              for (int j = 0; j < BK / UMMA_K; ++j) {
                const uint64_t a_hi = smem_desc(adesc_base, sa + j * a_kstep);
                const uint64_t a_lo = smem_desc(adesc_base, sa_lo + j * a_kstep);
                const uint64_t b_hi = smem_desc(bdesc_base, sb + j * b_kstep);
                const uint64_t b_lo = smem_desc(bdesc_base, sb_lo + j * b_kstep);
                if (p.dbg & 6) {                                                       // timing experiments
                  if (p.dbg & 2) umma_tf32_n<NCTA>(d_tmem, a_hi, b_hi, idesc, (kb > kb0) || (j > 0));
                  continue;
                }
                umma_tf32_n<NCTA>(d_tmem, a_lo, b_hi, idesc, (kb > kb0) || (j > 0));   // small terms first
                umma_tf32_n<NCTA>(d_tmem, a_hi, b_lo, idesc, 1u);
                umma_tf32_n<NCTA>(d_tmem, a_hi, b_hi, idesc, 1u);
              }
              umma_commit_n<NCTA>(empty_bar(stage));       // both rings' slots reusable (in both CTAs) once these
              umma_commit_n<NCTA>(lofree_bar(lo));         // MMAs retire
              if (++stage == R) stage = 0;
              if (++lo == L) { lo = 0; lphase ^= 1; }
            }
            umma_commit_n<NCTA>(tfull_bar(acc));           // chunk accumulator complete (both CTAs)
          }
        }
      }
    } else if (warp >= 4 || (!PRO && p.mix)) {
      // ===================== converters (warps 4..7; in mix mode warps 2..7 = 192 threads) =====================
      const bool wide = !PRO && p.mix;
      const int ct = wide ? threadIdx.x - 64 : threadIdx.x - 128;
      int stage = 0, lo = 0;
      uint32_t phase = 0, lphase = 0;
      for (int w = unit; w < num_work; w += num_units) {
        for (int kb = work_kb0(w), kbe = work_kb1(w); kb < kbe; ++kb) {
          mbar_wait(lofree_bar(lo), lphase ^ 1, 5);          // lo slot free (usually long since)
          mbar_wait(full_bar(stage), phase, 3);              // raw bytes landed
          uint8_t* sa = smem_gen + stage * C::RAW_STAGE;
          uint8_t* sb = sa + C::A_TILE;
          uint8_t* sy = sb + C::B_TILE;                      // prologue only: Y tile
          uint8_t* sa_lo = smem_gen + R * C::RAW_STAGE + lo * C::LO_STAGE;
          uint8_t* sb_lo = sa_lo + C::A_TILE;
          if (p.dbg & 8) {
            // timing experiment (wrong results): signal first, convert afterwards - the MMAs never wait for the
            // conversion, which still uses the shared-memory port and the issue slots
            __syncwarp();
            if (lane == 0) {
              if (PAIR) mbar_arrive_cluster(mapa_shared(conv_bar(lo), 0));
              else mbar_arrive(conv_bar(lo));
            }
          }
          if (p.dbg & 1) {
            // timing experiment: hand the stage over untouched
          } else {
            if (!PRO && p.mix) {
              if (A_MN) convert_tile_mix_mn<BM, 192>(sa, sa_lo, sa_lo + C::A_TILE / 2, ct);
              else convert_tile_mix_k<C::A_TILE / 16, 192>(sa, sa_lo, sa_lo + C::A_TILE / 2, ct);
              if (B_MN) convert_tile_mix_mn<C::BN_LOCAL, 192>(sb, sb_lo, sb_lo + C::B_TILE / 2, ct);
              else convert_tile_mix_k<C::B_TILE / 16, 192>(sb, sb_lo, sb_lo + C::B_TILE / 2, ct);
            } else {
            if (PRO) convert_tile_prologue<C::A_TILE / 16>(sa, sy, sa_lo, ct, !p.write_hi);
            else if (p.write_hi) convert_tile<C::A_TILE / 16, true>(sa, sa_lo, ct);
            else convert_tile<C::A_TILE / 16, false>(sa, sa_lo, ct);
            if (p.write_hi) convert_tile<C::B_TILE / 16, true>(sb, sb_lo, ct);
            else convert_tile<C::B_TILE / 16, false>(sb, sb_lo, ct);
            }
          }
          fence_proxy_async();          // generic-proxy writes -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0 && !(p.dbg & 8)) {
            if (PAIR) mbar_arrive_cluster(mapa_shared(conv_bar(lo), 0));      // the LEADER's barrier
            else mbar_arrive(conv_bar(lo));
          }
          if (++stage == R) { stage = 0; phase ^= 1; }
          if (++lo == L) { lo = 0; lphase ^= 1; }
        }
      }
    }
  } else {
    // ===================== accumulate + epilogue (warps 8..15) =====================
    reg_inc<REGS_EPILOGUE>();
    constexpr int COLS = C::EPI_COLS;
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may touch
    const int half = (warp - 8) >> 2;          // which half of the accumulator columns
    const uint32_t lane_base = uint32_t(quad * 32) << 16;
    float run[COLS];
    uint32_t cc = 0;
    for (int w = unit; w < num_work; w += num_units) {
      int mb, nb;
      work_mn(w, mb, nb);
      const long long m0 = (long long)mb * (BM * NCTA) + (long long)cta_rank * BM;
      const long long n0 = (long long)nb * BN;
      const int kb_begin = work_kb0(w), kb_end = work_kb1(w);
      for (int kb0 = kb_begin; kb0 < kb_end; kb0 += chunk_kb, ++cc) {
        const int acc = cc & 1;
        if (kb0 + chunk_kb >= kb_end && (p.epi_tanh || p.accumulate)) {
          // last chunk of the tile: its MMAs take a few microseconds - pull this thread's row of the epilogue's DRAM
          // operand (the tanh' source, or the old C) into L2 now, so that the epilogue's loads are L2 hits
          const long long prow = m0 + quad * 32 + lane;
          if (prow < p.M) {
            const float* src = p.epi_tanh ? p.epi_tanh + prow * p.N
                                          : p.C + (long long)work_split(w) * p.M * p.N + prow * p.N;
#pragma unroll
            for (int c = 0; c < COLS; c += 32) {
              const long long n = n0 + half * COLS + c;
              if (n < p.N) asm volatile("prefetch.global.L2 [%0];" ::"l"(src + n));
            }
          }
        }
        mbar_wait(tfull_bar(acc), (cc >> 1) & 1, 4);
        tc_fence_after();
        const uint32_t taddr = tmem_base + lane_base + acc * BN + half * COLS;
#pragma unroll
        for (int c = 0; c < COLS / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          if (kb0 == kb_begin) {
#pragma unroll
            for (int j = 0; j < 32; ++j) run[c * 32 + j] = __uint_as_float(r[j]);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) run[c * 32 + j] += __uint_as_float(r[j]);   // IEEE RN fold
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {                                   // buffer back to the (leader's) MMA warp
          if (PAIR) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));
          else mbar_arrive(tempty_bar(acc));
        }
      }
      // ---- tile epilogue: (mix-mode scale), bias, tanh, tanh' of the producer, (accumulate), store.  One row per thread,
      // 128 B per 32 columns.  The global operand of the epilogue - the tanh' source (dX), the old C (accumulate) or the bias
      // (forward) - is loaded EB vectors at a time BEFORE any arithmetic or store of the batch: written as one loop the
      // compiler had to keep load -> math -> store order per vector (the stores may alias the next load), i.e. 32
      // serialised global-memory round trips per tile - ~38 us for the dX epilogue, during which these warps do not
      // drain TMEM chunks and the MMA warp stalls after two chunks (ncu: tensor pipe 73 % active in dX against 81 % in dW).
      const bool mix_scale = p.mix || !p.write_hi;   // centred split (every mode but write_hi): accumulator scaled by S
      const long long row = m0 + quad * 32 + lane;
      if (row < p.M) {
        float* crow = p.C + (long long)work_split(w) * p.M * p.N + row * p.N;   // split partials are stacked
        const float* erow = p.epi_tanh ? p.epi_tanh + row * p.N : nullptr;
        const long long nb0 = n0 + half * COLS;
        constexpr int EB = 4;                                        // vectors per batch
        const int emode = p.epi_tanh ? 1 : (p.accumulate ? 2 : (p.bias ? 3 : 0));   // which operand is batched
#pragma unroll
        for (int j0 = 0; j0 < COLS; j0 += 4 * EB) {
          float4 aux[EB];
#pragma unroll
          for (int u = 0; u < EB; ++u) {
            const long long n = nb0 + j0 + 4 * u;
            if (n < p.N) {                     // N % 4 == 0 (checked on the host)
              if (emode == 1) aux[u] = __ldg(reinterpret_cast<const float4*>(erow + n));
              else if (emode == 2) aux[u] = *reinterpret_cast<const float4*>(crow + n);
              else if (emode == 3) aux[u] = __ldg(reinterpret_cast<const float4*>(p.bias + n));
            }
          }
#pragma unroll
          for (int u = 0; u < EB; ++u) {
            const int j = j0 + 4 * u;
            const long long n = nb0 + j;
            if (n < p.N) {
              float4 v = make_float4(run[j], run[j + 1], run[j + 2], run[j + 3]);
              if (mix_scale) { v.x *= kMixScale; v.y *= kMixScale; v.z *= kMixScale; v.w *= kMixScale; }
              if (p.bias) {
                const float4 b = emode == 3 ? aux[u] : __ldg(reinterpret_cast<const float4*>(p.bias + n));
                v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
              }
              if (p.act == MT_ACT_TANH) { v.x = tanhf(v.x); v.y = tanhf(v.y); v.z = tanhf(v.z); v.w = tanhf(v.w); }
              if (emode == 1) {
                const float4 t = aux[u];
                v.x = __fmul_rn(v.x, __fsub_rn(1.f, __fmul_rn(t.x, t.x)));
                v.y = __fmul_rn(v.y, __fsub_rn(1.f, __fmul_rn(t.y, t.y)));
                v.z = __fmul_rn(v.z, __fsub_rn(1.f, __fmul_rn(t.z, t.z)));
                v.w = __fmul_rn(v.w, __fsub_rn(1.f, __fmul_rn(t.w, t.w)));
              }
              float4* dst = reinterpret_cast<float4*>(crow + n);
              if (p.accumulate) {
                const float4 o = emode == 2 ? aux[u] : *dst;
                v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
              }
              if (p.store_cs) __stcs(dst, v);
              else *dst = v;
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();   // nobody leaves while the peer may still touch its smem / TMEM
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_n<NCTA>(tmem_base, C::TMEM_COLS);
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D fp32 tensor map: dim0 = contiguous extent `inner`, dim1 = `outer` rows `ld` elements apart.
// Encoded maps are kept in a small cache keyed by everything that goes into them: a training loop launches the same
// GEMMs on the same pool blocks step after step, and three driver calls per launch were a measurable part of the
// launch floor of small problems.
struct MapKey {
  const float* base;
  long long inner, outer, ld;
  int box_inner, box_outer, mn;
  bool operator==(const MapKey& o) const {
    return base == o.base && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner &&
           box_outer == o.box_outer && mn == o.mn;
  }
};
struct MapSlot { MapKey key; CUtensorMap map; bool used; };
static MapSlot g_map_cache[64];
static int g_map_next = 0;
static int g_map_cache_on = 1;

int make_map(CUtensorMap* m, const float* base, long long inner, long long outer, long long ld, int box_inner,
             int box_outer, bool mn_major) {
  const MapKey key{base, inner, outer, ld, box_inner, box_outer, mn_major ? 1 : 0};
  if (g_map_cache_on)
    for (const MapSlot& sl : g_map_cache)
      if (sl.used && sl.key == key) { *m = sl.map; return MT_OK; }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(MT_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(MT_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) inner=%lld outer=%lld ld=%lld box=%dx%d", (int)r,
                inner, outer, ld, box_inner, box_outer);
  if (g_map_cache_on) {
    MapSlot& sl = g_map_cache[g_map_next];
    g_map_next = (g_map_next + 1) % 64;
    sl.key = key;
    sl.map = *m;
    sl.used = true;
  }
  return MT_OK;
}

// host-mapped word the device writes before trapping on a barrier timeout (shared by both tcgen05 translation units:
// each has its own copy of the device-side pointer variable, see mt_tc_common.cuh)
static int* g_dbg_host = nullptr;
static int* g_dbg_dev = nullptr;
int* debug_word_device() {
  if (g_dbg_dev) return g_dbg_dev;
  if (cudaHostAlloc((void**)&g_dbg_host, 4 * sizeof(int), cudaHostAllocMapped) != cudaSuccess) { g_dbg_host = nullptr; return nullptr; }
  memset(g_dbg_host, 0, 4 * sizeof(int));
  if (cudaHostGetDevicePointer((void**)&g_dbg_dev, g_dbg_host, 0) != cudaSuccess) g_dbg_dev = nullptr;
  return g_dbg_dev;
}
static int ensure_debug_word() {
  static bool set = false;
  if (set) return MT_OK;
  int* dev = debug_word_device();
  if (!dev) return fail(MT_ERR_CUDA, "gemm_splitfp32: no host-mapped debug word");
  MT_CUDA(cudaMemcpyToSymbol(g_dbg_word, &dev, sizeof(dev)));
  set = true;
  return MT_OK;
}

// per-launch event timing for the bench roofline (mt_prof_*)
struct ProfRec { cudaEvent_t a, b; double flops; double unit_flops; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
static std::vector<cudaEvent_t> g_prof_pool;
static cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}

// the same bookkeeping for the mid-size kernel (mt_gemm_tc_mid.cu): token = index into g_prof, -1 when profiling is off
int prof_begin(double flops, double unit_flops) {
  if (!g_prof_on) return -1;
  ProfRec rec{};
  rec.a = prof_event();
  rec.b = prof_event();
  rec.flops = flops;
  rec.unit_flops = unit_flops;
  cudaEventRecord(rec.a, g().stream);
  g_prof.push_back(rec);
  return (int)g_prof.size() - 1;
}
void prof_end(int token) {
  if (token >= 0 && token < (int)g_prof.size()) cudaEventRecord(g_prof[token].b, g().stream);
}

// knobs (diagnostics / tuning; defaults are what the measurements in DESIGN.md chose)
static int g_write_hi = 0;    // 0: the MMA's own truncation of the raw fp32 bits is `hi` (measured exact)
static int g_chunk_kb = 8;    // 8 k-blocks = K 256 per TMEM accumulation chunk

static int g_dbg_flags = 0;   // Params::dbg
static int g_splitk = 1;      // allow split-K for badly quantised tile counts
static int g_lo_stages = 3;   // lo ring depth; the raw ring gets whatever else fits (or g_raw_stages if > 0)
static int g_raw_stages = 0;
static int g_fwd_mix = 0, g_fwd_chunk_kb = 4;     // >= 0: override mix / chunk_kb for GEMMs flagged `forward`; the default is
                                                  // precision level 1 ("higher", see mt_gemm_set_precision)
static int g_mix = 1;         // 1 (default): hi.hi on TF32, the two first-order corrections as BF16 products; 0: 3xTF32
static int g_group_n = 0;     // tile order: 0 = choose per problem (see launch), > 0 = that many n-blocks per column group,
                              // < 0 = plain row-major order (all n-blocks in one group)
static int g_hints = 6;       // bit 0: A loads evict_first, bit 1: streaming C stores, bit 2: B loads evict_last (see launch)
static int g_max_units = 0;   // > 0: cap the persistent grid at this many units (CTAs or CTA pairs): leaves SMs free for a
                              // concurrent collective kernel (experiment knob, tools/dp_ab.py)
static int g_pair = 1;        // 1: CTA pairs (cta_group::2) for N > 128 when the shape allows; 0: single CTAs only

// what the precision policy (mt_gemm_set_precision and the diagnostic knobs) asks of one GEMM
void precision_for(bool forward, int* mix, int* chunk_kb, int* dbg) {
  const int want_mix = (forward && g_fwd_mix >= 0) ? g_fwd_mix : g_mix;
  const int want_chunk = (forward && g_fwd_chunk_kb >= 0) ? g_fwd_chunk_kb : g_chunk_kb;
  *mix = (want_mix && !g_write_hi) ? 1 : 0;
  *chunk_kb = want_chunk;
  *dbg = g_dbg_flags;
}

template <int BN, bool A_MN, bool B_MN, int NCTA, bool PRO>
static int launch(const GemmArgs& a, long long lda, long long ldb) {
  using C = Cfg<BN, NCTA, PRO>;
  CUtensorMap ma, mb, mt;
  int rc;
  if (A_MN) rc = make_map(&ma, a.A, a.M, a.K, lda, 32, BK, true);
  else rc = make_map(&ma, a.A, a.K, a.M, lda, BK, BM, false);
  if (rc) return rc;
  if (B_MN) rc = make_map(&mb, a.B, a.N, a.K, ldb, 32, BK, true);
  else rc = make_map(&mb, a.B, a.K, a.N, ldb, BK, C::BN_LOCAL, false);
  if (rc) return rc;
  if (a.tanh_src) {
    if (A_MN) rc = make_map(&mt, a.tanh_src, a.M, a.K, lda, 32, BK, true);
    else rc = make_map(&mt, a.tanh_src, a.K, a.M, lda, BK, BM, false);
    if (rc) return rc;
  } else {
    mt = ma;
  }
  Params p;
  p.C = a.C;
  p.bias = a.bias;
  p.epi_tanh = a.epi_tanh;
  p.M = a.M; p.N = a.N; p.K = a.K;
  p.act = a.act;
  p.accumulate = a.accumulate;
  p.m_blocks = (int)ceil_div(a.M, BM * NCTA);
  p.n_blocks = (int)ceil_div(a.N, BN);
  p.k_blocks = (int)ceil_div(a.K, BK);
  const int want_mix = (a.forward && g_fwd_mix >= 0) ? g_fwd_mix : g_mix;
  const int want_chunk = (a.forward && g_fwd_chunk_kb >= 0) ? g_fwd_chunk_kb : g_chunk_kb;
  p.chunk_kb = want_chunk > 0 ? want_chunk : p.k_blocks;
  p.write_hi = g_write_hi;
  p.mix = (want_mix && !PRO && !g_write_hi) ? 1 : 0;
  p.dbg = g_dbg_flags;
  {  // ring depths: lo ring g_lo_stages deep (default 3), raw ring takes the rest of the 227 KB
    int l = g_lo_stages, r = g_raw_stages;
    const int room = C::SMEM_MAX - C::AUX;
    if (l < 1) l = 1;
    if (l > C::MAX_STAGES) l = C::MAX_STAGES;
    while (l > 1 && room - l * C::LO_STAGE < 2 * C::RAW_STAGE) --l;
    const int fit = (room - l * C::LO_STAGE) / C::RAW_STAGE;
    if (r <= 0 || r > fit) r = fit;
    if (r > C::MAX_STAGES) r = C::MAX_STAGES;
    if (r < 1) return fail(MT_ERR_INVALID, "gemm_splitfp32: no room for a raw stage");
    p.raw_stages = r;
    p.lo_stages = l;
  }
  // split-K when the tile count quantises badly over the 148/NCTA persistent units (dW: 256 tiles on 74 pairs = 4
  // waves at 86 %): pick the smallest S <= 8 whose wave efficiency reaches 95 %, keeping >= 64 k-blocks per split
  const long long tiles = (long long)p.m_blocks * p.n_blocks;
  long long unit_count = g().sm_count / NCTA;
  if (g_max_units > 0) unit_count = std::min<long long>(unit_count, std::max(1, g_max_units * 2 / NCTA));
  int S = 1;
  if (g_splitk && tiles < 8 * unit_count) {
    double best = 0.0;
    // multi-wave problems keep >= 64 k-blocks per split (the fold pass must stay negligible); a problem that does not
    // even fill one wave is latency-bound on its serial K loop (~1.1 us per k-block, tools/engine_crossover.py), so idle
    // units take K slices down to one accumulation chunk (8 k-blocks) each
    const int min_kb = tiles < unit_count ? 8 : 64;
    for (int s = 1; s <= 8; ++s) {
      if (s > 1 && p.k_blocks / s < min_kb) break;
      const long long work = tiles * s;
      if (s > 1 && tiles < unit_count && work > unit_count) break;      // single-wave case: never spill into a 2nd wave
      const double eff = (double)work / (double)(ceil_div(work, unit_count) * unit_count);
      if (eff > best + 0.02) { best = eff; S = s; }
      if (best >= 0.95) break;
    }
  }
  p.kb_per_split = (int)ceil_div(p.k_blocks, S);
  p.splits = (int)ceil_div(p.k_blocks, p.kb_per_split);
  {
    // Tile order (work_mn in the kernel).  A B panel is BN rows x this split's K.  If ALL of B's panels fit in a third
    // of L2 the plain order is best: B stays resident, A and C stream through exactly once.  Otherwise cut the n-blocks
    // into column groups about sqrt(units) wide: a wave of U concurrent tiles then touches ~2*sqrt(U) panels instead of
    // n_blocks + U/n_blocks, the group's B panels are what stays hot in L2, and A is re-read once per group.
    const double b_panel = (double)BN * (double)p.kb_per_split * BK * 4.0;
    int gn = p.n_blocks;
    if (g_group_n > 0) gn = std::min(g_group_n, p.n_blocks);
    else if (g_group_n == 0 && p.n_blocks * b_panel > (double)g().l2_bytes / 3.0) {
      int root = 1;
      while ((long long)(root + 1) * (root + 1) <= unit_count) ++root;         // floor(sqrt(units)): 8 for 74 pairs
      const int groups = (int)ceil_div(p.n_blocks, root);
      gn = (int)ceil_div(p.n_blocks, groups);                                   // equal-width groups
    }
    p.group_n = gn < 1 ? 1 : gn;
    // L2 priorities: the B panels of the current column group are re-read by every wave -> evict_last; A is read by the
    // group_n concurrent tiles of its row and then not again before a whole group has gone by -> evict_first.  When K is
    // so long that not even one group of B panels fits (dW: K = batch rows), nothing is reused across waves: no hints.
    const bool b_reused = p.group_n * b_panel <= (double)g().l2_bytes / 2.0 && p.m_blocks * p.group_n > unit_count;
    p.hint_a = ((g_hints & 1) && b_reused) ? 0x12F0000000000000ull : 0ull;
    p.hint_b = ((g_hints & 4) && b_reused) ? 0x14F0000000000000ull : 0ull;
    // C streams out (evict-first) when it is far larger than L2 and nobody re-reads it soon (split-K partials are folded
    // right away: they keep the default policy)
    p.store_cs = ((g_hints & 2) && p.splits == 1 && (double)a.M * a.N * 4.0 > (double)g().l2_bytes / 2.0) ? 1 : 0;
  }
  PoolScratch scratch;
  float* partial = nullptr;
  if (p.splits > 1) {
    rc = scratch.alloc(sizeof(float) * (size_t)p.splits * a.M * a.N);
    if (rc) return rc;
    partial = scratch.f32();
    p.C = partial;
    p.bias = nullptr;
    p.epi_tanh = nullptr;
    p.act = MT_ACT_NONE;
    p.accumulate = 0;
  }
  const int smem_bytes = C::smem_bytes(p.raw_stages, p.lo_stages);
  rc = ensure_debug_word();
  if (rc) return rc;
  auto kern = gemm_splitfp32_kernel<BN, A_MN, B_MN, NCTA, PRO>;
  static bool attr_set = false;
  if (!attr_set) {
    MT_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_MAX));
    attr_set = true;
  }
  const int units = (int)std::min<long long>(tiles * p.splits, unit_count);
  ProfRec rec{};
  if (g_prof_on) {
    rec.a = prof_event();
    rec.b = prof_event();
    rec.flops = 2.0 * (double)a.M * (double)a.N * (double)a.K;
    // bf16-rate units per logical multiply-add: a TF32 product costs 2, a BF16 product 1 (mix: 2 + 1 + 1, all-TF32: 6)
    rec.unit_flops = rec.flops * (p.mix ? 4.0 : 6.0);
    cudaEventRecord(rec.a, g().stream);
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(units * NCTA));
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = g().stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = NCTA;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, kern, ma, mb, mt, p);
  if (g_prof_on) {
    cudaEventRecord(rec.b, g().stream);
    g_prof.push_back(rec);
  }
  if (le != cudaSuccess) return fail(MT_ERR_CUDA, "gemm_splitfp32 launch failed: %s", cudaGetErrorString(le));
  g().launches++;
  if (g().trace_on) {
    char label[56];
    snprintf(label, sizeof(label), "tc_gemm %lldx%lldx%lld %c%c s%d", a.M, a.N, a.K, A_MN ? 'm' : 'k', B_MN ? 'm' : 'k', p.splits);
    trace_mark(label, g().stream, 0);
  }
  if (partial) {
    rc = splitk_finalize(a.C, partial, p.splits, a.M, a.N, a.bias, a.act, a.accumulate, a.epi_tanh);
    if (rc) return rc;
  }
  return MT_OK;
}

template <int BN, int NCTA, bool PRO>
static int dispatch_major2(const GemmArgs& a, bool a_mn, bool b_mn, long long lda, long long ldb) {
  if (!a_mn && !b_mn) return launch<BN, false, false, NCTA, PRO>(a, lda, ldb);
  if (!a_mn && b_mn) return launch<BN, false, true, NCTA, PRO>(a, lda, ldb);
  if (a_mn && !b_mn) return launch<BN, true, false, NCTA, PRO>(a, lda, ldb);
  return launch<BN, true, true, NCTA, PRO>(a, lda, ldb);
}
template <int BN, int NCTA>
static int dispatch_major(const GemmArgs& a, bool a_mn, bool b_mn, long long lda, long long ldb) {
  return a.tanh_src ? dispatch_major2<BN, NCTA, true>(a, a_mn, b_mn, lda, ldb)
                    : dispatch_major2<BN, NCTA, false>(a, a_mn, b_mn, lda, ldb);
}

}  // namespace tc

namespace mt {

int gemm_tc_try(const GemmArgs& a) {
  // operand classification
  bool a_mn, b_mn;
  long long lda, ldb;
  if (a.a_sk == 1 && a.a_sm >= a.K) { a_mn = false; lda = a.a_sm; }
  else if (a.a_sm == 1 && a.a_sk >= a.M) { a_mn = true; lda = a.a_sk; }
  else return MT_ERR_UNSUPPORTED;
  if (a.b_sk == 1 && a.b_sn >= a.K) { b_mn = false; ldb = a.b_sn; }
  else if (a.b_sn == 1 && a.b_sk >= a.N) { b_mn = true; ldb = a.b_sk; }
  else return MT_ERR_UNSUPPORTED;
  // TMA: 16 B aligned bases and row pitches; epilogue: float4 rows of C / bias
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (!al16(a.A) || !al16(a.B) || !al16(a.C) || (a.bias && !al16(a.bias)) || (a.tanh_src && !al16(a.tanh_src)) ||
      (a.epi_tanh && !al16(a.epi_tanh)))
    return MT_ERR_UNSUPPORTED;
  if (lda % 4 || ldb % 4 || a.N % 4) return MT_ERR_UNSUPPORTED;
  // the mid-size regime has its own kernel (no tanh' prologue flavour: callers hand it an explicit dZ)
  if (gemm_tc_mid_wants(a)) {
    if (!a.tanh_src) {
      const int rc = gemm_tc_mid_try(a);
      if (rc != MT_ERR_UNSUPPORTED || g().gemm_engine_force == 2) return rc;
    } else if (g().gemm_engine_force == 2) {
      return MT_ERR_UNSUPPORTED;
    }
  }
  // worthwhile shapes only: small / skinny problems are SIMT territory (HBM- or latency-bound)
  if (g().gemm_engine_force != 1) {
    if (a.M < 128 || a.N < 64 || a.K < 32) return MT_ERR_UNSUPPORTED;
    // measured crossover (tools/engine_crossover.py, profiles/r01_engine_crossover.jsonl): a tile's serial K loop plus
    // the 512-thread / 227 KB / TMEM launch cost ~12 us + ~1 us per k-block, so below 2^28 multiply-adds the SIMT tile
    // (64x64 / 32x32 tiles spread over the SMs) finishes first - by 1.5-2.4x at 2^26 - and from 2^28 up tcgen05 wins
    if (a.M * a.N * a.K < (1ll << 28)) return MT_ERR_UNSUPPORTED;
  } else if (a.M < 1 || a.N < 4 || a.K < 1) {
    return MT_ERR_UNSUPPORTED;
  }
  if (ceil_div(a.M, tc::BM) * ceil_div(a.N, 128) > 0x7fffffffLL) return MT_ERR_UNSUPPORTED;
  if (a.N <= 128) return tc::dispatch_major<128, 1>(a, a_mn, b_mn, lda, ldb);
  // CTA pairs need enough 256-row tiles to keep 74 pairs busy; small M stays on single CTAs
  const bool pair = tc::g_pair && a.M >= 256;
  return pair ? tc::dispatch_major<256, 2>(a, a_mn, b_mn, lda, ldb) : tc::dispatch_major<256, 1>(a, a_mn, b_mn, lda, ldb);
}

}  // namespace mt

MT_API int mt_prof_enable(int on) {
  tc::g_prof_on = on != 0;
  return MT_OK;
}
MT_API int mt_prof_reset(void) {
  for (auto& r : tc::g_prof) {
    tc::g_prof_pool.push_back(r.a);
    tc::g_prof_pool.push_back(r.b);
  }
  tc::g_prof.clear();
  return MT_OK;
}
static int prof_totals(uint64_t* launches, double* ms_total, double* flops_total, double* unit_flops_total) {
  MT_REQUIRE_INIT();
  MT_CUDA(cudaStreamSynchronize(g().stream));
  double ms = 0.0, fl = 0.0, ufl = 0.0;
  for (auto& r : tc::g_prof) {
    float t = 0.f;
    MT_CUDA(cudaEventElapsedTime(&t, r.a, r.b));
    ms += t;
    fl += r.flops;
    ufl += r.unit_flops;
  }
  if (launches) *launches = tc::g_prof.size();
  if (ms_total) *ms_total = ms;
  if (flops_total) *flops_total = fl;
  if (unit_flops_total) *unit_flops_total = ufl;
  return MT_OK;
}
MT_API int mt_prof_gemm(uint64_t* launches, double* ms_total, double* flops_total) {
  return prof_totals(launches, ms_total, flops_total, nullptr);
}
MT_API int mt_prof_gemm_units(uint64_t* launches, double* ms_total, double* flops_total, double* unit_flops_total) {
  return prof_totals(launches, ms_total, flops_total, unit_flops_total);
}

// Precision policy of the tensor-core GEMMs (public: mt_gemm_set_precision in the header).  Per-GEMM error against
// float64 (4096^3, rms relative; tests/gemm_accuracy.py, profiles/r02_gemm_accuracy.jsonl) and what it costs:
//   0 "high"    TF32 hi.hi + two BF16 corrections, centred remainders, TMEM chunks of K = 256:            1.4e-6
//               (the fast mode: opt-in, like TF32 in torch)
//   1 "higher"  (DEFAULT) forward GEMMs as the north_star's 3xTF32 - all three products on TF32 - with chunks of
//               K = 128; the backward GEMMs (dW, dX) as "high":                                           9.0e-7 fwd
//               (a deep, high-gain net amplifies FORWARD rounding ~100x into its gradients; backward hardly matters:
//               config #5 at full width sits 1.3e-4 / 7.5e-5 / 5.3e-5 from the NumPy oracle at levels 0 / 1 / 2 - the
//               oracle itself is 4e-5 from float64 there).  Level 1 is the cheapest level that keeps EVERY BASELINE
//               config inside the north_star's 1e-4 of the oracle end to end;           config #2 step +10-12 % over "high"
//   2 "highest" all-TF32 split, chunks of K = 64, everywhere:  4.8e-7 (NumPy / OpenBLAS SGEMM: 3.7e-7)    step +40 %
MT_API int mt_gemm_set_precision(int level) {
  MT_CHECK_ARG(level >= 0 && level <= 2, "mt_gemm_set_precision: level must be 0 (high), 1 (higher) or 2 (highest)");
  tc::g_mix = level == 2 ? 0 : 1;
  tc::g_chunk_kb = level == 2 ? 2 : 8;
  tc::g_fwd_mix = level == 1 ? 0 : -1;
  tc::g_fwd_chunk_kb = level == 1 ? 4 : -1;
  return MT_OK;
}

// diagnostic knobs (tests / hardware bring-up).
//   write_hi : 1 = converters store a round-to-nearest hi explicitly, 0 (default) = the tensor core's own
//              truncation of the raw fp32 operand is hi.
//   chunk_kb : k-blocks (of 32) accumulated in TMEM between flushes into the fp32 register sums; 0 = whole K.
extern "C" __attribute__((visibility("default"))) int mt_gemm_tc_set_write_hi(int v) {
  tc::g_write_hi = v ? 1 : 0;
  return MT_OK;
}
// after a trapped launch: {0x7100|role, block, thread, parity}; role 0 producer, 1 MMA(tempty), 2 MMA(conv),
// 3 converter, 4 epilogue.  All zero when no wait ever timed out.
extern "C" __attribute__((visibility("default"))) const int* mt_gemm_tc_debug_word(void) { return tc::g_dbg_host; }
//   pair     : 1 (default) = CTA pairs (tcgen05 cta_group::2) for wide problems, 0 = single-CTA tiles only.
extern "C" __attribute__((visibility("default"))) int mt_gemm_tc_set_rings(int raw_stages, int lo_stages) {
  tc::g_raw_stages = raw_stages;
  tc::g_lo_stages = lo_stages;
  return MT_OK;
}
extern "C" __attribute__((visibility("default"))) int mt_gemm_tc_set_splitk(int v) {
  tc::g_splitk = v ? 1 : 0;
  return MT_OK;
}
extern "C" __attribute__((visibility("default"))) int mt_gemm_tc_set_dbg(int v) {
  tc::g_dbg_flags = v;
  return MT_OK;
}
//   group_n  : tile order - 0 (default) chosen per problem, n > 0 that many n-blocks per column group, < 0 row-major.
//   hints    : bit 0 = A loads evict_first, bit 1 = streaming stores of C, bit 2 = B loads evict_last.
extern "C" __attribute__((visibility("default"))) int mt_gemm_tc_set_group_n(int v) {
  tc::g_group_n = v;
  return MT_OK;
}
extern "C" __attribute__((visibility("default"))) int mt_gemm_tc_set_hints(int v) {
  tc::g_hints = v;
  return MT_OK;
}
extern "C" __attribute__((visibility("default"))) int mt_gemm_tc_set_pair(int v) {
  tc::g_pair = v ? 1 : 0;
  return MT_OK;
}
extern "C" __attribute__((visibility("default"))) int mt_gemm_tc_set_chunk_kb(int v) {
  tc::g_chunk_kb = v < 0 ? 0 : v;
  return MT_OK;
}
//   fwd      : precision of the FORWARD GEMMs only (mt_linear_fwd_f32): mix / chunk_kb as above, -1 = same as the rest.
extern "C" __attribute__((visibility("default"))) int mt_gemm_tc_set_fwd(int mix, int chunk_kb) {
  tc::g_fwd_mix = mix;
  tc::g_fwd_chunk_kb = chunk_kb;
  return MT_OK;
}
//   max_units: cap of the persistent grid in CTA PAIRS (single-CTA kernels get twice as many CTAs); 0 = every SM.
extern "C" __attribute__((visibility("default"))) int mt_gemm_tc_set_max_units(int v) {
  tc::g_max_units = v < 0 ? 0 : v;
  return MT_OK;
}
extern "C" __attribute__((visibility("default"))) int mt_gemm_tc_set_mix(int v) {
  tc::g_mix = v ? 1 : 0;
  return MT_OK;
}
