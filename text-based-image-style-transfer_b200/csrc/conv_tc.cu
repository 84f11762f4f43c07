// tcgen05 implicit-GEMM convolution for the VGG-19 trunk (3x3 pad 1, and 1x1 for the Gram backward).
//
// Replaces, for the hot path of the reference, the library kernels behind
//   multi_style_transfer/helper_functions.py:94-101 (Vgg19.forward: nn.Conv2d + ReLU + MaxPool2d)
// and their autograd data-gradients (run_style_transfer.py:140, loss.backward()).
//
// Tiling: one CTA tile = 16 x 8 output pixels (M = 128) x BLOCK_N output channels.  For every 64-channel slice of
// the input, TMA loads ONE 18 x 10 x 64 halo patch (out-of-bounds -> zero = padding) into 128B-swizzled shared memory;
// the nine filter taps are nine views of that patch: the A-operand descriptor starts (dr * 10 + ds) rows further and
// steps 10 rows between 8-row groups (SBO = 1280 B).  tcgen05.mma applies the 128B swizzle to absolute shared-memory
// addresses, so a descriptor may start on any 128-byte row (measured: profiles/r01_umma_descriptor_probe.log).
// Only the [BLOCK_N x 64] weight slice is streamed per tap.  One elected thread issues four tcgen05.mma (K = 16
// each) per (slice, tap), accumulating in TMEM.  The kernel is persistent with two TMEM accumulator stages, so the
// epilogue of tile i overlaps the main loop of tile i+1.  Warp roles: 0 = TMA producer, 1 = MMA issuer,
// 2 = TMEM allocator, 4..7 = epilogue.
//
// Why the halo matters: the main loop is bound by shared-memory bandwidth (128 B/clk per SM, shared by the TMA
// writes and the tensor core's operand reads).  Re-loading the patch per tap costs 16 KB of writes per 4 MMAs;
// the halo costs 23 KB per 36.
#include "conv_tc.cuh"
#include "common.cuh"
#include "conv_epilogue.cuh"

#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>

namespace nst {

static constexpr int TILE_H = 16;
static constexpr int TILE_W = 8;
static constexpr int BLOCK_M = TILE_H * TILE_W;  // 128
static constexpr int BLOCK_K = 64;               // 64 x 16-bit = one 128-byte swizzle row
static constexpr int UMMA_K = 16;
static constexpr int HALO_ROWS = (TILE_H + 2) * (TILE_W + 2);     // 180 pixels
static constexpr int HALO_TX_BYTES = HALO_ROWS * BLOCK_K * 2;     // 23040
static constexpr int HALO_STAGE_BYTES = 23 * 1024;                // keeps the next stage 1024-byte aligned
static constexpr int FLAT_TX_BYTES = BLOCK_M * BLOCK_K * 2;       // 1x1 convolution: no halo
static constexpr int SMEM_BUDGET = 196608;  // bytes of operand staging per CTA
#ifndef NST_HALO64
#define NST_HALO64 3
#endif
#ifndef NST_HALO64_PAIR
#define NST_HALO64_PAIR 3
#endif
#ifndef NST_HALO128_PAIR_DGRAD
#define NST_HALO128_PAIR_DGRAD NST_HALO128_PAIR
#endif
#ifndef NST_HALO16
#define NST_HALO16 2
#endif
#ifndef NST_HALO128_PAIR
#define NST_HALO128_PAIR 4
#endif

#ifndef NST_DGRAD_BUDGET
#define NST_DGRAD_BUDGET SMEM_BUDGET
#endif
template <int BLOCK_N, bool PAIR = false, int MODE = CONV_FWD>
struct ConvCfg {
  // operand staging of this instantiation: the data gradients' epilogues read their masks / routing bytes / seeds through L1,
  // which is what shared memory leaves of the SM's 256 KB
  static constexpr int BUDGET = (MODE == CONV_DGRAD && BLOCK_N >= 64) ? NST_DGRAD_BUDGET : SMEM_BUDGET;
  // warps 0..3: TMA producer, MMA issuer, TMEM allocator, spare; then the epilogue warps.  Eight of them (two per TMEM
  // lane quarter, each taking half of the accumulator columns): the epilogue is a latency chain (tcgen05.ld -> math ->
  // scattered 16-byte stores) and one warp per scheduler cannot hide it.
  static constexpr int EPI_WARPS = BLOCK_N >= 64 ? 8 : 4;
  static constexpr int NUM_THREADS = 128 + 32 * EPI_WARPS;
  // Patches in flight.  r01 (single CTAs, the Gram backward still a launch of its own): two; deeper rings (5 at N = 64, 8 at
  // N = 16) made the 64-channel data gradients slower inside the captured step (conv1_2: 39 -> 58 us).  Round 2, as CTA pairs,
  // every build on the same box inside the step (profiles/r03_patch_stage_sweep.log): a pair CTA stages half of the weight
  // rows, so a weight stage is half as large and the ring holds five of them at N = 128 beside FOUR patches (conv class 412 ->
  // 406 us per evaluation; six patches leave two weight stages and starve the weight stream: 459 us); at N = 64 three patches
  // (426 -> 419 us: the data gradient of conv1_2 loads two operands per tile - the patch and the tap tile of the folded Gram
  // backward - and waited 10 of its 30 us for them with two stages).  conv1_1's gradient (N = 16) keeps two: its two MMA
  // issuers own one patch stage each.
  static constexpr int HALO_STAGES = (BLOCK_N == 64) ? (PAIR ? NST_HALO64_PAIR : NST_HALO64)
                                                     : ((BLOCK_N == 128 && PAIR) ? (MODE == CONV_DGRAD ? NST_HALO128_PAIR_DGRAD : NST_HALO128_PAIR) : (BLOCK_N == 16 ? NST_HALO16 : 2));
  // Filter taps per weight stage.  One pipeline iteration (barrier wait, fence, commit) costs a few hundred cycles of
  // the issuing thread and every tcgen05.mma about 45 (profiles/r01_mma_issue_rate.log); with one tap (4 MMAs) per
  // iteration the main loop was issue-bound at ~480 cycles per tap for every N <= 128.  Three taps per stage amortise it.
  static constexpr int TPS = BLOCK_N >= 256 ? 1 : (BLOCK_N >= 64 ? 3 : 9);
  static constexpr int B_TILE_BYTES = BLOCK_N * BLOCK_K * 2;   // one tap (both CTAs of a pair together)
  static constexpr int B_STAGE_BYTES = TPS * B_TILE_BYTES / (PAIR ? 2 : 1);   // a CTA of a pair stages half of the rows
  static constexpr int B_STAGES_FIT = (BUDGET - HALO_STAGES * HALO_STAGE_BYTES) / B_STAGE_BYTES;
  static constexpr int B_STAGES = B_STAGES_FIT > 6 ? 6 : B_STAGES_FIT;
  static constexpr int TMEM_COLS = 2 * BLOCK_N < 32 ? 32 : 2 * BLOCK_N;  // 32 / 128 / 256 / 512: powers of two >= 32
  // data gradient with a folded Gram backward: two more accumulator stages (columns 2N .. 4N)
  static constexpr int TMEM_COLS_SEED = (BLOCK_N >= 64 && BLOCK_N <= 128) ? 4 * BLOCK_N : TMEM_COLS;
  static constexpr int OPERAND_BYTES = HALO_STAGES * HALO_STAGE_BYTES + B_STAGES * B_STAGE_BYTES;
  static constexpr int BIAS_BYTES = 2 * BLOCK_N * 4;  // one bias slice per accumulator stage
  // staging for the epilogue's TMA stores: two 2 KB buffers per epilogue warp, 1024-byte aligned
  static constexpr int EPI_BYTES = EPI_WARPS * 2 * EPI_STAGE_BYTES;
  static constexpr int MISC_BYTES = ((256 /*barriers*/ + BIAS_BYTES + 1023) / 1024) * 1024;
  static constexpr int SMEM_BYTES = OPERAND_BYTES + 1024 /*align slack*/ + MISC_BYTES + EPI_BYTES;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
// PAIR: two CTAs on the two SMs of a TPC (cluster of 2) work on two neighbouring pixel tiles of the same N tile with
// tcgen05.mma.cta_group::2 (M = 256): each CTA stages its own patch and HALF of the weight rows, the tensor cores of both SMs
// read every weight half once.  Per K = 16 step a CTA's shared memory then delivers (128 + N/2) x 32 B instead of
// (128 + N) x 32 B - 48 cycles at N = 128, below the 64 cycles of math; alone an SM needs all 64 and has nothing left for
// the TMA writes of the weight stream (measured 81 cycles per MMA) - and one issuing thread feeds two tensor cores, which
// halves the per-MMA issue cost that bounds the N = 64 layers.  The leader (cluster rank 0) issues every MMA; both CTAs keep
// their own producer and epilogue warps.
template <int BLOCK_N, int MODE, bool TMA_OUT, bool PAIR>
__global__ void __launch_bounds__(ConvCfg<BLOCK_N, PAIR, MODE>::NUM_THREADS, 1) conv_tc_kernel(const __grid_constant__ ConvParams p) {
  using Cfg = ConvCfg<BLOCK_N, PAIR, MODE>;
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle needs 1024-byte aligned stages
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + Cfg::HALO_STAGES * HALO_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OPERAND_BYTES);
  uint64_t* afull_bar = bars;
  uint64_t* aempty_bar = afull_bar + Cfg::HALO_STAGES;
  uint64_t* bfull_bar = aempty_bar + Cfg::HALO_STAGES;
  uint64_t* bempty_bar = bfull_bar + Cfg::B_STAGES;
  uint64_t* tfull_bar = bempty_bar + Cfg::B_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* sbias = reinterpret_cast<float*>(smem + Cfg::OPERAND_BYTES + 256);
  uint8_t* sEpi = smem + Cfg::OPERAND_BYTES + Cfg::MISC_BYTES;  // 1024-byte aligned (operand stages are multiples of 1024)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // persistent workers: CTAs, or CTA pairs
  const int cta_rank = PAIR ? static_cast<int>(cluster_ctarank()) : 0;
  const int worker = PAIR ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int workers = PAIR ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  constexpr int B_HALF_DIV = PAIR ? 2 : 1;   // weight rows of a stage held by this CTA: BLOCK_N / B_HALF_DIV
  const bool dbg = NST_DBG_PTR(p) != nullptr && blockIdx.x == 0;
  const long long life0 = NST_DBG_PTR(p) != nullptr ? clock64() : 0;
#define NST_STAMP(slot, cond)                                        \
  do {                                                               \
    if (dbg && (cond)) NST_DBG_PTR(p)[slot] = clock64();                      \
  } while (0)
  // wait accounting of CTA 0 (debug): slot 8 MMA waits for the patch, 9 for weights, 10 for a free accumulator stage,
  // 11 epilogue (warp 4) waits for the accumulator, 12 producer waits for a free patch stage, 13 for a free weight stage
  long long wacc0 = 0, wacc1 = 0, wacc2 = 0;
#define NST_WAIT(acc, stmt)                   \
  do {                                        \
    if (dbg) {                                \
      const long long w0__ = clock64();       \
      stmt;                                   \
      acc += clock64() - w0__;                \
    } else {                                  \
      stmt;                                   \
    }                                         \
  } while (0)
  NST_STAMP(0, threadIdx.x == 0);
  if (NST_TL_PTR(p) != nullptr && threadIdx.x == 0) atomicMin(&NST_TL_PTR(p)[0], globaltimer_ns());
  // Programmatic dependent launch: the next kernel in the stream may be scheduled once every CTA of this grid has said so
  // (its CTAs run their own setup, then block in griddepcontrol.wait until this grid has completed and flushed).  A CTA says
  // so when its MMA thread has issued the main loop of its LAST tile - not at kernel entry: a dependent CTA needs a whole SM
  // (shared memory, registers), so released early it could only land on SMs this grid leaves idle (20 of 148 for the
  // 128-tile layers at 512^2) and would sit there in griddepcontrol.wait for the whole launch, keeping the side stream's
  // Gram work off exactly the SMs it is planned for (gram.cu, GramParams::max_ctas).

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA);
    tma_prefetch_desc(&p.tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::HALO_STAGES; ++s) {
      mbar_init(&afull_bar[s], 1);
      mbar_init(&aempty_bar[s], 1);
    }
    for (int s = 0; s < Cfg::B_STAGES; ++s) {
      mbar_init(&bfull_bar[s], 1);
      mbar_init(&bempty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      // one arrival per epilogue warp; the leader of a pair also collects its peer's (the accumulator stage of BOTH CTAs
      // is rewritten by the leader's next MMA)
      mbar_init(&tempty_bar[s], Cfg::EPI_WARPS * (PAIR ? 2 : 1));
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    if constexpr (PAIR) {
      tmem_alloc_2sm(tmem_slot, MODE == CONV_DGRAD ? Cfg::TMEM_COLS_SEED : Cfg::TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, MODE == CONV_DGRAD ? Cfg::TMEM_COLS_SEED : Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  // pair: the peer's barriers must exist before a remote arrival, a multicast commit or a TMA completion reaches them
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above touched only kernel parameters and on-chip state; from here on this grid reads and writes tensors
  // produced by earlier kernels: wait for the grid(s) it depends on (no-op without the launch attribute).  The producer
  // warp waits later: the network weights are written by nobody inside a step, so its first weight stages are requested
  // before the wait (they come from DRAM - the L2 is full of activations and optimizer history by then - and would
  // otherwise arrive after the first patch, which the previous launch has just left in L2).
  if (warp != 0) asm volatile("griddepcontrol.wait;" ::: "memory");
  NST_STAMP(1, threadIdx.x == 0);
  if (NST_TL_PTR(p) != nullptr && threadIdx.x == 128) atomicMin(&NST_TL_PTR(p)[2], globaltimer_ns());

  const int k_slices = p.K / BLOCK_K;
  const int pad = p.taps == 9 ? 1 : 0;
  const int halo_w = TILE_W + 2 * pad;                     // pixels per patch row
  const uint32_t a_tx = p.taps == 9 ? HALO_TX_BYTES : FLAT_TX_BYTES;
  const int sp_tiles = p.tiles_w * p.tiles_h;
  // work items: (N tile, pixel tile), or (N tile, two consecutive pixel tiles) for a pair - this CTA takes the tile of its
  // rank; with an odd tile count the last pair's second tile lies below the image (all loads zero-fill, nothing is stored)
  // Several images per launch (p.batch > 1): the tensors are [batch][H][W][C]; an item also names its image.  Pairs never
  // straddle two images.
  const int img_items = PAIR ? (sp_tiles + 1) >> 1 : sp_tiles;
  const int sp_items = img_items * p.batch;
  const int num_items = sp_items * p.tiles_n;
  auto item_coords = [&](int item, int& nt, int& b, int& th, int& tw) {
    nt = item / sp_items;
    int r = item - nt * sp_items;
    b = 0;
    if (p.batch > 1) {
      b = r / img_items;
      r -= b * img_items;
    }
    const int sp = PAIR ? 2 * r + cta_rank : r;
    if (PAIR && sp >= sp_tiles) {
      th = p.tiles_h;
      tw = 0;
    } else {
      th = sp / p.tiles_w;
      tw = sp - th * p.tiles_w;
    }
  };
  const int tps = p.taps == 9 ? Cfg::TPS : 1;              // taps per weight stage (the weight tensor map's box depth)
  // A layer with 64 input channels has ONE 64-channel slice: its whole weight set (nine taps) fits the weight ring.  It
  // is then loaded once per CTA and stays resident - re-streaming it for every tile (72 KB next to a 23 KB patch at
  // N = 64) made the 64-channel layers L2-bandwidth bound (148 SMs x ~96 KB per microsecond).  All tiles of a launch
  // must then use the same weights: one N tile only.
  const bool seed = MODE == CONV_DGRAD && Cfg::TMEM_COLS_SEED == 4 * BLOCK_N && p.seed_k > 0;
  const int seed_slices = seed ? p.seed_k / BLOCK_K : 0;
  // dbg_flags bit 2 (timing experiment, wrong results): never re-stream the weights - how fast is the main loop without
  // the TMA writes of the B operand competing with the tensor core's shared-memory reads?
  const bool b_resident = p.taps == 9 && 9 / Cfg::TPS <= Cfg::B_STAGES && !seed &&
                          ((k_slices == 1 && p.tiles_n == 1) || NST_DBG_FLAG(p, 4));
  // one-slice tiles with resident weights are issued by two threads (see the MMA issuer)
  const bool dual_issue = !PAIR && b_resident && k_slices == 1 && Cfg::HALO_STAGES == 2 && p.dual_issue != 0;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (elect_one()) {
      int as = 0, bs = 0;
      uint32_t aphase = 0, bphase = 0;
      // weight stages of the first tile's first slice, requested ahead of griddepcontrol.wait (not for the 1x1 Gram backward:
      // its B operand is written by the Gram kernel right before it)
      int pre_b = 0;
      // pair: the leader's barrier counts the bytes of both CTAs' loads (a weight stage is BLOCK_N rows either way, a patch
      // stage two patches); the peer only issues its loads, which report to the leader's barrier
      const bool expects = !PAIR || cta_rank == 0;
      const uint32_t a_expect = PAIR ? 2u * a_tx : a_tx;
      const int n_rank = PAIR ? cta_rank * (BLOCK_N / 2) : 0;
      auto load_a = [&](const CUtensorMap* tm, int stage, uint32_t bytes, int c0, int c1, int c2, int c3) {
        if (expects) mbar_arrive_expect_tx(&afull_bar[stage], bytes);
        if constexpr (PAIR) tma_load_4d_2sm(sA + stage * HALO_STAGE_BYTES, tm, mapa_u32(smem_u32(&afull_bar[stage]), 0), c0, c1, c2, c3);
        else tma_load_4d(sA + stage * HALO_STAGE_BYTES, tm, &afull_bar[stage], c0, c1, c2, c3);
      };
      auto load_b = [&](const CUtensorMap* tm, int stage, uint32_t bytes, int c0, int c1, int c2) {
        if (expects) mbar_arrive_expect_tx(&bfull_bar[stage], bytes);
        if constexpr (PAIR) tma_load_3d_2sm(sB + stage * Cfg::B_STAGE_BYTES, tm, mapa_u32(smem_u32(&bfull_bar[stage]), 0), c0, c1 + n_rank, c2);
        else tma_load_3d(sB + stage * Cfg::B_STAGE_BYTES, tm, &bfull_bar[stage], c0, c1, c2);
      };
      if (MODE != CONV_SCALE && worker < num_items) {
        const int n0 = (worker / sp_items) * BLOCK_N;
        for (int tap = 0; tap < p.taps && pre_b < Cfg::B_STAGES; tap += tps, ++pre_b) {
          load_b(&p.tmB, bs, static_cast<uint32_t>(tps) * Cfg::B_TILE_BYTES, 0, n0, tap);
          if (++bs == Cfg::B_STAGES) {
            bs = 0;
            bphase ^= 1u;
          }
        }
      }
      asm volatile("griddepcontrol.wait;" ::: "memory");
      for (int tile = worker; tile < num_items; tile += workers) {
        int nt, b, th, tw;
        item_coords(tile, nt, b, th, tw);
        const int h0 = th * TILE_H, w0 = tw * TILE_W, n0 = nt * BLOCK_N;
        const int b_dh = p.b_per_image ? b : 0;   // Gram-backward operand: one matrix per image in the third map dimension
        for (int ks = 0; ks < k_slices; ++ks) {
          NST_WAIT(wacc0, mbar_wait(&aempty_bar[as], aphase ^ 1u));
          load_a(&p.tmA, as, a_expect, ks * BLOCK_K, w0 - pad, h0 - pad, b);
          if (++as == Cfg::HALO_STAGES) {
            as = 0;
            aphase ^= 1u;
          }
          if (b_resident && (tile != worker || ks > 0)) continue;
          for (int tap = 0; tap < p.taps; tap += tps) {
            if (pre_b > 0) {   // already in flight (first tile, first slice)
              --pre_b;
              continue;
            }
            NST_WAIT(wacc1, mbar_wait(&bempty_bar[bs], bphase ^ 1u));
            load_b(&p.tmB, bs, static_cast<uint32_t>(tps) * Cfg::B_TILE_BYTES, ks * BLOCK_K, n0, tap + (MODE == CONV_SCALE ? b_dh : 0));
            if (++bs == Cfg::B_STAGES) {
              bs = 0;
              bphase ^= 1u;
            }
          }
        }
        // folded Gram backward: the tile of the tap (no halo) and the matching slice of dh, one 64-channel slice at a time
        for (int ks2 = 0; ks2 < seed_slices; ++ks2) {
          NST_WAIT(wacc0, mbar_wait(&aempty_bar[as], aphase ^ 1u));
          load_a(&p.tmA2, as, (PAIR ? 2u : 1u) * FLAT_TX_BYTES, ks2 * BLOCK_K, w0, h0, b);
          if (++as == Cfg::HALO_STAGES) {
            as = 0;
            aphase ^= 1u;
          }
          NST_WAIT(wacc1, mbar_wait(&bempty_bar[bs], bphase ^ 1u));
          load_b(&p.tmB2, bs, Cfg::B_TILE_BYTES, ks2 * BLOCK_K, n0, b_dh);
          if (++bs == Cfg::B_STAGES) {
            bs = 0;
            bphase ^= 1u;
          }
        }
      }
      if constexpr (PAIR) {
        // The peer never issues an MMA: it releases the dependent launch here, when its last load is on its way (the leader's
        // MMA thread, which comes later, decides).  Then both producers stay until every stage they filled has been released:
        // the leader's commits arrive at this CTA's barriers, which must still exist.
#ifndef NST_EXP_LATE_TRIGGER   // experiment builds only (profiles/r03_concurrent_plans_hang.log): dependents released at CTA exit
        if (cta_rank != 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
        for (int s = 0; s < Cfg::HALO_STAGES; ++s) {
          mbar_wait(&aempty_bar[as], aphase ^ 1u);
          if (++as == Cfg::HALO_STAGES) {
            as = 0;
            aphase ^= 1u;
          }
        }
        if (!b_resident) {
          for (int s = 0; s < Cfg::B_STAGES; ++s) {
            mbar_wait(&bempty_bar[bs], bphase ^ 1u);
            if (++bs == Cfg::B_STAGES) {
              bs = 0;
              bphase ^= 1u;
            }
          }
        }
      }
      if (dbg) {
        NST_DBG_PTR(p)[12] = wacc0;
        NST_DBG_PTR(p)[13] = wacc1;
      }
    }
  } else if ((warp == 1 || (warp == 3 && dual_issue)) && (!PAIR || cta_rank == 0) && elect_one()) {
    // ===================== MMA issuer (one thread; two for layers whose weights are resident) =====================
    // The thread is chosen with elect.sync in a warp-uniform branch (not `lane == 0`): only then does the compiler know that
    // a single lane runs the uniform-datapath UTCHMMA / UTCBAR instructions and emits them back to back; with a divergent
    // predicate every tcgen05.mma was wrapped in an ELECT / BRA.U.ANY serialisation loop (~45 cycles per MMA).
    // The issue rate of tcgen05.mma is set by the scalar instructions this thread executes between two of them
    // (measured: 45 cycles per MMA with constant descriptors, 76 with a few extra index operations,
    // profiles/r01_mma_issue_rate.log), so the 64-bit descriptors are not rebuilt per MMA: their upper halves are
    // loop constants and the lower halves (address >> 4) advance by compile-time immediates.
    // Two issuers (dual_issue): one thread cannot issue a tcgen05.mma more often than every ~47 cycles
    // (profiles/r01_mma_issue_rate_elect.log: 47 cycles per MMA at N = 16, 51 at N = 64 in a bare loop; 70 - 79 inside this
    // kernel with its barrier traffic), which is more than the 32 cycles an M128 x N64 MMA occupies the tensor core.  Where a
    // tile is one 64-channel slice with resident weights (conv1_2 forward, conv1_1's data gradient) the tiles of a CTA
    // are independent streams: warp 1 takes the CTA's even tiles (patch stage 0, accumulator stage 0), warp 3 the odd ones
    // (stages 1), each with its own barriers - the producer and the epilogue already alternate the stages tile by tile.
    const int issuer = warp == 3 ? 1 : 0;
    const int tile_step = dual_issue ? 2 * workers : workers;
    int as = dual_issue ? issuer : 0, bs = 0, ts = dual_issue ? issuer : 0;
    uint32_t aphase = 0, bphase = 0, tphase = 0;
    const int first_tile = worker + issuer * workers;
    const uint32_t sbo = static_cast<uint32_t>(halo_w) * 128u;  // bytes between 8-pixel groups of the A operand
    const uint32_t idesc = p.idesc;
    const uint32_t a_hi = static_cast<uint32_t>(umma_desc_sw128(0, 16, sbo) >> 32);
    const uint32_t b_hi = static_cast<uint32_t>(umma_desc_sw128(0, 16, 1024) >> 32);
    const uint32_t lbo_lo = static_cast<uint32_t>(umma_desc_sw128(0, 16, 0) & 0xffffffffu);  // LBO field, address 0
    const bool conv3x3 = p.taps == 9;
    long long tl_c0 = 0;
    unsigned long long tl_g0 = 0;
    auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t id, uint32_t acc) {
      if constexpr (PAIR) umma_f16_2sm(d, da, db, id, acc);
      else umma_f16(d, da, db, id, acc);
    };
    auto commit = [&](uint64_t* bar) {
      if constexpr (PAIR) umma_commit_2sm(bar);
      else umma_commit(bar);
    };
    for (int tile = first_tile; tile < num_items; tile += tile_step) {
      NST_WAIT(wacc2, mbar_wait(&tempty_bar[ts], tphase ^ 1u));
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(ts * BLOCK_N);
      uint32_t accumulate = 0;
      for (int ks = 0; ks < k_slices; ++ks) {
        NST_WAIT(wacc0, mbar_wait(&afull_bar[as], aphase));
        NST_STAMP(2, ks == 0 && tile == worker);
        if (NST_TL_PTR(p) != nullptr && ks == 0 && tile == first_tile) {
          const unsigned long long now = globaltimer_ns();
          atomicMin(&NST_TL_PTR(p)[4], now);
          atomicMax(&NST_TL_PTR(p)[5], now);
          tl_c0 = clock64();
          tl_g0 = now;
        }
        const uint32_t a_lo0 = lbo_lo | (smem_u32(sA + as * HALO_STAGE_BYTES) >> 4);
        if (conv3x3) {
          // nine taps = nine row shifts of the patch: (dr * 10 + ds) rows of 128 B = (dr * 10 + ds) * 8 descriptor units
#pragma unroll 1
          for (int g = 0; g < 9 / Cfg::TPS; ++g) {
            // resident weights: stage g holds filter row(s) g for the whole launch; only the first tile waits for them
            if (b_resident) bs = g;
            if (!b_resident || tile == first_tile) NST_WAIT(wacc1, mbar_wait(&bfull_bar[bs], bphase));
            tc_fence_after();
            const uint32_t b_lo0 = lbo_lo | (smem_u32(sB + bs * Cfg::B_STAGE_BYTES) >> 4);
            // TPS = 1: tap g;  TPS = 3: taps 3g .. 3g+2 (one filter row);  TPS = 9: all taps
            const uint32_t a_lo_g = Cfg::TPS == 1 ? a_lo0 + static_cast<uint32_t>((g / 3) * (TILE_W + 2) + g % 3) * 8u
                                                  : a_lo0 + static_cast<uint32_t>(g) * ((TILE_W + 2) * 8u);
#pragma unroll
            for (int tt = 0; tt < Cfg::TPS; ++tt) {
              const uint32_t a_tap = Cfg::TPS == 9 ? static_cast<uint32_t>(((tt / 3) * (TILE_W + 2) + tt % 3) * 8)
                                                   : static_cast<uint32_t>(tt * 8);
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                // NST_DBG_FLAG bit 3 (instrumented build, wrong results): every tap reads the un-shifted view - what do the
                // shifted (not 1024-byte aligned) A views cost the tensor core's shared-memory reads?
                const uint32_t a_view = NST_DBG_FLAG(p, 8) ? a_lo0 : a_lo_g + a_tap;
                const uint64_t da = (static_cast<uint64_t>(a_hi) << 32) | (a_view + static_cast<uint32_t>(k * 2));
                const uint64_t db = (static_cast<uint64_t>(b_hi) << 32) |
                                    (b_lo0 + static_cast<uint32_t>(tt * (Cfg::B_TILE_BYTES / B_HALF_DIV / 16) + k * 2));
                mma(d_tmem, da, db, idesc, accumulate);
                accumulate = 1u;
              }
            }
            if (b_resident) continue;
            commit(&bempty_bar[bs]);  // frees the weight stage when its MMAs retire
            if (++bs == Cfg::B_STAGES) {
              bs = 0;
              bphase ^= 1u;
            }
          }
        } else {
          // 1x1 convolution: one weight tile per 64-channel slice, no shift
          mbar_wait(&bfull_bar[bs], bphase);
          tc_fence_after();
          const uint32_t b_lo0 = lbo_lo | (smem_u32(sB + bs * Cfg::B_STAGE_BYTES) >> 4);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = (static_cast<uint64_t>(a_hi) << 32) | (a_lo0 + static_cast<uint32_t>(k * 2));
            const uint64_t db = (static_cast<uint64_t>(b_hi) << 32) | (b_lo0 + static_cast<uint32_t>(k * 2));
            mma(d_tmem, da, db, idesc, accumulate);
            accumulate = 1u;
          }
          commit(&bempty_bar[bs]);
          if (++bs == Cfg::B_STAGES) {
            bs = 0;
            bphase ^= 1u;
          }
        }
        commit(&aempty_bar[as]);  // ... and the patch after its last tap
        if (dual_issue) {
          aphase ^= 1u;   // this issuer's own patch stage, next use
        } else if (++as == Cfg::HALO_STAGES) {
          as = 0;
          aphase ^= 1u;
        }
      }
      if (seed) {
        // second accumulator of this stage: seed[pixel, n] = sum_k tap[pixel, k] * dh[n, k]  (fp16 operands)
        const uint32_t d2_tmem = tmem_base + static_cast<uint32_t>((2 + ts) * BLOCK_N);
        const uint32_t a_hi2 = static_cast<uint32_t>(umma_desc_sw128(0, 16, TILE_W * 128u) >> 32);
        const uint32_t idesc2 = p.idesc2;
        uint32_t acc2 = 0;
        for (int ks2 = 0; ks2 < seed_slices; ++ks2) {
          NST_WAIT(wacc0, mbar_wait(&afull_bar[as], aphase));
          NST_WAIT(wacc1, mbar_wait(&bfull_bar[bs], bphase));
          tc_fence_after();
          const uint32_t a_lo0 = lbo_lo | (smem_u32(sA + as * HALO_STAGE_BYTES) >> 4);
          const uint32_t b_lo0 = lbo_lo | (smem_u32(sB + bs * Cfg::B_STAGE_BYTES) >> 4);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = (static_cast<uint64_t>(a_hi2) << 32) | (a_lo0 + static_cast<uint32_t>(k * 2));
            const uint64_t db = (static_cast<uint64_t>(b_hi) << 32) | (b_lo0 + static_cast<uint32_t>(k * 2));
            mma(d2_tmem, da, db, idesc2, acc2);
            acc2 = 1u;
          }
          commit(&bempty_bar[bs]);
          if (++bs == Cfg::B_STAGES) {
            bs = 0;
            bphase ^= 1u;
          }
          commit(&aempty_bar[as]);
          if (++as == Cfg::HALO_STAGES) {
            as = 0;
            aphase ^= 1u;
          }
        }
      }
      commit(&tfull_bar[ts]);  // accumulator(s) complete
      NST_STAMP(3, tile == worker);
      if (NST_TL_PTR(p) != nullptr && tile + tile_step >= num_items) {
        const unsigned long long now = globaltimer_ns();
        atomicMax(&NST_TL_PTR(p)[6], now);
        // SM clock over this CTA's main loops: (SM cycles << 32) | nanoseconds, worker 0 only
        if (worker == 0) NST_TL_PTR(p)[7] = (static_cast<unsigned long long>(clock64() - tl_c0) << 32) | ((now - tl_g0) & 0xffffffffull);
      }
      if (dual_issue) {
        tphase ^= 1u;     // this issuer's own accumulator stage, next use
      } else if (++ts == 2) {
        ts = 0;
        tphase ^= 1u;
      }
    }
#ifdef NST_EXP_LATE_TRIGGER
    if constexpr (!PAIR)
#endif
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if constexpr (PAIR) {
      // The peer's epilogue warps release the accumulator stages of the last tiles with REMOTE arrivals on this CTA's
      // barriers, and nobody would wait for them: the teardown barrier below orders execution, not those writes.  If this CTA
      // left first, an arrival still in flight could land in the shared memory of the NEXT CTA scheduled on this SM - on the
      // freshly initialised barrier it keeps at the same offset (with several plans stepping on their own streams that next
      // CTA starts right away: one run in five of four concurrent 256 x 256 frames hung).  So the leader collects them.
      for (int s = 0; s < 2; ++s) {
        mbar_wait(&tempty_bar[ts], tphase ^ 1u);
        if (++ts == 2) {
          ts = 0;
          tphase ^= 1u;
        }
      }
    }
    if (dbg) {
      NST_DBG_PTR(p)[8] = wacc0;
      NST_DBG_PTR(p)[9] = wacc1;
      NST_DBG_PTR(p)[10] = wacc2;
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 4) >> 2;       // which half of the accumulator columns (0 when there are four epilogue warps)
    constexpr int COLS = BLOCK_N / (Cfg::EPI_WARPS / 4);  // columns per warp
    const int col0 = half * COLS;
    const int t = q * 32 + lane;
    const int hl = t / TILE_W, wl = t % TILE_W;
    const int et = threadIdx.x - 128;       // index among the epilogue threads
    EpiStore est;
    est.buf = sEpi + (warp - 4) * (2 * EPI_STAGE_BYTES);
    est.next = 0;
    est.lane = lane;
    // compile-time: the direct-store instantiation must not carry the staging code (register pressure in the data gradient)
    EpiStore* const st = TMA_OUT ? &est : nullptr;
    int ts = 0;
    uint32_t tphase = 0;
    float alpha = 0.f;
    if (MODE == CONV_SCALE || seed) alpha = __ldg(p.alpha);
    // Data-gradient epilogue operands (ReLU mask + tap seed, or pool routing bytes; for conv1_1 the pixel-term gradient)
    // do not depend on the accumulator.  All chunks of a tile are requested together when they fit in registers
    // (<= 4 chunks), and for narrow tiles (<= 2 chunks, and conv1_1) the NEXT tile's operands are requested before the
    // current tile is processed: with ~1 us of tensor work per tile a global-memory latency per tile (cold: inside the
    // step these operands were written by earlier launches, they are not L2-warm as in a re-run) is the whole epilogue.
    constexpr int DG_CHUNKS = COLS / DG_CH;
    constexpr int DG_PRE = DG_CHUNKS <= 4 ? DG_CHUNKS : 1;
    constexpr bool XT = MODE == CONV_DGRAD_PIX || (MODE == CONV_DGRAD && DG_PRE == DG_CHUNKS && DG_PRE <= 2);
    DgradAux aux[DG_PRE];
    DgradAux aux_x[XT ? DG_PRE : 1];
    DgradAux aux_nxt;
    float gp[3] = {0.f, 0.f, 0.f}, gp_x[3] = {0.f, 0.f, 0.f};
    // h_ addresses the [batch * H][W][C] view of the tensors (image b's rows follow image b - 1's); valid_ is the pixel's
    // place inside its own image
    auto coords = [&](int tile_, int& h_, int& w_, int& n0_, bool& valid_, int& b_) {
      int nt_, th_, tw_;
      item_coords(tile_, nt_, b_, th_, tw_);
      h_ = th_ * TILE_H + hl;
      w_ = tw_ * TILE_W + wl;
      n0_ = nt_ * BLOCK_N;
      valid_ = h_ < p.H && w_ < p.W;
      h_ += b_ * p.H;
    };
    auto request = [&](int tile_, DgradAux* a_, float* g_) {
      int h_, w_, n0_, b_;
      bool valid_;
      coords(tile_, h_, w_, n0_, valid_, b_);
      if constexpr (MODE == CONV_DGRAD) {
#pragma unroll
        for (int c = 0; c < DG_PRE; ++c) dgrad_aux_load(p, a_[c], h_, w_, n0_ + col0 + c * DG_CH, valid_);
      }
      if constexpr (MODE == CONV_DGRAD_PIX) {
        if (valid_ && p.grad_pix != nullptr) {
          const size_t HW_ = static_cast<size_t>(p.H) * p.W;
          const size_t o_ = static_cast<size_t>(h_) * p.W + w_;
#pragma unroll
          for (int c = 0; c < 3; ++c) g_[c] = __ldg(p.grad_pix + c * HW_ + o_);
        }
      }
    };
    if constexpr (XT) {
      if (worker < num_items) request(worker, aux, gp);
    }
    const uint32_t tempty_leader = PAIR ? mapa_u32(smem_u32(&tempty_bar[0]), 0) : 0u;   // the leader's barriers (8 bytes apart)
    for (int tile = worker; tile < num_items; tile += workers) {
      int h, w, n0, bimg;
      bool valid;
      coords(tile, h, w, n0, valid, bimg);
      if (p.batch > 1 && (MODE == CONV_SCALE || seed)) alpha = __ldg(p.alpha + bimg * p.alpha_stride);
      est.w0 = w - wl;            // this warp's 8 x 4 pixel block of the tile, inside its image
      est.h0 = h - bimg * p.H - hl + q * 4;
      est.b = bimg;
      if constexpr (MODE == CONV_FWD) {
        for (int j = et; j < BLOCK_N; j += Cfg::EPI_WARPS * 32) sbias[ts * BLOCK_N + j] = __ldg(p.bias + n0 + j);
        asm volatile("bar.sync 1, %0;" ::"n"(Cfg::EPI_WARPS * 32) : "memory");  // the epilogue warps only
      }
      if constexpr (XT) {
        if (tile + workers < num_items) request(tile + workers, aux_x, gp_x);
      } else if constexpr (MODE == CONV_DGRAD) {
        request(tile, aux, gp);
      }
      NST_WAIT(wacc0, mbar_wait(&tfull_bar[ts], tphase));
      tc_fence_after();
      NST_STAMP(4, threadIdx.x == 128 && tile == worker);
      if (NST_TL_PTR(p) != nullptr && threadIdx.x == 128 && tile + workers >= num_items) atomicMax(&NST_TL_PTR(p)[3], globaltimer_ns());
      const uint32_t taddr =
          tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(ts * BLOCK_N + col0);
      if constexpr (MODE == CONV_DGRAD_PIX) {
        // conv1_1: 3 of the 16 accumulator columns are image channels
        uint32_t r[16];
        tmem_ld16(taddr, r);
        tmem_ld_wait();
        if (valid) {
          const size_t HW = static_cast<size_t>(p.H) * p.W;
          const size_t o = static_cast<size_t>(h) * p.W + w;
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float g = __uint_as_float(r[c]) * p.inv_std[c] + gp[c];  // gp stays 0 without pixel terms
            p.out_pix[c * HW + o] = g;
          }
        }
      } else if constexpr (MODE == CONV_DGRAD) {
        if constexpr (DG_PRE == DG_CHUNKS) {
#pragma unroll
          for (int c = 0; c < DG_CHUNKS; ++c) {
            uint32_t r[DG_CH];
            tmem_ld16(taddr + c * DG_CH, r);
            tmem_ld_wait();
            float v[DG_CH];
#pragma unroll
            for (int j = 0; j < DG_CH; ++j) v[j] = __uint_as_float(r[j]);
            if (seed) {
              uint32_t r2[DG_CH];
              tmem_ld16(taddr + 2 * BLOCK_N + c * DG_CH, r2);
              tmem_ld_wait();
              float sv[DG_CH];
#pragma unroll
              for (int j = 0; j < DG_CH; ++j) sv[j] = alpha * __uint_as_float(r2[j]);
              epilogue_dgrad(p, v, h, w, n0 + col0 + c * DG_CH, valid, aux[c], st, sv);
            } else {
              epilogue_dgrad(p, v, h, w, n0 + col0 + c * DG_CH, valid, aux[c], st);
            }
          }
        } else {
#pragma unroll 1
          for (int c = 0; c < DG_CHUNKS; ++c) {
            if (c + 1 < DG_CHUNKS) dgrad_aux_load(p, aux_nxt, h, w, n0 + col0 + (c + 1) * DG_CH, valid);
            uint32_t r[DG_CH];
            tmem_ld16(taddr + c * DG_CH, r);
            tmem_ld_wait();
            float v[DG_CH];
#pragma unroll
            for (int j = 0; j < DG_CH; ++j) v[j] = __uint_as_float(r[j]);
            if (seed) {
              uint32_t r2[DG_CH];
              tmem_ld16(taddr + 2 * BLOCK_N + c * DG_CH, r2);
              tmem_ld_wait();
              float sv[DG_CH];
#pragma unroll
              for (int j = 0; j < DG_CH; ++j) sv[j] = alpha * __uint_as_float(r2[j]);
              epilogue_dgrad(p, v, h, w, n0 + col0 + c * DG_CH, valid, aux[0], st, sv);
            } else {
              epilogue_dgrad(p, v, h, w, n0 + col0 + c * DG_CH, valid, aux[0], st);
            }
            aux[0] = aux_nxt;
          }
        }
      } else {
#pragma unroll 1
        for (int c = 0; c < COLS / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          const int n = n0 + col0 + c * 32;
          if constexpr (MODE == CONV_FWD) epilogue_fwd(p, v, h, w, n, valid, lane, sbias + ts * BLOCK_N + col0 + c * 32, st);
          if constexpr (MODE == CONV_SCALE) epilogue_scale(p, v, h, w, n, valid, alpha, st);
        }
      }
      if constexpr (XT) {
#pragma unroll
        for (int c = 0; c < DG_PRE; ++c) aux[c] = aux_x[c];
#pragma unroll
        for (int c = 0; c < 3; ++c) gp[c] = gp_x[c];
      }
      tc_fence_before();
      __syncwarp();
      NST_STAMP(5, threadIdx.x == 128 && tile == worker);
      if (lane == 0) {
        if constexpr (PAIR) mbar_arrive_cluster(tempty_leader + static_cast<uint32_t>(ts) * 8u);
        else mbar_arrive(&tempty_bar[ts]);
      }
      if (++ts == 2) {
        ts = 0;
        tphase ^= 1u;
      }
    }
  }

  if (dbg && threadIdx.x == 128) NST_DBG_PTR(p)[11] = wacc0;
  if (TMA_OUT && warp >= 4 && lane == 0) bulk_wait_all();  // the staging buffers must outlive the stores reading them
  tc_fence_before();
  // pair: neither CTA may leave (shared memory, barriers, tensor memory) while the other still refers to it
  if constexpr (PAIR) cluster_sync_relaxed(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_2sm(tmem_base, MODE == CONV_DGRAD ? Cfg::TMEM_COLS_SEED : Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, MODE == CONV_DGRAD ? Cfg::TMEM_COLS_SEED : Cfg::TMEM_COLS);
  }
  NST_STAMP(6, threadIdx.x == 64);
  if (NST_DBG_PTR(p) != nullptr && threadIdx.x == 64) NST_DBG_PTR(p)[16 + blockIdx.x] = clock64() - life0;  // lifetime of every CTA
  if (NST_TL_PTR(p) != nullptr && threadIdx.x == 64) atomicMax(&NST_TL_PTR(p)[1], globaltimer_ns());
#undef NST_STAMP
#undef NST_WAIT
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || ptr == nullptr) return nullptr;
    fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  }
  return fn;
}

int make_tmap_act(CUtensorMap* out, const void* base, int H, int W, int C, int box_c, int box_w, int box_h, int batch) {
  auto fn = get_encode_fn();
  if (!fn) return -1;
  // always 4-D (C, W, H, image): with one image the last coordinate is 0
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(batch < 1 ? 1 : batch)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W) * C * 2, static_cast<cuuint64_t>(H) * W * C * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  // FLOAT16 and BFLOAT16 only differ for NaN fill; both are 2-byte element moves
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 1000;
}

int make_tmap_out(CUtensorMap* out, void* base, int H, int W, int C, int box_c, int box_w, int box_h, int batch) {
  auto fn = get_encode_fn();
  if (!fn) return -1;
  // always 4-D (C, W, H, image): a box that overhangs the bottom of its image is clipped there, not written into the next one
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(W), static_cast<cuuint64_t>(H),
                        static_cast<cuuint64_t>(batch < 1 ? 1 : batch)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(W) * C * 2, static_cast<cuuint64_t>(H) * W * C * 2};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUtensorMapSwizzle sw = box_c * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                                  : (box_c * 2 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  if (box_c * 2 != 128 && box_c * 2 != 64 && box_c * 2 != 32) return -2;
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                  CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 1000;
}

int conv_taps_per_stage(int block_n, int taps) {
  if (taps != 9) return 1;
  return block_n >= 256 ? ConvCfg<256>::TPS : (block_n >= 128 ? ConvCfg<128>::TPS : (block_n >= 64 ? ConvCfg<64>::TPS : ConvCfg<16>::TPS));
}

bool conv_use_pair(int mode, int block_n, int taps, int K, int N) {
  // NST_PAIR: 0 = never; otherwise a bit mask: bit 0 forward, bit 1 data gradient, bit 2 also the forward layers whose whole
  // weight set stays resident (one 64-channel slice, one N tile: conv1_2).  Those are bound by their epilogue, not by the
  // operand stream, and lose as a pair (two CTAs in lock step per tile: 34.7 vs 31.9 us inside the step at 512^2).
  static const int mask = getenv("NST_PAIR") ? atoi(getenv("NST_PAIR")) : 7;
  if (taps != 9 || (block_n != 64 && block_n != 128)) return false;
  if (mode == CONV_FWD) return (mask & 1) != 0 && ((mask & 4) != 0 || !(K == BLOCK_K && N == block_n));
  if (mode == CONV_DGRAD) return (mask & 2) != 0;
  return false;
}

int make_tmap_wgt(CUtensorMap* out, const void* base, int taps, int N, int K, int box_n, bool pair) {
  auto fn = get_encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(N), static_cast<cuuint64_t>(taps)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(K) * 2, static_cast<cuuint64_t>(N) * K * 2};
  // a CTA of a pair loads half of the N tile's rows per stage
  cuuint32_t box[3] = {static_cast<cuuint32_t>(BLOCK_K), static_cast<cuuint32_t>(pair ? box_n / 2 : box_n),
                       static_cast<cuuint32_t>(conv_taps_per_stage(box_n, taps))};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 1000;
}

// Picks the N tile that minimises the modelled kernel time: (tiles per CTA) x (main loop of one tile) + the epilogue of
// the last tile (earlier epilogues overlap the next tile's main loop).  Measured constants (profiles/r01_mma_issue_rate.log,
// r01_conv_phases_v3_issue_loop.log): a K=16 step costs max(tensor floor 128 N / 256, ~48 issue cycles); the epilogue of a
// 128 x N tile about 20 cycles per output column.  Wide tiles are cheaper per FLOP; narrow tiles fill the 148 SMs on small
// images and let the epilogue of the first half of a pixel tile hide behind the main loop of the second half.
int conv_block_n(int N, int H, int W, int k_total, int num_sms) {
  const int sp = ((W + TILE_W - 1) / TILE_W) * ((H + TILE_H - 1) / TILE_H);
  if (num_sms < 1) num_sms = 148;
  const int ksteps = k_total / UMMA_K;
  int best = 64;
  long best_cost = -1;
  const int cand[3] = {256, 128, 64};
  for (int i = 0; i < 3; ++i) {
    if (cand[i] > N || N % cand[i] != 0) continue;
    const long tiles = static_cast<long>(sp) * (N / cand[i]);
    const long per_cta = (tiles + num_sms - 1) / num_sms;
    const long step = cand[i] / 2 > 48 ? cand[i] / 2 : 48;
    const long cost = per_cta * ksteps * step + 20L * cand[i];
    if (best_cost < 0 || cost < best_cost) {
      best_cost = cost;
      best = cand[i];
    }
  }
  return best;
}

void conv_finalize_params(ConvParams& p, int mode) {
  static const bool single_issue = getenv("NST_SINGLE_ISSUE") != nullptr;
  p.dual_issue = single_issue ? 0 : 1;
  const int bn = p.block_n;
  if (p.batch < 1) p.batch = 1;
  p.tiles_w = (p.W + TILE_W - 1) / TILE_W;
  p.tiles_h = (p.H + TILE_H - 1) / TILE_H;
  p.tiles_n = p.N / bn;
  p.num_tiles = p.tiles_w * p.tiles_h * p.tiles_n * p.batch;
  p.idesc = umma_idesc_f16(p.pair ? 2 * BLOCK_M : BLOCK_M, bn, (mode == CONV_DGRAD || mode == CONV_DGRAD_PIX) ? 1 : 0, 0, 0);
}

int conv_grid_ctas(const ConvParams& p, int num_sms) {
  if (!p.pair) return p.num_tiles < num_sms ? p.num_tiles : num_sms;
  const int items = ((p.tiles_w * p.tiles_h + 1) / 2) * p.tiles_n * (p.batch < 1 ? 1 : p.batch);
  const int pairs = items < num_sms / 2 ? items : num_sms / 2;
  return 2 * pairs;
}

template <int BLOCK_N, int MODE, bool TMA_OUT, bool PAIR>
static cudaError_t launch_one_t(const ConvParams& p, int num_sms, cudaStream_t stream) {
  using Cfg = ConvCfg<BLOCK_N, PAIR, MODE>;
  const int grid = conv_grid_ctas(p, num_sms);
  static const bool pdl = getenv("NST_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(Cfg::NUM_THREADS);
  // the staging buffers of the TMA-store epilogue sit at the end of the allocation: a launch that stores directly does not
  // reserve them - shared memory and L1 share one array, and the data-gradient epilogue's 16-byte loads of 128-byte lines
  // live on L1 hits (34 KB less L1 cost conv1_2's data gradient 17 us inside the step)
  cfg.dynamicSmemBytes = Cfg::SMEM_BYTES - (TMA_OUT ? 0 : Cfg::EPI_BYTES);
  cfg.stream = stream;
  cudaLaunchAttribute attr[3];
  int na = 0;
  // NST_KERNEL_PRIO: the tensor-core launches get the highest launch priority, the optimizer's streaming passes the lowest
  // (lbfgs.cu), whatever the priority of the stream: when work of two frames shares the GPU an SM that frees up goes to a
  // waiting convolution CTA first
  static const int prio = [] {
    int least = 0, greatest = 0;
    cudaDeviceGetStreamPriorityRange(&least, &greatest);
    return getenv("NST_KERNEL_PRIO") != nullptr ? greatest : 1000;
  }();
  if (prio != 1000) {
    attr[na].id = cudaLaunchAttributePriority;
    attr[na].val.priority = prio;
    ++na;
  }
  if (PAIR) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  // A pair launch of a plan that shares the GPU with other plans' launches carries no programmatic-launch attribute
  // (ConvParams::no_pdl_pair, nst_plan_set_shared_gpu): two chains of cluster launches, each released early by its
  // predecessor, running beside each other on different streams hung about one run in four (four 256 x 256 frames in flight);
  // without the attribute on the pairs - or without pairs, or without any programmatic launch - no run did.  A single
  // chain keeps it: it is worth 5 % of the step (conv class 409 -> 461 us without).
  static const bool pdl_pairs = getenv("NST_NO_PDL_PAIR") == nullptr;
  if (pdl && (!PAIR || (pdl_pairs && !p.no_pdl_pair))) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, conv_tc_kernel<BLOCK_N, MODE, TMA_OUT, PAIR>, p);
}
template <int BLOCK_N, int MODE, bool TMA_OUT>
static cudaError_t launch_one_p(const ConvParams& p, int num_sms, cudaStream_t stream) {
  if constexpr ((MODE == CONV_FWD || MODE == CONV_DGRAD) && (BLOCK_N == 64 || BLOCK_N == 128)) {
    if (p.pair) return launch_one_t<BLOCK_N, MODE, TMA_OUT, true>(p, num_sms, stream);
  }
  if (p.pair) return cudaErrorInvalidValue;
  return launch_one_t<BLOCK_N, MODE, TMA_OUT, false>(p, num_sms, stream);
}
template <int BLOCK_N, int MODE>
static cudaError_t launch_one(const ConvParams& p, int num_sms, cudaStream_t stream) {
  if constexpr (MODE == CONV_DGRAD_PIX) {
    return launch_one_p<BLOCK_N, MODE, false>(p, num_sms, stream);
  } else {
    return p.tma_out ? launch_one_p<BLOCK_N, MODE, true>(p, num_sms, stream) : launch_one_p<BLOCK_N, MODE, false>(p, num_sms, stream);
  }
}

template <int BLOCK_N, int MODE>
static cudaError_t init_one() {
  cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, MODE, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       ConvCfg<BLOCK_N, false, MODE>::SMEM_BYTES);
  if constexpr (MODE != CONV_DGRAD_PIX) {
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, MODE, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               ConvCfg<BLOCK_N, false, MODE>::SMEM_BYTES);
  }
  if constexpr ((MODE == CONV_FWD || MODE == CONV_DGRAD) && (BLOCK_N == 64 || BLOCK_N == 128)) {
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, MODE, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               ConvCfg<BLOCK_N, true, MODE>::SMEM_BYTES);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(conv_tc_kernel<BLOCK_N, MODE, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               ConvCfg<BLOCK_N, true, MODE>::SMEM_BYTES);
  }
  return e;
}
template <int MODE>
static cudaError_t init_mode() {
  cudaError_t e = init_one<256, MODE>();
  if (e == cudaSuccess) e = init_one<128, MODE>();
  if (e == cudaSuccess) e = init_one<64, MODE>();
  return e;
}
cudaError_t conv_tc_init() {
  cudaError_t e = init_mode<CONV_FWD>();
  if (e == cudaSuccess) e = init_mode<CONV_DGRAD>();
  if (e == cudaSuccess) e = init_mode<CONV_SCALE>();
  if (e == cudaSuccess) e = init_one<16, CONV_DGRAD_PIX>();
  return e;
}

template <int MODE>
static cudaError_t launch_mode(const ConvParams& p, int num_sms, cudaStream_t stream) {
  switch (p.block_n) {
    case 256: return launch_one<256, MODE>(p, num_sms, stream);
    case 128: return launch_one<128, MODE>(p, num_sms, stream);
    default: return launch_one<64, MODE>(p, num_sms, stream);
  }
}

cudaError_t launch_conv_tc(const ConvParams& p, int mode, int num_sms, cudaStream_t stream) {
  if (mode == CONV_DGRAD_PIX) {
    if (p.K % BLOCK_K != 0 || p.N != 16 || p.block_n != 16 || p.num_tiles <= 0) return cudaErrorInvalidValue;
    return launch_one<16, CONV_DGRAD_PIX>(p, num_sms, stream);
  }
  if (p.K % BLOCK_K != 0 || p.N % 64 != 0 || p.num_tiles <= 0) return cudaErrorInvalidValue;
  switch (mode) {
    case CONV_FWD: return launch_mode<CONV_FWD>(p, num_sms, stream);
    case CONV_DGRAD: return launch_mode<CONV_DGRAD>(p, num_sms, stream);
    case CONV_SCALE: return launch_mode<CONV_SCALE>(p, num_sms, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace nst
