// A whole chain of VGG-19 convolution layers as ONE persistent tcgen05 launch with tile-level dataflow between layers.
//
// Replaces, for the hot path of the reference, the same library kernels as conv_tc.cu
//   multi_style_transfer/helper_functions.py:94-101 (Vgg19.forward: nn.Conv2d + ReLU + MaxPool2d) and the autograd
//   data-gradient chain behind run_style_transfer.py:140 (loss.backward()),
// but without a kernel boundary per layer.  Measured inside the captured step (profiles/r01_timeline_512_per_layer.log):
// a per-layer launch costs ~7 us of fixed time (launch gap, barrier / tensor-memory set-up, first operand latency, the
// epilogue of the last tile, a partially filled last wave) around main loops of 5..11 us - 40 % of the chain.
//
// Here the work list is every 128-pixel x BLOCK_N tile of every layer of the chain, in layer order; tile g runs on CTA
// g mod gridDim.x, so the tiles of layer L+1 start on the SMs that layer L no longer fills.  A tile of layer L+1 needs
// only the tiles of layer L that cover its 18 x 10 input patch: every finished tile bumps a counter of its spatial
// position (release), and the TMA producer of a dependent tile polls the (at most 4 x 4) counters it needs (acquire)
// before it issues the first patch load.  Every CTA walks its tiles in work-list order and a tile depends only on
// earlier entries of the list, so the wait graph is acyclic; all CTAs are co-resident (grid <= number of SMs, one CTA
// per SM).  Pipeline state (operand rings, the two accumulator stages in tensor memory) simply carries over from one
// layer to the next.  The main loop, the operand layouts and the epilogues are those of conv_tc.cu.
//
// Warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = scout (dependency polling, runs ahead of the producer; also allocates
// tensor memory), 3 = publisher, 4..11 = epilogue.
//
// Ordering: the epilogue writes activations with ordinary (generic proxy) stores, the consumer reads them with TMA (async
// proxy).  Writer: stores -> mbarrier arrive (epilogue warps) -> mbarrier wait, fence.proxy.async, __threadfence,
// red.release.gpu (publisher thread; the release is cumulative over the stores it observed through the barrier).
// Reader: ld.acquire.gpu (scout) -> shared-memory count -> fence.proxy.async, cp.async.bulk.tensor (producer thread).
#include "conv_chain.cuh"
#include "common.cuh"
#include "conv_epilogue.cuh"

#include <stdio.h>
#include <stdlib.h>

namespace nst {

static constexpr int TILE_H = CONV_TILE_H;  // 16
static constexpr int TILE_W = CONV_TILE_W;  // 8
static constexpr int BLOCK_K = 64;
static constexpr int UMMA_K = 16;
static constexpr int MAX_N = CHAIN_MAX_BLOCK_N;                   // 128
static constexpr int HALO_ROWS = (TILE_H + 2) * (TILE_W + 2);     // 180 pixels
static constexpr int HALO_TX_BYTES = HALO_ROWS * BLOCK_K * 2;     // 23040
static constexpr int HALO_STAGE_BYTES = 23 * 1024;
static constexpr int FLAT_TX_BYTES = TILE_H * TILE_W * BLOCK_K * 2;
static constexpr int HALO_STAGES = 3;
static constexpr int TPS = 3;                                     // filter taps per weight stage (one filter row)
static constexpr int B_STAGE_BYTES = TPS * MAX_N * BLOCK_K * 2;   // 49152
static constexpr int B_STAGES = 3;
static constexpr int OPERAND_BYTES = HALO_STAGES * HALO_STAGE_BYTES + B_STAGES * B_STAGE_BYTES;  // 218112
static constexpr int EPI_WARPS = 8;
static constexpr int NUM_THREADS = 128 + 32 * EPI_WARPS;
static constexpr int TMEM_COLS = 2 * MAX_N;
static constexpr int BIAS_BYTES = 2 * MAX_N * 4;
static constexpr int SMEM_BYTES = OPERAND_BYTES + 1024 + 256 + BIAS_BYTES;

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(int* p, int v) {
  asm volatile("red.release.gpu.global.add.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

struct TileCoord {
  int layer, th, tw, nt;
};

// work-list entry g -> (layer, spatial tile, channel tile); `li` is a cursor that only moves forward
__device__ __forceinline__ void locate(const ChainLayer* __restrict__ layers, int n_layers, int g, int& li, TileCoord& t) {
  while (li + 1 < n_layers && g >= __ldg(&layers[li + 1].item_base)) ++li;
  const ChainLayer& L = layers[li];
  const int tile = g - __ldg(&L.item_base);
  const int tiles_n = __ldg(&L.c.tiles_n), tiles_w = __ldg(&L.c.tiles_w);
  const int sp = tile / tiles_n;
  t.layer = li;
  t.nt = tile - sp * tiles_n;
  t.th = sp / tiles_w;
  t.tw = sp - t.th * tiles_w;
}

__global__ void __launch_bounds__(NUM_THREADS, 1)
conv_chain_kernel(const ChainLayer* __restrict__ layers, int n_layers, int total_items, long long* __restrict__ dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;
  uint8_t* sB = smem + HALO_STAGES * HALO_STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OPERAND_BYTES);
  uint64_t* afull_bar = bars;
  uint64_t* aempty_bar = afull_bar + HALO_STAGES;
  uint64_t* bfull_bar = aempty_bar + HALO_STAGES;
  uint64_t* bempty_bar = bfull_bar + B_STAGES;
  uint64_t* tfull_bar = bempty_bar + B_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint64_t* pfull_bar = tempty_bar + 2;    // epilogue -> publisher: the tile's stores have been issued
  uint64_t* pempty_bar = pfull_bar + 2;    // publisher -> epilogue: the slot's tile has been published
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pempty_bar + 2);
  volatile int* deps_ok = reinterpret_cast<volatile int*>(tmem_slot + 1);  // tiles of this CTA whose inputs are complete
  float* sbias = reinterpret_cast<float*>(smem + OPERAND_BYTES + 256);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // debug accounting (dbg == nullptr in production): SM cycles per CTA spent in the waits of every role, 16 slots per CTA
  //  0 kernel start  1 kernel end  2 producer: dependency wait  3 producer: ring full  4 MMA: operands not landed
  //  5 MMA: accumulator stage busy  6 epilogue: accumulator not ready  7 epilogue: publish (fences + counter)
  //  8 epilogue: tap-seed acquire  9 items
  long long dbg_acc0 = 0, dbg_acc1 = 0, dbg_acc2 = 0;
#define DBG_T0() const long long dbg_t0 = dbg != nullptr ? clock64() : 0
#define DBG_ADD(acc) do { if (dbg != nullptr) acc += clock64() - dbg_t0; } while (0)
  if (dbg != nullptr && threadIdx.x == 0) dbg[blockIdx.x * 16 + 0] = clock64();

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < HALO_STAGES; ++s) {
      mbar_init(&afull_bar[s], 1);
      mbar_init(&aempty_bar[s], 1);
    }
    for (int s = 0; s < B_STAGES; ++s) {
      mbar_init(&bfull_bar[s], 1);
      mbar_init(&bempty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], EPI_WARPS);
      mbar_init(&pfull_bar[s], EPI_WARPS);
      mbar_init(&pempty_bar[s], 1);
    }
    *deps_ok = 0;
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp == 0) {
    // ===================== TMA producer (one elected lane) =====================
    if (elect_one()) {
      int as = 0, bs = 0, li = 0, ord = 0;
      uint32_t aphase = 0, bphase = 0;
      for (int g = blockIdx.x; g < total_items; g += gridDim.x, ++ord) {
        TileCoord t;
        locate(layers, n_layers, g, li, t);
        const ChainLayer& L = layers[li];
        const int taps = __ldg(&L.c.taps);
        const int pad = taps == 9 ? 1 : 0;
        const int h0 = t.th * TILE_H, w0 = t.tw * TILE_W, n0 = t.nt * __ldg(&L.c.block_n);
        // the scout warp runs ahead through the work list and counts the tiles whose inputs are complete
        {
          DBG_T0();
          if (*deps_ok <= ord) {
            const long long t0 = clock64();
            while (*deps_ok <= ord) {
              if (clock64() - t0 > 4000000000ll) __trap();
            }
          }
          __threadfence_block();
          fence_proxy_async();
          DBG_ADD(dbg_acc0);
        }
        const CUtensorMap* tmA = &L.c.tmA;
        const CUtensorMap* tmB = &L.c.tmB;
        const int k_slices = __ldg(&L.c.K) / BLOCK_K;
        const uint32_t a_tx = taps == 9 ? HALO_TX_BYTES : FLAT_TX_BYTES;
        const int tps = taps == 9 ? TPS : 1;
        const uint32_t b_tx = static_cast<uint32_t>(tps) * static_cast<uint32_t>(__ldg(&L.c.block_n)) * BLOCK_K * 2;
        for (int ks = 0; ks < k_slices; ++ks) {
          {
            DBG_T0();
            mbar_wait(&aempty_bar[as], aphase ^ 1u);
            DBG_ADD(dbg_acc1);
          }
          mbar_arrive_expect_tx(&afull_bar[as], a_tx);
          tma_load_3d(sA + as * HALO_STAGE_BYTES, tmA, &afull_bar[as], ks * BLOCK_K, w0 - pad, h0 - pad);
          if (++as == HALO_STAGES) {
            as = 0;
            aphase ^= 1u;
          }
          for (int tap = 0; tap < taps; tap += tps) {
            {
              DBG_T0();
              mbar_wait(&bempty_bar[bs], bphase ^ 1u);
              DBG_ADD(dbg_acc1);
            }
            mbar_arrive_expect_tx(&bfull_bar[bs], b_tx);
            tma_load_3d(sB + bs * B_STAGE_BYTES, tmB, &bfull_bar[bs], ks * BLOCK_K, n0, tap);
            if (++bs == B_STAGES) {
              bs = 0;
              bphase ^= 1u;
            }
          }
        }
      }
      if (dbg != nullptr) {
        dbg[blockIdx.x * 16 + 2] = dbg_acc0;
        dbg[blockIdx.x * 16 + 3] = dbg_acc1;
      }
    }
  } else if (warp == 2) {
    // ===================== scout: dependency wait, ahead of the producer (whole warp) =====================
    // Polling a completion counter is an L2 round trip even when it is already satisfied; on the producer's path that
    // latency sat between two tiles of every dependent layer.  The scout polls, the producer reads a shared-memory count.
    int li = 0, ord = 0;
    for (int g = blockIdx.x; g < total_items; g += gridDim.x, ++ord) {
      TileCoord t;
      locate(layers, n_layers, g, li, t);
      const ChainLayer& L = layers[li];
      const int h0 = t.th * TILE_H, w0 = t.tw * TILE_W;
      DBG_T0();
#pragma unroll 1
      for (int d = 0; d < 2; ++d) {
        const int dl = __ldg(&L.dep_layer[d]);
        if (dl < 0) continue;
        const ChainLayer& P = layers[dl];
        const int rpt = __ldg(&L.dep_rpt[d]), cpt = __ldg(&L.dep_cpt[d]), halo = __ldg(&L.dep_halo[d]);
        const int pth = __ldg(&P.c.tiles_h), ptw = __ldg(&P.c.tiles_w), need = __ldg(&P.c.tiles_n);
        int r0 = (h0 - halo) / rpt, r1 = (h0 + TILE_H - 1 + halo) / rpt;
        int c0 = (w0 - halo) / cpt, c1 = (w0 + TILE_W - 1 + halo) / cpt;
        if (h0 - halo < 0) r0 = 0;
        if (w0 - halo < 0) c0 = 0;
        if (r1 > pth - 1) r1 = pth - 1;
        if (c1 > ptw - 1) c1 = ptw - 1;
        if (r0 > r1) r0 = r1;
        if (c0 > c1) c0 = c1;
        const int nc = c1 - c0 + 1, cnt = (r1 - r0 + 1) * nc;
        const int* done = P.done;
        for (int base = 0; base < cnt; base += 32) {
          const int i = base + lane;
          const int* flag = i < cnt ? done + (r0 + i / nc) * ptw + (c0 + i % nc) : nullptr;
          bool ok = flag == nullptr || ld_acquire(flag) >= need;
          if (!__all_sync(0xffffffffu, ok)) {
            const long long t0 = clock64();
            do {
              if (!ok) {
                __nanosleep(32);
                ok = ld_acquire(flag) >= need;
              }
              if (clock64() - t0 > 4000000000ll) __trap();  // a lost completion becomes an error, not a hung GPU
            } while (!__all_sync(0xffffffffu, ok));
          }
        }
      }
      __syncwarp();
      if (dbg != nullptr && lane == 0) {
        const long long dt = clock64() - dbg_t0;
        dbg_acc2 += dt;
        atomicAdd(reinterpret_cast<unsigned long long*>(dbg + 16 * gridDim.x + 4 * li), static_cast<unsigned long long>(dt));
      }
      if (lane == 0) {
        __threadfence_block();
        *deps_ok = ord + 1;
      }
    }
    if (dbg != nullptr && lane == 0) dbg[blockIdx.x * 16 + 10] = dbg_acc2;
  } else if (warp == 3) {
    // ===================== publisher: makes finished tiles visible to the other SMs (one elected lane) =====================
    // The epilogue warps only issue their stores and arrive on a barrier; this thread waits for them, fences once for the
    // whole CTA (release cumulativity covers the stores it observed through the barrier) and bumps the tile's counter.
    // When the next tile's stores have been issued as well, one fence covers both.
    if (elect_one()) {
      int ts = 0, li = 0;
      uint32_t tphase = 0;
      int g = blockIdx.x;
      while (g < total_items) {
        TileCoord t;
        locate(layers, n_layers, g, li, t);
        int* flag0 = layers[li].done + t.th * __ldg(&layers[li].c.tiles_w) + t.tw;
        mbar_wait(&pfull_bar[ts], tphase);
        const int g1 = g + gridDim.x;
        const int ts1 = ts ^ 1;
        const uint32_t tphase1 = ts == 1 ? tphase ^ 1u : tphase;
        const bool two = g1 < total_items && mbar_try_wait(smem_u32(&pfull_bar[ts1]), tphase1);
        fence_proxy_async();
        __threadfence();
        red_release_add(flag0, 1);
        mbar_arrive(&pempty_bar[ts]);
        g = g1;
        ts = ts1;
        tphase = tphase1;
        if (two) {
          locate(layers, n_layers, g, li, t);
          red_release_add(layers[li].done + t.th * __ldg(&layers[li].c.tiles_w) + t.tw, 1);
          mbar_arrive(&pempty_bar[ts]);
          g += gridDim.x;
          if (ts == 1) tphase ^= 1u;
          ts ^= 1;
        }
      }
    }
  } else if (warp == 1 && elect_one()) {
    // ===================== MMA issuer (one elected thread; see conv_tc.cu) =====================
    int as = 0, bs = 0, ts = 0, li = 0;
    uint32_t aphase = 0, bphase = 0, tphase = 0;
    const uint32_t b_hi = static_cast<uint32_t>(umma_desc_sw128(0, 16, 1024) >> 32);
    const uint32_t lbo_lo = static_cast<uint32_t>(umma_desc_sw128(0, 16, 0) & 0xffffffffu);
    for (int g = blockIdx.x; g < total_items; g += gridDim.x) {
      while (li + 1 < n_layers && g >= __ldg(&layers[li + 1].item_base)) ++li;
      const ChainLayer& L = layers[li];
      const int k_slices = __ldg(&L.c.K) / BLOCK_K;
      const bool conv3x3 = __ldg(&L.c.taps) == 9;
      const uint32_t idesc = __ldg(&L.c.idesc);
      const uint32_t btile16 = static_cast<uint32_t>(__ldg(&L.c.block_n)) * (BLOCK_K * 2 / 16);  // one tap, in 16-byte units
      const uint32_t sbo = conv3x3 ? (TILE_W + 2) * 128u : TILE_W * 128u;
      const uint32_t a_hi = static_cast<uint32_t>(umma_desc_sw128(0, 16, sbo) >> 32);
      {
        DBG_T0();
        mbar_wait(&tempty_bar[ts], tphase ^ 1u);
        DBG_ADD(dbg_acc1);
      }
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(ts * MAX_N);
      uint32_t accumulate = 0;
      for (int ks = 0; ks < k_slices; ++ks) {
        {
          DBG_T0();
          mbar_wait(&afull_bar[as], aphase);
          DBG_ADD(dbg_acc0);
        }
        const uint32_t a_lo0 = lbo_lo | (smem_u32(sA + as * HALO_STAGE_BYTES) >> 4);
        if (conv3x3) {
#pragma unroll 1
          for (int gq = 0; gq < 3; ++gq) {
            {
              DBG_T0();
              mbar_wait(&bfull_bar[bs], bphase);
              DBG_ADD(dbg_acc0);
            }
            tc_fence_after();
            const uint32_t b_lo0 = lbo_lo | (smem_u32(sB + bs * B_STAGE_BYTES) >> 4);
            const uint32_t a_lo_g = a_lo0 + static_cast<uint32_t>(gq) * ((TILE_W + 2) * 8u);  // filter row gq
#pragma unroll
            for (int tt = 0; tt < TPS; ++tt) {
#pragma unroll
              for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                const uint64_t da = (static_cast<uint64_t>(a_hi) << 32) | (a_lo_g + static_cast<uint32_t>(tt * 8 + k * 2));
                const uint64_t db = (static_cast<uint64_t>(b_hi) << 32) | (b_lo0 + static_cast<uint32_t>(tt) * btile16 + static_cast<uint32_t>(k * 2));
                umma_f16(d_tmem, da, db, idesc, accumulate);
                accumulate = 1u;
              }
            }
            umma_commit(&bempty_bar[bs]);
            if (++bs == B_STAGES) {
              bs = 0;
              bphase ^= 1u;
            }
          }
        } else {
          mbar_wait(&bfull_bar[bs], bphase);
          tc_fence_after();
          const uint32_t b_lo0 = lbo_lo | (smem_u32(sB + bs * B_STAGE_BYTES) >> 4);
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = (static_cast<uint64_t>(a_hi) << 32) | (a_lo0 + static_cast<uint32_t>(k * 2));
            const uint64_t db = (static_cast<uint64_t>(b_hi) << 32) | (b_lo0 + static_cast<uint32_t>(k * 2));
            umma_f16(d_tmem, da, db, idesc, accumulate);
            accumulate = 1u;
          }
          umma_commit(&bempty_bar[bs]);
          if (++bs == B_STAGES) {
            bs = 0;
            bphase ^= 1u;
          }
        }
        umma_commit(&aempty_bar[as]);
        if (++as == HALO_STAGES) {
          as = 0;
          aphase ^= 1u;
        }
      }
      umma_commit(&tfull_bar[ts]);
      if (++ts == 2) {
        ts = 0;
        tphase ^= 1u;
      }
    }
    if (dbg != nullptr) {
      dbg[blockIdx.x * 16 + 4] = dbg_acc0;
      dbg[blockIdx.x * 16 + 5] = dbg_acc1;
    }
  } else if (warp >= 4) {
    // ===================== epilogue (eight warps: TMEM lane quarter x column half) =====================
    const int q = warp & 3;
    const int half = (warp - 4) >> 2;
    const int tq = q * 32 + lane;
    const int hl = tq / TILE_W, wl = tq % TILE_W;
    const int et = threadIdx.x - 128;
    int ts = 0, li = 0;
    uint32_t tphase = 0;
    for (int g = blockIdx.x; g < total_items; g += gridDim.x) {
      TileCoord t;
      locate(layers, n_layers, g, li, t);
      const ChainLayer& L = layers[li];
      const ConvParams& p = L.c;
      const int mode = __ldg(&L.mode);
      const int bn = __ldg(&p.block_n);
      const int cols = bn >> 1;  // accumulator columns of this warp
      const int col0 = half * cols;
      const int h = t.th * TILE_H + hl, w = t.tw * TILE_W + wl, n0 = t.nt * bn;
      const bool valid = h < p.H && w < p.W;
      DgradAux aux_cur, aux_nxt;
      float alpha = 0.f;
      if (mode == CONV_FWD) {
        for (int j = et; j < bn; j += EPI_WARPS * 32) sbias[ts * MAX_N + j] = __ldg(p.bias + n0 + j);
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_WARPS * 32) : "memory");
      } else if (mode == CONV_DGRAD) {
        // the mask / routing bytes come from the forward pass (an earlier kernel); the tap seed was produced earlier in
        // this launch and the producer warp has already acquired its completion counter for this tile - but that acquire
        // was made by another warp, so order this warp's reads after it with its own acquire of the same counter
        const int dl = __ldg(&L.dep_layer[1]);
        if (dl >= 0) {
          const ChainLayer& P = layers[dl];
          const int* flag = P.done + t.th * __ldg(&P.c.tiles_w) + t.tw;
          const int need = __ldg(&P.c.tiles_n);
          DBG_T0();
          const long long t0 = clock64();
          while (ld_acquire(flag) < need) {
            __nanosleep(64);
            if (clock64() - t0 > 4000000000ll) __trap();
          }
          DBG_ADD(dbg_acc2);
        }
        dgrad_aux_load(p, aux_cur, h, w, n0 + col0, valid);
      } else {
        alpha = __ldg(p.alpha);
      }
      {
        DBG_T0();
        mbar_wait(&tfull_bar[ts], tphase);
        DBG_ADD(dbg_acc0);
      }
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(ts * MAX_N + col0);
      if (mode == CONV_DGRAD) {
        const int chunks = cols / DG_CH;
#pragma unroll 1
        for (int c = 0; c < chunks; ++c) {
          if (c + 1 < chunks) dgrad_aux_load(p, aux_nxt, h, w, n0 + col0 + (c + 1) * DG_CH, valid);
          uint32_t r[DG_CH];
          tmem_ld16(taddr + c * DG_CH, r);
          tmem_ld_wait();
          float v[DG_CH];
#pragma unroll
          for (int j = 0; j < DG_CH; ++j) v[j] = __uint_as_float(r[j]);
          epilogue_dgrad(p, v, h, w, n0 + col0 + c * DG_CH, valid, aux_cur);
          aux_cur = aux_nxt;
        }
      } else {
        const int chunks = cols / 32;
#pragma unroll 1
        for (int c = 0; c < chunks; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          const int n = n0 + col0 + c * 32;
          if (mode == CONV_FWD) {
            epilogue_fwd(p, v, h, w, n, valid, lane, sbias + ts * MAX_N + col0 + c * 32);
          } else {
            epilogue_scale(p, v, h, w, n, valid, alpha);
          }
        }
      }
      // accumulator stage free again
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[ts]);
      const int pslot = ts;
      const uint32_t pphase = tphase;
      if (++ts == 2) {
        ts = 0;
        tphase ^= 1u;
      }
      // hand the tile to the publisher warp: the slot must have been published two tiles ago (it always has, bar a
      // publisher that fell behind), then one arrival per epilogue warp says "stores issued"
      if (lane == 0) {
        DBG_T0();
        mbar_wait(&pempty_bar[pslot], pphase ^ 1u);
        mbar_arrive(&pfull_bar[pslot]);
        DBG_ADD(dbg_acc1);
      }
      if (dbg != nullptr && et == 0) dbg[blockIdx.x * 16 + 9] += 1;
    }
    if (dbg != nullptr && et == 0) {
      dbg[blockIdx.x * 16 + 6] = dbg_acc0;
      dbg[blockIdx.x * 16 + 7] = dbg_acc1;
      dbg[blockIdx.x * 16 + 8] = dbg_acc2;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
  if (dbg != nullptr && threadIdx.x == 0) dbg[blockIdx.x * 16 + 1] = clock64();
#undef DBG_T0
#undef DBG_ADD
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int chain_block_n(int N) { return N >= CHAIN_MAX_BLOCK_N ? CHAIN_MAX_BLOCK_N : 64; }

cudaError_t conv_chain_init() {
  return cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
}

cudaError_t launch_conv_chain(const ChainLayer* layers_dev, int n_layers, int total_items, int num_sms, cudaStream_t stream,
                              long long* dbg) {
  if (n_layers < 1 || total_items < 1) return cudaErrorInvalidValue;
  const int grid = total_items < num_sms ? total_items : num_sms;
  static const bool pdl = getenv("NST_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, conv_chain_kernel, layers_dev, n_layers, total_items, dbg);
}

}  // namespace nst
