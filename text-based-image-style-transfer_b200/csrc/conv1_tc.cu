// conv1_1 forward (3 -> 64 channels, fused with normalize()) on tcgen05 tensor cores at fp32-class accuracy.
//
// Replaces, for the hot path of the reference, the first nn.Conv2d of
//   multi_style_transfer/helper_functions.py:94-101 (Vgg19.forward) together with
//   multi_style_transfer/style_transfer_losses.py:9-28 (normalize).
//
// The layer is 0.5 % of the FLOPs but writes the largest activation (2 x 33.5 MB at 512^2): it is bandwidth bound, and
// on CUDA cores the 453 M multiply-adds kept it at 36 us against a 10 us HBM floor.  K = 27 is padded to 32 and the
// operands are split into fp16 high and low parts, x = xh + xl, w = wh + wl (xl, wl = what fp16 rounding dropped), so
// that  x . w  ~=  xh . wh + xl . wh + xh . wl  with fp32 accumulation - a relative error of ~2^-21 per product, i.e. the
// accuracy class of an fp32 FMA chain, not of an fp16 convolution (rounding the first layer's operands to
// 10-bit mantissas is what sets the Gram error of every layer; this does not).
//
// There is no TMA-able operand: the input is [3][H][W] fp32.  Four builder warps (one thread per pixel of the 16 x 8 tile)
// normalise a 18 x 10 x 3 patch, split it and write the im2col rows straight into the 128-byte-swizzled K-major layout
// tcgen05.mma reads:  A1 = [xh | xl] (K = 64),  A2 = [xh] (K = 32);  the weights  B1 = [wh | wh],  B2 = [wl]  are built once
// per CTA.  Six MMAs (M128 x N64 x K16) per tile.  Epilogue = the forward epilogue of conv_tc.cu (bias, pre-ReLU tap,
// ReLU, activation; both outputs through swizzled shared memory + TMA stores).  Persistent, two TMEM accumulator stages.
#include "conv_tc.cuh"
#include "common.cuh"
#include "conv_epilogue.cuh"
#include "pixel.cuh"

#include <stdlib.h>

namespace nst {

static constexpr int C1_TILE_H = CONV_TILE_H;  // 16
static constexpr int C1_TILE_W = CONV_TILE_W;  // 8
static constexpr int C1_N = 64;
static constexpr int C1_A1_BYTES = 128 * 128;              // 128 pixels x [xh(32) | xl(32)] fp16
static constexpr int C1_A2_BYTES = 128 * 128;              // 128 pixels x [xh(32) | unused]
static constexpr int C1_STAGE_BYTES = C1_A1_BYTES + C1_A2_BYTES;
// im2col stages in flight.  Measured inside the step on one box (bench.py, two repetitions): 2 / 3 / 4 stages -> 30.7 / 32.5 /
// 35.3 us per launch: the builder warps read the image through L1, which shares its 256 KB with shared memory - every 32 KB
// stage is 32 KB less cache for the 3x overlapping patch reads.
#ifndef NST_C1_STAGES
#define NST_C1_STAGES 2
#endif
static constexpr int C1_STAGES = NST_C1_STAGES;
static constexpr int C1_B_BYTES = 2 * C1_N * 128;          // B1, B2: 64 rows x 128 B each
static constexpr int C1_PATCH = 3 * (C1_TILE_H + 2) * (C1_TILE_W + 2);  // 540 normalised input values
static constexpr int C1_PATCH_PER_THREAD = (C1_PATCH + 127) / 128;      // 5
static constexpr int C1_EPI_WARPS = 8;
static constexpr int C1_THREADS = 256 + 32 * C1_EPI_WARPS;  // warps 0..3 control, 4..7 builders, 8..15 epilogue
static constexpr int C1_MISC_BYTES = 2048;                  // barriers + bias
static constexpr int C1_SMEM_BYTES = C1_STAGES * C1_STAGE_BYTES + C1_B_BYTES + 1024 + C1_MISC_BYTES + 2 * 4 * C1_PATCH +
                                     1024 + C1_EPI_WARPS * 8192;

// 16-byte piece `chunk` (0..7) of row `row` in a K-major 128-byte-swizzled operand whose 8-row groups are 1024 B apart
__device__ __forceinline__ uint4* sw128_piece(uint8_t* base, int row, int chunk) {
  return reinterpret_cast<uint4*>(base + row * 128 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ void split_h(float v, __half& hi, __half& lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}
__device__ __forceinline__ uint32_t pack2(__half a, __half b) {
  return static_cast<uint32_t>(__half_as_ushort(a)) | (static_cast<uint32_t>(__half_as_ushort(b)) << 16);
}

__global__ void __launch_bounds__(C1_THREADS, 1)
conv1_tc_kernel(const __grid_constant__ ConvParams p, const float* __restrict__ x, const float* __restrict__ wgt, PixelConsts pc) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* sA = smem;                                     // C1_STAGES x (A1, A2)
  uint8_t* sB = smem + C1_STAGES * C1_STAGE_BYTES;        // B1 then B2
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + C1_B_BYTES);
  uint64_t* afull_bar = bars;                             // builders -> MMA (4 arrivals: one per builder warp)
  uint64_t* aempty_bar = afull_bar + C1_STAGES;           // MMA -> builders
  uint64_t* tfull_bar = aempty_bar + C1_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* sbias = reinterpret_cast<float*>(sB + C1_B_BYTES + 256);  // [2][64]
  float* spatch = reinterpret_cast<float*>(sB + C1_B_BYTES + C1_MISC_BYTES);  // [2][540]
  uint8_t* sEpi = sB + C1_B_BYTES + C1_MISC_BYTES + 2 * 4 * C1_PATCH;
  sEpi += (1024u - (smem_u32(sEpi) & 1023u)) & 1023u;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmO0);
    tma_prefetch_desc(&p.tmO1);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < C1_STAGES; ++s) {
      mbar_init(&afull_bar[s], 4);
      mbar_init(&aempty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);  // the four epilogue warps of the group that owns the stage
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  // weights: B1 = [wh | wh], B2 = [wl | -]; row n = output channel, k = c*9 + r*3 + s (27, zero-padded to 32)
  for (int i = threadIdx.x; i < C1_N * 4; i += C1_THREADS) {
    const int n = i >> 2, q = i & 3;  // 16-byte piece q = k 8q .. 8q+7
    __half hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = 8 * q + e;
      const float w = k < 27 ? __ldg(wgt + n * 27 + k) : 0.f;
      split_h(w, hi[e], lo[e]);
    }
    const uint4 uh = make_uint4(pack2(hi[0], hi[1]), pack2(hi[2], hi[3]), pack2(hi[4], hi[5]), pack2(hi[6], hi[7]));
    const uint4 ul = make_uint4(pack2(lo[0], lo[1]), pack2(lo[2], lo[3]), pack2(lo[4], lo[5]), pack2(lo[6], lo[7]));
    *sw128_piece(sB, n, q) = uh;
    *sw128_piece(sB, n, 4 + q) = uh;
    *sw128_piece(sB + C1_N * 128, n, q) = ul;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  asm volatile("griddepcontrol.wait;" ::: "memory");

  const int tiles_w = p.tiles_w;
  const int num_tiles = p.num_tiles;

  if (warp == 1) {
    // ===================== MMA issuer =====================
    if (elect_one()) {
      int as = 0, ts = 0;
      uint32_t aphase = 0, tphase = 0;
      const uint32_t idesc = umma_idesc_f16(128, C1_N, 0, 0, 0);
      const uint32_t hi32 = static_cast<uint32_t>(umma_desc_sw128(0, 16, 1024) >> 32);
      const uint32_t lbo_lo = static_cast<uint32_t>(umma_desc_sw128(0, 16, 0) & 0xffffffffu);
      const uint32_t b1_lo = lbo_lo | (smem_u32(sB) >> 4);
      const uint32_t b2_lo = lbo_lo | (smem_u32(sB + C1_N * 128) >> 4);
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        mbar_wait(&tempty_bar[ts], tphase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(ts * C1_N);
        mbar_wait(&afull_bar[as], aphase);
        tc_fence_after();
        const uint32_t a1_lo = lbo_lo | (smem_u32(sA + as * C1_STAGE_BYTES) >> 4);
        const uint32_t a2_lo = lbo_lo | (smem_u32(sA + as * C1_STAGE_BYTES + C1_A1_BYTES) >> 4);
#pragma unroll
        for (int k = 0; k < 4; ++k)  // [xh | xl] . [wh | wh]
          umma_f16(d_tmem, (static_cast<uint64_t>(hi32) << 32) | (a1_lo + k * 2), (static_cast<uint64_t>(hi32) << 32) | (b1_lo + k * 2),
                   idesc, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 2; ++k)  // xh . wl
          umma_f16(d_tmem, (static_cast<uint64_t>(hi32) << 32) | (a2_lo + k * 2), (static_cast<uint64_t>(hi32) << 32) | (b2_lo + k * 2),
                   idesc, 1u);
        umma_commit(&aempty_bar[as]);
        umma_commit(&tfull_bar[ts]);
        if (++as == C1_STAGES) {
          as = 0;
          aphase ^= 1u;
        }
        if (++ts == 2) {
          ts = 0;
          tphase ^= 1u;
        }
      }
    }
  } else if (warp >= 4 && warp < 8) {
    // ===================== builders: one thread = one pixel of the tile =====================
    const int t = threadIdx.x - 128;
    const int hl = t / C1_TILE_W, wl = t % C1_TILE_W;
    const size_t HW = static_cast<size_t>(p.H) * p.W;
    // The input patch of a tile is fetched TWO tiles ahead (two register sets, the loop is unrolled by two so that they
    // swap roles without copies): with ~0.3 us of work per tile a single tile of look-ahead left every iteration waiting
    // for a global-memory round trip (measured: 3 us per tile).
    float ra[C1_PATCH_PER_THREAD], rb[C1_PATCH_PER_THREAD];
    // normalised patch value e of a tile (zero padding lives in the normalised domain)
    auto fetch = [&](int tile, float (&dst)[C1_PATCH_PER_THREAD]) {
      if (tile >= num_tiles) return;
      const int w0 = (tile % tiles_w) * C1_TILE_W, h0 = (tile / tiles_w) * C1_TILE_H;
#pragma unroll
      for (int q = 0; q < C1_PATCH_PER_THREAD; ++q) {
        const int i = t + q * 128;
        dst[q] = 0.f;
        if (i < C1_PATCH) {
          const int c = i / ((C1_TILE_H + 2) * (C1_TILE_W + 2));
          const int rem = i - c * ((C1_TILE_H + 2) * (C1_TILE_W + 2));
          const int sy = rem / (C1_TILE_W + 2), sx = rem - sy * (C1_TILE_W + 2);
          const int hh = h0 + sy - 1, ww = w0 + sx - 1;
          // raw value only: any arithmetic here would make the warp wait for the load at once.  Outside the image the
          // channel mean stands in, which normalises to the zero padding of the normalised domain.
          dst[q] = pc.mean[c];
          if (hh >= 0 && hh < p.H && ww >= 0 && ww < p.W) {
            asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(dst[q]) : "l"(x + c * HW + static_cast<size_t>(hh) * p.W + ww));  // predicated load, no select
          }
        }
      }
    };
    auto stash = [&](int buf, const float (&src)[C1_PATCH_PER_THREAD]) {
#pragma unroll
      for (int q = 0; q < C1_PATCH_PER_THREAD; ++q) {
        const int i = t + q * 128;
        if (i < C1_PATCH) {
          const int c = i / ((C1_TILE_H + 2) * (C1_TILE_W + 2));
          spatch[buf * C1_PATCH + i] = (src[q] - pc.mean[c]) / pc.stdv[c];  // normalize(), style_transfer_losses.py:9-28
        }
      }
    };
    int as = 0;
    uint32_t aphase = 0;
    const int stride = gridDim.x;
    // builds the rows of `tile` from patch buffer `buf`, then stashes `next_regs` (the patch of tile + stride) into the
    // other buffer and starts fetching tile + 2 * stride into `far_regs`' place... (the caller orders the register sets)
    const bool dbg = NST_DBG_PTR(p) != nullptr && blockIdx.x == 0 && t == 0;
    long long acc_build = 0, acc_wait = 0, acc_write = 0, acc_sync = 0;
    auto build = [&](int buf) {
      const long long c0 = dbg ? clock64() : 0;
      const float* pt = spatch + buf * C1_PATCH;
      __half hi[32], lo[32];
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        if (k < 27) {
          const int c = k / 9, r = (k % 9) / 3, s2 = k % 3;
          split_h(pt[(c * (C1_TILE_H + 2) + hl + r) * (C1_TILE_W + 2) + wl + s2], hi[k], lo[k]);
        } else {
          hi[k] = __ushort_as_half(0);
          lo[k] = __ushort_as_half(0);
        }
      }
      const long long c1 = dbg ? clock64() : 0;
      if (lane == 0) mbar_wait(&aempty_bar[as], aphase ^ 1u);
      __syncwarp();
      const long long c2 = dbg ? clock64() : 0;
      uint8_t* a1 = sA + as * C1_STAGE_BYTES;
      uint8_t* a2 = a1 + C1_A1_BYTES;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint4 uh = make_uint4(pack2(hi[8 * q], hi[8 * q + 1]), pack2(hi[8 * q + 2], hi[8 * q + 3]), pack2(hi[8 * q + 4], hi[8 * q + 5]),
                                    pack2(hi[8 * q + 6], hi[8 * q + 7]));
        const uint4 ul = make_uint4(pack2(lo[8 * q], lo[8 * q + 1]), pack2(lo[8 * q + 2], lo[8 * q + 3]), pack2(lo[8 * q + 4], lo[8 * q + 5]),
                                    pack2(lo[8 * q + 6], lo[8 * q + 7]));
        *sw128_piece(a1, t, q) = uh;
        *sw128_piece(a1, t, 4 + q) = ul;
        *sw128_piece(a2, t, q) = uh;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to tcgen05.mma
      __syncwarp();
      if (lane == 0) mbar_arrive(&afull_bar[as]);
      if (dbg) {
        const long long c3 = clock64();
        acc_build += c1 - c0;
        acc_wait += c2 - c1;
        acc_write += c3 - c2;
      }
      if (++as == C1_STAGES) {
        as = 0;
        aphase ^= 1u;
      }
    };
    const int t0 = blockIdx.x;
    fetch(t0, ra);
    fetch(t0 + stride, rb);
    if (t0 < num_tiles) stash(0, ra);
    asm volatile("bar.sync 3, 128;" ::: "memory");
    for (int tile = t0; tile < num_tiles; tile += 2 * stride) {
      // even step: patch of `tile` is in buffer 0, rb holds tile + stride, ra is free
      fetch(tile + 2 * stride, ra);
      build(0);
      const long long c4 = dbg ? clock64() : 0;
      if (tile + stride < num_tiles) stash(1, rb);
      asm volatile("bar.sync 3, 128;" ::: "memory");
      if (dbg) acc_sync += clock64() - c4;
      if (tile + stride >= num_tiles) break;
      // odd step: patch of tile + stride is in buffer 1, ra holds tile + 2 stride, rb is free
      fetch(tile + 3 * stride, rb);
      build(1);
      if (tile + 2 * stride < num_tiles) stash(0, ra);
      asm volatile("bar.sync 3, 128;" ::: "memory");
    }
    if (dbg) {
      NST_DBG_PTR(p)[0] = acc_build;   // patch -> split rows (registers)
      NST_DBG_PTR(p)[1] = acc_wait;    // waiting for a free operand stage
      NST_DBG_PTR(p)[2] = acc_write;   // swizzled shared-memory writes + proxy fence + arrive
      NST_DBG_PTR(p)[3] = acc_sync;    // stash of the prefetched patch + builder barrier (even steps only)
    }
  } else if (warp >= 8) {
    // ===================== epilogue =====================
    // Two groups of four warps alternate tiles (group g owns accumulator stage g); a warp owns 32 pixels x all 64 channels,
    // so a pixel's output is one full 128-byte line: both outputs are staged as 32 rows x 128 B (128-byte swizzle) and leave
    // as one 4 KB TMA store each - half the box rows of the 32-channel stores of conv_tc.cu, and whole lines.
    const int ew = warp - 8;
    const int q = ew & 3;
    const int g = ew >> 2;
    const int t = q * 32 + lane;
    uint8_t* buf_tap = sEpi + ew * 8192;
    uint8_t* buf_act = buf_tap + 4096;
    uint32_t tphase = 0;
    const int et = threadIdx.x - 256;
    if (et < C1_N) sbias[et] = __ldg(p.bias + et);
    asm volatile("bar.sync 1, %0;" ::"n"(C1_EPI_WARPS * 32) : "memory");
    const bool skip = NST_DBG_FLAG(p, 1);
    int k = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++k) {
      if ((k & 1) != g) continue;
      const int th = tile / tiles_w, tw = tile - th * tiles_w;
      const long long e0 = (NST_DBG_PTR(p) != nullptr && blockIdx.x == 0 && ew == 0 && lane == 0) ? clock64() : 0;
      mbar_wait(&tfull_bar[g], tphase);
      tphase ^= 1u;
      tc_fence_after();
      if (e0) NST_DBG_PTR(p)[4] += clock64() - e0;  // epilogue group 0: waiting for the accumulator
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(g * C1_N);
      uint32_t r0[32], r1[32];
      tmem_ld32(taddr, r0);
      tmem_ld32(taddr + 32, r1);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[g]);  // the accumulator is in registers: release the stage before the stores
      if (lane == 0) bulk_wait_read<0>();           // the previous tile's stores have finished reading the buffers
      __syncwarp();
      const int x7 = t & 7;
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(hh == 0 ? r0[j] : r1[j]) + sbias[32 * hh + j];
        uint4 u[4];
        pack_h32(v, u);
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(buf_tap + lane * 128 + (((4 * hh + j) ^ (lane & 7)) << 4)) = u[j];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : 0.f;
        pack_h32(v, u);
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint4*>(buf_act + lane * 128 + (((4 * hh + j) ^ (lane & 7)) << 4)) = u[j];
      }
      (void)x7;
      fence_async_smem();
      __syncwarp();
      if (lane == 0 && !skip) {
        const int w0 = tw * C1_TILE_W, h0 = th * C1_TILE_H + q * 4;
        if (p.out_tap != nullptr) tma_store_4d(&p.tmO0, buf_tap, 0, w0, h0, 0);
        if (p.out_act != nullptr) tma_store_4d(&p.tmO1, buf_act, 0, w0, h0, 0);
        bulk_commit();
      }
    }
    if (lane == 0) bulk_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

cudaError_t conv1_tc_init() {
  return cudaFuncSetAttribute(conv1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C1_SMEM_BYTES);
}

// p: forward ConvParams of conv1_1 (H, W, N = 64, bias, out_tap / out_act with their TMA output maps, tiles_*)
cudaError_t launch_conv1_tc(const ConvParams& p, const float* x, const float* w, PixelConsts pc, int num_sms, cudaStream_t stream) {
  if (p.N != C1_N || !p.tma_out || p.num_tiles <= 0 || p.pool) return cudaErrorInvalidValue;
  const int grid = p.num_tiles < num_sms ? p.num_tiles : num_sms;
  static const bool pdl = getenv("NST_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(C1_THREADS);
  cfg.dynamicSmemBytes = C1_SMEM_BYTES;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, conv1_tc_kernel, p, x, w, pc);
}

}  // namespace nst
