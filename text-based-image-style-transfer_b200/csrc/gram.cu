// Gram matrices on tcgen05 with split-K, fused with the style-loss reduction.
//
// Replaces, on the hot path of the reference,
//   gram_matrix   multi_style_transfer/style_transfer_losses.py:70-95   (bmm(X, X^T) / (b*c*h*w))
//   style_loss    multi_style_transfer/style_transfer_losses.py:98-146  (MSELoss(mean) per layer, / #layers)
// The feature map is NHWC fp16, i.e. a [HW x C] matrix with C contiguous, so F^T F contracts over rows:
// both MMA operands are "MN-major" views of the same 128B-swizzled tile (64 pixels x 64 channels per
// TMA box).  Layers with few K chunks are finished by the Gram kernel itself (one CTA per block pair: G - T, the fp16
// operand of the backward GEMM and the sum of squares straight from tensor memory); for the others split-K partial tiles
// go to a workspace and one small kernel reduces them in a fixed order (deterministic).  The per-layer MSE is summed from
// the per-block sums by the loss assembly (pixel.cu).
#include "gram.cuh"
#include <stdlib.h>
#include "common.cuh"

#include <cudaTypedefs.h>

namespace nst {

static constexpr int G_THREADS = 256;
static constexpr int G_KCHUNK = 64;                       // pixels per pipeline stage
static constexpr int G_BOX_BYTES = G_KCHUNK * 128;        // 64 pixels x 64 channels x 2 B
static constexpr int G_STAGE_BYTES = 4 * G_BOX_BYTES;     // A: 2 boxes, B: up to 2 boxes
static constexpr int G_STAGES = 5;
// + one box of padding: with 64 channels the M = 128 operand's second 64-channel box does not exist and the tensor core
// reads whatever follows the first one (those accumulator rows are never stored); behind the last packed chunk of the last
// stage that must still be inside the allocation
static constexpr int G_SMEM_BYTES = G_STAGES * G_STAGE_BYTES + G_BOX_BYTES + 1024 + 256;
static constexpr int G_TMEM_COLS = 256;                       // two accumulator stages of 128 columns
static constexpr int GRAM_FUSED_MAX_CHUNKS = 32;            // up to 2048 pixels: one CTA per block pair, finished in place

__device__ __forceinline__ int gram_find_layer_by_item(const GramParams& p, int item) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < GRAM_MAX_LAYERS; ++i)
    if (i < p.num_layers && item >= p.L[i].item0) l = i;
  return l;
}
__device__ __forceinline__ int gram_find_layer_by_finblk(const GramParams& p, int blk) {
  int l = 0;
#pragma unroll
  for (int i = 1; i < GRAM_MAX_LAYERS; ++i)
    if (i < p.num_layers && blk >= p.L[i].fin_blk0) l = i;
  return l;
}
__device__ __forceinline__ void gram_pair_to_blocks(int pair, int nblk, int& bi, int& bj) {
  bi = 0;
  int rowlen = nblk;
  while (pair >= rowlen) {
    pair -= rowlen;
    ++bi;
    --rowlen;
  }
  bj = bi + pair;
}

// What one work item (block pair x K split of one layer) is
struct GramItem {
  int l, local, pair, bi, bj, c_begin, c_end, a_boxes, b_boxes, nb, kmul;
  int img;       // image of a batched launch (GramParams::batch > 1): items [img * num_items, (img + 1) * num_items)
};
__device__ __forceinline__ GramItem gram_item(const GramParams& p, int item) {
  GramItem it;
  it.img = 0;
  if (p.batch > 1) {
    it.img = item / p.num_items;
    item -= it.img * p.num_items;
  }
  it.l = gram_find_layer_by_item(p, item);

  const GramLayer& L = p.L[it.l];
  it.local = item - L.item0;
  it.pair = it.local / L.splits;
  const int split = it.local - it.pair * L.splits;
  gram_pair_to_blocks(it.pair, L.nblk, it.bi, it.bj);
  it.c_begin = static_cast<int>((static_cast<long long>(split) * L.chunks) / L.splits);
  it.c_end = static_cast<int>((static_cast<long long>(split + 1) * L.chunks) / L.splits);
  it.a_boxes = L.C >= 128 ? 2 : 1;
  it.b_boxes = it.bi == it.bj ? 0 : L.bn / 64;
  // A stage holds 32 KB: four 64-pixel x 64-channel boxes.  An off-diagonal block of a >= 128-channel layer needs all four
  // for one 64-pixel K chunk (A: 2, B: 2); a diagonal block needs two and conv1_1 (64 channels) one - those pack two / four
  // consecutive K chunks into a stage, so that every kind of item keeps the same number of bytes in flight (conv1_1's Gram
  // is a pure 33 MB read; with one 8 KB box per stage it was latency bound and set the duration of the whole launch).
  it.nb = it.a_boxes + it.b_boxes;   // boxes per 64-pixel chunk
  it.kmul = 4 / it.nb;               // chunks per stage: 4, 2 or 1
  return it;
}

// Persistent: CTA b takes items b, b + gridDim.x, ...  With gridDim.x = num_items (a launch that has the GPU to itself) every
// CTA runs one item; a launch that shares the GPU with the convolution chain is given only as many CTAs as that chain leaves
// idle (GramParams::max_ctas) and walks the item list, so that it never takes an SM a convolution tile is waiting for.
// Two accumulator stages in tensor memory: the epilogue of item i overlaps the main loop of item i + 1.
__global__ void __launch_bounds__(G_THREADS, 1) gram_partial_kernel(const __grid_constant__ GramParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G_STAGES * G_STAGE_BYTES + G_BOX_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + G_STAGES;
  uint64_t* tfull_bar = bars + 2 * G_STAGES;   // [2] accumulator stage complete
  uint64_t* tempty_bar = tfull_bar + 2;        // [2] accumulator stage read out by the four epilogue warps
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);
  float* s_sq = reinterpret_cast<float*>(tmem_slot + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int l = 0; l < p.num_layers; ++l) tma_prefetch_desc(&p.tm[l]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < G_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4);
    }
    mbar_fence_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, G_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // programmatic dependent launch: everything above touched only on-chip state; the taps come from the previous launch
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // A launch that has the GPU to itself lets its dependent start scheduling right away.  A launch that runs beside the
  // convolution chain (max_ctas > 0, side stream) must not: its dependent is the finalize kernel, whose thousands of small
  // blocks would sit in griddepcontrol.wait on every SM for as long as this kernel runs - and a convolution CTA needs a whole
  // SM's registers, so the chain's next layer could not be scheduled (measured: conv4_4 100 us instead of 18).  There the
  // dependent is released at the end.
  if (p.max_ctas == 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;
  const int total_items = p.num_items * (p.batch > 1 ? p.batch : 1);

  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
        const GramItem it = gram_item(p, item);
        const CUtensorMap* tm = &p.tm[it.l];
        for (int c = it.c_begin; c < it.c_end; c += it.kmul) {
          const int n = it.c_end - c < it.kmul ? it.c_end - c : it.kmul;
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(n * it.nb * G_BOX_BYTES));
          for (int j = 0; j < n; ++j) {
            uint8_t* sa = smem + stage * G_STAGE_BYTES + j * it.nb * G_BOX_BYTES;
            uint8_t* sb = sa + it.a_boxes * G_BOX_BYTES;
            for (int b = 0; b < it.a_boxes; ++b)
              tma_load_3d(sa + b * G_BOX_BYTES, tm, &full_bar[stage], it.bi * 128 + b * 64, (c + j) * G_KCHUNK, it.img);
            for (int b = 0; b < it.b_boxes; ++b)
              tma_load_3d(sb + b * G_BOX_BYTES, tm, &full_bar[stage], it.bj * 128 + b * 64, (c + j) * G_KCHUNK, it.img);
          }
          if (++stage == G_STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    int stage = 0, ts = 0;
    uint32_t phase = 0, tphase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const GramItem it = gram_item(p, item);
      const GramLayer& L = p.L[it.l];
      const uint32_t idesc = umma_idesc_f16(128, L.bn, 0, 1, 1);
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(ts * 128);
      mbar_wait(&tempty_bar[ts], tphase ^ 1u);
      tc_fence_after();
      for (int c = it.c_begin; c < it.c_end; c += it.kmul) {
        const int n = it.c_end - c < it.kmul ? it.c_end - c : it.kmul;
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          for (int j = 0; j < n; ++j) {
            const uint32_t a_addr = smem_u32(smem + stage * G_STAGE_BYTES + j * it.nb * G_BOX_BYTES);
            const uint32_t b_addr = it.bi == it.bj ? a_addr : a_addr + it.a_boxes * G_BOX_BYTES;
#pragma unroll
            for (int k = 0; k < G_KCHUNK / 16; ++k) {
              // MN-major: LBO = distance between the two 64-channel boxes, SBO = 8 pixel rows
              const uint64_t da = umma_desc_sw128(a_addr + k * 16 * 128, G_BOX_BYTES, 1024);
              const uint64_t db = umma_desc_sw128(b_addr + k * 16 * 128, G_BOX_BYTES, 1024);
              umma_f16(d_tmem, da, db, idesc, (c > it.c_begin || j > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
          if (c + it.kmul >= it.c_end) umma_commit(&tfull_bar[ts]);
        }
        __syncwarp();
        if (++stage == G_STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      if (++ts == 2) {
        ts = 0;
        tphase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    int ts = 0;
    uint32_t tphase = 0;
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const GramItem it = gram_item(p, item);
      const GramLayer& L = p.L[it.l];
      const int bi = it.bi, bj = it.bj;
      // per-image outputs of a batched launch: [batch][C][C] operands / targets, [batch][GRAM_MAX_LAYERS] scalars
      const size_t cc_off = static_cast<size_t>(it.img) * L.C * L.C;
      const int sc_off = it.img * p.scal_stride;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(ts * 128);
      const bool row_ok = bi * 128 + row < L.C;
      if (L.fused && L.target != nullptr) {
        // ---- the layer is finished here: d = G - T, fp16 backward operand d * dh_scale (both triangles), sum of d^2.
        // Everything that does not depend on the accumulator is fetched before waiting for it.
        const int gi = bi * 128 + row;
        const float scale = L.dh_scale != nullptr ? __ldg(L.dh_scale + sc_off) : 1.f;
        const float* trow = L.target + static_cast<size_t>(it.img) * L.target_stride + static_cast<size_t>(row_ok ? gi : 0) * L.C + bj * 128;
        float sq = 0.f;
        const float w = bi == bj ? 1.f : 2.f;   // an off-diagonal block stands for its mirror image too
        // target row, one 32-column chunk ahead of its use: the first chunk is requested before the accumulator is waited for
        float4 t4[8], t4n[8];
        if (row_ok) {
#pragma unroll
          for (int j = 0; j < 8; ++j) t4n[j] = __ldg(reinterpret_cast<const float4*>(trow) + j);
        }
        mbar_wait(&tfull_bar[ts], tphase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < L.bn / 32; ++c) {
#pragma unroll
          for (int j = 0; j < 8; ++j) t4[j] = t4n[j];
          if (row_ok && c + 1 < L.bn / 32) {
#pragma unroll
            for (int j = 0; j < 8; ++j) t4n[j] = __ldg(reinterpret_cast<const float4*>(trow + (c + 1) * 32) + j);
          }
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          if (row_ok) {
            const float* tt = reinterpret_cast<const float*>(t4);
            __align__(16) __half hv[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int gj = bj * 128 + c * 32 + j;
              float d = __uint_as_float(r[j]) * L.inv_norm - tt[j];
              if (gj >= L.C) d = 0.f;
              sq = fmaf(w * d, d, sq);
              hv[j] = __float2half_rn(d * scale);
            }
            if (L.dh != nullptr) {
              __half* drow = L.dh + cc_off + static_cast<size_t>(gi) * L.C + bj * 128 + c * 32;
#pragma unroll
              for (int j = 0; j < 4; ++j) *(reinterpret_cast<uint4*>(drow) + j) = *(reinterpret_cast<const uint4*>(hv) + j);
              if (bi != bj) {
                // mirror image: consecutive lanes are consecutive rows gi, so every store instruction of the warp is contiguous
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const int gj = bj * 128 + c * 32 + j;
                  if (gj < L.C) L.dh[cc_off + static_cast<size_t>(gj) * L.C + gi] = hv[j];
                }
              }
            }
          }
        }
        tc_fence_before();
        sq = warp_sum(sq);
        if (lane == 0) {
          mbar_arrive(&tempty_bar[ts]);
          s_sq[q] = sq;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // the four epilogue warps
        if (warp == 4 && lane == 0) {
          p.fin_part[it.img * p.num_fin_blocks + L.fin_blk0 + it.pair] = ((s_sq[0] + s_sq[1]) + s_sq[2]) + s_sq[3];
          if (it.pair == 0 && L.alpha != nullptr) L.alpha[sc_off] = L.grad_coef / scale;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");   // s_sq is rewritten by the next item
      } else {
        mbar_wait(&tfull_bar[ts], tphase);
        tc_fence_after();
        float* dst = p.ws + static_cast<size_t>(it.img) * p.ws_img + L.ws_off + (static_cast<size_t>(it.local) * 128 + row) * L.bn;
#pragma unroll 1
        for (int c = 0; c < L.bn / 32; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          if (row_ok) {
            // 32-byte stores: the lanes of a warp are whole rows apart, so a 16-byte store would write half sectors
#pragma unroll
            for (int j = 0; j < 4; ++j)
              asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst + c * 32 + 8 * j), "r"(r[8 * j]), "r"(r[8 * j + 1]),
                           "r"(r[8 * j + 2]), "r"(r[8 * j + 3]), "r"(r[8 * j + 4]), "r"(r[8 * j + 5]), "r"(r[8 * j + 6]), "r"(r[8 * j + 7])
                           : "memory");
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[ts]);
      }
      if (++ts == 2) {
        ts = 0;
        tphase ^= 1u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (p.max_ctas != 0) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, G_TMEM_COLS);
  }
}

// element of the upper-triangular block list handled by a thread: 64 consecutive elements per block, the split-K
// partials of each element are summed by 4 threads (splits s = q mod 4) and combined in a fixed order
struct FinElem {
  int gi, gj;
  bool ok, offdiag;
  float g;
};
__device__ __forceinline__ FinElem gram_fin_elem(const GramParams& p, const GramLayer& L, int blk_in_layer, float* s_part, int img) {
  FinElem e;
  const int tile = 128 * L.bn;
  const int QL = L.fin_q;                       // 1 or 4
  const int FIN_ELEMS = GRAM_FIN_THREADS / QL;  // elements per block
  const int el = threadIdx.x % FIN_ELEMS, q = threadIdx.x / FIN_ELEMS;
  const int id = blk_in_layer * FIN_ELEMS + el;
  e.ok = id < L.pairs * tile;
  e.gi = e.gj = 0;
  e.offdiag = false;
  e.g = 0.f;
  int pair = 0, li = 0, lj = 0;
  if (e.ok) {
    pair = id / tile;
    const int rem = id - pair * tile;
    li = rem / L.bn;
    lj = rem - li * L.bn;
    int bi, bj;
    gram_pair_to_blocks(pair, L.nblk, bi, bj);
    e.gi = bi * 128 + li;
    e.gj = bj * 128 + lj;
    e.offdiag = bi != bj;
    if (e.gi >= L.C || e.gj >= L.C) e.ok = false;
  }
  float a0 = 0.f, a1 = 0.f;
  if (e.ok) {
    const float* src = p.ws + static_cast<size_t>(img) * p.ws_img + L.ws_off + (static_cast<size_t>(pair) * L.splits * 128 + li) * L.bn + lj;
    int s = q;
    for (; s + QL < L.splits; s += 2 * QL) {
      a0 += src[static_cast<size_t>(s) * tile];
      a1 += src[static_cast<size_t>(s + QL) * tile];
    }
    if (s < L.splits) a0 += src[static_cast<size_t>(s) * tile];
  }
  if (QL == 1) {
    e.g = (a0 + a1) * L.inv_norm;
    return e;
  }
  s_part[q * FIN_ELEMS + el] = a0 + a1;
  __syncthreads();
  if (q == 0 && e.ok) {
    const float acc = ((s_part[el] + s_part[FIN_ELEMS + el]) + s_part[2 * FIN_ELEMS + el]) + s_part[3 * FIN_ELEMS + el];
    e.g = acc * L.inv_norm;
  }
  if (q != 0) e.ok = false;
  __syncthreads();
  return e;
}

// Split-K layers: reduce the partial tiles in a fixed order, G (no target) or d = G - T, the fp16 backward operand
// d * dh_scale (both triangles) and the per-block sum of d^2.  Fused layers (GramLayer::fused) have nothing left to do.
__global__ void __launch_bounds__(GRAM_FIN_THREADS) gram_finalize_kernel(const __grid_constant__ GramParams p) {
  __shared__ float scratch[GRAM_FIN_THREADS / 32];
  __shared__ float s_part[GRAM_FIN_THREADS];
  int img = 0, blk = blockIdx.x;
  if (p.batch > 1) {
    img = blk / p.num_fin_blocks;
    blk -= img * p.num_fin_blocks;
  }
  const int l = gram_find_layer_by_finblk(p, blk);
  const GramLayer& L = p.L[l];
  const size_t cc_off = static_cast<size_t>(img) * L.C * L.C;
  const int sc_off = img * p.scal_stride;
  const float scale = (L.dh_scale != nullptr && !L.fused) ? __ldg(L.dh_scale + sc_off) : 1.f;  // set with the target: not written by the previous launch
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (L.fused && L.target != nullptr) return;
  FinElem e = gram_fin_elem(p, L, blk - L.fin_blk0, s_part, img);
  float sq = 0.f;
  if (e.ok) {
    float d = e.g;
    if (L.target != nullptr) d -= L.target[static_cast<size_t>(img) * L.target_stride + static_cast<size_t>(e.gi) * L.C + e.gj];
    if (L.gram_out != nullptr) {
      L.gram_out[static_cast<size_t>(e.gi) * L.C + e.gj] = d;
      if (e.offdiag) L.gram_out[static_cast<size_t>(e.gj) * L.C + e.gi] = d;
    }
    if (L.dh != nullptr) {
      const __half h = __float2half_rn(d * scale);
      L.dh[cc_off + static_cast<size_t>(e.gi) * L.C + e.gj] = h;
      if (e.offdiag) L.dh[cc_off + static_cast<size_t>(e.gj) * L.C + e.gi] = h;
    }
    sq = e.offdiag ? 2.f * d * d : d * d;
  }
  sq = block_sum(sq, scratch);
  if (threadIdx.x == 0) {
    p.fin_part[blockIdx.x] = sq;
    if (blk == L.fin_blk0 && L.alpha != nullptr) L.alpha[sc_off] = L.grad_coef / scale;
  }
}

// dh_scale of a target: largest power of two s with max|T| * s <= 64 (1 for an all-zero target)
__global__ void __launch_bounds__(256) gram_target_scale_kernel(const float* __restrict__ t, int cc, float* __restrict__ out) {
  __shared__ float scratch[8];
  float m = 0.f;
  for (int i = threadIdx.x; i < cc; i += 256) m = fmaxf(m, fabsf(t[i]));
  m = block_max(m, scratch);
  if (threadIdx.x == 0) {
    float s = 1.f;
    if (m > 0.f && isfinite(m)) {
      int e;
      frexpf(m, &e);            // m = f * 2^e, f in [0.5, 1)  ->  m * 2^(6 - e) in [32, 64)
      s = ldexpf(1.f, 6 - e);
    }
    *out = s;
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int make_tmap_feat(CUtensorMap* out, const void* base, int HW, int C, int batch) {
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || ptr == nullptr)
    return -1;
  auto fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  // (C, HW, image): rows behind the last pixel of an image zero-fill, they are not the next image's first pixels
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(HW), static_cast<cuuint64_t>(batch < 1 ? 1 : batch)};
  cuuint64_t strides[2] = {static_cast<cuuint64_t>(C) * 2, static_cast<cuuint64_t>(HW) * C * 2};
  cuuint32_t box[3] = {64, static_cast<cuuint32_t>(G_KCHUNK), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 1000;
}

size_t gram_plan(GramParams& p, int target_ctas) {
  // Work is balanced by operand bytes, not by (block pair, K chunk) units: a diagonal block of a >= 128-channel layer
  // loads two 8 KB boxes per chunk, an off-diagonal one four, conv1_1 (64 channels) a single one - with equal unit counts
  // the conv4_1 items moved three times the bytes of the conv1_1 items and set the duration of the whole launch.
  long long total_boxes = 0;
  long long layer_boxes[GRAM_MAX_LAYERS];
  for (int l = 0; l < p.num_layers; ++l) {
    GramLayer& L = p.L[l];
    L.nblk = (L.C + 127) / 128;
    L.pairs = L.nblk * (L.nblk + 1) / 2;
    L.bn = L.C < 128 ? L.C : 128;
    L.chunks = (L.HW + G_KCHUNK - 1) / G_KCHUNK;
    const int a_boxes = L.C >= 128 ? 2 : 1;
    const int off_pairs = L.pairs - L.nblk;
    // boxes per K chunk over all pairs of the layer: diagonal pairs load A only, off-diagonal pairs A and B
    layer_boxes[l] = static_cast<long long>(L.chunks) * (static_cast<long long>(L.nblk) * a_boxes + static_cast<long long>(off_pairs) * (a_boxes + L.bn / 64));
    total_boxes += layer_boxes[l];
  }
  const long long per_cta = (total_boxes + target_ctas - 1) / target_ctas;  // boxes one CTA should move
  int item = 0, fin = 0;
  size_t ws = 0;
  for (int l = 0; l < p.num_layers; ++l) {
    GramLayer& L = p.L[l];
    // items of this layer = pairs x splits, each moving layer_boxes / (pairs x splits) boxes on average
    long long s = (layer_boxes[l] + per_cta * L.pairs - 1) / ((per_cta > 0 ? per_cta : 1) * L.pairs);
    if (s < 1) s = 1;
    if (s > L.chunks) s = L.chunks;
    // Few K chunks (deep layers at moderate resolution: conv5_1 at 512^2 has 16): splitting them ~15 ways made every CTA run
    // four MMAs and then write a 64 KB partial tile for a finalize launch to gather again.  One CTA per block pair instead
    // keeps the accumulator in tensor memory to the end and finishes the layer in the same launch (GramLayer::fused).
    if (L.chunks <= GRAM_FUSED_MAX_CHUNKS) s = 1;
    L.splits = static_cast<int>(s);
    L.fused = L.splits == 1 ? 1 : 0;
    L.item0 = item;
    item += L.pairs * L.splits;
    L.fin_blk0 = fin;
    L.fin_q = L.splits >= 16 ? 4 : 1;
    const int fin_elems = GRAM_FIN_THREADS / L.fin_q;
    // a fused layer with a target needs one sum-of-squares slot per block pair; without a target (nst_plan_tap_gram: G itself
    // is wanted) it goes through the workspace and the finalize kernel like a split layer with one split
    L.fin_blocks = (L.pairs * 128 * L.bn + fin_elems - 1) / fin_elems;
    fin += L.fin_blocks;
    L.ws_off = ws;
    ws += static_cast<size_t>(L.pairs) * L.splits * 128 * L.bn;
  }
  p.num_items = item;
  p.num_fin_blocks = fin;
  return ws;
}

cudaError_t gram_init() {
  return cudaFuncSetAttribute(gram_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM_BYTES);
}

// The three launches are chained with programmatic dependent launch: each kernel's blocks are scheduled while the previous
// kernel drains and block in griddepcontrol.wait until it has completed, which takes the launch latency of two tiny
// kernels off the critical path between the deepest forward convolution and the first data gradient.
template <typename... Args>
static cudaError_t launch_pdl(void (*kernel)(Args...), int grid, int block, size_t smem, cudaStream_t stream, const GramParams& p) {
  static const bool pdl = getenv("NST_NO_PDL") == nullptr;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, p);
}

static bool gram_needs_finalize(const GramParams& p) {
  for (int l = 0; l < p.num_layers; ++l)
    if (!(p.L[l].fused && p.L[l].target != nullptr)) return true;
  return false;
}
int gram_launches(const GramParams& p) { return gram_needs_finalize(p) ? 2 : 1; }

cudaError_t launch_gram(const GramParams& p, cudaStream_t stream) {
  const int nb = p.batch > 1 ? p.batch : 1;
  const int items = p.num_items * nb;
  const int grid = (p.max_ctas > 0 && p.max_ctas < items) ? p.max_ctas : items;
  cudaError_t e = launch_pdl(gram_partial_kernel, grid, G_THREADS, G_SMEM_BYTES, stream, p);
  if (e != cudaSuccess || !gram_needs_finalize(p)) return e;
  return launch_pdl(gram_finalize_kernel, p.num_fin_blocks * nb, GRAM_FIN_THREADS, 0, stream, p);
}

cudaError_t launch_gram_target_scale(const float* target, int cc, float* scale_out, cudaStream_t stream) {
  gram_target_scale_kernel<<<1, 256, 0, stream>>>(target, cc, scale_out);
  return cudaGetLastError();
}

}  // namespace nst
