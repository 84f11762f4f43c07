// Gram matrices on tcgen05 with split-K, fused with the style-loss reduction.
//
// Replaces, on the hot path of the reference,
//   gram_matrix   multi_style_transfer/style_transfer_losses.py:70-95   (bmm(X, X^T) / (b*c*h*w))
//   style_loss    multi_style_transfer/style_transfer_losses.py:98-146  (MSELoss(mean) per layer, / #layers)
// The feature map is NHWC fp16, i.e. a [HW x C] matrix with C contiguous, so F^T F contracts over rows:
// both MMA operands are "MN-major" views of the same 128B-swizzled tile (64 pixels x 64 channels per
// TMA box).  Split-K partial tiles go to a workspace; two small kernels reduce them in a fixed order
// (deterministic), form G - T, the per-layer MSE and the scaled fp16 operand of the backward GEMM.
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
static constexpr int G_TMEM_COLS = 128;

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

__global__ void __launch_bounds__(G_THREADS, 1) gram_partial_kernel(const __grid_constant__ GramParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G_STAGES * G_STAGE_BYTES + G_BOX_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + G_STAGES;
  uint64_t* done_bar = bars + 2 * G_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(done_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int item = blockIdx.x;
  const int l = gram_find_layer_by_item(p, item);
  const GramLayer& L = p.L[l];
  const int local = item - L.item0;
  const int pair = local / L.splits;
  const int split = local - pair * L.splits;
  int bi, bj;
  gram_pair_to_blocks(pair, L.nblk, bi, bj);
  const int c_begin = static_cast<int>((static_cast<long long>(split) * L.chunks) / L.splits);
  const int c_end = static_cast<int>((static_cast<long long>(split + 1) * L.chunks) / L.splits);
  const int a_boxes = L.C >= 128 ? 2 : 1;
  const int b_boxes = bi == bj ? 0 : L.bn / 64;
  const CUtensorMap* tm = &p.tm[l];

  if (warp == 0 && lane == 0) tma_prefetch_desc(tm);
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < G_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(done_bar, 1);
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
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  // A stage holds 32 KB: four 64-pixel x 64-channel boxes.  An off-diagonal block of a >= 128-channel layer needs all four
  // for one 64-pixel K chunk (A: 2, B: 2); a diagonal block needs two and conv1_1 (64 channels) one - those pack two / four
  // consecutive K chunks into a stage, so that every kind of item keeps the same number of bytes in flight (conv1_1's Gram
  // is a pure 33 MB read; with one 8 KB box per stage it was latency bound and set the duration of the whole launch).
  const int nb = a_boxes + b_boxes;        // boxes per 64-pixel chunk
  const int kmul = 4 / nb;                 // chunks per stage: 4, 2 or 1
  if (warp == 0) {
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int c = c_begin; c < c_end; c += kmul) {
        const int n = c_end - c < kmul ? c_end - c : kmul;
        mbar_wait(&empty_bar[stage], phase ^ 1u);
        mbar_arrive_expect_tx(&full_bar[stage], static_cast<uint32_t>(n * nb * G_BOX_BYTES));
        for (int j = 0; j < n; ++j) {
          uint8_t* sa = smem + stage * G_STAGE_BYTES + j * nb * G_BOX_BYTES;
          uint8_t* sb = sa + a_boxes * G_BOX_BYTES;
          for (int b = 0; b < a_boxes; ++b)
            tma_load_2d(sa + b * G_BOX_BYTES, tm, &full_bar[stage], bi * 128 + b * 64, (c + j) * G_KCHUNK);
          for (int b = 0; b < b_boxes; ++b)
            tma_load_2d(sb + b * G_BOX_BYTES, tm, &full_bar[stage], bj * 128 + b * 64, (c + j) * G_KCHUNK);
        }
        if (++stage == G_STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t idesc = umma_idesc_f16(128, L.bn, 0, 1, 1);
    for (int c = c_begin; c < c_end; c += kmul) {
      const int n = c_end - c < kmul ? c_end - c : kmul;
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        for (int j = 0; j < n; ++j) {
          const uint32_t a_addr = smem_u32(smem + stage * G_STAGE_BYTES + j * nb * G_BOX_BYTES);
          const uint32_t b_addr = bi == bj ? a_addr : a_addr + a_boxes * G_BOX_BYTES;
#pragma unroll
          for (int k = 0; k < G_KCHUNK / 16; ++k) {
            // MN-major: LBO = distance between the two 64-channel boxes, SBO = 8 pixel rows
            const uint64_t da = umma_desc_sw128(a_addr + k * 16 * 128, G_BOX_BYTES, 1024);
            const uint64_t db = umma_desc_sw128(b_addr + k * 16 * 128, G_BOX_BYTES, 1024);
            umma_f16(tmem_base, da, db, idesc, (c > c_begin || j > 0 || k > 0) ? 1u : 0u);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (c + kmul >= c_end) umma_commit(done_bar);
      }
      __syncwarp();
      if (++stage == G_STAGES) {
        stage = 0;
        phase ^= 1u;
      }
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    const int row = q * 32 + lane;
    mbar_wait(done_bar, 0);
    tc_fence_after();
    float* dst = p.ws + L.ws_off + (static_cast<size_t>(local) * 128 + row) * L.bn;
    const bool row_ok = bi * 128 + row < L.C;
#pragma unroll 1
    for (int c = 0; c < L.bn / 32; ++c) {
      uint32_t r[32];
      tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, r);
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
  }

  tc_fence_before();
  __syncthreads();
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
__device__ __forceinline__ FinElem gram_fin_elem(const GramParams& p, const GramLayer& L, int blk_in_layer, float* s_part) {
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
    const float* src = p.ws + L.ws_off + (static_cast<size_t>(pair) * L.splits * 128 + li) * L.bn + lj;
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

// pass 1: reduce split-K partials, G or G - T, per-block (sum of squares, max abs)
__global__ void __launch_bounds__(GRAM_FIN_THREADS) gram_finalize1_kernel(const __grid_constant__ GramParams p) {
  __shared__ float scratch[GRAM_FIN_THREADS / 32];
  __shared__ float s_part[GRAM_FIN_THREADS];
  const int l = gram_find_layer_by_finblk(p, blockIdx.x);
  const GramLayer& L = p.L[l];
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  FinElem e = gram_fin_elem(p, L, blockIdx.x - L.fin_blk0, s_part);
  float sq = 0.f, mx = 0.f;
  if (e.ok) {
    float d = e.g;
    if (L.target != nullptr) d -= L.target[static_cast<size_t>(e.gi) * L.C + e.gj];
    L.gram_out[static_cast<size_t>(e.gi) * L.C + e.gj] = d;
    if (e.offdiag) L.gram_out[static_cast<size_t>(e.gj) * L.C + e.gi] = d;
    sq = e.offdiag ? 2.f * d * d : d * d;
    mx = fabsf(d);
  }
  sq = block_sum(sq, scratch);
  mx = block_max(mx, scratch);
  if (threadIdx.x == 0) {
    p.fin_part[2 * blockIdx.x] = sq;
    p.fin_part[2 * blockIdx.x + 1] = mx;
  }
}

// pass 2: per-layer loss, scale, fp16 backward operand
__global__ void __launch_bounds__(GRAM_FIN_THREADS) gram_finalize2_kernel(const __grid_constant__ GramParams p) {
  __shared__ float scratch[GRAM_FIN_THREADS / 32];
  __shared__ float s_bcast[2];
  const int l = gram_find_layer_by_finblk(p, blockIdx.x);
  const GramLayer& L = p.L[l];
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (L.target == nullptr) return;
  float sq = 0.f, mx = 0.f;
  for (int b = threadIdx.x; b < L.fin_blocks; b += GRAM_FIN_THREADS) {
    sq += p.fin_part[2 * (L.fin_blk0 + b)];
    mx = fmaxf(mx, p.fin_part[2 * (L.fin_blk0 + b) + 1]);
  }
  sq = block_sum(sq, scratch);
  mx = block_max(mx, scratch);
  if (threadIdx.x == 0) {
    s_bcast[0] = sq;
    s_bcast[1] = mx;
  }
  __syncthreads();
  sq = s_bcast[0];
  mx = s_bcast[1];
  if (blockIdx.x == L.fin_blk0 && threadIdx.x == 0) {
    *L.loss = sq / (static_cast<float>(L.C) * static_cast<float>(L.C));
    *L.alpha = L.grad_coef * mx;
  }
  if (L.dh == nullptr) return;
  const float inv = mx > 0.f ? 1.f / mx : 0.f;
  const int tile = 128 * L.bn;
  const int FIN_ELEMS = GRAM_FIN_THREADS / L.fin_q;
  if (threadIdx.x >= FIN_ELEMS) return;
  const int id = (blockIdx.x - L.fin_blk0) * FIN_ELEMS + threadIdx.x;
  if (id >= L.pairs * tile) return;
  const int pair = id / tile;
  const int rem = id - pair * tile;
  const int li = rem / L.bn, lj = rem - li * L.bn;
  int bi, bj;
  gram_pair_to_blocks(pair, L.nblk, bi, bj);
  const int gi = bi * 128 + li, gj = bj * 128 + lj;
  if (gi >= L.C || gj >= L.C) return;
  const float d = L.gram_out[static_cast<size_t>(gi) * L.C + gj] * inv;
  L.dh[static_cast<size_t>(gi) * L.C + gj] = __float2half_rn(d);
  if (bi != bj) L.dh[static_cast<size_t>(gj) * L.C + gi] = __float2half_rn(d);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int make_tmap_feat(CUtensorMap* out, const void* base, int HW, int C) {
  void* ptr = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || ptr == nullptr)
    return -1;
  auto fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(HW)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(C) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(G_KCHUNK)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
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
    L.splits = static_cast<int>(s);
    L.item0 = item;
    item += L.pairs * L.splits;
    L.fin_blk0 = fin;
    L.fin_q = L.splits >= 16 ? 4 : 1;
    const int fin_elems = GRAM_FIN_THREADS / L.fin_q;
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

cudaError_t launch_gram(const GramParams& p, cudaStream_t stream) {
  cudaError_t e = launch_pdl(gram_partial_kernel, p.num_items, G_THREADS, G_SMEM_BYTES, stream, p);
  if (e != cudaSuccess) return e;
  e = launch_pdl(gram_finalize1_kernel, p.num_fin_blocks, GRAM_FIN_THREADS, 0, stream, p);
  if (e != cudaSuccess) return e;
  return launch_pdl(gram_finalize2_kernel, p.num_fin_blocks, GRAM_FIN_THREADS, 0, stream, p);
}

}  // namespace nst
