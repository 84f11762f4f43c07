// Interface of the Gram-matrix kernels (gram.cu): G = F^T F / (C*H*W) on tcgen05 with split-K,
// fused with the style-MSE reduction and the preparation of the backward operand.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace nst {

static constexpr int GRAM_MAX_LAYERS = 5;
static constexpr int GRAM_FIN_THREADS = 256;

struct GramLayer {
  int C;          // channels
  int HW;         // pixels
  float inv_norm; // 1 / (b * c * h * w) of gram_matrix (style_transfer_losses.py:84-93); c may be < C when C is padded
  int nblk;       // ceil(C / 128) channel blocks
  int pairs;      // nblk * (nblk + 1) / 2 upper-triangular block pairs
  int bn;         // N of the MMA: min(C, 128)
  int splits;     // split-K factor over pixels
  int chunks;     // ceil(HW / 64) 64-pixel K chunks
  int item0;      // first work item of this layer
  int fin_blk0;   // first finalize block of this layer
  int fin_blocks; // finalize blocks of this layer
  int fin_q;      // threads that share the split-K sum of one element in finalize1: 4 when there are many splits, else 1
  int fused;      // splits == 1: the Gram kernel finishes the layer itself (G - T, backward operand, sum of squares) straight
                  // from tensor memory - no workspace round trip, no finalize launch; its fin blocks are its block pairs
  size_t ws_off;  // float offset of this layer's partial tiles in the workspace
  // finalize inputs / outputs
  size_t target_stride; // batched launch: floats between the targets of consecutive images (0: one target for all)
  const float* target;  // [C,C] target Gram or nullptr (then `gram_out` receives G itself)
  float* gram_out;      // [C,C] G (if target == nullptr) or G - T
  __half* dh;           // [C,C] (G - T) * dh_scale as fp16 (backward operand), or nullptr
  // Device scalar: the power of two that maps max|T| to 64 (gram_target_scale).  It depends on the TARGET only, so the
  // operand can be written in the same pass that forms G - T (r01 normalised by max|G - T|, which needs a second pass over
  // the matrix and a launch of its own).  fp16 keeps 10 bits for every |G - T| between 1e-6 and 1e3 times max|T|.
  const float* dh_scale;
  float* alpha;         // device scalar: coefficient of the backward 1x1 convolution = grad_coef / dh_scale
  float grad_coef;      // 4 * w_style / (num_layers * C^3 * HW)
};

struct GramParams {
  CUtensorMap tm[GRAM_MAX_LAYERS];  // 2-D maps (C, HW) of the fp16 NHWC feature maps
  GramLayer L[GRAM_MAX_LAYERS];
  int num_layers;
  int num_items;
  // Several images per launch (batch > 1): the feature maps are [batch][HW][C] (3-D tensor maps (C, HW, image): the last 64-pixel
  // chunk of an image zero-fills behind ITS last pixel), the item list is repeated per image; per-image outputs: dh [batch][C][C], dh_scale / alpha [batch][scal_stride],
  // fin_part [batch][num_fin_blocks], ws [batch][ws_img].
  int batch;
  int scal_stride;
  size_t ws_img;
  int max_ctas;     // 0: one CTA per item; else the Gram kernel runs persistent on at most this many CTAs (see gram.cu)
  int num_fin_blocks;
  float* ws;        // split-K partial tiles [item][128][bn]
  float* fin_part;  // [num_fin_blocks] partial sums of (G - T)^2; summed per layer in a fixed order by the loss assembly
};

int make_tmap_feat(CUtensorMap* out, const void* base, int HW, int C, int batch = 1);
// fills nblk/pairs/bn/splits/chunks/item0/fin_* /ws_off and the totals; returns workspace floats needed
size_t gram_plan(GramParams& p, int target_ctas);
cudaError_t launch_gram(const GramParams& p, cudaStream_t stream);
// number of kernels launch_gram enqueues for this parameter set (1 when every layer is fused, else 2)
int gram_launches(const GramParams& p);
// dh_scale of a [C,C] target: 2^(6 - ceil(log2(max|T|))), 1 for an all-zero target (device scalar out)
cudaError_t launch_gram_target_scale(const float* target, int cc, float* scale_out, cudaStream_t stream);
// sets the kernel attributes (opt-in shared memory); call once per process before any capture
cudaError_t gram_init();

}  // namespace nst
