// Interface of the tcgen05 implicit-GEMM 3x3 / 1x1 convolution kernel (conv_tc.cu).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

// Instrumentation accessors: in the product build they are compile-time constants, so every stamp, wait counter and
// wrong-result experiment below folds away; -DNST_INSTRUMENT (tools/build.py --instrument) turns them into the fields of
// ConvParams above.
#ifdef NST_INSTRUMENT
#define NST_DBG_PTR(p) ((p).dbg)
#define NST_TL_PTR(p) ((p).tl)
#define NST_DBG_FLAG(p, bit) ((((p).dbg_flags) & (bit)) != 0)
#else
#define NST_DBG_PTR(p) (static_cast<long long*>(nullptr))
#define NST_TL_PTR(p) (static_cast<unsigned long long*>(nullptr))
#define NST_DBG_FLAG(p, bit) (false)
#endif

namespace nst {

// output-pixel tile of one CTA (M = 128): activation tensor maps use a (64, W+2, H+2) box for 3x3, (64, W, H) for 1x1
static constexpr int CONV_TILE_H = 16;
static constexpr int CONV_TILE_W = 8;

enum ConvMode : int {
  CONV_FWD = 0,    // fp16 operands; epilogue: +bias, optional pre-ReLU tap store, ReLU, optional 2x2 max-pool
  CONV_DGRAD = 1,  // bf16 operands; epilogue: ReLU mask or max-pool routing of the previous layer, + tap gradient
  CONV_SCALE = 2,  // fp16 operands; epilogue: out = alpha * acc as bf16 (Gram backward as a 1x1 convolution)
  CONV_DGRAD_PIX = 3,  // bf16 operands, N = 16 (3 used): conv1_1's data gradient; epilogue: / std + pixel-term gradient -> fp32 [3,H,W]
};

// One launch = one convolution layer as an implicit GEMM:
//   D[pixel, n] = sum_{tap, k} A[pixel + offset(tap), k] * B[tap][n][k]
// A is an NHWC activation tensor read through a 3-D TMA map (C, W, H) whose out-of-bounds
// zero fill implements the padding; B is [taps][N][K] (K contiguous) read through a 3-D TMA map.
struct ConvParams {
  CUtensorMap tmA;
  CUtensorMap tmB;
  // outputs through TMA stores (tma_out != 0): tmO0 = pre-ReLU tap (forward) / gradient (data gradient, Gram backward),
  // tmO1 = post-ReLU activation of a layer that is not pooled; boxes of 8 x 4 pixels x 32 channels (64-byte swizzle) or
  // x 16 channels (32-byte swizzle, data gradient), 16 x 8 pixels for the pool-routing scatter
  CUtensorMap tmO0;
  CUtensorMap tmO1;
  int tma_out;
  // CONV_DGRAD with the Gram backward of the produced layer folded in (seed_k > 0): after the 3x3 main loop the tile
  // also accumulates tap[pixel, :] x dh[n, :] (1x1, fp16, K = seed_k) into a second accumulator, and the epilogue adds
  // alpha * that to the masked gradient - instead of a separate 1x1 launch that writes the seed tensor and an epilogue
  // that reads it back.  tmA2 = the layer's pre-ReLU tap [H][W][C] (box 64 x 8 x 16), tmB2 = dh [1][C][C].
  CUtensorMap tmA2;
  CUtensorMap tmB2;
  int seed_k;
  uint32_t idesc2;
  int H, W;        // pixel grid of the GEMM M dimension
  int K, N;        // channels contracted per tap, output channels
  int taps;        // 9 (3x3, pad 1) or 1 (1x1)
  int block_n;     // N tile: 64, 128 or 256 (conv_block_n)
  int tiles_w, tiles_h, tiles_n, num_tiles;
  uint32_t idesc;
  // Several images per launch: every activation / gradient / routing tensor is [batch][H][W][C] (image stride = one image),
  // tmA / tmA2 are 4-D maps (C, W, H, batch) so that the zero fill at the image border still is the padding; outputs, masks
  // and seeds read / written by single threads are addressed through the stacked [batch * H][W][C] view (H and W multiples of 16:
  // 2 x 2 windows never straddle two images at any resolution); the output maps tmO0 / tmO1 are 4-D like tmA, so a TMA store
  // box that overhangs the bottom of its image (deep layers: fewer than 16 rows) is clipped there.  batch == 1: exactly the single-image launch.
  int batch;
  int alpha_stride;  // floats between the alpha scalars of consecutive images
  int b_per_image;   // != 0: tmB (CONV_SCALE) / tmB2 (folded Gram backward) hold one [N][K] matrix per image in their third dimension
  int pair;        // != 0: the launch runs as CTA pairs (cluster of 2, tcgen05 cta_group::2, M = 256): conv_use_pair; set BEFORE
                   // make_tmap_wgt (the weight box is half an N tile) and conv_finalize_params (instruction descriptor)
  int no_pdl_pair; // != 0: a pair launch is issued without the programmatic-dependent-launch attribute (plans that step beside other
                   // plans on the same GPU: nst_plan_set_shared_gpu)
  int dual_issue;  // != 0: one-slice layers with resident weights are issued by two threads (conv_tc.cu); NST_SINGLE_ISSUE clears it
  // ---- CONV_FWD
  const float* bias;   // [N]
  __half* out_tap;     // [H,W,N] pre-ReLU or nullptr
  __half* out_act;     // [H,W,N] post-ReLU (pool == 0) or [H/2,W/2,N] pooled (pool == 1); nullptr = skip
  uint8_t* out_route;  // [H/2,W/2,N] arg-max position 0..3 in the window, 4 = pooled value <= 0 (pool == 1)
  int pool;
  // ---- CONV_DGRAD
  const __half* mask_act;    // [H,W,N] post-ReLU activation whose gradient is produced (route == nullptr)
  const uint8_t* route;      // [H,W,N] routing bytes of the pooled layer whose gradient is produced
  const __nv_bfloat16* addend;  // [H,W,N] gradient injected at this tap, or nullptr
  __nv_bfloat16* out_grad;   // [H,W,N], or [Hup,Wup,N] when route != nullptr
  int Hup, Wup;
  // ---- CONV_SCALE
  const float* alpha;  // device scalar
  // ---- CONV_DGRAD_PIX
  const float* grad_pix;  // [3,H,W] gradient of the TV + edge terms, or nullptr
  float* out_pix;         // [3,H,W] fp32 gradient w.r.t. the raw image
  float inv_std[3];       // d normalize / d x
#ifdef NST_INSTRUMENT
  // ---- instrumented build only (libnst_b200_instr.so, tools/): never part of the product library
  // phase timestamps of CTA 0: see tools/conv_phases.py
  long long* dbg;
  // launch span inside a captured step: tl[0] = earliest CTA start, tl[1] = latest CTA end, both %globaltimer ns
  // (atomicMin / atomicMax by every CTA): tools/conv_timeline.py
  unsigned long long* tl;
  // timing experiments that produce WRONG results: bit 0 = skip the epilogue's global stores, bit 1 = skip its global
  // loads, bit 2 = never re-stream the weights, bit 3 = every filter tap reads the un-shifted patch view
  int dbg_flags;
#endif
};

// tensor-map builders (driver entry point resolved at run time; no link-time dependency on libcuda)
int make_tmap_act(CUtensorMap* out, const void* base, int H, int W, int C, int box_c, int box_w, int box_h, int batch = 1);
int make_tmap_wgt(CUtensorMap* out, const void* base, int taps, int N, int K, int box_n, bool pair = false);
// whether a launch of this mode / N tile / filter size runs as CTA pairs (NST_PAIR=0 turns them off)
bool conv_use_pair(int mode, int block_n, int taps, int K, int N);
// CTAs of the launch (persistent: at most one per SM; pairs: an even number)
int conv_grid_ctas(const ConvParams& p, int num_sms);
// output map for the epilogue's TMA stores: [H][W][C] 16-bit tensor, box (box_c, box_w, box_h); box_c * 2 = 128, 64 or 32 bytes
int make_tmap_out(CUtensorMap* out, void* base, int H, int W, int C, int box_c, int box_w, int box_h, int batch = 1);
// filter taps the kernel loads per weight stage for this N tile (the depth of the weight tensor map's box)
int conv_taps_per_stage(int block_n, int taps);

// picks the N tile for a layer of N output channels on an H x W pixel grid contracting k_total = taps * K values
int conv_block_n(int N, int H, int W, int k_total, int num_sms);
// fills tiles_* / idesc from H, W, K, N, taps and mode
void conv_finalize_params(ConvParams& p, int mode);
// sets the kernel attributes (opt-in shared memory) of every instantiation; call once per process
cudaError_t conv_tc_init();
// enqueue
cudaError_t launch_conv_tc(const ConvParams& p, int mode, int num_sms, cudaStream_t stream);

// conv1_1 forward on tensor cores (conv1_tc.cu): p = forward ConvParams of conv1_1 (N = 64, TMA output maps), x = [3][H][W]
// fp32 image, w = [64][27] fp32 weights.  PixelConsts: pixel.cuh.
struct PixelConsts;
cudaError_t conv1_tc_init();
cudaError_t launch_conv1_tc(const ConvParams& p, const float* x, const float* w, PixelConsts pc, int num_sms, cudaStream_t stream);

}  // namespace nst
