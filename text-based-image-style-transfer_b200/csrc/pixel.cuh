// Interface of the bandwidth-bound pixel-space / elementwise kernels (pixel.cu).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace nst {

static constexpr int PIX_THREADS = 128;

struct PixelConsts {
  float mean[3];
  float stdv[3];
};

// number of thread blocks (= partial-sum slots) the pixel kernels use for an H x W image
int pixel_blocks(int H, int W);
int content_blocks(size_t numel);

// edge target: central differences of the grayscale of the NORMALISED content image
//   (run_style_transfer.py:73-74 -> helper_functions.py:104-113, style_transfer_losses.py:177-204)
cudaError_t launch_edge_target(const float* content /*[3,H,W]*/, float* tedge /*[2,H,W]*/, int H, int W,
                               PixelConsts pc, cudaStream_t s);

// TV loss on the normalised image + edge loss on the raw image, and their gradient w.r.t. the raw image
//   (run_style_transfer.py:126-136; style_transfer_losses.py:149-174, 177-225)
cudaError_t launch_pixel_losses(const float* x /*[3,H,W]*/, const float* tedge, float* grad_pix /*[3,H,W]*/,
                                double* tv_part, double* edge_part, int H, int W, PixelConsts pc, float w_tv,
                                float w_edge, cudaStream_t s);

// conv1_1 forward (3 -> 64 channels) on CUDA cores in fp32, fused with normalize():
//   writes the pre-ReLU tap and the post-ReLU activation as NHWC fp16
cudaError_t launch_conv1_fwd(const float* x /*[3,H,W]*/, const float* w /*[64,3,3,3]*/, const float* b /*[64]*/,
                             __half* out_tap, __half* out_act, int H, int W, PixelConsts pc, cudaStream_t s);

// conv1_1 data-gradient (64 -> 3 channels) + d normalize + TV/edge gradient -> flat gradient [3,H,W] fp32
cudaError_t launch_conv1_dgrad(const __nv_bfloat16* gy /*[H,W,64]*/, const float* w /*[64,3,3,3]*/,
                               const float* grad_pix /*[3,H,W] or nullptr*/, float* grad /*[3,H,W]*/, int H, int W,
                               PixelConsts pc, cudaStream_t s);

// content loss: per-block partial sums of (y - yc)^2 and the tap gradient gcoef * (y - yc) as bf16
//   (gcoef = 2 w_c / (numel * num_content_layers); accumulate != 0 adds to an existing tap gradient)
//   (style_transfer_losses.py:31-67)
cudaError_t launch_content_loss(const __half* y, const float* yc, __nv_bfloat16* addend, double* part, size_t numel,
                                float gcoef, int accumulate, cudaStream_t s);

// scalars: out[0] total, [1] w_c*content, [2] w_s*style, [3] w_tv*tv, [4] w_e*edge, [5..9] per-layer Gram MSE
struct LossAssembleArgs {
  const double* tv_part;
  int n_tv;
  const double* edge_part;
  int n_edge;
  const double* content_part;
  int n_content;
  // per style layer: per-block sums of (G - T)^2 written by the Gram kernels (gram.cu); MSE_l = sum / C_l^2
  const float* style_fin[5];
  int style_fin_n[5];
  float style_inv_cc[5];
  int num_style;
  float w_style, w_content, w_tv, w_edge;
  double tv_norm;       // 1 / (3 H W)
  double edge_norm;     // 1 / ((H-2)(W-2))
  double content_norm;  // 1 / num_content_layers (the per-block partial sums already carry 1 / numel of their layer)
  float* out;           // [16]: 0 total, 1..4 weighted c/s/tv/e, 5..9 per-layer Gram MSE, 10..13 unweighted c/s/tv/e
  // closure bookkeeping (run_style_transfer.py:143): *counter += 1 and one trace row per evaluation,
  // unless *stop_flag != 0 (the L-BFGS step already terminated; the evaluation is a no-op replay)
  int* counter;
  const int* stop_flag;
  float* trace;  // [trace_cap][5] or nullptr
  int trace_cap;
};
cudaError_t launch_loss_assemble(const LossAssembleArgs& a, cudaStream_t s);

// layout / precision conversion helpers (setup and tests)
cudaError_t launch_nhwc_half_to_nchw_float(const __half* in, float* out, int H, int W, int C, cudaStream_t s);
cudaError_t launch_nchw_float_to_nhwc_half(const float* in, __half* out, int H, int W, int C, cudaStream_t s);
// as above with an output row stride of Cp >= C halfs (extra channels are left untouched)
cudaError_t launch_nchw_float_to_nhwc_half_padded(const float* in, __half* out, int H, int W, int C, int Cp,
                                                  cudaStream_t s);
cudaError_t launch_nhwc_bf16_to_nchw_float(const __nv_bfloat16* in, float* out, int H, int W, int C, cudaStream_t s);
cudaError_t launch_nchw_float_to_nhwc_bf16(const float* in, __nv_bfloat16* out, int H, int W, int C, cudaStream_t s);

// weight repacking: torch [Cout,Cin,3,3] fp32 -> forward operand [9][Cout][Cin] fp16 and
// data-gradient operand [9][Cin][Cout] bf16 (taps flipped)
cudaError_t launch_pack_weights(const float* w, __half* w_fwd, __nv_bfloat16* w_bwd, int Cout, int Cin, cudaStream_t s);

// conv1_1 data-gradient operand: torch [64,3,3,3] fp32 -> [9][16][64] bf16 (taps flipped, image channels padded 3 -> 16)
cudaError_t launch_pack_weights_conv1_bwd(const float* w, __nv_bfloat16* w_bwd, cudaStream_t s);

// content target: yc = float(y) * gate[c] (gate == nullptr -> 1)   (run_style_transfer.py:13-25, 94-96)
cudaError_t launch_make_content_target(const __half* y, const float* gate, float* yc, size_t pixels, int C,
                                       cudaStream_t s);

// channel attention gate: sigmoid(relu(W2 relu(W1 avgpool(y))))   (ChannelAttention.py:23-40)
cudaError_t launch_channel_gate(const __half* y, const float* w1 /*[C/r,C]*/, const float* w2 /*[C,C/r]*/,
                                float* pooled /*[C] scratch*/, float* gate /*[C]*/, size_t pixels, int C, int Cr,
                                cudaStream_t s);

// StyleMixer: out = (1-wgt) * resize(a) + wgt * resize(b), bilinear align_corners=True   (StyleMixer.py:25-38)
cudaError_t launch_style_mix(const __half* a, int Ha, int Wa, const __half* b, int Hb, int Wb, __half* out, int Ho,
                             int Wo, int C, float wgt, cudaStream_t s);

}  // namespace nst
