// Interface of the mask-compositing kernel (mask.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nst {

static constexpr int MASK_MAX_K = 127;  // largest Gaussian kernel size (the staged tile then is 158 x 158 bytes)

// fills w[0..k-1] with the 8-bit fixed-point Gaussian kernel of cv2.GaussianBlur(uint8, (k, k), 0); k odd, 1 <= k <= MASK_MAX_K
int mask_gaussian_weights(int k, int* w);
// content, style, out: [H][W][C] uint8; mask: [H][W] bytes (non-zero = take the style image); k = 0: hard selection,
// else odd Gaussian kernel size
cudaError_t launch_mask_composite(const uint8_t* content, const uint8_t* style, const uint8_t* mask, uint8_t* out, int H, int W, int C,
                                  int k, cudaStream_t s);

}  // namespace nst
