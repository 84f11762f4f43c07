// Interface of the multi-plane (depth-binned) split / merge kernels (depth.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nst {

static constexpr int MIP_MAX_PLANES = 16;  // the reference's slider: 2..10 (app.py:971)

struct MipBins {
  int n;
  int dmin, range;             // uint8 depth: min and max - min of the map (the normalisation of util.py:27)
  double lo[MIP_MAX_PLANES];   // create_bins (util.py:38-50): inclusive on both sides (util.py:30)
  double hi[MIP_MAX_PLANES];
};

// image [pixels][C] uint8 -> out [n][pixels][C]: plane i keeps the pixels whose normalised depth lies in [lo_i, hi_i], others 0.
// depth: uint8 map (normalised in the kernel exactly like numpy: (d - min) / (max - min) in fp64) or, with depth_f64 != 0, the
// already normalised fp64 map.
cudaError_t launch_mip_split(const uint8_t* image, const void* depth, int depth_f64, size_t pixels, int C, const MipBins& bins,
                             uint8_t* out, cudaStream_t s);
// planes [n][pixels][3] uint8 -> out [pixels][3]: sum over i of (plane i masked with bin i), uint8 wrap-around like numpy's +=
cudaError_t launch_mip_merge(const uint8_t* planes, const void* depth, int depth_f64, size_t pixels, const MipBins& bins, uint8_t* out,
                             cudaStream_t s);

}  // namespace nst
