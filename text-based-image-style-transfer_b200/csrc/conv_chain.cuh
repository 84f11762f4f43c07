// Interface of the chained tcgen05 convolution kernel (conv_chain.cu): every layer of a forward or backward chain of
// the VGG-19 trunk in one persistent launch, with tile-level dependencies between layers.
#pragma once
#include "conv_tc.cuh"

namespace nst {

static constexpr int CHAIN_MAX_BLOCK_N = 128;  // N tile of a chained layer: 128, or 64 for 64-channel outputs

// One entry of the work list = one convolution layer (ConvParams as for the per-layer kernel, block_n <= 128).
struct ChainLayer {
  ConvParams c;
  int mode;        // CONV_FWD, CONV_DGRAD or CONV_SCALE
  int item_base;   // index of this layer's first tile in the work list (tiles: spatial-major, channel tile fastest)
  int* done;       // [tiles_h * tiles_w] channel tiles finished per spatial tile (zeroed before the launch)
  // Up to two producers inside the same launch.  [0]: the layer that writes this layer's A operand (halo 1);
  // [1] (data gradient only): the Gram-backward layer that writes the tap seed this layer's epilogue adds (halo 0).
  // One producer tile covers dep_rpt x dep_cpt pixels of what this layer reads: 16 x 8 at the same resolution,
  // 8 x 4 behind a max-pool (forward), 32 x 16 behind its routing (backward).
  int dep_layer[2];  // index into the chain, -1 = none
  int dep_rpt[2], dep_cpt[2], dep_halo[2];
};

int chain_block_n(int N);
cudaError_t conv_chain_init();
// `done` counters of all layers must be zero when the kernel starts (the caller memsets them earlier in the stream)
// dbg (nullptr in production): 16 x gridDim.x SM-cycle counters of where every role of every CTA waited (tools/chain_waits.py)
cudaError_t launch_conv_chain(const ChainLayer* layers_dev, int n_layers, int total_items, int num_sms, cudaStream_t stream,
                              long long* dbg = nullptr);

}  // namespace nst
