// Interface of the bandwidth-bound L-BFGS vector kernels (lbfgs.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "lbfgs_ctl.h"

namespace nst {

static constexpr int LB_THREADS = 256;
static constexpr int LB_VEC_PER_THREAD = 3;                        // float4 per thread per vector
static constexpr int LB_RING = 4;                                  // stored pairs in flight per block (shared-memory stages)
static constexpr int LB_MAX_VEC_PER_BLOCK = LB_THREADS * LB_VEC_PER_THREAD;
static constexpr int LB_PART_STRIDE = NST_LBFGS_SLOTS * NST_LBFGS_NDOT + NST_LBFGS_NSCAL;  // floats per block in pass-1 partials
static constexpr int LB_CTL_THREADS = 512;
// controller: work arrays + one mbarrier
static constexpr int LB_CTL_SMEM = (NST_CTL_WORK_DOUBLES + 2) * 8;

struct LbfgsBuffers {
  int n_pad;        // vector length, multiple of 4; elements >= n are zero everywhere
  int nblocks;      // pass-1 / pass-2 grid
  int vec_per_blk;  // float4 per block (<= LB_MAX_VEC_PER_BLOCK)
  float* x;         // [n_pad] image being optimised (flat [3,H,W])
  float* g;         // [n_pad] gradient of the latest evaluation
  float* g_prev;    // [n_pad]
  float* d;         // [n_pad] direction
  // stored pairs, chunk-major: [nblocks][SLOTS][2 (s, y)][vec_per_blk * 4].  A block streams ONE contiguous region
  // (~1 MB) instead of touching 2 x 101 vectors that are megabytes apart: the 0.64 GB history otherwise thrashes the
  // TLB once more than ~60 pairs are stored (measured: pass 1 took 381 us at m = 100 against 75 us at m = 57).
  float* hist;      // s = torch old_stps, y = torch old_dirs
  float* part;      // [LB_PART_STRIDE][nblocks] pass-1 per-block partial dots, output-major
  float* td_part;   // [nblocks] pass-2 per-block max|t d|
  double* dots;     // [SLOTS*NDOT] reduced
  double* scal;     // [NSCAL] reduced
  double* R;        // [SLOTS][SLOTS] s_i . y_j (upper triangle in age order)
  double* YY;       // [SLOTS][SLOTS] y_i . y_j
  NstLbfgsCtl* ctl;
  const float* eval_loss;  // device scalar written by the evaluation (total loss)
};

// picks nblocks / vec_per_blk for a vector of n_pad floats
void lbfgs_plan(LbfgsBuffers& b, int num_sms);
// floats in LbfgsBuffers::hist
size_t lbfgs_hist_floats(const LbfgsBuffers& b);
// sets the controller kernel's opt-in shared memory; call once per process
cudaError_t lbfgs_init();
// clears ctl.stop at step() entry
cudaError_t launch_lbfgs_step_begin(const LbfgsBuffers& b, cudaStream_t s);
// the four launches of one iteration, individually (timing) ...
cudaError_t launch_lbfgs_pass1(const LbfgsBuffers& b, cudaStream_t s);
cudaError_t launch_lbfgs_reduce(const LbfgsBuffers& b, cudaStream_t s);
cudaError_t launch_lbfgs_control(const LbfgsBuffers& b, int mode, cudaStream_t s);
cudaError_t launch_lbfgs_pass2(const LbfgsBuffers& b, cudaStream_t s);
// ... and together: pass 1 + reduction + controller + pass 2 = one L-BFGS iteration after an evaluation
cudaError_t launch_lbfgs_iteration(const LbfgsBuffers& b, int mode, cudaStream_t s);

}  // namespace nst
