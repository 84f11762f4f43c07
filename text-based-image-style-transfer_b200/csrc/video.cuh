// Interface of the video back-end kernel (video.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace nst {

static constexpr int VIDEO_MAX_INTERP = 15;  // the reference's slider stops at 5 (app.py:953)

// frames: [F][H][W][3] uint8 RGB; out: [(F - 1) * (n_interp + 1) + 1][H][W][3] uint8 BGR: every input frame with its
// channels swapped, and between two consecutive frames n_interp cross-dissolved ones (cv2.addWeighted arithmetic)
cudaError_t launch_video_assemble(const uint8_t* frames, int F, size_t pixels, int n_interp, uint8_t* out, cudaStream_t s);

}  // namespace nst
