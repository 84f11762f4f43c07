"""CPU restatement of the video back end of apply_video_process - TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's CPU leg).

Follows /root/reference/app.py:800-806 (RGB -> BGR of every processed frame) and app.py:820-840 (cross-dissolve: between
consecutive frames, n frames cv2.addWeighted(prev, 1 - alpha, frame, alpha, 0) with alpha = (i + 1) / (n + 1); `prev` is
the previous REAL frame, because it is re-read from final_frames[-1] before the interpolated ones are appended).

cv2.addWeighted is third-party (opencv-python, unpinned in requirements.txt; 4.13.0 installed here).  Its published 8-bit
algorithm (modules/core/src/arithm.simd.hpp, op_add_weighted): alpha, beta, gamma are converted to fp32, the result is
v_fma(a, alpha, v_fma(b, beta, gamma)) in fp32 - with gamma = 0 that is fma(a, alpha, fl(b * beta)) - rounded to nearest
even and saturated.  `add_weighted_u8` restates that in numpy (the fma through exact fp64 products); it is pinned to
cv2.addWeighted itself on all 65 536 byte pairs for every alpha the reference's slider can produce (tests/test_video_oracle.py)
and to golden frame lists written by running the reference's loop with cv2 (tests/golden/make_golden_video.py).
"""
import numpy as np


def add_weighted_u8(a: np.ndarray, alpha: float, b: np.ndarray, beta: float) -> np.ndarray:
    al = np.float64(np.float32(alpha))
    be = np.float64(np.float32(beta))
    inner = (b.astype(np.float64) * be).astype(np.float32)            # fl32(b * beta): the product is exact in fp64
    v = (a.astype(np.float64) * al + inner.astype(np.float64)).astype(np.float32)  # fma: a * alpha exact (8 + 24 bits), sum exact in fp64
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def assemble_frames(frames_rgb, number_of_interpolations=0):
    """frames_rgb: sequence of (H, W, 3) uint8 RGB arrays -> list of BGR arrays, app.py:800-840."""
    frames = [np.ascontiguousarray(np.asarray(f)[:, :, ::-1]) for f in frames_rgb]   # :803 COLOR_RGB2BGR
    n = number_of_interpolations
    if not n:                                                                        # :822, :841-842
        return frames
    final = [frames[0]]                                                              # :824
    for frame in frames[1:]:
        prev = final[-1]                                                             # :828
        for i in range(n):
            alpha = (i + 1) / (n + 1)                                                # :831-832
            final.append(add_weighted_u8(prev, 1 - alpha, frame, alpha))             # :835
        final.append(frame)                                                          # :838
    return final
