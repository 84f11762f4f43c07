"""Golden frame lists for the video back end (tests/golden/video_assemble.npz).

The reference's code for this step is inline in apply_video_process (app.py cannot be imported here: gradio and model
downloads at import time), so this script runs the statements of /root/reference/app.py:800-806 and 820-840 as they stand
there, with the same cv2 calls (cv2.cvtColor(..., cv2.COLOR_RGB2BGR), cv2.addWeighted(prev, 1 - alpha, frame, alpha, 0)),
on small seeded frames.  Run in the build container:  python tests/golden/make_golden_video.py
"""
import os

import cv2
import numpy as np


def reference_final_frames(pil_like_rgb_frames, interpolation_slider_input):
    frames = []
    for processed_frame in pil_like_rgb_frames:
        processed_frame = cv2.cvtColor(np.array(processed_frame), cv2.COLOR_RGB2BGR)
        frames.append(processed_frame)
    number_of_interpolations = interpolation_slider_input
    if number_of_interpolations:
        final_frames = [frames[0]]
        if len(frames) > 1:
            for frame in frames[1:]:
                prev_frame = final_frames[-1]
                for i in range(number_of_interpolations):
                    alpha = (i + 1) / (number_of_interpolations + 1)
                    interpolated = cv2.addWeighted(prev_frame, 1 - alpha, frame, alpha, 0)
                    final_frames.append(interpolated)
                final_frames.append(frame)
    else:
        final_frames = frames
    return final_frames


def main():
    rng = np.random.default_rng(2024)
    out = {}
    cases = [(5, 16, 24, 0), (4, 16, 24, 1), (3, 20, 16, 2), (4, 9, 7, 3), (3, 12, 20, 5), (1, 8, 8, 3), (2, 33, 31, 5), (3, 16, 16, 4)]
    for n, (F, H, W, k) in enumerate(cases):
        frames = rng.integers(0, 256, (F, H, W, 3), dtype=np.uint8)
        final = np.stack(reference_final_frames(list(frames), k), 0)
        out["frames_%d" % n], out["k_%d" % n], out["final_%d" % n] = frames, np.int64(k), final
    out["n"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "video_assemble.npz"), **out)


if __name__ == "__main__":
    main()
