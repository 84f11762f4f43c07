"""The video back-end oracle (oracle/video_oracle.py) against OpenCV itself and against the golden frame lists written by
the reference's statements (tests/golden/make_golden_video.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle import video_oracle as V

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "video_assemble.npz")


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 6, 9, 15])
def test_add_weighted_matches_cv2_on_every_byte_pair(n):
    cv2 = pytest.importorskip("cv2")
    a, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    for i in range(n):
        alpha = (i + 1) / (n + 1)
        assert np.array_equal(V.add_weighted_u8(a, 1 - alpha, b, alpha), cv2.addWeighted(a, 1 - alpha, b, alpha, 0)), (n, i)


def test_assemble_matches_reference_golden():
    g = np.load(GOLD)
    for n in range(int(g["n"])):
        final = V.assemble_frames(list(g["frames_%d" % n]), int(g["k_%d" % n]))
        assert np.array_equal(np.stack(final, 0), g["final_%d" % n]), n


def test_frame_count_and_real_frames_are_channel_swaps():
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 256, (4, 6, 5, 3), dtype=np.uint8)
    for k in (0, 1, 5):
        final = V.assemble_frames(list(frames), k)
        assert len(final) == 3 * (k + 1) + 1
        for f in range(4):
            assert np.array_equal(final[f * (k + 1)], frames[f][:, :, ::-1])
