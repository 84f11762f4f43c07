"""The mask-compositing oracle (oracle/mask_oracle.py) against cv2 itself and against golden vectors written by the
unmodified reference function (tests/golden/make_golden_mask.py).  Everything here is bit-exact: it is byte arithmetic."""
import numpy as np
import pytest

from conftest import golden
from oracle import mask_oracle as M


@pytest.mark.parametrize("k", list(range(1, 64, 2)))
def test_gaussian_blur_matches_cv2_bit_for_bit(k):
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(k)
    for trial, shape in enumerate([(41, 57), (64, 33), (5, 9)]):
        img = ((rng.random(shape) > 0.5).astype(np.uint8) * 255) if trial < 2 else rng.integers(0, 256, shape, dtype=np.uint8)
        if min(shape) * 2 - 2 < k // 2:
            continue
        assert np.array_equal(M.gaussian_blur_u8(img, k), cv2.GaussianBlur(img, (k, k), 0)), (k, shape)


def test_kernel_properties():
    for k in range(1, 100, 2):
        w = M.gaussian_kernel_fixed(k)
        assert sum(w) == 256 and w == w[::-1] and min(w) >= 0
    assert M.gaussian_kernel_fixed(5) == [16, 64, 96, 64, 16]


def test_against_reference_golden():
    g = golden("mask_composite")
    for n in range(int(g["n"])):
        out = M.segmentation_style_transfer(g["content_%d" % n], g["style_%d" % n], g["mask_%d" % n], int(g["k_%d" % n]))
        assert out.dtype == np.uint8 and np.array_equal(out, g["out_%d" % n]), n


def test_library_kernel_weights_match_the_oracle(built_libs):
    """The host-side weight computation of the library (csrc/mask.cu, needs no GPU) equals the oracle's for every size."""
    import ctypes as C
    import nst_b200
    lib = nst_b200._lib.load()
    for k in range(1, 128, 2):
        w = (C.c_int * k)()
        assert lib.nst_mask_gaussian_weights(k, w) == 0
        assert list(w) == M.gaussian_kernel_fixed(k), k
    assert lib.nst_mask_gaussian_weights(4, (C.c_int * 4)()) < 0
