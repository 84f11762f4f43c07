#!/usr/bin/env python3
"""Golden vectors of the mask compositing step, written by the UNMODIFIED reference function
/root/reference/text/segmentation_style_transfer.py (needs cv2 and PIL, present in this image).
Run from the repo root:  python tests/golden/make_golden_mask.py   ->  tests/golden/mask_composite.npz"""
import importlib.util
import os

import numpy as np
from PIL import Image

spec = importlib.util.spec_from_file_location("ref_seg", "/root/reference/text/segmentation_style_transfer.py")
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

rng = np.random.default_rng(7)


def blob_mask(H, W, seed):
    r = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:H, 0:W]
    m = np.zeros((H, W), bool)
    for _ in range(4):
        cy, cx, rad = r.integers(0, H), r.integers(0, W), r.integers(min(H, W) // 8 + 1, min(H, W) // 3 + 2)
        m |= (yy - cy) ** 2 + (xx - cx) ** 2 < rad ** 2
    return m


cases = {}
# (content HxW, style HxW, edge_smoothing)
specs = [((48, 64), (48, 64), 5), ((37, 53), (37, 53), 0), ((40, 56), (44, 60), 7), ((45, 61), (41, 57), 4),
         ((64, 48), (64, 48), 21), ((9, 7), (9, 7), 13), ((50, 50), (50, 50), 1), ((33, 47), (33, 47), 63)]
for n, (cs, ss, k) in enumerate(specs):
    content = rng.integers(0, 256, cs + (3,), dtype=np.uint8)
    style = rng.integers(0, 256, ss + (3,), dtype=np.uint8)
    mask = blob_mask(cs[0], cs[1], 100 + n)
    out = np.asarray(ref.segmentation_style_transfer(Image.fromarray(content), Image.fromarray(style), mask, edge_smoothing=k))
    cases["content_%d" % n], cases["style_%d" % n], cases["mask_%d" % n] = content, style, mask
    cases["k_%d" % n], cases["out_%d" % n] = np.int64(k), out
cases["n"] = np.int64(len(specs))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "mask_composite.npz"), **cases)
print("wrote", len(specs), "cases")
