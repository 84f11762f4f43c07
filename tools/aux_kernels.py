#!/usr/bin/env python3
"""One call of each kernel around the loop (mask compositing, video frame list, plane split / merge) at the bench sizes.
Without arguments: warm-up + CUDA-event timing.  With --once: exactly one launch each (the form captured under ncu:
  ncu --set full --clock-control none -k regex:'mask_composite_fast|video_assemble_kernel|mip_split_rgb16|mip_merge_rgb16' \
      -o gpurun_out/aux python tools/aux_kernels.py --once)."""
import os
import sys
from importlib import import_module

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
once = "--once" in sys.argv
dev = torch.device("cuda:0")
seg = import_module("text-based-image-style-transfer_b200.text.segmentation_style_transfer")
video = import_module("text-based-image-style-transfer_b200.video")
U = import_module("text-based-image-style-transfer_b200.components.style_transfer_depth.util")
g = torch.Generator(device="cpu").manual_seed(0)

S = 8192
content = torch.randint(0, 256, (S, S, 3), dtype=torch.uint8, generator=g).to(dev)
style = torch.randint(0, 256, (S, S, 3), dtype=torch.uint8, generator=g).to(dev)
yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
mask = (((yy - S / 2) ** 2 + (xx - S / 3) ** 2) < (S / 2.5) ** 2).to(dev)
frames = torch.randint(0, 256, (64, 720, 1280, 3), dtype=torch.uint8, generator=g).to(dev)
H2, W2, n = 8192, 4096, 4
img = torch.randint(0, 256, (H2, W2, 3), dtype=torch.uint8, generator=g).to(dev)
y2, x2 = np.mgrid[0:H2, 0:W2]
depth = ((np.sin(y2 / 300.0) + np.cos(x2 / 450.0) + 2) * 63).astype(np.uint8)
bins = U.create_bins(n)
d_dev = torch.from_numpy(depth).to(dev)
lo_hi = (int(depth.min()), int(depth.max()))
planes = None


def run(name, fn, nbytes):
    reps = 1 if once else 10
    out = None
    if not once:
        for _ in range(3):
            out = fn()   # assigned: both result buffers the timed loop alternates between exist before it starts
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("%-16s %8.3f ms  %8.1f GB/s on %d bytes" % (name, ms, nbytes / ms / 1e6, nbytes))
    return out


run("mask_composite", lambda: seg.composite_tensors(content, style, mask, 5), 7 * S * S)
run("video_assemble", lambda: video.assemble_frames(frames, 3), (64 + 253) * 720 * 1280 * 3)
planes = run("mip_split", lambda: U.split_planes(img, d_dev, bins, depth_range=lo_hi), (4 + 3 * n) * H2 * W2)
run("mip_merge", lambda: U.merge_planes(planes, d_dev, bins, depth_range=lo_hi), 7 * H2 * W2)
