#!/usr/bin/env python3
"""Golden vectors of the depth-aware / multi-plane variant (tests/golden/depth_mip.npz), written by the UNMODIFIED reference
(/root/reference/components/style_transfer_depth) on the CPU.  Build container only.

Shim, no source edits: the one of make_golden.py (matplotlib stub, random-init vgg19 under seed 1234) plus
`transformers.pipeline` replaced by a stub that returns a fixed synthetic depth map as a PIL 'L' image (the
Depth-Anything weights cannot be downloaded).  StyleA3.num_steps is lowered through the attribute (40 -> 60 evaluations per
plane) to keep the run short.  Also stores mask_image_depth / reconstruct_mip_image on small random inputs, including depth
maps whose values sit exactly on bin edges (pixels in two planes, uint8 wrap-around in the sum).
"""
import os
import sys

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import nst_oracle as O  # noqa: E402  (inputs only: synth_image)
import make_golden as MG  # noqa: E402


def synth_depth(h, w, seed):
    """a smooth field in 0..255: large-scale structure like a depth map"""
    z = O.synth_image(h, w, seed, beta=3.0)[:, :, 0].astype(np.float64)
    z = (z - z.min()) / max(z.max() - z.min(), 1e-9)
    return np.round(z * 255).astype(np.uint8)


def main():
    MG.install_shim()
    torch.set_num_threads(1)
    import transformers
    depth_box = {}

    def fake_pipeline(task=None, model=None, **kw):
        return lambda image: {"depth": Image.fromarray(depth_box["depth"])}

    transformers.pipeline = fake_pipeline
    from components.style_transfer_depth.style_transfer_depth import DepthStyle
    from components.style_transfer_depth import util as U

    out = {}
    # ---- full multi-plane run
    H, W, n = 48, 64, 3
    content = O.synth_image(H, W, 11)
    style = O.synth_image(40, 56, 12)
    depth_box["depth"] = synth_depth(H, W, 13)
    ds = DepthStyle(device="cpu")
    ds.style_model.num_steps = 40
    ds.style_model.print_iter = 10 ** 9
    losses = []
    orig_backward = torch.Tensor.backward

    def backward(self, *a, **k):
        losses.append(float(self.detach()))
        return orig_backward(self, *a, **k)

    torch.Tensor.backward = backward
    try:
        final, stylized = ds.style_MIP(Image.fromarray(content), Image.fromarray(style), n)
    finally:
        torch.Tensor.backward = orig_backward
    per = len(losses) // n
    out.update(mip_content=content, mip_style=style, mip_depth=depth_box["depth"], mip_n=np.int64(n), mip_num_steps=np.int64(40),
               mip_final=np.asarray(final), mip_stylized=np.stack([np.asarray(s) for s in stylized], 0),
               mip_losses=np.array(losses, dtype=np.float64).reshape(n, per),
               mip_planes=np.stack([np.asarray(p) for p in ds.depth_split(Image.fromarray(content), n)], 0))
    # ---- the byte functions alone
    rng = np.random.default_rng(5)
    cases = []
    for (h, w, n_, kind) in ((16, 20, 2, "u8"), (9, 7, 3, "u8"), (12, 16, 4, "edges"), (8, 8, 2, "flat"), (10, 12, 5, "f32"), (16, 16, 10, "edges2"),
                             (6, 10, 3, "gray")):
        img = rng.integers(0, 256, (h, w) if kind == "gray" else (h, w, 3), dtype=np.uint8)
        if kind in ("u8", "gray"):
            d = rng.integers(0, 256, (h, w), dtype=np.uint8)
        elif kind == "edges":
            d = rng.integers(0, 5, (h, w), dtype=np.uint8) + 7       # range 4: values exactly on the edges of 2 and 4 bins
        elif kind == "edges2":
            d = rng.integers(0, 11, (h, w), dtype=np.uint8) * 20     # range 200, 10 bins
        elif kind == "flat":
            d = np.full((h, w), 9, dtype=np.uint8)                   # max == min: nan, nothing selected
        else:
            d = rng.random((h, w)).astype(np.float32) * 3 - 1
        with np.errstate(all="ignore"):
            planes = [np.asarray(p) for p in U.generate_mip_layers(Image.fromarray(img), d, n_)]
            styl = [Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)) for _ in range(n_)]
            merged = np.asarray(U.reconstruct_mip_image(styl, d, n_))
        k = len(cases)
        out.update({"b_img_%d" % k: img, "b_depth_%d" % k: d, "b_n_%d" % k: np.int64(n_), "b_planes_%d" % k: np.stack(planes, 0),
                    "b_styl_%d" % k: np.stack([np.asarray(s) for s in styl], 0), "b_merged_%d" % k: merged})
        cases.append(kind)
    out["b_count"] = np.int64(len(cases))
    np.savez_compressed(os.path.join(HERE, "depth_mip.npz"), **out)
    print("wrote depth_mip.npz:", per, "evaluations per plane; final losses", out["mip_losses"][:, -1])


if __name__ == "__main__":
    main()
