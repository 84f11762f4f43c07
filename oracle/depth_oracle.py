"""CPU restatement of the depth-aware / multi-plane variant - TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's CPU leg).

Follows /root/reference/components/style_transfer_depth:
  util.py:9-35   mask_image_depth     util.py:38-50  create_bins     util.py:53-66  generate_mip_layers
  util.py:69-88  reconstruct_mip_image
  Style_a3.py:62-166 (_run_style_transfer: the loop of multi_style_transfer with one style, restated by oracle/nst_oracle.py),
  Style_a3.py:45 (vgg_std[0] = 0.485), :168-193 (style_transfer: w_style = 5e5 * e^(strength - 1 / strength)),
  style_transfer_depth.py:61-90 (process_mip_layers, style_MIP).
The depth term (Style_a3.py:142-146) goes through save_image -> PIL -> the depth model -> image_loader: it has no gradient
path, so it adds a constant to the loss value only; it is left out here, as in the product (the model cannot be downloaded).

Pinned to the reference itself: tests/golden/make_golden_depth.py runs the unmodified DepthStyle.style_MIP / depth_split with
a stub depth estimator and random-init VGG weights, and stores planes, per-plane loss traces, stylised planes and the merged
image (tests/golden/depth_mip.npz); tests/test_depth_oracle.py holds this file to those."""
import numpy as np

from . import nst_oracle as O

VGG_STD_A3 = (0.485, 0.224, 0.225)   # Style_a3.py:45


def mask_image_depth(image, depth, thresholds):
    image = np.asarray(image)
    depth = np.asarray(depth)
    if len(depth.shape) > 2:
        raise ValueError("The depth map (image2) must be a single-channel image.")
    with np.errstate(all="ignore"):
        depth = (depth - np.min(depth)) / (np.max(depth) - np.min(depth))        # :27
    lo, hi = thresholds
    mask = (depth >= lo) & (depth <= hi)                                          # :30
    out = np.copy(image)
    out[~mask] = 0                                                                # :32-33
    return out


def create_bins(n):
    edges = np.linspace(0, 1, n + 1)
    return [[edges[i], edges[i + 1]] for i in range(n)]


def generate_mip_layers(image, depth, n):
    return [mask_image_depth(image, depth, b) for b in create_bins(n)]


def reconstruct_mip_image(stylized, depth, n):
    bins = create_bins(n)
    mip = np.zeros(np.asarray(stylized[0]).shape[:2] + (3,), dtype=np.uint8)
    for i in range(n):
        mip += mask_image_depth(stylized[i], depth, bins[i])                      # :85-87 (uint8: wraps)
    return mip


def style_weight(strength):
    return 5e5 if strength < 0 else 5e5 * (np.e ** (strength - 1 / strength))     # Style_a3.py:184-187


def style_transfer(weights, biases, style_u8, content_u8, strength=1, num_steps=400, max_evals=None):
    """StyleA3.style_transfer on uint8 arrays -> (uint8 image, OracleResult)."""
    res = O.run_oracle(weights, biases, content_u8, [style_u8], num_steps, w_style=style_weight(strength), w_content=1, w_tv=2e1,
                       w_edge=2e1, std=VGG_STD_A3, max_evals=max_evals)
    return O.to_u8(res.image), res


def style_mip(weights, biases, image_u8, style_u8, depth, n=2, num_steps=400):
    """DepthStyle.style_MIP -> (merged uint8 image, stylised planes, per-plane OracleResults)."""
    planes = generate_mip_layers(image_u8, depth, n)
    outs, results = [], []
    for ind, p in enumerate(planes):
        img, res = style_transfer(weights, biases, style_u8, p, 1 - ind / len(planes), num_steps)   # style_transfer_depth.py:72
        outs.append(img)
        results.append(res)
    return reconstruct_mip_image(outs, depth, n), outs, results
