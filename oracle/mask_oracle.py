"""CPU restatement (numpy, integer / fp64 arithmetic) of the reference's mask compositing - TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product never does.

Restates  text/segmentation_style_transfer.py:5-94  (segmentation_style_transfer, _edge_smoothing): the step that follows
the style-transfer loop in six of the reference's eleven call sites (app.py:203,318,407,512,...): the stylised image is
blended into the original through a segmentation mask whose edges are softened by cv2.GaussianBlur.

Third-party arithmetic not under /root/reference: OpenCV (requirements.txt:7 `opencv-python`, unpinned; 4.13.0 in this image).
cv2.GaussianBlur(uint8, (k, k), sigmaX=0) is restated from OpenCV's published algorithm (modules/imgproc/src/smooth.dispatch.cpp):
  * kernel: for k <= 7 the fixed table small_gaussian_tab, else exp(-x^2 / (2 sigma^2)) normalised, sigma = 0.3((k-1)/2 - 1) + 0.8
    (getGaussianKernelBitExact);
  * 8-bit fixed point with error diffusion on the first half, centre tap = 256 - 2 * sum (getGaussianKernelFixedPoint_ED);
  * separable, BORDER_REFLECT_101, horizontal pass in 8.8 fixed point, vertical pass in 16.16, result (v + 2^15) >> 16.
Pinned against cv2 itself for every odd k <= 63 and against golden vectors written by the UNMODIFIED reference function
(tests/golden/make_golden_mask.py -> tests/golden/mask_composite.npz): tests/test_mask_oracle.py.
"""
import math

import numpy as np

_SMALL = {1: [1.0], 3: [0.25, 0.5, 0.25], 5: [0.0625, 0.25, 0.375, 0.25, 0.0625],
          7: [0.03125, 0.109375, 0.21875, 0.28125, 0.21875, 0.109375, 0.03125]}


def gaussian_kernel_fixed(k: int):
    """The 8-bit fixed-point kernel (integers summing to 256) cv2.GaussianBlur uses for uint8 images, sigma = 0."""
    assert k >= 1 and k % 2 == 1
    if k <= 7:
        g = _SMALL[k]
    else:
        sigma = 0.3 * ((k - 1) * 0.5 - 1) + 0.8
        scale2x = -0.5 / (sigma * sigma)
        e = [math.exp(scale2x * (i - (k - 1) * 0.5) ** 2) for i in range(k)]
        s = sum(e)
        g = [v / s for v in e]
    n2 = k // 2
    err, tot = 0.0, 0
    res = [0] * k
    for i in range(n2):
        adj = g[i] * 256.0 + err
        v0 = int(np.rint(adj))
        err = adj - v0
        res[i] = res[k - 1 - i] = v0
        tot += v0
    res[n2] = 256 - 2 * tot
    return res


def _reflect101(i, n):
    if n == 1:
        return np.zeros_like(i)
    period = 2 * n - 2
    i = np.mod(i, period)
    return np.where(i >= n, period - i, i)


def gaussian_blur_u8(img: np.ndarray, k: int) -> np.ndarray:
    """cv2.GaussianBlur(img, (k, k), 0) for a 2-D uint8 image, bit for bit."""
    w = gaussian_kernel_fixed(k)
    r = k // 2
    H, W = img.shape
    src = img.astype(np.int64)
    cols = _reflect101(np.arange(-r, W + r), W)
    rows = _reflect101(np.arange(-r, H + r), H)
    p = src[rows][:, cols]
    tmp = np.zeros((H + 2 * r, W), np.int64)
    for j in range(k):
        tmp += w[j] * p[:, j:j + W]
    out = np.zeros((H, W), np.int64)
    for i in range(k):
        out += w[i] * tmp[i:i + H, :]
    return ((out + 32768) >> 16).astype(np.uint8)


def edge_smoothing(content: np.ndarray, style: np.ndarray, mask: np.ndarray, blur_strength: int = 5) -> np.ndarray:
    """_edge_smoothing, segmentation_style_transfer.py:59-94 (fp64 blend, truncation to uint8)."""
    if blur_strength % 2 != 1:
        blur_strength += 1                                                         # :76-77
    num = np.where(mask, 1, 0).astype(np.uint8) * 255                              # :80
    blurred = gaussian_blur_u8(num, blur_strength)                                 # :83
    nb = blurred / 255.0                                                           # :86
    return (content * (1 - nb[..., None]) + style * nb[..., None]).astype(np.uint8)  # :91


def crop_to_common(content: np.ndarray, style: np.ndarray, mask: np.ndarray):
    """Centre crop of the larger of the two images (and of the mask with the content), :28-45."""
    c_H, c_W = content.shape[:2]
    s_H, s_W = style.shape[:2]
    if c_H < s_H:
        o = (s_H - c_H) // 2
        style = style[o:o + c_H]
    elif s_H < c_H:
        o = (c_H - s_H) // 2
        content, mask = content[o:o + s_H], mask[o:o + s_H]
    if c_W < s_W:
        o = (s_W - c_W) // 2
        style = style[:, o:o + c_W]
    elif s_W < c_W:
        o = (c_W - s_W) // 2
        content, mask = content[:, o:o + s_W], mask[:, o:o + s_W]
    return content, style, mask


def segmentation_style_transfer(content: np.ndarray, style: np.ndarray, mask: np.ndarray, edge_smoothing_k=5) -> np.ndarray:
    """segmentation_style_transfer, :5-57, on uint8 HWC arrays (the PIL conversions of :24-25,55 are the identity on them)."""
    content, style, mask = crop_to_common(content, style, mask)
    if edge_smoothing_k:
        return edge_smoothing(content, style, mask, blur_strength=edge_smoothing_k)
    m = np.repeat(mask[:, :, None], content.shape[2], axis=2)
    return np.where(m > 0, style, content)                                         # :52
