"""Mirror of text/segmentation_style_transfer.py (same function name, signature and PIL in / PIL out), on the GPU.

The reference merges the stylised image into the original on the host (numpy + cv2.GaussianBlur) right after
run_multi_style_transfer (app.py:203,318,407,512); here it is one CUDA kernel (csrc/mask.cu, C entry nst_mask_composite),
bit-exact with the reference, so that a pipeline can keep the stylised image on the device (`composite_tensors`)."""
import ctypes as C

import numpy as np
import torch
from PIL import Image

from .. import _lib
from .._lib import check
from ..engine import _require_cuda


def composite_tensors(content: torch.Tensor, style: torch.Tensor, mask: torch.Tensor, edge_smoothing: int = 5) -> torch.Tensor:
    """content, style: (H, W, C) uint8 CUDA tensors of the same shape, mask: (H, W) bool / uint8 CUDA tensor -> (H, W, C) uint8."""
    dev = _require_cuda(content.device)
    if content.dtype != torch.uint8 or style.dtype != torch.uint8 or content.shape != style.shape or content.dim() != 3:
        raise ValueError("content and style must be uint8 (H, W, C) tensors of the same shape")
    H, W, Cc = content.shape
    if tuple(mask.shape) != (H, W):
        raise ValueError("mask must be (H, W)")
    content, style = content.contiguous(), style.to(dev).contiguous()
    mask = mask.to(dev)
    if mask.dtype == torch.bool:
        m8 = mask.contiguous().view(torch.uint8)     # one byte per element already: no copy
    elif mask.dtype == torch.uint8:
        m8 = mask.contiguous()                        # the kernel tests for non-zero
    else:
        m8 = (mask != 0).contiguous().view(torch.uint8)
    out = torch.empty_like(content)
    lib = _lib.load()
    with torch.cuda.device(dev):
        check(lib.nst_mask_composite(C.c_void_p(content.data_ptr()), C.c_void_p(style.data_ptr()), C.c_void_p(m8.data_ptr()),
                                     int(H), int(W), int(Cc), int(edge_smoothing), C.c_void_p(out.data_ptr()),
                                     C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out


def segmentation_style_transfer(content_image, style_image, segmentation_mask, edge_smoothing=5, device="cuda"):
    """Same contract as the reference (segmentation_style_transfer.py:5-57): where the mask is True the pixel comes from the
    style image, elsewhere from the content image; `edge_smoothing` (made odd) is the Gaussian kernel size that softens
    the mask edges, 0 / None = hard selection.  Images of different size are centre-cropped to the smaller one."""
    content = np.array(content_image)
    style = np.array(style_image)
    mask = np.asarray(segmentation_mask)
    c_H, c_W = content.shape[:2]
    s_H, s_W = style.shape[:2]
    if c_H < s_H:                                   # :31-37
        o = (s_H - c_H) // 2
        style = style[o:o + c_H]
    elif s_H < c_H:
        o = (c_H - s_H) // 2
        content, mask = content[o:o + s_H], mask[o:o + s_H]
    if c_W < s_W:                                   # :39-45
        o = (s_W - c_W) // 2
        style = style[:, o:o + c_W]
    elif s_W < c_W:
        o = (c_W - s_W) // 2
        content, mask = content[:, o:o + s_W], mask[:, o:o + s_W]
    dev = _require_cuda(device)
    k = int(edge_smoothing) if edge_smoothing else 0
    out = composite_tensors(torch.from_numpy(np.ascontiguousarray(content)).to(dev), torch.from_numpy(np.ascontiguousarray(style)).to(dev),
                            torch.from_numpy(np.ascontiguousarray(mask > 0)).to(dev), k)
    return Image.fromarray(out.cpu().numpy())
