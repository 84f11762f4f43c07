"""Mirror of components/style_transfer_depth/util.py:9-101: depth-bin masking, plane generation and reconstruction (same names,
same PIL / numpy contracts), computed on the GPU (csrc/depth.cu: nst_mip_split, nst_mip_merge), plus image_loader / save_image.

The losses and the Vgg19 class that file copies from multi_style_transfer (util.py:104-348) live in
..multi_style_transfer and are re-exported here under the same names."""
import ctypes as C

import numpy as np
import torch
from PIL import Image

from ... import _lib
from ...engine import _require_cuda
from ...multi_style_transfer.helper_functions import Vgg19, to_grayscale                                    # noqa: F401
from ...multi_style_transfer.run_style_transfer import PIL_to_tensor, tensor_to_PIL
from ...multi_style_transfer.style_transfer_losses import (content_loss, edge_loss, get_gradient_imgs, gram_matrix,  # noqa: F401
                                                           normalize, style_loss, total_variation_loss)

_DEVICE = "cuda"


def _depth_args(depth, dev, depth_range=None):
    """The depth map as the kernels take it: uint8 maps go up as they are with their min / max (normalised in the kernel exactly
    like util.py:27), any other dtype is normalised here with the reference's own expression and goes up as fp64.
    A uint8 CUDA tensor is used where it is (depth_range = its (min, max) if the caller knows them, else one reduction)."""
    if isinstance(depth, torch.Tensor) and depth.is_cuda:
        if depth.dtype != torch.uint8 or depth.dim() != 2:
            raise ValueError("a device depth map must be a (H, W) uint8 tensor")
        lo, hi = depth_range if depth_range is not None else (int(depth.min()), int(depth.max()))
        return depth.to(dev).contiguous(), 0, int(lo), int(hi)
    depth = np.asarray(depth)
    if len(depth.shape) > 2:
        raise ValueError("The depth map (image2) must be a single-channel image.")          # util.py:24-25
    if depth.dtype == np.uint8:
        return torch.from_numpy(np.array(depth)).to(dev), 0, int(depth.min()), int(depth.max())
    with np.errstate(all="ignore"):
        norm = (depth - np.min(depth)) / (np.max(depth) - np.min(depth))                      # util.py:27
    return torch.from_numpy(np.ascontiguousarray(norm, dtype=np.float64)).to(dev), 1, 0, 0


def _bounds(bins):
    lo = (C.c_double * len(bins))(*[float(b[0]) for b in bins])
    hi = (C.c_double * len(bins))(*[float(b[1]) for b in bins])
    return lo, hi


def split_planes(image_u8: torch.Tensor, depth, bins, device=None, depth_range=None) -> torch.Tensor:
    """image_u8: (H, W, C) uint8 CUDA tensor -> (len(bins), H, W, C): plane i keeps the pixels whose normalised depth is in bins[i]."""
    dev = _require_cuda(device or image_u8.device)
    image_u8 = image_u8.to(dev).contiguous()
    H, W, Cc = image_u8.shape
    d, f64, dmin, dmax = _depth_args(depth, dev, depth_range)
    if tuple(d.shape) != (H, W):
        raise IndexError("boolean index did not match indexed array: depth %s vs image %s" % (tuple(d.shape), (H, W)))
    out = torch.empty((len(bins), H, W, Cc), dtype=torch.uint8, device=dev)
    lo, hi = _bounds(bins)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().nst_mip_split(C.c_void_p(image_u8.data_ptr()), C.c_void_p(d.data_ptr()), f64, dmin, dmax, int(H), int(W),
                                             int(Cc), len(bins), lo, hi, C.c_void_p(out.data_ptr()),
                                             C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out


def merge_planes(planes_u8: torch.Tensor, depth, bins, device=None, depth_range=None) -> torch.Tensor:
    """planes_u8: (n, H, W, 3) uint8 CUDA tensor -> (H, W, 3): the planes masked with their bins and added up in uint8."""
    dev = _require_cuda(device or planes_u8.device)
    planes_u8 = planes_u8.to(dev).contiguous()
    n, H, W, Cc = planes_u8.shape
    if Cc != 3 or n != len(bins):
        raise ValueError("planes must be (len(bins), H, W, 3)")
    d, f64, dmin, dmax = _depth_args(depth, dev, depth_range)
    if tuple(d.shape) != (H, W):
        raise IndexError("boolean index did not match indexed array: depth %s vs image %s" % (tuple(d.shape), (H, W)))
    out = torch.empty((H, W, 3), dtype=torch.uint8, device=dev)
    lo, hi = _bounds(bins)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().nst_mip_merge(C.c_void_p(planes_u8.data_ptr()), C.c_void_p(d.data_ptr()), f64, dmin, dmax, int(H), int(W),
                                             n, lo, hi, C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out


def mask_image_depth(image1, depth, thresholds):
    """util.py:9-35: the image with every pixel whose normalised depth is outside [min, max] zeroed (PIL image)."""
    image1 = np.asarray(image1)
    a = image1 if image1.ndim == 3 else image1[:, :, None]
    out = split_planes(torch.from_numpy(np.array(a)).to(_DEVICE), depth, [thresholds])[0].cpu().numpy()
    return Image.fromarray(out if image1.ndim == 3 else out[:, :, 0])


def create_bins(n):
    """util.py:38-50"""
    bin_edges = np.linspace(0, 1, n + 1)
    return [[bin_edges[i], bin_edges[i + 1]] for i in range(n)]


def generate_mip_layers(image, depth, n):
    """util.py:53-66: one masked PIL image per depth bin (one kernel launch for all of them)."""
    image = np.asarray(image)
    a = image if image.ndim == 3 else image[:, :, None]
    planes = split_planes(torch.from_numpy(np.array(a)).to(_DEVICE), depth, create_bins(n)).cpu().numpy()
    return [Image.fromarray(p if image.ndim == 3 else p[:, :, 0]) for p in planes]


def reconstruct_mip_image(stylized_images, depth, n):
    """util.py:69-88: the stylised planes masked with their bins and summed (uint8, wrapping where bins share an edge)."""
    planes = np.stack([np.asarray(im) for im in stylized_images[:n]], 0)
    return Image.fromarray(merge_planes(torch.from_numpy(planes).to(_DEVICE), depth, create_bins(n)).cpu().numpy())


def image_loader(img, device="cuda"):
    """util.py:91-94: ToTensor()(img).unsqueeze(0)"""
    return PIL_to_tensor(img).to(device, torch.float)


def save_image(tensor):
    """util.py:97-101"""
    return tensor_to_PIL(tensor.detach().cpu().clone().squeeze(0))
