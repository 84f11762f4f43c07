"""Mirror of multi_style_transfer/style_transfer_losses.py: the same seven functions, same argument meaning,
computed by the kernels of libnst_b200.so on CUDA tensors.  They return plain tensors (no autograd graph):
inside the optimisation loop the fused closure of the library produces losses and gradient together
(engine.Plan.eval); these functions exist so that callers and tests can evaluate the individual terms
exactly as they would with the reference."""
import ctypes as C
from typing import List

import torch
from torch import Tensor

from .. import _lib
from ..engine import _f32c, _ptr, _require_cuda, _stream_ptr, gram_chw
from .._lib import check, f3
from .helper_functions import *  # noqa: F401,F403  (the reference re-exports its helpers the same way)
from .StyleMixer import StyleMixer


def _dev(t: Tensor):
    return _require_cuda(t.device)


def _scalar(device):
    return torch.empty((), device=device, dtype=torch.float32)


def normalize(img, mean, std):
    """ Z-normalizes an image tensor (b, 3, h, w) with per-channel mean / std   (reference :9-28)."""
    dev = _dev(img)
    x = _f32c(img, dev)
    b, c, h, w = x.shape
    if c != 3:
        raise RuntimeError("The size of tensor a (%d) must match the size of tensor b (3) at non-singleton dimension 1" % c)
    m = [float(v) for v in torch.as_tensor(mean).reshape(-1).cpu()]
    s = [float(v) for v in torch.as_tensor(std).reshape(-1).cpu()]
    y = torch.empty_like(x)
    with torch.cuda.device(dev):
        check(_lib.load().nst_normalize(_ptr(x), _ptr(y), b * c, c, h, w, f3(m), f3(s), _stream_ptr(dev)))
    return y


def _gray(img):
    dev = _dev(img)
    x = _f32c(img, dev)
    b, c, h, w = x.shape
    if b != 1:
        raise _lib.NstError("to_grayscale: only batch size 1 is implemented on this path")
    y = torch.empty((1, 1, h, w), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        check(_lib.load().nst_grayscale(_ptr(x), _ptr(y), c, h, w, _stream_ptr(dev)))
    return y


def _mse(a: Tensor, b: Tensor) -> Tensor:
    dev = _dev(a)
    if tuple(a.shape) != tuple(b.shape):
        raise RuntimeError("The size of tensor a %s must match the size of tensor b %s" % (tuple(a.shape), tuple(b.shape)))
    x, y = _f32c(a, dev), _f32c(b, dev)
    out = _scalar(dev)
    with torch.cuda.device(dev):
        check(_lib.load().nst_mse(_ptr(x), _ptr(y), x.numel(), _ptr(out), _stream_ptr(dev)))
    return out


def content_loss(input_features, content_features, content_layers):
    """ Content loss of Gatys et al. 2016: mean over `content_layers` of the feature MSE   (reference :31-67)."""
    loss = 0.0
    for layer in content_layers:
        loss = loss + _mse(input_features[layer], content_features[layer].detach())
    return loss / len(content_layers)


def gram_matrix(x):
    """ Normalised Gram matrix X X^T / (b c h w) of a (b, c, h, w) feature tensor   (reference :70-95)."""
    return gram_chw(x)


def style_loss(input_features, style_features: List[Tensor], style_layers, style_img_weight):
    """ Style loss: mean over `style_layers` of the MSE between Gram matrices; with two style feature dicts the
    target is the Gram of their StyleMixer blend   (reference :98-146)."""
    loss = 0.0
    for layer in style_layers:
        g_in = gram_matrix(input_features[layer])
        if len(style_features) == 1:
            g_style = gram_matrix(style_features[0][layer])
        else:
            mixer = StyleMixer([feats[layer] for feats in style_features], style_img_weight)
            g_style = gram_matrix(mixer.mix())
        loss = loss + _mse(g_in, g_style)
    return loss / len(style_layers)


def total_variation_loss(y):
    """ Total variation over the spatial dimensions, normalised by c*h*w   (reference :149-174)."""
    dev = _dev(y)
    x = _f32c(y, dev)
    b, c, h, w = x.shape
    out = _scalar(dev)
    with torch.cuda.device(dev):
        check(_lib.load().nst_total_variation(_ptr(x), b * c, h, w, _ptr(out), _stream_ptr(dev)))
    return out * b if b != 1 else out


def get_gradient_imgs(img):
    """ Central-difference gradient images of a grayscale (1,1,H,W) tensor -> (1,2,H-2,W-2), x first   (reference :177-204)."""
    dev = _dev(img)
    x = _f32c(img, dev)
    if x.dim() != 4 or x.shape[0] != 1 or x.shape[1] != 1:
        raise _lib.NstError("get_gradient_imgs expects a (1,1,H,W) tensor")
    h, w = int(x.shape[2]), int(x.shape[3])
    full = torch.empty((1, 2, h, w), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        check(_lib.load().nst_edge_images(_ptr(x), 1, h, w, _ptr(full), _stream_ptr(dev)))
    return full[:, :, 1:h - 1, 1:w - 1].contiguous()


def edge_loss(img1, img2):
    """ Mean of the per-direction MSEs between two (1,2,H,W) gradient-image tensors   (reference :207-225)."""
    dx = _mse(img1[:, 0, :, :], img2[:, 0, :, :])
    dy = _mse(img1[:, 1, :, :], img2[:, 1, :, :])
    return (dx + dy) / 2
