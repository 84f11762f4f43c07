"""Mirror of multi_style_transfer/helper_functions.py: Vgg19 feature extractor, to_grayscale, seed_everything.

Vgg19 keeps the reference contract (helper_functions.py:44-101): constructed from content and style layer
names plus a device, `forward(x)` takes a NORMALISED (1,3,H,W) image and returns a dict
{layer name: PRE-ReLU feature (1,C,h,w) fp32}, `layer_names` lists them in network order, unknown names raise.
The convolutions run in libnst_b200.so (tcgen05 implicit GEMM, fp16 operands, fp32 accumulation).
"""
import random
from typing import Callable, List, Optional, Tuple

import numpy as np
import torch

from ..engine import CONV_NAMES, Net, Plan, _require_cuda

# Source of the VGG-19 convolution parameters.  Default: torchvision's ImageNet weights exactly as the
# reference requests them (helper_functions.py:47).  Offline (tests, bench) a provider returning
# random-init weights is installed with set_vgg_weight_provider().
_provider: Optional[Callable[[], Tuple[List[torch.Tensor], List[torch.Tensor]]]] = None
_provider_tag = 0
_net_cache = {}


def _torchvision_imagenet():
    import torchvision.models as models
    feats = models.vgg19(weights=models.VGG19_Weights.IMAGENET1K_V1).features
    convs = [m for m in feats if isinstance(m, torch.nn.Conv2d)]
    return [c.weight.detach() for c in convs], [c.bias.detach() for c in convs]


def set_vgg_weight_provider(fn):
    """fn() -> (weights, biases): lists of torch-layout conv parameters for conv1_1.. in order."""
    global _provider, _provider_tag
    _provider = fn
    _provider_tag += 1
    _net_cache.clear()


def get_net(device) -> Net:
    device = _require_cuda(device)
    key = (str(device), _provider_tag)
    net = _net_cache.get(key)
    if net is None:
        ws, bs = (_provider or _torchvision_imagenet)()
        net = Net(list(ws), list(bs), device)
        _net_cache[key] = net
    return net


def seed_everything(seed):
    np.random.seed(seed)
    torch.manual_seed(seed)
    random.seed(seed)


class Vgg19(torch.nn.Module):
    def __init__(self, content_layers, style_layers, device):
        super().__init__()
        wanted = list(content_layers) + list(style_layers)
        if any(n not in CONV_NAMES for n in wanted):
            raise Exception('Not all layers provided in content_layes and/or style_layers exist.')
        self.device = _require_cuda(device)
        self.net = get_net(self.device)
        self.layer_names = [n for n in CONV_NAMES if n in set(wanted)]
        if CONV_NAMES.index(self.layer_names[-1]) >= self.net.n_conv:
            raise Exception('Not all layers provided in content_layes and/or style_layers exist.')
        self._plans = {}

    def plan_for(self, H, W) -> Plan:
        p = self._plans.get((H, W))
        if p is None:
            if len(self._plans) >= 4:
                self._plans.pop(next(iter(self._plans))).close()
            p = Plan(self.net, H, W, self.layer_names)  # identity normalisation: forward() takes normalised input
            self._plans[(H, W)] = p
        return p

    def forward(self, x):
        if x.dim() != 4 or x.shape[0] != 1 or x.shape[1] != 3:
            raise ValueError("Vgg19.forward expects a (1,3,H,W) tensor")
        p = self.plan_for(int(x.shape[2]), int(x.shape[3]))
        p.features(x)
        return {name: p.get_tap(name) for name in self.layer_names}


def to_grayscale(img):
    """helper_functions.py:104-113: channel mean of a (1,3,H,W) tensor -> (1,1,H,W)."""
    from . import style_transfer_losses as L
    return L._gray(img)
