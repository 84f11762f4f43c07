"""Mirror of multi_style_transfer/ChannelAttention.py (reference :3-40): squeeze-and-excitation style gate
x * sigmoid(relu(fc2(relu(fc1(avgpool(x)))))), two bias-free linear layers with PyTorch's default init.
The reference's constructor calls `super()._init_()` (:11) and therefore raises AttributeError as shipped;
this mirror implements the intended module."""
import torch
from torch import nn

from .. import _lib
from ..engine import _f32c, _ptr, _require_cuda, _stream_ptr
from .._lib import check


class ChannelAttention(nn.Module):
    def __init__(self, in_channels, reduction_ratio=2):
        super().__init__()
        self.channels = in_channels
        self.reduction_ratio = reduction_ratio
        self.fc1 = nn.Linear(self.channels, self.channels // self.reduction_ratio, bias=False)
        self.fc2 = nn.Linear(self.channels // self.reduction_ratio, self.channels, bias=False)

    def forward(self, x):
        dev = _require_cuda(x.device)
        xc = _f32c(x, dev)
        if xc.shape[0] != 1 or xc.shape[1] != self.channels:
            raise _lib.NstError("ChannelAttention: expected a (1,%d,H,W) tensor" % self.channels)
        w1, w2 = _f32c(self.fc1.weight, dev), _f32c(self.fc2.weight, dev)
        y = torch.empty_like(xc)
        with torch.cuda.device(dev):
            check(_lib.load().nst_channel_attention_chw(_ptr(xc), self.channels, int(xc.shape[2]), int(xc.shape[3]),
                                                        _ptr(w1), _ptr(w2), self.reduction_ratio, _ptr(y),
                                                        _stream_ptr(dev)))
        return y
