"""Mirror of multi_style_transfer/StyleMixer.py (reference :7-38): both feature maps are resized with
bilinear interpolation (align_corners=True) to (H1 + H2 // 2, W1 + W2 // 2) - the reference's operator
precedence - and blended as (1 - w) * style1 + w * style2."""
from typing import List

import torch
from torch import Tensor

from .. import _lib
from ..engine import _f32c, _ptr, _require_cuda, _stream_ptr
from .._lib import check


class StyleMixer():
    def __init__(self, style_feat_list: List[Tensor], style_feat_weight: float):
        self.style_feat_list = [style_feat_list[0].detach(), style_feat_list[1].detach()]
        # weights sum to 1; `style_feat_weight` belongs to the SECOND style (reference :23)
        self.style_feat_weights = [1 - style_feat_weight, style_feat_weight]

    def mix(self):
        a, b = self.style_feat_list
        dev = _require_cuda(a.device)
        a, b = _f32c(a, dev), _f32c(b, dev)
        if a.shape[0] != 1 or b.shape[0] != 1 or a.shape[1] != b.shape[1]:
            raise _lib.NstError("StyleMixer: expected two (1,C,H,W) feature tensors with equal C")
        c = int(a.shape[1])
        ha, wa, hb, wb = int(a.shape[2]), int(a.shape[3]), int(b.shape[2]), int(b.shape[3])
        out = torch.empty((1, c, ha + hb // 2, wa + wb // 2), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            check(_lib.load().nst_style_mix_tensors(_ptr(a), ha, wa, _ptr(b), hb, wb, c,
                                                    float(self.style_feat_weights[1]), _ptr(out), _stream_ptr(dev)))
        return out
