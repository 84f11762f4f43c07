#!/usr/bin/env python3
"""Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference
(/root/reference/multi_style_transfer) on CPU.  Run in the build container only (the GPU box has no
/root/reference); the .npz files it writes are committed.

Shim, no source edits:
  * stub `matplotlib` / `matplotlib.pyplot` (helper_functions.py:3 imports it, only save_image uses it);
  * torchvision.models.vgg19 ignores `weights=` and builds random-init weights under torch.manual_seed(1234)
    with the caller's RNG state saved and restored (pretrained weights cannot be downloaded);
  * nn.Module._init_ = nn.Module.__init__ so that ChannelAttention.py:11's typo does not raise.
Instrumentation: Tensor.backward is wrapped to record every closure's loss and the gradient it produced.
"""
import os
import sys
import types

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import nst_oracle as O  # noqa: E402  (inputs only: synth_image)


def install_shim():
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    import torchvision
    orig = torchvision.models.vgg19

    def vgg19(weights=None, **kw):
        state = torch.random.get_rng_state()
        torch.manual_seed(1234)
        m = orig(weights=None, **kw)
        torch.random.set_rng_state(state)
        return m

    torchvision.models.vgg19 = vgg19
    torch.nn.Module._init_ = torch.nn.Module.__init__
    sys.path.insert(0, REF)


def run_reference(content_u8, styles_u8, num_steps, weights, style_img_weight=0.5, channel_attention=False,
                  random_init=False):
    from multi_style_transfer.run_style_transfer import run_multi_style_transfer
    losses, grads, iterates = [], [], []
    orig_backward = torch.Tensor.backward
    orig_clamp = torch.Tensor.clamp_

    def backward(self, *a, **k):
        losses.append(float(self.detach()))
        return orig_backward(self, *a, **k)

    def clamp_(self, *a, **k):
        r = orig_clamp(self, *a, **k)
        if isinstance(self, torch.nn.Parameter):
            iterates.append(self.detach().clone())
        return r

    torch.Tensor.backward = backward
    torch.Tensor.clamp_ = clamp_
    try:
        mean = torch.tensor(O.VGG_MEAN)
        std = torch.tensor(O.VGG_STD)
        imgs = [Image.fromarray(s) for s in styles_u8]
        out = run_multi_style_transfer(mean, std, Image.fromarray(content_u8), num_steps, random_init,
                                       weights["w_style"], weights["w_content"], weights["w_tv"], weights["w_edge"],
                                       imgs[0], imgs[1] if len(imgs) > 1 else None, style_img_weight,
                                       print_iter=10 ** 9, channel_attention=channel_attention, device="cpu")
    finally:
        torch.Tensor.backward = orig_backward
        torch.Tensor.clamp_ = orig_clamp
    return np.asarray(out), np.array(losses, dtype=np.float64), iterates


def step0_reference(content_u8, styles_u8, weights, style_img_weight=0.5, channel_attention=False):
    """Per-term losses, per-layer Grams / Gram MSEs and the pixel gradient of the reference's closure at
    x = content, computed with the reference's own functions."""
    from multi_style_transfer.run_style_transfer import (PIL_to_tensor, channel_att_per_chosen_layers)
    from multi_style_transfer import style_transfer_losses as L
    from multi_style_transfer.helper_functions import Vgg19, seed_everything, to_grayscale
    seed_everything(101)
    mean = torch.tensor(O.VGG_MEAN)
    std = torch.tensor(O.VGG_STD)
    cl, sl = ['conv4_2'], ['conv1_1', 'conv2_1', 'conv3_1', 'conv4_1', 'conv5_1']
    content = PIL_to_tensor(Image.fromarray(content_u8))
    styles = [L.normalize(PIL_to_tensor(Image.fromarray(s)), mean, std) for s in styles_u8]
    model = Vgg19(cl, sl, "cpu")
    normed_content = L.normalize(content, mean, std)
    target_grad = L.get_gradient_imgs(to_grayscale(normed_content)).detach()
    with torch.no_grad():
        sfeats = [model(s) for s in styles]
        cfeats = model(normed_content)
    if channel_attention:
        cfeats = channel_att_per_chosen_layers(cfeats, cl, device="cpu")
    x = torch.nn.Parameter(content.clone())
    normed = L.normalize(x, mean, std)
    feats = model(normed)
    c = weights["w_content"] * L.content_loss(feats, cfeats, cl)
    s = weights["w_style"] * L.style_loss(feats, sfeats, sl, style_img_weight)
    tv = weights["w_tv"] * L.total_variation_loss(normed)
    e = weights["w_edge"] * L.edge_loss(target_grad, L.get_gradient_imgs(to_grayscale(x)))
    total = torch.tensor([0.]) + s + c + tv + e
    total.backward()
    out = dict(total=float(total), content=float(c), style=float(s), tv=float(tv), edge=float(e),
               grad=x.grad.detach().numpy().copy())
    with torch.no_grad():
        for name in sl:
            g = L.gram_matrix(feats[name])[0]
            if len(sfeats) == 1:
                t = L.gram_matrix(sfeats[0][name])[0]
            else:
                from multi_style_transfer.StyleMixer import StyleMixer
                t = L.gram_matrix(StyleMixer([f[name] for f in sfeats], style_img_weight).mix())[0]
            out["gram_mse_" + name] = float(((g - t) ** 2).mean())
            # Grams up to 128x128 are stored whole; larger ones as their leading 64x64 block plus moments
            out["gram_" + name] = g[:128, :128].numpy().copy() if g.shape[0] <= 128 else g[:64, :64].numpy().copy()
            out["gram_fro_" + name] = float(g.norm())
            out["target_fro_" + name] = float(t.norm())
            out["feat_rms_" + name] = float(feats[name].pow(2).mean().sqrt())
        out["content_feat_rms"] = float(cfeats['conv4_2'].pow(2).mean().sqrt())
        out["content_feat_sample"] = cfeats['conv4_2'][0, :8, :4, :4].detach().numpy().copy()
    return out


CASES = {
    # name: (content (h, w, seed), [style (h, w, seed)], num_steps, style_img_weight, channel_attention)
    "single_64": ((64, 64, 0), [(64, 64, 1)], 50, 0.5, False),
    "odd_50x38": ((50, 38, 3), [(44, 60, 4)], 20, 0.5, False),
    "mix_ca_48x40": ((48, 40, 5), [(40, 56, 6), (64, 48, 7)], 20, 0.3, True),
}


def main():
    install_shim()
    torch.set_num_threads(1)  # single-thread summation order: the most reproducible reference
    for name, (cspec, sspecs, steps, wgt, ca) in CASES.items():
        content = O.synth_image(*cspec)
        styles = [O.synth_image(*s) for s in sspecs]
        img, losses, iterates = run_reference(content, styles, steps, O.APP_WEIGHTS, wgt, ca)
        s0 = step0_reference(content, styles, O.APP_WEIGHTS, wgt, ca)
        arrays = dict(final_u8=img, loss_trace=losses, x_eval5=iterates[5].numpy(), x_last=iterates[-1].numpy(),
                      n_evals=np.array(len(losses)))
        for k, v in s0.items():
            arrays["s0_" + k] = np.asarray(v)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **arrays)
        print(name, "evals", len(losses), "loss", losses[0], "->", losses[-1], os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
