#!/usr/bin/env python3
"""Step-0 parity (loss terms, pixel gradient) of the CUDA path against the CPU oracle at sizes the unit tests do not cover
(720p video frames, odd sizes whose pooled levels are odd).  Usage: python tools/size_sweep.py [HxW ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import nst_b200  # noqa: E402,F401
from nst_b200 import synth  # noqa: E402
from oracle import nst_oracle as O  # noqa: E402
from importlib import import_module  # noqa: E402

hf = import_module("text-based-image-style-transfer_b200.multi_style_transfer.helper_functions")
rst = import_module("text-based-image-style-transfer_b200.multi_style_transfer.run_style_transfer")
ws, bs = synth.vgg19_random_weights(1234, 13)
hf.set_vgg_weight_provider(lambda: (ws, bs))
sizes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(720, 1280), (333, 517), (130, 66)]
style_u8 = synth.synth_image(256, 256, 1)
for H, W in sizes:
    content = synth.synth_image(H, W, 0)
    c = O.to_tensor_u8(content)
    x = (c + 0.03 * torch.randn(c.shape, generator=torch.Generator().manual_seed(5))).clamp(0, 1)
    sess = rst.StyleTransferSession(synth.VGG_MEAN, synth.VGG_STD, (H, W), [O.to_tensor_u8(style_u8).cuda()], device="cuda",
                                    **synth.APP_WEIGHTS)
    sess.prepare(c.cuda(), trace_capacity=64)
    with torch.cuda.stream(sess.stream):
        losses, grad = sess.plan.eval(x.cuda())
    t0 = time.time()
    co = O.ClosureOracle(ws, bs, c, [O.to_tensor_u8(style_u8)], **O.APP_WEIGHTS)
    ref = co.evaluate(x)
    got = losses.cpu().tolist()
    terms = {k: abs(got[i] - ref[k]) / (abs(ref[k]) + 1e-12) for i, k in enumerate(("total", "content", "style", "tv", "edge"))}
    gerr = float((grad.cpu() - ref["grad"]).norm() / ref["grad"].norm())
    n = sess.run(20)
    tr = sess.trace()[:, 0]
    print("%4dx%-4d  rel err: %s  grad %.2e | 40 evaluations: loss %.5f -> %.5f, finite %s (oracle %.0f s)" % (
        H, W, " ".join("%s %.1e" % kv for kv in terms.items()), gerr, float(tr[0]), float(tr[n - 1]), bool(torch.isfinite(tr[:n]).all()),
        time.time() - t0))
    assert all(v < 1e-3 for v in terms.values()) and gerr < 1e-3
    sess.close()
print("size sweep ok")
