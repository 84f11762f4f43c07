#!/usr/bin/env python3
"""Where do the chained convolution launches wait?  Runs the forward and the backward chain once each with the wait
accounting of conv_chain.cu and prints, per role, the mean / max over the CTAs in microseconds.
Usage: python tools/chain_waits.py [size]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import nst_b200  # noqa: E402
from nst_b200 import synth  # noqa: E402
from importlib import import_module  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
hf = import_module("text-based-image-style-transfer_b200.multi_style_transfer.helper_functions")
rst = import_module("text-based-image-style-transfer_b200.multi_style_transfer.run_style_transfer")
ws, bs = synth.vgg19_random_weights(1234, 13)
hf.set_vgg_weight_provider(lambda: (ws, bs))
img = lambda seed: torch.from_numpy(synth.synth_image(S, S, seed)).permute(2, 0, 1).float().div(255).unsqueeze(0).cuda()  # noqa: E731
sess = rst.StyleTransferSession(synth.VGG_MEAN, synth.VGG_STD, (S, S), [img(1)], device="cuda", **synth.APP_WEIGHTS)
sess.prepare(img(0))
sess.run(40)
lib = nst_b200._lib.load()
st = C.c_void_p(sess.stream.cuda_stream)
names = ["start", "end", "producer: waits for the scout", "producer: operand ring full", "MMA: operands not landed",
         "MMA: accumulator stage busy", "epilogue: accumulator not ready", "epilogue: hand-off to publisher", "epilogue: seed acquire", "tiles",
         "scout: dependency wait"]
MHZ = 1965.0
for which, tag in ((0, "forward chain"), (1, "backward chain")):
    buf = (C.c_longlong * (16 * 160 + 256))()
    for rep in range(2):
        n = lib.nst_plan_chain_waits(sess.plan.handle, which, buf, 160, st)
        nst_b200._lib.check(min(n, 0))
    flat = np.array(list(buf), dtype=np.int64)
    a = flat[:16 * n].reshape(n, 16)
    per_layer = flat[16 * n:16 * n + 256].reshape(64, 4)
    life = (a[:, 1] - a[:, 0]) / MHZ
    print("%s: %d CTAs, CTA lifetime mean %.1f us, max %.1f us" % (tag, n, life.mean(), life.max()))
    for k in list(range(2, 9)) + [10]:
        v = a[:, k] / MHZ
        print("  %-34s mean %7.1f us   max %7.1f us" % (names[k], v.mean(), v.max()))
    print("  tiles per CTA: %d .. %d" % (a[:, 9].min(), a[:, 9].max()))
    print("  scout wait per chain layer (mean over CTAs, us): " + " ".join("%.1f" % (x / MHZ / n) for x in per_layer[:, 0] if x > 0 or True)[:400])
