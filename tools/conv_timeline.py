#!/usr/bin/env python3
"""Where do the convolution launches sit inside the captured L-BFGS step?  Launch spans (earliest CTA start .. latest
CTA end, %globaltimer) of every tcgen05 convolution launch of ONE evaluation in the middle of a CUDA-graph step,
with the gaps between consecutive launches.  Usage: python tools/conv_timeline.py [size]"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import nst_b200  # noqa: E402
nst_b200._lib.load_instrumented()  # tools/build.py --instrument: the product library has no stamps / timing experiments
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
sess.run(100)   # history fills up; clocks settle
lib = nst_b200._lib.load()
st = C.c_void_p(sess.stream.cuda_stream)
buf = (C.c_ulonglong * 384)()
nst_b200._lib.check(lib.nst_plan_timeline(sess.plan.handle, 1, None, st))
sess.run(20)    # re-capture with the slots armed
nst_b200._lib.check(lib.nst_plan_timeline(sess.plan.handle, 1, None, st))
sess.run(20)
nst_b200._lib.check(lib.nst_plan_timeline(sess.plan.handle, 0, buf, st))
v = list(buf)
spans = []
for slot in range(48):
    t0, t1, tw, ta, tf0, tf1, tl, ck = v[8 * slot:8 * slot + 8]
    if t1 == 0:
        continue
    tag = ("fwd", "dgrad", "gram_bwd")[slot // 16]
    spans.append((t0, t1, "%s %d" % (tag, slot % 16), tw, ta, tf0, tf1, tl, ck))
spans.sort()
base = spans[0][0]
print("%-12s %10s %10s %10s %10s | %s" % ("launch", "start us", "end us", "span us", "gap us",
      "previous launch's last CTA end -> dependency wait returns -> first patch of the first CTA (.. of the last CTA) -> last MMA issued -> last accumulator complete -> last CTA end"))
prev_end = None
tot = {"fwd": 0.0, "dgrad": 0.0, "gram_bwd": 0.0}
for t0, t1, name, tw, ta, tf0, tf1, tl, ck in spans:
    gap = (t0 - prev_end) / 1e3 if prev_end is not None else 0.0
    extra = ""
    if prev_end is not None and tw != 2 ** 64 - 1 and ta != 0:
        extra = " | %6.1f %6.1f (%4.1f) %6.1f %6.1f %6.1f" % ((tw - prev_end) / 1e3, (tf0 - tw) / 1e3, (tf1 - tw) / 1e3, (tl - tf0) / 1e3, (ta - tl) / 1e3, (t1 - ta) / 1e3)
        if ck & 0xffffffff:
            extra += "  SM clock %4.0f MHz" % ((ck >> 32) * 1e3 / (ck & 0xffffffff))
    print("%-12s %10.1f %10.1f %10.1f %10.1f%s" % (name, (t0 - base) / 1e3, (t1 - base) / 1e3, (t1 - t0) / 1e3, gap, extra))
    tot[name.split()[0]] += (t1 - t0) / 1e3
    prev_end = max(prev_end, t1) if prev_end is not None else t1
print("sum of spans:", {k: round(x, 1) for k, x in tot.items()}, "first start -> last end: %.1f us" % ((max(s[1] for s in spans) - base) / 1e3))
