#!/usr/bin/env python3
"""Where does a convolution launch spend its time?  Phase timestamps (SM clock) of CTA 0 of every conv layer."""
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
c = img(0)
sess.prepare(c)
sess.run(20)
lib = nst_b200._lib.load()
buf = (C.c_longlong * 176)()
names = ["setup", "first operands", "main loop (issue)", "drain -> acc ready", "epilogue", "teardown"]
print("%-10s %s  total | waits of CTA 0 (us): MMA patch / weights / acc stage, epilogue acc, producer patch / weight stage" % ("launch", "  ".join("%18s" % n for n in names)))
with torch.cuda.stream(sess.stream):
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        nst_b200._lib.check(lib.nst_plan_conv_phases(sess.plan.handle, 0, 0, buf, C.c_void_p(sess.stream.cuda_stream)))
    print("conv1_1 forward (tensor cores), CTA 0, us: build %.1f  wait stage %.1f  write+fence %.1f  stash+barrier %.1f | epilogue group 0 waits for accumulator %.1f" % tuple(x / 1.965e3 for x in list(buf)[:5]))
    for mode, tag in ((0, "fwd"), (1, "dgrad")):
        for conv in range(0 if mode == 1 else 1, 13):
            for rep in range(2):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(sess.stream)
                nst_b200._lib.check(lib.nst_plan_conv_phases(sess.plan.handle, conv, mode, buf, C.c_void_p(sess.stream.cuda_stream)))
                e1.record(sess.stream)
                e1.synchronize()
                wall = e0.elapsed_time(e1) * 1e3
            t = list(buf)[:7]
            d = [(t[i + 1] - t[i]) / 1.965e3 for i in range(6)]
            w = [x / 1.965e3 for x in list(buf)[8:14]]
            life = [x / 1.965e3 for x in list(buf)[16:176] if x > 0]
            print("%-10s %s  %6.1f us | %5.1f %5.1f %5.1f  %5.1f  %5.1f %5.1f | CTA lifetimes: %d CTAs min %.1f mean %.1f max %.1f us | first CTA start -> last CTA end (globaltimer) %.1f us" % (
                "%s %d" % (tag, conv), "  ".join("%15.1f us" % v for v in d), (t[6] - t[0]) / 1.965e3, *w, len(life), min(life), sum(life) / len(life), max(life), (list(buf)[15] - list(buf)[14]) / 1e3))
