#!/usr/bin/env python3
"""Phase times of the L-BFGS controller launch (one thread block, latency bound) at full history.
Usage: python tools/ctl_phases.py [size]"""
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
sess.run(140)   # history full (100 pairs)
lib = nst_b200._lib.load()
buf = (C.c_longlong * 8)()
nst_b200._lib.check(lib.nst_lbfgs_ctl_clocks(sess.plan.handle, buf, C.c_void_p(sess.stream.cuda_stream)))
t = list(buf)
names = ["stage matrices (2 x 80 KB bulk copy) + stop tests", "accept pair, initialise rows", "loop 1 (newest -> oldest)",
         "y.r initialisation (block parallel)", "loop 2 (oldest -> newest)", "coefficients, step length, g.d"]
for i, n in enumerate(names):
    print("%-52s %6.2f us" % (n, (t[i + 1] - t[i]) / 1.965e3))
print("%-52s %6.2f us" % ("total inside the kernel", (t[6] - t[0]) / 1.965e3))
