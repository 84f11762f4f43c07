#!/usr/bin/env python3
"""Small profiling target: BASELINE configs[1] (512x512 pair), three optimizer.step() graph launches = 60 closure evaluations
with the history filling up.  Meant to run under ncu (profiles/README.md gives the two command lines): the full bench.py
launches ~30 000 kernels, far too many to replay.  Usage: python tools/ncu_target.py [size] [steps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import nst_b200  # noqa: E402,F401
from nst_b200 import synth  # noqa: E402
from importlib import import_module  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 512
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 3
hf = import_module("text-based-image-style-transfer_b200.multi_style_transfer.helper_functions")
rst = import_module("text-based-image-style-transfer_b200.multi_style_transfer.run_style_transfer")
ws, bs = synth.vgg19_random_weights(1234, 13)
hf.set_vgg_weight_provider(lambda: (ws, bs))
img = lambda seed: torch.from_numpy(synth.synth_image(S, S, seed)).permute(2, 0, 1).float().div(255).unsqueeze(0).cuda()  # noqa: E731
sess = rst.StyleTransferSession(synth.VGG_MEAN, synth.VGG_STD, (S, S), [img(1)], device="cuda", **synth.APP_WEIGHTS)
sess.prepare(img(0), trace_capacity=64 + 20 * STEPS)
for _ in range(STEPS):
    sess.step()
st = sess.status()
print("evaluations", st.closure_calls, "loss", st.loss, "stored pairs", st.hist_len)
sess.close()
