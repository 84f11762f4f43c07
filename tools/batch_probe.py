#!/usr/bin/env python3
"""Batch head against the single-image plan, evaluation by evaluation: loss terms of the first evaluation and the first
update (x1 - x0 is the first gradient up to a scalar) of every member.  Usage: python tools/batch_probe.py [H W batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import nst_b200  # noqa: E402
from nst_b200 import synth, engine  # noqa: E402
from importlib import import_module  # noqa: E402

H = int(sys.argv[1]) if len(sys.argv) > 1 else 64
W = int(sys.argv[2]) if len(sys.argv) > 2 else 80
B = int(sys.argv[3]) if len(sys.argv) > 3 else 3
hf = import_module("text-based-image-style-transfer_b200.multi_style_transfer.helper_functions")
rst = import_module("text-based-image-style-transfer_b200.multi_style_transfer.run_style_transfer")
ws, bs = synth.vgg19_random_weights(1234, 13)
hf.set_vgg_weight_provider(lambda: (ws, bs))
img = lambda h, w, seed: torch.from_numpy(synth.synth_image(h, w, seed)).permute(2, 0, 1).float().div(255).unsqueeze(0).cuda()  # noqa: E731
s0 = rst.StyleTransferSession(synth.VGG_MEAN, synth.VGG_STD, (H, W), [img(56, 72, 5)], device="cuda", **synth.APP_WEIGHTS)
cs = [img(H, W, 70 + k) for k in range(B)]
with torch.cuda.stream(s0.stream):
    bp = engine.BatchPlan(s0.net, H, W, rst.STYLE_LAYERS, rst.CONTENT_LAYERS, B, mean=s0.mean, std=s0.std)
    bp.set_weights(*s0.weights)
    bp.set_style_targets(s0.style_targets)
    for m, c in zip(bp.members, cs):
        m.features(c)
        for name in m.content_layers:
            m.set_content_target(name, m, None)
        m.set_edge_target(c)
        m.lbfgs_init(c, 64)
    for n_evals in (2, 20):
        for m, c in zip(bp.members, cs):
            m.lbfgs_init(c, 64)
        bp.head.lbfgs_partial_step(n_evals)
        s0.stream.synchronize()
        rows = []
        for k, m in enumerate(bp.members):
            rows.append((m.lbfgs_trace(64).double(), m.lbfgs_x().cpu().double()))
        for k in range(B):
            s0.prepare(cs[k], trace_capacity=64)
            s0.plan.lbfgs_partial_step(n_evals)
            s0.stream.synchronize()
            tr1 = s0.plan.lbfgs_trace(64).double()
            x1 = s0.plan.lbfgs_x().cpu().double()
            tr, x = rows[k]
            d = x - cs[k].cpu().double()
            d1 = x1 - cs[k].cpu().double()
            print("evals %2d member %d: loss terms of evaluation 0 (total c s tv e) rel dev %s | last eval total %.3e | update rel dev %.3e (|update| %.3e)"
                  % (n_evals, k, " ".join("%.1e" % float(abs(tr[0, j] - tr1[0, j]) / max(abs(tr1[0, j]), 1e-30)) for j in range(5)),
                     float(abs(tr[-1, 0] - tr1[-1, 0]) / abs(tr1[-1, 0])), float((d - d1).norm() / d1.norm()), float(d1.norm())))
bp.close()
s0.close()
