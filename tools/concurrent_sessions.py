#!/usr/bin/env python3
"""Do independent optimisations that share a GPU fill each other's launch ramps and tails?
K sessions (own plan, own stream, own CUDA graph) are stepped round-robin from one host thread; reports the aggregate
closure evaluations per second for K = 1, 2, 3 at 512x512 and 1280x720.
  python tools/concurrent_sessions.py [--sizes 512x512,720x1280] [--ks 1,2,3] [--iters 8]"""
import argparse
import os
import sys
import time
from importlib import import_module

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nst_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="512x512,720x1280")
ap.add_argument("--ks", default="1,2,3")
ap.add_argument("--iters", type=int, default=8)
args = ap.parse_args()
hf = import_module("text-based-image-style-transfer_b200.multi_style_transfer.helper_functions")
rst = import_module("text-based-image-style-transfer_b200.multi_style_transfer.run_style_transfer")
ws, bs = synth.vgg19_random_weights(1234, 13)
hf.set_vgg_weight_provider(lambda: (ws, bs))
dev = torch.device("cuda:0")
style = torch.from_numpy(synth.synth_image(512, 512, 1)).permute(2, 0, 1).float().div(255).unsqueeze(0).to(dev)
for size in args.sizes.split(","):
    H, W = (int(v) for v in size.split("x"))
    for K in (int(k) for k in args.ks.split(",")):
        sessions, targets = [], None
        for k in range(K):
            s = rst.StyleTransferSession(synth.VGG_MEAN, synth.VGG_STD, (H, W), [style], device=dev, style_targets=targets, **synth.APP_WEIGHTS)
            targets = s.style_targets
            if K > 1:
                s.plan.set_shared_gpu(True)   # several chains of pair launches side by side: no programmatic launch on them
            content = torch.from_numpy(synth.synth_image(H, W, 100 + k)).permute(2, 0, 1).float().div(255).unsqueeze(0).to(dev)
            s.prepare(content, None, False, trace_capacity=0)
            sessions.append(s)
        for _ in range(2):
            for s in sessions:
                s.step()
            for s in sessions:
                s.status()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.iters):
            for s in sessions:
                s.step()
            for s in sessions:
                s.status()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("%-9s K=%d  %8.1f evaluations/s aggregate  (%.3f ms per evaluation per session)" %
              (size, K, K * 20 * args.iters / dt, 1e3 * dt / (20 * args.iters)), flush=True)
        for s in sessions:
            s.close()
