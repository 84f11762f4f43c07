#!/usr/bin/env python3
"""How accurate is the gradient of the CUDA closure where the VGG terms dominate?  Teacher-forced on the oracle's own
trajectory (pixel terms off): per iterate the gradient norm, the relative error of the whole gradient, of the style part and
of the content part alone, and - as the yardstick - the relative difference between the oracle in fp32 and in fp64 at the same
iterate.  Near a stationary point |g| shrinks while the parts stay large, so the error RELATIVE TO |g| grows for any finite
precision; the last column puts it against |g| of the first evaluation instead.
Usage: python tools/grad_error_probe.py [size]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import nst_b200  # noqa: E402
from oracle import nst_oracle as O  # noqa: E402
from importlib import import_module  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 96
hf = import_module("text-based-image-style-transfer_b200.multi_style_transfer.helper_functions")
rst = import_module("text-based-image-style-transfer_b200.multi_style_transfer.run_style_transfer")
ws, bs = O.vgg19_random_weights(1234, 13)
hf.set_vgg_weight_provider(lambda: (ws, bs))
content, style = O.synth_image(S, S, 0), O.synth_image(S, S, 1)
c, st = O.to_tensor_u8(content), O.to_tensor_u8(style)


def rel(a, b):
    return float((a.double().cpu() - b.double().cpu()).norm() / (b.double().cpu().norm() + 1e-300))


for tag, wts in (("pixel terms off", dict(w_style=5e5, w_content=1.0, w_tv=0.0, w_edge=0.0)),
                 ("vgg terms x100", dict(w_style=5e7, w_content=1e2, w_tv=2e1, w_edge=2e1)),
                 ("app.py weights", dict(O.APP_WEIGHTS))):
    ref = O.run_oracle(ws, bs, content, [style], 0, keep_iterates=True, **wts)
    parts = {"all": wts, "style": dict(wts, w_content=0.0, w_tv=0.0, w_edge=0.0), "content": dict(wts, w_style=0.0, w_tv=0.0, w_edge=0.0)}
    sess, cos, cos64 = {}, {}, {}
    ws64, bs64 = [w.double() for w in ws], [b.double() for b in bs]
    for name, w in parts.items():
        s = rst.StyleTransferSession(O.VGG_MEAN, O.VGG_STD, (S, S), [st.cuda()], w["w_style"], w["w_content"], w["w_tv"], w["w_edge"], 0.5, "cuda")
        s.prepare(c.cuda())
        sess[name] = s
        cos[name] = O.ClosureOracle(ws, bs, c, [st], **w)
        cos64[name] = O.ClosureOracle(ws64, bs64, c.double(), [st.double()], **w)
    print("== %s, %dx%d: loss %.4g -> %.4g over %d evaluations" % (tag, S, S, ref.losses[0][0], ref.losses[-1][0], ref.evals))
    print("%4s %11s | %9s %9s %9s | %9s %9s %9s | %9s" % ("eval", "|g|", "all", "style", "content", "all f32/64", "sty f32/64", "con f32/64", "all/|g0|"))
    g0 = None
    for k in (0, 1, 2, 3, 5, 7, 10, 12, 15, 19):
        xk = ref.iterates[k]
        row = {}
        for name in parts:
            r32 = cos[name].evaluate(xk)["grad"]
            r64 = cos64[name].evaluate(xk.double())["grad"]
            with torch.cuda.stream(sess[name].stream):
                _, g = sess[name].plan.eval(xk.cuda())
            row[name] = (rel(g, r64), rel(r32, r64), float(r64.norm()), float((g.double().cpu() - r64).norm()))
        if g0 is None:
            g0 = row["all"][2]
        print("%4d %11.4e | %9.2e %9.2e %9.2e | %9.2e %9.2e %9.2e | %9.2e" % (k, row["all"][2], row["all"][0], row["style"][0], row["content"][0],
                                                                       row["all"][1], row["style"][1], row["content"][1], row["all"][3] / g0))
    for s in sess.values():
        s.close()
