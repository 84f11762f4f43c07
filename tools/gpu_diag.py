#!/usr/bin/env python3
"""Stage-by-stage GPU diagnostic: every stage runs in its own process (a CUDA fault or a trapped mbarrier
watchdog in one stage must not poison the next) and prints error figures against the CPU oracle.
Usage: python tools/gpu_diag.py [--stages 0,1,2] [--size 64]     (on a B200 box, via gpurun)"""
import argparse
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def rel(a, b):
    import torch
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-300))


def setup(size_h, size_w, n_conv=13):
    import torch
    import nst_b200
    from oracle import nst_oracle as O
    ws, bs = O.vgg19_random_weights(1234, n_conv)
    net = nst_b200.Net(ws, bs, "cuda")
    content = O.synth_image(size_h, size_w, 0)
    style = O.synth_image(size_h, size_w, 1)
    return torch, nst_b200, O, ws, bs, net, content, style


def stage0(a):
    import torch
    import nst_b200
    lib = nst_b200._lib.load()
    print("abi", lib.nst_abi_version(), "device", torch.cuda.get_device_name(0), torch.cuda.get_device_capability(0))
    print("device_check", lib.nst_device_check(), lib.nst_last_error())
    torch, nst, O, ws, bs, net, content, style = setup(a.size, a.size)
    print("net ok", net.n_conv)


def stage1(a):
    torch, nst, O, ws, bs, net, content, style = setup(a.size, a.size + 6)
    c = O.to_tensor_u8(content)
    x = (c + 0.05 * torch.randn(c.shape, generator=torch.Generator().manual_seed(3))).clamp(0, 1)
    co = O.ClosureOracle(ws, bs, c, [O.to_tensor_u8(style)], w_style=0, w_content=0, w_tv=20.0, w_edge=20.0)
    ref = co.evaluate(x)
    plan = nst.Plan(net, a.size, a.size + 6, O.STYLE_LAYERS + O.CONTENT_LAYERS, O.STYLE_LAYERS, O.CONTENT_LAYERS, True,
                    O.VGG_MEAN, O.VGG_STD)
    plan.set_weights(0, 0, 20.0, 20.0)
    plan.set_edge_target(c.cuda())
    losses, grad = plan.eval(x.cuda())
    l = losses.cpu().tolist()
    print("pixel-only  total %.7f ref %.7f | tv %.7f ref %.7f | edge %.7f ref %.7f" % (l[0], ref["total"], l[3], ref["tv"], l[4], ref["edge"]))
    print("pixel-only  grad rel err %.3e" % rel(grad, ref["grad"]))


def stage2(a):
    torch, nst, O, ws, bs, net, content, style = setup(a.size, a.size - 14)
    H, W = a.size, a.size - 14
    c = O.to_tensor_u8(content)
    names = O.CONV_NAMES[:13]
    vgg = O.VggFeatures(ws, bs, names)
    ref = vgg(O.normalize(c, O.VGG_MEAN, O.VGG_STD))
    for upto in (1, 2, 3, 5, 9, 13):
        plan = nst.Plan(net, H, W, names[:upto], mean=O.VGG_MEAN, std=O.VGG_STD)
        plan.features(c.cuda())
        torch.cuda.synchronize()
        errs = ["%s %.2e" % (n, rel(plan.get_tap(n), ref[n])) for n in names[:upto]]
        print("taps upto %2d:" % upto, "  ".join(errs), flush=True)
        plan.close()


def stage3(a):
    import torch
    import nst_b200
    g = torch.Generator().manual_seed(0)
    for (C, H, W) in [(64, 16, 16), (64, 40, 52), (128, 32, 32), (256, 24, 20), (512, 16, 16), (512, 7, 5), (96, 20, 20),
                      (64, 256, 256)]:
        x = torch.randn((1, C, H, W), generator=g) * 0.3
        refg = torch.bmm(x.double().reshape(1, C, -1), x.double().reshape(1, C, -1).transpose(1, 2)) / (C * H * W)
        out = nst_b200.gram_chw(x.cuda())
        torch.cuda.synchronize()
        print("gram C=%d HW=%dx%d rel err %.3e  (fp16-rounded input: %.3e)" % (
            C, H, W, rel(out, refg),
            rel(out, torch.bmm(x.half().double().reshape(1, C, -1), x.half().double().reshape(1, C, -1).transpose(1, 2)) / (C * H * W))),
            flush=True)


def make_session(nst, O, ws, bs, content, styles, weights, mix_w=0.5, ca=False):
    import torch
    from importlib import import_module
    hf = import_module("text-based-image-style-transfer_b200.multi_style_transfer.helper_functions")
    hf.set_vgg_weight_provider(lambda: (ws, bs))
    rst = import_module("text-based-image-style-transfer_b200.multi_style_transfer.run_style_transfer")
    c = O.to_tensor_u8(content).cuda()
    sess = rst.StyleTransferSession(O.VGG_MEAN, O.VGG_STD, c.shape[2:], [O.to_tensor_u8(s).cuda() for s in styles],
                                    weights["w_style"], weights["w_content"], weights["w_tv"], weights["w_edge"], mix_w, "cuda")
    return sess, c


def stage4(a):
    torch, nst, O, ws, bs, net, content, style = setup(a.size, a.size)
    for wts in (O.APP_WEIGHTS, dict(w_style=5e5, w_content=1.0, w_tv=0.0, w_edge=0.0)):
        sess, c = make_session(nst, O, ws, bs, content, [style], wts)
        x = (c.cpu() + 0.05 * torch.randn(c.shape, generator=torch.Generator().manual_seed(3))).clamp(0, 1)
        co = O.ClosureOracle(ws, bs, c.cpu(), [O.to_tensor_u8(style)], **wts)
        ref = co.evaluate(x)
        torch.manual_seed(101)
        sess.prepare(c)
        with torch.cuda.stream(sess.stream):
            losses, grad = sess.plan.eval(x.cuda())
        l = losses.cpu().tolist()
        print("weights", wts)
        print("  total %.6f ref %.6f | content %.6e ref %.6e | style %.6e ref %.6e | tv %.6f ref %.6f | edge %.6f ref %.6f" % (
            l[0], ref["total"], l[1], ref["content"], l[2], ref["style"], l[3], ref["tv"], l[4], ref["edge"]))
        print("  gram mse ours", ["%.4e" % v for v in l[5:10]])
        print("  gram mse ref ", ["%.4e" % v for v in ref["gram_mse"]])
        print("  grad rel err %.3e  |grad| ours %.4e ref %.4e" % (rel(grad, ref["grad"]), float(grad.norm()), float(ref["grad"].norm())), flush=True)
        sess.close()


def stage5(a):
    torch, nst, O, ws, bs, net, content, style = setup(a.size, a.size)
    steps = a.steps
    t0 = time.time()
    ref = O.run_oracle(ws, bs, content, [style], steps, **O.APP_WEIGHTS)
    print("oracle %d evals in %.1fs: loss %.5f -> %.5f" % (ref.evals, time.time() - t0, ref.losses[0][0], ref.losses[-1][0]))
    sess, c = make_session(nst, O, ws, bs, content, [style], O.APP_WEIGHTS)
    sess.prepare(c, trace_capacity=ref.evals + 64)
    t0 = time.time()
    n = sess.run(steps)
    out = sess.result()
    dt = time.time() - t0
    tr = sess.trace()
    st = sess.status()
    print("ours   %d evals in %.3fs: loss %.5f -> %.5f  n_iter %d hist %d stop %d launches/step %d" % (
        n, dt, float(tr[0, 0]), float(tr[-1, 0]), st.n_iter, st.hist_len, st.stop, sess.plan.launches_per_step()))
    m = min(len(ref.losses), tr.shape[0])
    dev = max(abs(float(tr[i, 0]) - ref.losses[i][0]) / abs(ref.losses[i][0]) for i in range(m))
    print("loss curve max rel dev %.3e over %d evals; final PSNR vs oracle %.2f dB" % (dev, m, O.psnr(out.cpu(), ref.image)))
    for i in (0, 1, 2, 5, 10, 19, 20, 21, 39, 59):
        if i < m:
            print("  eval %3d ours %.6f ref %.6f" % (i, float(tr[i, 0]), ref.losses[i][0]))


def stage6(a):
    torch, nst, O, ws, bs, net, content, style = setup(64, 64)
    S = a.big
    content = O.synth_image(S, S, 0)
    style = O.synth_image(S, S, 1)
    sess, c = make_session(nst, O, ws, bs, content, [style], O.APP_WEIGHTS)
    print("plan bytes %.1f MB" % (sess.plan.bytes() / 1e6))
    sess.prepare(c)
    sess.run(40)
    torch.cuda.synchronize()
    sess.prepare(c)
    t0 = time.time()
    n = sess.run(a.big_steps)
    torch.cuda.synchronize()
    dt = time.time() - t0
    print("%dx%d: %d evals in %.3f s = %.1f evals/s (%.3f ms/eval), final loss %.5f" % (S, S, n, dt, n / dt, 1e3 * dt / n, sess.status().loss))


STAGES = {0: stage0, 1: stage1, 2: stage2, 3: stage3, 4: stage4, 5: stage5, 6: stage6}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stages", default="0,1,2,3,4,5,6")
    ap.add_argument("--stage", type=int, default=None)
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--big", type=int, default=512)
    ap.add_argument("--big-steps", type=int, default=100)
    ap.add_argument("--timeout", type=int, default=240)
    a = ap.parse_args()
    if a.stage is not None:
        STAGES[a.stage](a)
        return
    for s in [int(v) for v in a.stages.split(",")]:
        print("=" * 20, "stage", s, STAGES[s].__name__, flush=True)
        cmd = [sys.executable, os.path.abspath(__file__), "--stage", str(s), "--size", str(a.size), "--steps", str(a.steps),
               "--big", str(a.big), "--big-steps", str(a.big_steps)]
        try:
            r = subprocess.run(cmd, timeout=a.timeout, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
            print(r.stdout[-6000:])
            print("exit", r.returncode, flush=True)
        except subprocess.TimeoutExpired as e:
            print((e.stdout or b"")[-3000:] if isinstance(e.stdout, (bytes, str)) else "")
            print("TIMEOUT", flush=True)


if __name__ == "__main__":
    main()
