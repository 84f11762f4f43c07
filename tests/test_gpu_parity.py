"""Parity of the CUDA path (through the C ABI) with the CPU oracle and with the golden vectors written by the
reference itself.  Tolerances are the ones BASELINE.json's north_star states:
  * per-layer Gram matrices and loss values within 1e-3 relative at step 0,
  * the loss curve within 1e-2 relative over the run,
  * the final image at >= 40 dB PSNR against the reference.
The CUDA path computes the VGG trunk with fp16 operands / fp32 accumulation (forward) and bf16 operands (data
gradients); pixel-space terms and the optimizer are fp32."""
import importlib
import io
import contextlib
import json
import os

import numpy as np
import pytest
import torch

from conftest import golden

pytestmark = pytest.mark.gpu

GRAM_TOL = 1e-3
LOSS_TOL = 1e-3
CURVE_TOL = 1e-2
PSNR_MIN = 40.0
GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-300))


@pytest.fixture(scope="module")
def nst(built_libs, vgg_weights):
    pkg = importlib.import_module("text-based-image-style-transfer_b200")
    hf = importlib.import_module("text-based-image-style-transfer_b200.multi_style_transfer.helper_functions")
    hf.set_vgg_weight_provider(lambda: vgg_weights)
    return pkg


@pytest.fixture(scope="module")
def rst():
    return importlib.import_module("text-based-image-style-transfer_b200.multi_style_transfer.run_style_transfer")


def session(rst, O, content, styles, weights=None, mix_w=0.5):
    weights = weights or O.APP_WEIGHTS
    c = O.to_tensor_u8(content).cuda()
    s = rst.StyleTransferSession(O.VGG_MEAN, O.VGG_STD, c.shape[2:], [O.to_tensor_u8(x).cuda() for x in styles],
                                 weights["w_style"], weights["w_content"], weights["w_tv"], weights["w_edge"], mix_w, "cuda")
    return s, c


def noisy(c, seed=3):
    return (c.cpu() + 0.05 * torch.randn(c.shape, generator=torch.Generator().manual_seed(seed))).clamp(0, 1)


# ------------------------------------------------------------------------------------------------ step 0
@pytest.mark.parametrize("hw", [(64, 64), (50, 38), (96, 130)])
def test_vgg_taps_match_oracle(nst, oracle, vgg_weights, hw):
    """Vgg19.forward: six pre-ReLU taps, odd sizes exercise the floor of MaxPool2d and partial tiles."""
    O = oracle
    ws, bs = vgg_weights
    c = O.to_tensor_u8(O.synth_image(hw[0], hw[1], 11))
    names = O.STYLE_LAYERS + O.CONTENT_LAYERS
    ref = O.VggFeatures(ws, bs, names)(O.normalize(c, O.VGG_MEAN, O.VGG_STD))
    plan = nst.Plan(nst.multi_style_transfer.helper_functions.get_net("cuda"), hw[0], hw[1], names, mean=O.VGG_MEAN, std=O.VGG_STD)
    plan.features(c.cuda())
    for n in ref:
        got = plan.get_tap(n)
        assert got.shape == ref[n].shape
        assert rel(got, ref[n]) < 2e-3, n
        assert float(got.min()) < 0  # pre-ReLU
    plan.close()


@pytest.mark.parametrize("shape", [(64, 16, 16), (128, 33, 31), (256, 24, 20), (512, 16, 16), (512, 3, 2), (96, 20, 20)])
def test_gram_kernel(nst, shape):
    C, H, W = shape
    x = torch.randn((1, C, H, W), generator=torch.Generator().manual_seed(C + H)) * 0.3
    ref = torch.bmm(x.double().reshape(1, C, -1), x.double().reshape(1, C, -1).transpose(1, 2)) / (C * H * W)
    got = nst.gram_chw(x.cuda())
    assert got.shape == (1, C, C)
    assert rel(got, ref) < GRAM_TOL
    assert torch.equal(got, got.transpose(1, 2))           # exactly symmetric: one triangle is computed, then mirrored
    xh = x.half().double()
    exact = torch.bmm(xh.reshape(1, C, -1), xh.reshape(1, C, -1).transpose(1, 2)) / (C * H * W)
    assert rel(got, exact) < 5e-6                          # fp32 accumulation of fp16-rounded operands


@pytest.mark.parametrize("name", ["single_64", "odd_50x38", "mix_ca_48x40"])
def test_step0_against_reference_golden(nst, rst, oracle, name):
    """Per-term losses, per-layer Gram MSE and Gram matrices at x = content against the REFERENCE's own numbers."""
    from test_oracle import CASES
    O = oracle
    g = golden(name)
    cspec, sspecs, steps, wgt, ca = CASES[name]
    s, c = session(rst, O, O.synth_image(*cspec), [O.synth_image(*x) for x in sspecs], mix_w=wgt)
    torch.manual_seed(101)
    s.prepare(c, channel_attention=ca)
    with torch.cuda.stream(s.stream):
        losses, grad = s.plan.eval(c)
        grams = {n: s.plan.tap_gram(n) for n in O.STYLE_LAYERS}
    l = losses.cpu().tolist()
    for idx, term in ((0, "total"), (1, "content"), (2, "style"), (3, "tv"), (4, "edge")):
        assert l[idx] == pytest.approx(float(g["s0_" + term]), rel=LOSS_TOL, abs=1e-7), term
    for i, layer in enumerate(O.STYLE_LAYERS):
        assert l[5 + i] == pytest.approx(float(g["s0_gram_mse_" + layer]), rel=LOSS_TOL), layer
        blk = torch.from_numpy(g["s0_gram_" + layer])
        assert rel(grams[layer][0][:blk.shape[0], :blk.shape[1]], blk) < GRAM_TOL, layer
        assert float(grams[layer].norm()) == pytest.approx(float(g["s0_gram_fro_" + layer]), rel=GRAM_TOL)
        assert float(s.style_targets[layer].norm()) == pytest.approx(float(g["s0_target_fro_" + layer]), rel=GRAM_TOL)
    assert rel(grad, torch.from_numpy(g["s0_grad"])) < 1e-3
    s.close()


def test_gradient_teacher_forced(nst, rst, oracle, vgg_weights):
    """Same iterate in, loss and gradient out - with and without the pixel-space terms that dominate the total."""
    O = oracle
    ws, bs = vgg_weights
    content, style = O.synth_image(64, 80, 0), O.synth_image(72, 64, 1)
    for wts, tol in ((O.APP_WEIGHTS, 1e-3), (dict(w_style=5e5, w_content=1.0, w_tv=0.0, w_edge=0.0), 5e-2),
                     (dict(w_style=0.0, w_content=0.0, w_tv=20.0, w_edge=20.0), 1e-6)):
        s, c = session(rst, O, content, [style], wts)
        s.prepare(c)
        co = O.ClosureOracle(ws, bs, c.cpu(), [O.to_tensor_u8(style)], **wts)
        for seed in (3, 4):
            x = noisy(c, seed)
            ref = co.evaluate(x)
            with torch.cuda.stream(s.stream):
                losses, grad = s.plan.eval(x.cuda())
            assert float(losses[0]) == pytest.approx(ref["total"], rel=LOSS_TOL)
            assert rel(grad, ref["grad"]) < tol
            cos = float((grad.cpu().double().flatten() @ ref["grad"].double().flatten()) /
                        (grad.double().norm().cpu() * ref["grad"].double().norm()))
            assert cos > 1 - tol
        s.close()


def test_eval_is_deterministic(nst, rst, oracle):
    O = oracle
    s, c = session(rst, O, O.synth_image(64, 64, 0), [O.synth_image(64, 64, 1)])
    s.prepare(c)
    x = noisy(c).cuda()
    with torch.cuda.stream(s.stream):
        l1, g1 = s.plan.eval(x)
        l2, g2 = s.plan.eval(x)
    assert torch.equal(l1, l2) and torch.equal(g1, g2)
    s.close()


# ------------------------------------------------------------------------------------------------ trajectories
@pytest.mark.parametrize("name", ["single_64", "odd_50x38", "mix_ca_48x40"])
def test_trajectory_against_reference_golden(nst, rst, oracle, name):
    """run_multi_style_transfer end to end (PIL in, PIL out) against the reference's recorded run."""
    from PIL import Image
    from test_oracle import CASES
    O = oracle
    g = golden(name)
    cspec, sspecs, steps, wgt, ca = CASES[name]
    content = O.synth_image(*cspec)
    styles = [O.synth_image(*x) for x in sspecs]
    # session form first: gives the loss trace and the float image
    s, c = session(rst, O, content, styles, mix_w=wgt)
    torch.manual_seed(101)
    s.prepare(c, channel_attention=ca, trace_capacity=256)
    n = s.run(steps)
    assert n == int(g["n_evals"])
    tr = s.trace()[:, 0].double().numpy()
    ref = g["loss_trace"]
    assert np.abs(tr - ref).max() <= CURVE_TOL * np.abs(ref).max()
    assert np.all(np.abs(tr - ref) <= CURVE_TOL * np.abs(ref))
    x = s.result().cpu()
    assert O.psnr(x, torch.from_numpy(g["x_last"])) >= PSNR_MIN
    st = s.status()
    assert st.n_iter == n and st.closure_calls == n and st.stop == 0 and st.hist_len <= 100
    s.close()
    # drop-in form
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out = nst.run_multi_style_transfer(torch.tensor(O.VGG_MEAN), torch.tensor(O.VGG_STD), Image.fromarray(content), steps,
                                           False, style_img1=Image.fromarray(styles[0]),
                                           style_img2=Image.fromarray(styles[1]) if len(styles) > 1 else None,
                                           style_img_weight=wgt, print_iter=20, channel_attention=ca, device="cuda",
                                           **O.APP_WEIGHTS)
    text = buf.getvalue()
    assert "Channel attention enabled: " + str(ca) in text                 # run_style_transfer.py:92
    assert f"Reached iteration 20/{steps}" in text                        # :144-146
    assert isinstance(out, Image.Image) and out.size == (cspec[1], cspec[0]) and out.mode == "RGB"
    diff = np.abs(np.asarray(out).astype(int) - g["final_u8"].astype(int))
    assert diff.max() <= 12 and diff.mean() < 1.0
    assert np.array_equal(np.asarray(out), O.to_u8(x))                     # both forms run the same device loop


def test_pixel_only_trajectory_is_tight(nst, rst, oracle, vgg_weights):
    """With the VGG terms off the path is fp32 end to end: the device L-BFGS must track the oracle's closely."""
    O = oracle
    ws, bs = vgg_weights
    wts = dict(w_style=0.0, w_content=0.0, w_tv=20.0, w_edge=20.0)
    content, style = O.synth_image(48, 56, 0), O.synth_image(48, 56, 1)
    ref = O.run_oracle(ws, bs, content, [style], 30, **wts)
    s, c = session(rst, O, content, [style], wts)
    s.prepare(c, trace_capacity=128)
    assert s.run(30) == ref.evals == 40
    tr = s.trace()[:, 0].double().numpy()
    rl = np.array([l[0] for l in ref.losses])
    assert np.abs(tr[:10] - rl[:10]).max() <= 1e-5 * rl[0]
    assert np.abs(tr - rl).max() <= 5e-3 * rl[0]
    assert O.psnr(s.result().cpu(), ref.image) > 50.0
    s.close()


def test_random_init_and_num_steps_semantics(nst, rst, oracle):
    """num_steps counts closure evaluations between optimizer.step() calls (20 each); random_init starts from randn,
    which the first closure clamps (run_style_transfer.py:84,108-109)."""
    O = oracle
    s, c = session(rst, O, O.synth_image(32, 32, 0), [O.synth_image(32, 32, 1)])
    for n, want in ((0, 20), (19, 20), (20, 40), (45, 60)):
        s.prepare(c, trace_capacity=128)
        assert s.run(n) == want
    x0 = torch.randn(c.shape, device="cuda", generator=torch.Generator("cuda").manual_seed(5))
    s.prepare(c, x0, trace_capacity=64)
    assert s.run(0) == 20
    x = s.result()
    assert float(x.min()) >= 0.0 and float(x.max()) <= 1.0
    tr = s.trace()[:, 0]
    assert torch.isfinite(tr).all()
    s.close()


# ------------------------------------------------------------------------------------------------ function-level mirror
def test_function_level_mirrors(nst, oracle, vgg_weights):
    """The seven loss functions + Vgg19 / StyleMixer / ChannelAttention called the way the reference's closure calls them."""
    O = oracle
    ws, bs = vgg_weights
    L = importlib.import_module("text-based-image-style-transfer_b200.multi_style_transfer.style_transfer_losses")
    CA = importlib.import_module("text-based-image-style-transfer_b200.multi_style_transfer.ChannelAttention")
    mean, std = torch.tensor(O.VGG_MEAN), torch.tensor(O.VGG_STD)
    c = O.to_tensor_u8(O.synth_image(40, 52, 0))
    s1 = O.to_tensor_u8(O.synth_image(36, 44, 1))
    s2 = O.to_tensor_u8(O.synth_image(48, 40, 2))
    model = L.Vgg19(O.CONTENT_LAYERS, O.STYLE_LAYERS, "cuda")
    assert model.layer_names == ["conv1_1", "conv2_1", "conv3_1", "conv4_1", "conv4_2", "conv5_1"]
    normed = L.normalize(c.cuda(), mean, std)
    assert rel(normed, O.normalize(c, O.VGG_MEAN, O.VGG_STD)) < 1e-7
    feats = model(normed)
    sfeats = [model(L.normalize(s.cuda(), mean, std)) for s in (s1, s2)]
    ovgg = O.VggFeatures(ws, bs, O.CONTENT_LAYERS + O.STYLE_LAYERS)
    ofeats = ovgg(O.normalize(c, O.VGG_MEAN, O.VGG_STD))
    osf = [ovgg(O.normalize(s, O.VGG_MEAN, O.VGG_STD)) for s in (s1, s2)]
    for n in ofeats:
        assert rel(feats[n], ofeats[n]) < 2e-3
    assert float(L.total_variation_loss(normed)) == pytest.approx(float(O.total_variation_loss(O.normalize(c, O.VGG_MEAN, O.VGG_STD))), rel=1e-5)
    gi = L.get_gradient_imgs(L.to_grayscale(c.cuda()))
    ogi = O.get_gradient_imgs(O.to_grayscale(c))
    assert gi.shape == ogi.shape and rel(gi, ogi) < 1e-6
    gi2 = L.get_gradient_imgs(L.to_grayscale(s1.cuda()[:, :, :, :]))
    assert float(L.edge_loss(gi, gi * 0.5)) == pytest.approx(float(O.edge_loss(ogi, ogi * 0.5)), rel=1e-5)
    assert gi2.shape == (1, 2, 34, 42)
    assert float(L.content_loss(feats, sfeats[0] if False else {k: v * 0.9 for k, v in feats.items()}, O.CONTENT_LAYERS)) == \
        pytest.approx(float(O.content_loss(ofeats, {k: v * 0.9 for k, v in ofeats.items()}, O.CONTENT_LAYERS)), rel=3e-3)
    # single style: style_img_weight is ignored (style_transfer_losses.py:127-129)
    one = float(L.style_loss(feats, sfeats[:1], O.STYLE_LAYERS, 0.9))
    oone, _ = O.style_loss_from_targets(ofeats, O.style_targets(osf[:1], O.STYLE_LAYERS, 0.9), O.STYLE_LAYERS)
    assert one == pytest.approx(float(oone), rel=3e-3)
    two = float(L.style_loss(feats, sfeats, O.STYLE_LAYERS, 0.3))
    otwo, _ = O.style_loss_from_targets(ofeats, O.style_targets(osf, O.STYLE_LAYERS, 0.3), O.STYLE_LAYERS)
    assert two == pytest.approx(float(otwo), rel=3e-3)
    mixed = L.StyleMixer([sfeats[0]["conv2_1"], sfeats[1]["conv2_1"]], 0.3).mix()
    omixed = O.style_mix(osf[0]["conv2_1"], osf[1]["conv2_1"], 0.3)
    assert mixed.shape == omixed.shape == (1, 128, 18 + 24 // 2, 22 + 20 // 2)
    assert rel(mixed, omixed) < 2e-3
    torch.manual_seed(101)
    ca = CA.ChannelAttention(512)
    w1, w2 = ca.fc1.weight.detach(), ca.fc2.weight.detach()
    got = ca(feats["conv4_2"])
    want = O.channel_attention(ofeats["conv4_2"], w1, w2)
    assert rel(got, want) < 2e-3
    with pytest.raises(Exception):
        L.Vgg19(["conv9_9"], O.STYLE_LAYERS, "cuda")
    with pytest.raises(RuntimeError):
        L.normalize(torch.zeros(1, 4, 8, 8, device="cuda"), mean, std)     # RGBA input breaks the 3-channel broadcast


# ------------------------------------------------------------------------------------------------ full size: properties
def test_full_size_512_properties(nst, rst, oracle):
    """BASELINE config 2 (512x512, 300 steps): too slow for the CPU oracle inside a unit test, so size-independent
    properties: 320 evaluations exactly, monotone-ish decreasing loss, iterate in [0,1], Gram symmetry and
    quadratic scaling, bounded history, bit-reproducibility of the whole run."""
    O = oracle
    content, style = O.synth_image(512, 512, 0), O.synth_image(512, 512, 1)
    s, c = session(rst, O, content, [style])
    finals = []
    for rep in range(2):
        s.prepare(c, trace_capacity=400)
        assert s.run(300) == 320
        tr = s.trace()[:, 0].double().numpy()
        assert tr.shape[0] == 320 and np.isfinite(tr).all()
        assert tr[-1] < 0.7 * tr[0]
        assert (np.diff(tr) < 0).mean() > 0.9
        st = s.status()
        assert st.hist_len == 100 and st.n_iter == 320 and st.stop == 0
        x = s.result()
        assert float(x.min()) >= 0.0 and float(x.max()) <= 1.0
        finals.append((x.clone(), tr.copy()))
    assert torch.equal(finals[0][0], finals[1][0]) and np.array_equal(finals[0][1], finals[1][1])
    with torch.cuda.stream(s.stream):
        s.plan.features(c)
        g1 = s.plan.tap_gram("conv1_1")
        s.plan.features(c * 0.0 + 0.5)
        g_flat = s.plan.tap_gram("conv3_1")
    assert torch.equal(g1, g1.transpose(1, 2))
    ev = torch.linalg.eigvalsh(g1[0].double().cpu())
    assert float(ev.min()) > -1e-6 * float(ev.max())                       # Gram matrices are positive semi-definite
    assert torch.isfinite(g_flat).all()
    s.close()


# ------------------------------------------------------------------------------------------------ BASELINE configs
def test_config1_256_against_oracle(nst, rst, oracle, vgg_weights):
    """BASELINE configs[0]: 256x256 pair, the reference's own CPU-runnable case.  One optimizer.step() (20 evaluations)
    of the oracle takes seconds; loss curve and iterate are compared with the north_star tolerances."""
    O = oracle
    ws, bs = vgg_weights
    content, style = O.synth_image(256, 256, 0), O.synth_image(256, 256, 1)
    ref = O.run_oracle(ws, bs, content, [style], 0, **O.APP_WEIGHTS)
    s, c = session(rst, O, content, [style])
    s.prepare(c, trace_capacity=64)
    with torch.cuda.stream(s.stream):
        losses, _ = s.plan.eval(c)
    assert float(losses[0]) == pytest.approx(ref.losses[0][0], rel=LOSS_TOL)
    assert s.run(0) == ref.evals == 20
    tr = s.trace()[:, 0].double().numpy()
    rl = np.array([l[0] for l in ref.losses])
    assert np.all(np.abs(tr - rl) <= CURVE_TOL * np.abs(rl))
    assert O.psnr(s.result().cpu(), ref.image) >= PSNR_MIN
    s.close()


def test_config3_two_styles_channel_attention_512(nst, rst, oracle, vgg_weights):
    """BASELINE configs[2]: StyleMixer two-style blending + channel attention at 512x512.  Step-0 terms against the
    oracle (one evaluation at 512^2 is affordable), then the run's size-independent properties."""
    O = oracle
    ws, bs = vgg_weights
    content = O.synth_image(512, 512, 0)
    styles = [O.synth_image(512, 512, 1), O.synth_image(384, 640, 2)]
    s, c = session(rst, O, content, styles, mix_w=0.5)
    torch.manual_seed(101)
    s.prepare(c, channel_attention=True, trace_capacity=128)
    torch.manual_seed(101)
    co = O.ClosureOracle(ws, bs, c.cpu(), [O.to_tensor_u8(x) for x in styles], style_img_weight=0.5,
                         channel_attention_on=True, **O.APP_WEIGHTS)
    ref = co.evaluate(c.cpu(), need_grad=False)
    with torch.cuda.stream(s.stream):
        losses, _ = s.plan.eval(c)
    l = losses.cpu().tolist()
    for idx, term in ((0, "total"), (1, "content"), (2, "style"), (3, "tv"), (4, "edge")):
        assert l[idx] == pytest.approx(ref[term], rel=LOSS_TOL, abs=1e-7), term
    for i in range(5):
        assert l[5 + i] == pytest.approx(ref["gram_mse"][i], rel=LOSS_TOL)
    # mixed target of conv1_1 lives at (512 + 384 // 2) x (512 + 640 // 2) feature resolution (StyleMixer.py:31-32)
    for name in O.STYLE_LAYERS:
        assert rel(s.style_targets[name], co.style_t[name]) < GRAM_TOL, name
    # On this input the reference's own trajectory is NOT monotone: unit-step L-BFGS (no line search) overshoots at the
    # sixth evaluation (oracle, CPU fp32: 0.6717 0.6716 0.6714 0.6699 0.6298 3.90 1.40 5.01 ... 92.9 ... 2.9) and is
    # chaotic from there on.  The evaluations before the overshoot are compared with the oracle;
    # the overshoot itself must be reproduced, not "fixed".
    ref_run = O.run_oracle(ws, bs, content, styles, 10 ** 9, style_img_weight=0.5, channel_attention_on=True, max_evals=7,
                           **O.APP_WEIGHTS)
    rl = np.array([v[0] for v in ref_run.losses])
    assert s.run(60) == 80
    tr = s.trace()[:, 0].double().numpy()
    assert np.isfinite(tr).all()
    assert np.all(np.abs(tr[:5] - rl[:5]) <= CURVE_TOL * np.abs(rl[:5]))
    assert tr[5] > 3.0 * tr[4] and rl[5] > 3.0 * rl[4]
    assert abs(tr[5] - rl[5]) <= 0.1 * rl[5]
    s.close()


def test_config4_1024_high_resolution(nst, rst, oracle):
    """BASELINE configs[3]: 1024x1024 (memory / tiling stress: 2.5 GB of L-BFGS history, 1.2 GB of activations).
    The first evaluations are compared with the oracle's committed loss trace (tests/golden/config4_1024_losses.json,
    made by tests/golden/make_config4_1024.py); then properties: evaluation count, bounded iterate, symmetric Grams,
    bit-reproducible evaluation."""
    O = oracle
    content, style = O.synth_image(1024, 1024, 0), O.synth_image(1024, 1024, 1)
    s, c = session(rst, O, content, [style])
    assert s.plan.bytes() > 3e9
    s.prepare(c, trace_capacity=128)
    with torch.cuda.stream(s.stream):
        l1, g1 = s.plan.eval(c)
        l2, g2 = s.plan.eval(c)
    assert torch.equal(l1, l2) and torch.equal(g1, g2) and torch.isfinite(g1).all()
    assert s.run(40) == 60
    tr = s.trace()[:, 0].double().numpy()
    assert tr.shape[0] == 60 and np.isfinite(tr).all()
    # As at 512^2 with two styles, the reference's trajectory on this input is not monotone: unit-step L-BFGS overshoots
    # at the seventh evaluation (oracle: ... 0.3648 0.3481 14.40 3.60 18.8) and is chaotic afterwards.
    # Before the overshoot the curve must match to CURVE_TOL; the overshoot itself must be reproduced.
    with open(os.path.join(GOLDEN_DIR, "config4_1024_losses.json")) as f:
        rl = np.array(json.load(f)["total_loss"])
    assert np.all(np.abs(tr[:6] - rl[:6]) <= CURVE_TOL * np.abs(rl[:6]))
    assert tr[6] > 3.0 * tr[5] and abs(tr[6] - rl[6]) <= 0.1 * rl[6]
    x = s.result()
    assert float(x.min()) >= 0.0 and float(x.max()) <= 1.0
    with torch.cuda.stream(s.stream):
        g = s.plan.tap_gram("conv2_1")
    assert torch.equal(g, g.transpose(1, 2))
    s.close()


def test_config5_video_frames_match_single_image_runs(nst, rst, oracle):
    """BASELINE configs[4] in miniature: a stream of frames through FrameStyler (host uint8 in, host uint8 out, style
    targets hoisted) equals run_multi_style_transfer on each frame on its own - frames are independent."""
    from PIL import Image
    video = importlib.import_module("text-based-image-style-transfer_b200.video")
    O = oracle
    style = O.synth_image(64, 64, 1)
    frames = torch.stack([torch.from_numpy(O.synth_image(72, 96, 100 + k)) for k in range(3)])
    styler = video.FrameStyler(O.VGG_MEAN, O.VGG_STD, (72, 96), [O.to_tensor_u8(style).cuda()], num_steps=20, device="cuda",
                               **O.APP_WEIGHTS)
    out = video.run_sharded(frames, styler, "cuda")
    styler.close()
    assert out.shape == frames.shape and out.dtype == torch.uint8
    for k in range(3):
        with contextlib.redirect_stdout(io.StringIO()):
            single = nst.run_multi_style_transfer(torch.tensor(O.VGG_MEAN), torch.tensor(O.VGG_STD), Image.fromarray(frames[k].numpy()),
                                                  20, False, style_img1=Image.fromarray(style), device="cuda", **O.APP_WEIGHTS)
        assert np.array_equal(np.asarray(single), out[k].cpu().numpy())


# ------------------------------------------------------------------------------------------------ optional code paths
@pytest.mark.parametrize("env", [
    {"NST_DIRECT_STORES": "1", "NST_NO_SEED_FOLD": "1"},   # per-thread stores, Gram backward as separate 1x1 launches
    {"NST_NO_PDL": "1", "NST_NO_SIDE_STREAM": "1"},        # no programmatic dependent launch, single stream
    {"NST_PAIR": "0"},                                      # single-CTA convolutions (cta_group::1, two MMA issuers on one-slice layers)
    {"NST_PAIR": "2", "NST_CONV_SMS": "100", "NST_KERNEL_PRIO": "1"},   # pairs for the data gradients only, fewer SMs, launch priorities
], ids=["direct-stores-unfolded", "no-pdl-single-stream", "single-cta", "mixed-pairs-100-sms"])
def test_optional_paths_stay_parity_green(env):
    """The schedule / epilogue variants that are switched by environment variables (read when the library or a plan is
    created, hence a fresh process) must give the same answers: __graft_entry__.smoke() checks step-0 losses and
    gradient and a 20-evaluation run against the oracle."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    e = dict(os.environ)
    e.update(env)
    r = subprocess.run([sys.executable, "-c", "import __graft_entry__ as g; g.smoke()"], cwd=root, env=e,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "smoke ok" in r.stdout


# ------------------------------------------------------------------------------------------------ mask compositing
def test_mask_composite_bit_exact_against_reference_golden():
    """text/segmentation_style_transfer.py through the CUDA kernel: byte arithmetic, so bit-exact against the vectors the
    unmodified reference function wrote (tests/golden/make_golden_mask.py) - crops, even sizes, hard selection included."""
    from PIL import Image
    seg = importlib.import_module("text-based-image-style-transfer_b200.text.segmentation_style_transfer")
    g = golden("mask_composite")
    for n in range(int(g["n"])):
        out = seg.segmentation_style_transfer(Image.fromarray(g["content_%d" % n]), Image.fromarray(g["style_%d" % n]),
                                              g["mask_%d" % n], edge_smoothing=int(g["k_%d" % n]))
        assert np.array_equal(np.asarray(out), g["out_%d" % n]), n


@pytest.mark.parametrize("k", [0, 1, 3, 5, 6, 9, 11, 31, 63, 127])
def test_mask_composite_bit_exact_against_oracle(k):
    from oracle import mask_oracle as M
    seg = importlib.import_module("text-based-image-style-transfer_b200.text.segmentation_style_transfer")
    rng = np.random.default_rng(k)
    # widths that are multiples of 64 take the 64 x 64-tile kernel for k <= 9, the others the general one; the clean masks
    # (noise = 0) have tiles that are plain copies
    for H, W, noise in ((97, 131, 0.03), (33, 40, 0.03), (7, 5, 0.03), (256, 320, 0.03), (200, 192, 0.0), (3, 64, 0.03), (65, 128, 0.0),
                        (300, 448, 0.001)):
        content = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        style = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        yy, xx = np.mgrid[0:H, 0:W]
        mask = ((yy - H / 2) ** 2 + (xx - W / 3) ** 2 < (min(H, W) / 2.5) ** 2) ^ (rng.random((H, W)) < noise)
        want = M.segmentation_style_transfer(content, style, mask, k)
        got = seg.composite_tensors(torch.from_numpy(content).cuda(), torch.from_numpy(style).cuda(), torch.from_numpy(mask).cuda(), k)
        assert np.array_equal(got.cpu().numpy(), want), (k, H, W)


# ------------------------------------------------------------------------------------------------ video back end
def test_video_assemble_bit_exact_against_reference_golden():
    """app.py:800-840 (RGB -> BGR + cross-dissolve) through the CUDA kernel against the frame lists the reference's own
    statements wrote with cv2 (tests/golden/make_golden_video.py)."""
    video = importlib.import_module("text-based-image-style-transfer_b200.video")
    g = golden("video_assemble")
    for n in range(int(g["n"])):
        got = video.assemble_frames(torch.from_numpy(g["frames_%d" % n]).cuda(), int(g["k_%d" % n]))
        assert np.array_equal(got.cpu().numpy(), g["final_%d" % n]), n


@pytest.mark.parametrize("k", [0, 1, 2, 3, 4, 5, 6, 15])
def test_video_assemble_bit_exact_against_oracle(k):
    from oracle import video_oracle as V
    video = importlib.import_module("text-based-image-style-transfer_b200.video")
    rng = np.random.default_rng(k)
    # frame sizes with H * W a multiple of 16 take the 48-bytes-per-thread kernel, the others the per-pixel one
    for F, H, W in ((3, 64, 48), (2, 17, 23), (1, 32, 32), (5, 120, 160), (2, 1, 16), (4, 3, 5)):
        frames = rng.integers(0, 256, (F, H, W, 3), dtype=np.uint8)
        if F > 1:  # every byte pair the rounding could trip on: all 65 536 (a, b) combinations somewhere in the first two frames
            a, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
            cnt = min(H * W * 3, 65536)
            frames[0].reshape(-1)[:cnt] = a.reshape(-1)[:cnt]
            frames[1].reshape(-1)[:cnt] = b.reshape(-1)[:cnt]
        want = np.stack(V.assemble_frames(list(frames), k), 0)
        got = video.assemble_frames(torch.from_numpy(frames).cuda(), k)
        assert got.shape == want.shape and np.array_equal(got.cpu().numpy(), want), (k, F, H, W)


def test_video_assemble_all_byte_pairs():
    """all 65 536 (prev, frame) byte pairs for every slider value, vector kernel"""
    from oracle import video_oracle as V
    video = importlib.import_module("text-based-image-style-transfer_b200.video")
    a, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing="ij")
    frames = np.stack([np.repeat(a.reshape(-1), 3).reshape(128, 512, 3), np.repeat(b.reshape(-1), 3).reshape(128, 512, 3)], 0)
    for k in range(1, 16):
        want = np.stack(V.assemble_frames(list(frames), k), 0)
        got = video.assemble_frames(torch.from_numpy(frames).cuda(), k)
        assert np.array_equal(got.cpu().numpy(), want), k


# ------------------------------------------------------------------------------------------------ depth-aware / multi-plane variant
def test_mip_split_merge_bit_exact_against_reference_golden(nst):
    """components/style_transfer_depth/util.py byte functions through the CUDA kernels against what the unmodified reference
    functions returned (incl. depth values exactly on bin edges, a flat depth map = nan, fp32 depth, a grayscale image)."""
    U = importlib.import_module("text-based-image-style-transfer_b200.components.style_transfer_depth.util")
    from PIL import Image
    g = golden("depth_mip")
    for k in range(int(g["b_count"])):
        img, depth, n = g["b_img_%d" % k], g["b_depth_%d" % k], int(g["b_n_%d" % k])
        planes = np.stack([np.asarray(p) for p in U.generate_mip_layers(Image.fromarray(img), depth, n)], 0)
        assert np.array_equal(planes, g["b_planes_%d" % k]), k
        merged = U.reconstruct_mip_image([Image.fromarray(s) for s in g["b_styl_%d" % k]], depth, n)
        assert np.array_equal(np.asarray(merged), g["b_merged_%d" % k]), k
        one = U.mask_image_depth(Image.fromarray(img), depth, U.create_bins(n)[n - 1])
        assert np.array_equal(np.asarray(one), g["b_planes_%d" % k][n - 1]), k


@pytest.mark.parametrize("n", [2, 3, 7, 10, 16])
def test_mip_split_merge_bit_exact_against_oracle(nst, n):
    from oracle import depth_oracle as D
    U = importlib.import_module("text-based-image-style-transfer_b200.components.style_transfer_depth.util")
    rng = np.random.default_rng(n)
    for H, W, kind in ((64, 48, "u8"), (33, 31, "u8"), (128, 96, "edges"), (17, 4, "f64"), (40, 40, "edges")):
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        if kind == "u8":
            depth = rng.integers(0, 256, (H, W), dtype=np.uint8)
        elif kind == "edges":
            depth = (rng.integers(0, 2 * n + 1, (H, W)) * 3 + 5).astype(np.uint8)    # range 6n: every bin edge is hit exactly
        else:
            depth = rng.random((H, W)) * 5 - 2
        bins = D.create_bins(n)
        planes = U.split_planes(torch.from_numpy(img).cuda(), depth, bins)
        want = np.stack(D.generate_mip_layers(img, depth, n), 0)
        assert np.array_equal(planes.cpu().numpy(), want), (n, H, W, kind)
        if depth.dtype == np.uint8:   # the depth map handed over as a device tensor, with and without its range
            d_dev = torch.from_numpy(depth).cuda()
            assert torch.equal(U.split_planes(torch.from_numpy(img).cuda(), d_dev, bins), planes)
            assert torch.equal(U.split_planes(torch.from_numpy(img).cuda(), d_dev, bins, depth_range=(int(depth.min()), int(depth.max()))), planes)
        styl = rng.integers(0, 256, (n, H, W, 3), dtype=np.uint8)
        merged = U.merge_planes(torch.from_numpy(styl).cuda(), depth, bins)
        assert np.array_equal(merged.cpu().numpy(), D.reconstruct_mip_image(list(styl), depth, n)), (n, H, W, kind)


def test_style_mip_against_reference_golden(nst):
    """DepthStyle.style_MIP (n loops with per-plane style weight + split / merge) against the run of the unmodified reference:
    planes bit-exact, per-plane loss curves within 1e-2, stylised planes and merged image >= 40 dB."""
    from PIL import Image
    mod = importlib.import_module("text-based-image-style-transfer_b200.components.style_transfer_depth.style_transfer_depth")
    g = golden("depth_mip")
    n, steps = int(g["mip_n"]), int(g["mip_num_steps"])
    depth = g["mip_depth"]
    ds = mod.DepthStyle("cuda", depth_pipeline=lambda image: {"depth": Image.fromarray(depth)})
    ds.style_model.num_steps = steps
    ds.style_model.print_iter = 0
    try:
        planes = ds.depth_split(Image.fromarray(g["mip_content"]), n)
        assert np.array_equal(np.stack([np.asarray(p) for p in planes], 0), g["mip_planes"])
        with contextlib.redirect_stdout(io.StringIO()):
            final, stylized = ds.style_MIP(Image.fromarray(g["mip_content"]), Image.fromarray(g["mip_style"]), n)
        traces = ds.style_model.traces[-n:]
    finally:
        ds.close()

    def psnr_u8(a, b):
        mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
        return 99.0 if mse == 0 else 10 * np.log10(255.0 ** 2 / mse)

    for i in range(n):
        ours, ref = traces[i][:, 0].double().numpy(), g["mip_losses"][i]
        assert len(ours) == len(ref) == 20 * (steps // 20 + 1)
        assert abs(ours[0] - ref[0]) <= LOSS_TOL * abs(ref[0]), i
        assert np.abs(ours - ref).max() <= CURVE_TOL * np.abs(ref).max(), (i, np.abs(ours - ref).max() / np.abs(ref).max())
        assert psnr_u8(np.asarray(stylized[i]), g["mip_stylized"][i]) >= PSNR_MIN, i
    # pixels exactly on an inner bin edge are the wrapped sum of two planes (util.py:30,86): a one-level difference there can
    # show up as 255 - the merge itself is checked bit for bit above, so they are left out of the PSNR
    from oracle import depth_oracle as D
    shared = sum((np.stack(D.generate_mip_layers(np.ones_like(g["mip_content"]), depth, n), 0)[:, :, :, 0] > 0).astype(int)) > 1
    assert psnr_u8(np.asarray(final)[~shared], g["mip_final"][~shared]) >= PSNR_MIN


def test_stylise_frames_against_oracle(nst, oracle, vgg_weights):
    """video.stylise_frames = per-frame loop + frame list: real frames >= 40 dB against the oracle's loop, and the whole list
    bit-exact against the video oracle applied to OUR stylised frames (the cross-dissolve is byte arithmetic)."""
    from oracle import video_oracle as V
    O = oracle
    ws, bs = vgg_weights
    video = importlib.import_module("text-based-image-style-transfer_b200.video")
    H, W, steps, n = 48, 64, 0, 2
    frames = np.stack([O.synth_image(H, W, 40 + k) for k in range(3)], 0)
    style = O.synth_image(40, 56, 9)
    styler = video.FrameStyler(O.VGG_MEAN, O.VGG_STD, (H, W), [O.to_tensor_u8(style).cuda()], num_steps=steps, device="cuda",
                               **O.APP_WEIGHTS)
    try:
        out = video.stylise_frames(torch.from_numpy(frames), styler, n).cpu().numpy()
        own = [np.ascontiguousarray(out[k * (n + 1)][:, :, ::-1]) for k in range(3)]
    finally:
        styler.close()
    assert out.shape == (2 * (n + 1) + 1, H, W, 3)
    assert np.array_equal(out, np.stack(V.assemble_frames(own, n), 0))
    for k in range(3):
        ref = O.to_u8(O.run_oracle(ws, bs, frames[k], [style], steps, **O.APP_WEIGHTS).image)
        mse = np.mean((own[k].astype(np.float64) - ref.astype(np.float64)) ** 2)
        assert mse == 0 or 10 * np.log10(255.0 ** 2 / mse) >= PSNR_MIN, k


def test_concurrent_frames_are_bit_identical_to_one_at_a_time(nst, oracle):
    """FrameStyler(concurrent=3).process_block (nst_run_frames_host: three plans / streams stepped in lock step) returns, for
    every frame, exactly the bytes of the one-frame call - with a block that does not divide the frame count."""
    O = oracle
    video = importlib.import_module("text-based-image-style-transfer_b200.video")
    H, W = 40, 56
    frames = torch.from_numpy(np.stack([O.synth_image(H, W, 60 + k) for k in range(5)], 0))
    style = [O.to_tensor_u8(O.synth_image(48, 40, 8)).cuda()]
    for ca in (False, True):
        one = video.FrameStyler(O.VGG_MEAN, O.VGG_STD, (H, W), style, num_steps=20, channel_attention=ca, device="cuda", **O.APP_WEIGHTS)
        many = video.FrameStyler(O.VGG_MEAN, O.VGG_STD, (H, W), style, num_steps=20, channel_attention=ca, device="cuda", concurrent=3,
                                 **O.APP_WEIGHTS)
        try:
            want = torch.stack([one(k, frames[k]) for k in range(5)], 0)
            got = many.process_block(frames)
            assert torch.equal(got, want), ca
            assert not torch.equal(got[0], frames[0])
            assert torch.equal(video.run_sharded(frames, many, "cuda").cpu(), want)
        finally:
            one.close()
            many.close()


# ------------------------------------------------------------------------------------------------ run-level parity at full size
LARGE_CASES = ["nat512", "syn512", "boat512", "nat512_mix_ca", "dog1024", "dog720p", "nat1024", "nat720p"]


@pytest.mark.parametrize("name", LARGE_CASES)
def test_full_run_against_reference_golden_large(nst, rst, oracle, name):
    """Run-level parity on the configurations the performance is quoted on (BASELINE configs[1]..[4]), against goldens written
    by the UNMODIFIED reference (tests/golden/make_golden_large.py -> /root/reference/multi_style_transfer/
    run_style_transfer.py:27-159 on CPU): the loss of EVERY closure evaluation within 1e-2 relative, the final float image at
    >= 40 dB - north_star's run-level tolerances - and the evaluation count.  Each golden carries the reference's own
    self-noise (the same run with another CPU thread count); our deviation is printed beside it."""
    path = os.path.join(GOLDEN_DIR, "large_%s.npz" % name)
    if not os.path.exists(path):
        pytest.skip("golden %s not generated" % path)
    O = oracle
    g = np.load(path)
    content = g["content_u8"]
    styles = [g["style%d_u8" % i] for i in range(2) if "style%d_u8" % i in g.files]
    steps, wgt, ca = int(g["num_steps"]), float(g["style_img_weight"]), bool(g["channel_attention"])
    ref = g["loss_trace"]
    s, c = session(rst, O, content, styles, mix_w=wgt)
    s.prepare(c, channel_attention=ca, trace_capacity=len(ref) + 32)
    n = s.run(steps)
    assert n == int(g["n_evals"]) == len(ref)
    tr = s.trace()[:, 0].double().numpy()
    dev = np.abs(tr - ref) / np.abs(ref)
    x = s.result().cpu()
    p = O.psnr(x, torch.from_numpy(g["x_final"].astype(np.float32))) if "x_final" in g.files else float("nan")
    print("\n[large golden %s] %d evals, loss %.5f -> %.5f (reference %.5f -> %.5f); max loss-curve deviation %.2e at eval %d "
          "(reference self-noise %.2e); final image %.1f dB (reference self-noise %.1f dB)"
          % (name, n, tr[0], tr[-1], ref[0], ref[-1], dev.max(), int(dev.argmax()), float(g["self_loss_dev"]), p, float(g["self_psnr"])))
    st = s.status()
    assert st.closure_calls == n and st.stop == 0
    stable = int(g["stable_prefix"]) if "stable_prefix" in g.files else len(ref)
    if stable == len(ref):
        # the reference agrees with itself over the whole run: north_star's run-level tolerances apply to all of it
        assert dev.max() <= CURVE_TOL, (name, float(dev.max()), int(dev.argmax()))
        assert p >= PSNR_MIN, (name, p)
        if "final_u8" in g.files:
            d8 = np.abs(O.to_u8(x).astype(int) - g["final_u8"].astype(int))
            assert d8.mean() < 1.0, (name, float(d8.mean()))
    else:
        # On this input the REFERENCE parts from itself after `stable` evaluations (two CPU thread counts more than 1e-3 apart:
        # unit-step L-BFGS overshoots and the trajectory is chaotic from there).  The evaluations before that are
        # held to the tolerance; the overshoot must be reproduced, not avoided: the curve leaves the band where the reference's
        # does (within a factor of two of its loss at the first evaluation after the stable prefix).
        # Inside the stable prefix the trajectory is already amplifying perturbations (the two reference runs start 1e-7 apart
        # and are 1e-3 apart at its end); this path starts ~5e-6 from the reference (fp16 / bf16 trunk), i.e. ~50x the
        # reference's own starting distance, and is carried by the same dynamics.  The tolerance per evaluation is therefore
        # north_star's 1e-2, widened to 100x the reference's running self-deviation where that is larger.
        sd = np.maximum.accumulate(g["self_dev_trace"][:stable])
        tol = np.maximum(CURVE_TOL, 100.0 * sd)
        print("  reference is stable against itself for the first %d evaluations only (self deviation %.1e at evaluation %d)"
              % (stable, float(g["self_dev_trace"][stable]), stable))
        print("  deviation / tolerance on that prefix: " + " ".join("%.0e/%.0e" % (a, b) for a, b in zip(dev[:stable], tol)))
        assert stable >= 3
        assert np.all(dev[:stable] <= tol), (name, dev[:stable], tol)
        assert dev[:4].max() <= 2e-4          # before the first overshoot every run is the same run (dog1024: 1.0e-4, the step-0 accuracy of the fp16 trunk)
        assert 0.5 * ref[stable] <= tr[stable] <= 2.0 * ref[stable], (tr[stable], ref[stable])
        assert np.isfinite(tr).all() and float(x.min()) >= 0.0 and float(x.max()) <= 1.0
    s.close()


# ------------------------------------------------------------------------------------------------ VGG-dominant regime
@pytest.mark.parametrize("weights", [dict(w_style=5e5, w_content=1.0, w_tv=0.0, w_edge=0.0),
                                     dict(w_style=5e7, w_content=1e2, w_tv=2e1, w_edge=2e1)],
                         ids=["pixel-terms-off", "vgg-terms-x100"])
def test_vgg_dominated_gradient_and_short_trajectory(nst, rst, oracle, vgg_weights, weights):
    """With random-init VGG weights and app.py's loss weights the fp32 TV / edge terms carry ~99 % of the gradient, which hides
    the reduced-precision VGG path.  Here the VGG terms dominate (pixel terms off, or style / content weights x100 - the
    regime ImageNet weights put the reference in): teacher-forced on the oracle's own trajectory.

    What the numbers mean (tools/grad_error_probe.py, profiles/r02_grad_error_probe.log): the style part of the gradient is
    accurate to 3e-3 .. 7e-3 of itself; the content part only to ~9e-2 while x is still close to the content image, because
    its seed F(x) - F(content) is a difference of two fp16-path feature maps whose rounding noise does not cancel.  Relative
    to the CURRENT gradient the error therefore grows as the run approaches a stationary point (|g| falls 30x within 20
    evaluations with the pixel terms off) - as it does for any finite precision: the fp32 reference against fp64 jumps to
    1e-2 at the same iterates whenever a ReLU / max-pool / |.| branch flips.  The bound is therefore stated against the
    gradient norm of the first evaluation, 2e-2, and 5e-3 relative at the first evaluation itself."""
    O = oracle
    ws, bs = vgg_weights
    content, style = O.synth_image(96, 96, 0), O.synth_image(96, 96, 1)
    ref = O.run_oracle(ws, bs, content, [style], 0, keep_iterates=True, **weights)
    s, c = session(rst, O, content, [style], weights)
    s.prepare(c, trace_capacity=64)
    g0 = float(ref.grads[0].double().norm())
    rows = []
    for k in (0, 3, 7, 12, 19):
        xk = ref.iterates[k].cuda()
        with torch.cuda.stream(s.stream):
            losses, grad = s.plan.eval(xk)
        assert float(losses[0]) == pytest.approx(ref.losses[k][0], rel=LOSS_TOL), k
        err = float((grad.double().cpu() - ref.grads[k].double()).norm())
        rows.append((k, err / float(ref.grads[k].double().norm()), err / g0))
    print("\n[vgg-dominated %s] eval: error / |g_k|, error / |g_0|  " % (weights,) + "  ".join("%d: %.1e, %.1e" % r for r in rows))
    assert rows[0][1] < 5e-3, rows
    assert max(r[2] for r in rows) < 2e-2, rows
    assert s.run(0) == ref.evals == 20
    tr = s.trace()[:, 0].double().numpy()
    rl = np.array([l[0] for l in ref.losses])
    print("  loss, this path: " + " ".join("%.4g" % v for v in tr) + "\n  loss, reference: " + " ".join("%.4g" % v for v in rl))
    assert np.isfinite(tr).all() and tr[-1] < tr[0]
    if weights["w_tv"] > 0:
        # pixel terms present (even at 1 % of the gradient): the curvature pairs are resolved, the run tracks the reference
        assert np.all(np.abs(tr[:8] - rl[:8]) <= CURVE_TOL * np.abs(rl[:8])), np.abs(tr - rl) / rl
    else:
        # VGG terms ALONE: the first update moves every pixel by ~1 / (number of pixels) (lbfgs.py:454-457: t = 1 / |g|_1), far
        # below what fp16 activations resolve, so the first curvature pair y = g(x1) - g(x0) is rounding noise here while the
        # fp32 reference resolves it: the two runs part at the third evaluation (Precision).  The reference is
        # itself chaotic in this regime (two CPU thread counts are 21 dB apart after 20 evaluations), so a curve
        # tolerance cannot be asked; what must hold is the first step and a descending, finite run.
        assert np.all(np.abs(tr[:2] - rl[:2]) <= 1e-4 * np.abs(rl[:2]))
    s.close()


# ------------------------------------------------------------------------------------------------ tutorial-style compat surface
def test_compat_run_style_transfer_and_loss_modules(nst, rst, oracle, vgg_weights):
    """north_star's run_style_transfer(cnn, norm_mean, norm_std, content_img, style_img, input_img, num_steps, style_weight,
    content_weight) + ContentLoss / StyleLoss modules: the same device loop with the pixel terms off, started from input_img."""
    O = oracle
    ws, bs = vgg_weights
    content, style = O.synth_image(64, 64, 0), O.synth_image(64, 64, 1)
    c, st = O.to_tensor_u8(content).cuda(), O.to_tensor_u8(style).cuda()
    x0 = noisy(c, seed=5)
    wts = dict(w_style=1e6, w_content=1.0, w_tv=0.0, w_edge=0.0)
    # oracle: the reference closure with those weights, started from x0
    co = O.ClosureOracle(ws, bs, c.cpu(), [st.cpu()], **wts)
    ref0 = co.evaluate(x0)
    inp = x0.clone().cuda()
    out = rst.run_style_transfer(None, O.VGG_MEAN, O.VGG_STD, c, st, inp, num_steps=0, style_weight=1e6, content_weight=1)
    assert out.shape == c.shape and float(out.min()) >= 0.0 and float(out.max()) <= 1.0
    assert torch.equal(out, inp)                                            # the tutorial's parameter is updated in place
    # a cnn given as a module: torchvision-style features built from the same weights
    import torchvision
    feats = torchvision.models.vgg19(weights=None).features
    convs = [m for m in feats if isinstance(m, torch.nn.Conv2d)]
    with torch.no_grad():
        for m, w, b in zip(convs, ws, bs):
            m.weight.copy_(w)
            m.bias.copy_(b)
    out2 = rst.run_style_transfer(feats, O.VGG_MEAN, O.VGG_STD, c, st, x0.clone().cuda(), num_steps=0, style_weight=1e6, content_weight=1)
    assert torch.equal(out, out2)
    # the loss after 20 evaluations went down from the oracle's step-0 value
    s, _ = session(rst, O, content, [style], wts)
    s.prepare(c, x0.cuda())
    with torch.cuda.stream(s.stream):
        l0, _ = s.plan.eval(x0.cuda())
        l1, _ = s.plan.eval(out)
    assert float(l0[0]) == pytest.approx(ref0["total"], rel=LOSS_TOL)
    assert float(l1[0]) < float(l0[0])
    s.close()
    # modules: transparent layers that record the per-layer terms
    L = importlib.import_module("text-based-image-style-transfer_b200.multi_style_transfer.style_transfer_losses")
    f_in = torch.randn(1, 64, 24, 20, generator=torch.Generator().manual_seed(1)).cuda()
    f_t = torch.randn(1, 64, 24, 20, generator=torch.Generator().manual_seed(2)).cuda()
    cl, sl = rst.ContentLoss(f_t), rst.StyleLoss(f_t)
    assert cl(f_in) is f_in and sl(f_in) is f_in
    assert float(cl.loss) == pytest.approx(float(torch.nn.functional.mse_loss(f_in, f_t)), rel=1e-5)
    want = torch.nn.functional.mse_loss(O.gram_matrix(f_in.cpu()), O.gram_matrix(f_t.cpu()))
    assert float(sl.loss) == pytest.approx(float(want), rel=5e-3)


# ------------------------------------------------------------------------------------------------ apply_video_process mirror
def test_apply_video_process_mirror(nst, rst, oracle, tmp_path):
    """app.py:742-864 through the mirror: a tiny synthetic clip -> JPEG round trip -> FrameStyler -> device assembly; the final
    BGR frame list equals (a) run_multi_style_transfer on every JPEG-round-tripped frame and (b) the cv2 assembly oracle."""
    cv2 = pytest.importorskip("cv2")
    from PIL import Image
    from oracle import video_oracle as V
    A = importlib.import_module("text-based-image-style-transfer_b200.app_video")
    O = oracle
    H, W, F = 48, 64, 3
    path = str(tmp_path / "in.mp4")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 10.0, (W, H))
    assert wr.isOpened()
    for k in range(F):
        wr.write(np.ascontiguousarray(O.synth_image(H, W, 200 + k)[:, :, ::-1]))
    wr.release()
    style = Image.fromarray(O.synth_image(40, 56, 1))
    got = []
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        out_path = A.apply_video_process(path, ["Style Transfer"], 0.5, 2, style, device="cuda", num_steps=20,
                                         output_video_filepath=str(tmp_path / "out.mp4"), frame_sink=got.append)
    assert out_path and os.path.exists(out_path) and os.path.getsize(out_path) > 0
    assert "Finished processing frame: 3 out of" in buf.getvalue() and "Current output video temporarily saved at" in buf.getvalue()
    final = got[0]
    assert final.shape == ((F - 1) * 3 + 1, H, W, 3)
    frames, fps, _ = A.read_frames(path, jpeg=True)
    singles = []
    for f in frames:
        with contextlib.redirect_stdout(io.StringIO()):
            singles.append(np.asarray(nst.run_multi_style_transfer(torch.tensor(O.VGG_MEAN), torch.tensor(O.VGG_STD), Image.fromarray(f), 20,
                                                                   False, style_img1=style, device="cuda", **O.APP_WEIGHTS)))
    want = np.stack(V.assemble_frames(singles, 2), 0)
    assert np.array_equal(final, want)
    cap = cv2.VideoCapture(out_path)
    assert int(cap.get(cv2.CAP_PROP_FRAME_COUNT)) == final.shape[0]
    assert cap.get(cv2.CAP_PROP_FPS) == pytest.approx(A.output_fps(fps, 2, 0.5), abs=0.5)
    cap.release()
    # effects outside the hot path raise instead of being skipped
    with pytest.raises(nst.NstError):
        A.apply_video_process(path, ["Pixel Art", "Style Transfer"], input_style=style, device="cuda")


# ------------------------------------------------------------------------------------------------ several content layers
def test_content_loss_over_layers_of_different_size(nst, rst, oracle, vgg_weights):
    """content_loss is the MEAN over the content layers of the per-layer MSEs (style_transfer_losses.py:53-65).  With layers
    of different resolution every layer's partial sums must carry its own 1 / numel - r01 normalised all of them by the
    first layer's size.  Loss and gradient against the oracle with content on conv3_2 (1/4 resolution, 256 channels) and
    conv4_2 (1/8, 512), one of them weighted up so that a wrong normalisation cannot hide."""
    O = oracle
    ws, bs = vgg_weights
    content, style = O.synth_image(64, 80, 0), O.synth_image(64, 80, 1)
    layers = ["conv3_2", "conv4_2"]
    wts = dict(w_style=5e5, w_content=50.0, w_tv=2e1, w_edge=2e1)
    c = O.to_tensor_u8(content).cuda()
    s = rst.StyleTransferSession(O.VGG_MEAN, O.VGG_STD, c.shape[2:], [O.to_tensor_u8(style).cuda()], wts["w_style"], wts["w_content"],
                                 wts["w_tv"], wts["w_edge"], 0.5, "cuda", content_layers=layers)
    s.prepare(c)
    co = O.ClosureOracle(ws, bs, c.cpu(), [O.to_tensor_u8(style)], content_layers=layers, **wts)
    x = noisy(c, seed=9)
    ref = co.evaluate(x)
    with torch.cuda.stream(s.stream):
        losses, grad = s.plan.eval(x.cuda())
    l = losses.cpu().tolist()
    assert l[1] == pytest.approx(ref["content"], rel=LOSS_TOL)
    assert l[0] == pytest.approx(ref["total"], rel=LOSS_TOL)
    assert rel(grad, ref["grad"]) < 2e-3
    s.close()
