"""Pins the CPU oracle (oracle/nst_oracle.py) against outputs of the reference itself: tests/golden/*.npz were
written by tests/golden/make_golden.py, which runs /root/reference/multi_style_transfer unmodified.  The
reference ships no tests or golden vectors of its own."""
import numpy as np
import pytest
import torch

from conftest import golden
from oracle import nst_oracle as O

CASES = {
    "single_64": ((64, 64, 0), [(64, 64, 1)], 50, 0.5, False),
    "odd_50x38": ((50, 38, 3), [(44, 60, 4)], 20, 0.5, False),
    "mix_ca_48x40": ((48, 40, 5), [(40, 56, 6), (64, 48, 7)], 20, 0.3, True),
}


def _inputs(name):
    cspec, sspecs, steps, wgt, ca = CASES[name]
    return O.synth_image(*cspec), [O.synth_image(*s) for s in sspecs], steps, wgt, ca


@pytest.mark.parametrize("name", list(CASES))
def test_step0_terms_gradient_and_grams_match_reference(name, vgg_weights):
    ws, bs = vgg_weights
    g = golden(name)
    content, styles, steps, wgt, ca = _inputs(name)
    torch.manual_seed(101)
    co = O.ClosureOracle(ws, bs, O.to_tensor_u8(content), [O.to_tensor_u8(s) for s in styles], style_img_weight=wgt,
                         channel_attention_on=ca, **O.APP_WEIGHTS)
    r = co.evaluate(O.to_tensor_u8(content))
    for term in ("total", "content", "style", "tv", "edge"):
        assert r[term] == pytest.approx(float(g["s0_" + term]), rel=2e-5, abs=1e-9), term
    ref_grad = torch.from_numpy(g["s0_grad"])
    assert float((r["grad"] - ref_grad).norm() / ref_grad.norm()) < 1e-5
    for i, layer in enumerate(O.STYLE_LAYERS):
        assert r["gram_mse"][i] == pytest.approx(float(g["s0_gram_mse_" + layer]), rel=1e-4)
        blk = g["s0_gram_" + layer]
        ours = r["grams"][layer][0][:blk.shape[0], :blk.shape[1]].numpy()
        assert np.abs(ours - blk).max() <= 1e-5 * np.abs(blk).max()
        assert float(r["grams"][layer].norm()) == pytest.approx(float(g["s0_gram_fro_" + layer]), rel=1e-5)
        assert float(co.style_t[layer].norm()) == pytest.approx(float(g["s0_target_fro_" + layer]), rel=1e-5)
    sample = co.content_t["conv4_2"][0, :8, :4, :4].numpy()
    assert np.abs(sample - g["s0_content_feat_sample"]).max() <= 1e-5 * (np.abs(g["s0_content_feat_sample"]).max() + 1e-6)


@pytest.mark.parametrize("name", list(CASES))
def test_trajectory_matches_reference(name, vgg_weights):
    """Full loop: evaluation count exactly, loss curve and final image within the reference's own self-noise
    (the golden run used 1 thread, this one uses the default thread count: SURVEY appendix A.3)."""
    ws, bs = vgg_weights
    g = golden(name)
    content, styles, steps, wgt, ca = _inputs(name)
    res = O.run_oracle(ws, bs, content, styles, steps, style_img_weight=wgt, channel_attention_on=ca, keep_iterates=True,
                       **O.APP_WEIGHTS)
    assert res.evals == int(g["n_evals"]) == 20 * (steps // 20 + 1)
    ours = np.array([l[0] for l in res.losses])
    ref = g["loss_trace"]
    assert abs(ours[0] - ref[0]) <= 1e-5 * abs(ref[0])
    assert np.abs(ours - ref).max() <= 1e-2 * np.abs(ref).max()
    assert O.psnr(res.iterates[5], torch.from_numpy(g["x_eval5"])) > 80.0
    assert O.psnr(res.image, torch.from_numpy(g["x_last"])) > 40.0
    # uint8 output: truncation like ToPILImage; trajectories this close differ by at most a few grey levels
    diff = np.abs(O.to_u8(res.image).astype(int) - g["final_u8"].astype(int))
    assert diff.max() <= 8 and diff.mean() < 0.5


def test_lbfgs_restatement_equals_torch_optim():
    """LbfgsOracle is torch.optim.LBFGS (no line search) statement by statement: identical iterates on CPU."""
    torch.manual_seed(0)
    n = 30
    A = torch.randn(n, n)
    A = A @ A.T / n + 0.3 * torch.eye(n)
    b = torch.randn(n)

    def f(x):
        return 0.5 * x @ (A @ x) - b @ x + 0.05 * (x[1:] - x[:-1]).abs().sum()

    x0 = torch.rand(n)
    p = torch.nn.Parameter(x0.clone())
    opt = torch.optim.LBFGS([p])

    def closure_t():
        opt.zero_grad()
        with torch.no_grad():
            p.clamp_(0, 1)
        loss = f(p)
        loss.backward()
        return loss

    o = O.LbfgsOracle(x0.clone())

    def closure_o():
        o.x.clamp_(0, 1)
        x = o.x.detach().clone().requires_grad_(True)
        loss = f(x)
        loss.backward()
        return float(loss.detach()), x.grad

    for _ in range(7):  # 140 evaluations: exercises the history pop at 100 pairs
        opt.step(closure_t)
        o.step(closure_o)
        assert torch.equal(p.detach(), o.x)
    st = opt.state[p]
    assert st["n_iter"] == o.n_iter and st["func_evals"] == o.func_evals and len(st["old_dirs"]) == len(o.old_dirs)


def test_num_steps_counts_closure_evaluations(vgg_weights):
    """num_steps=N runs 20 * (N // 20 + 1) evaluations (run_style_transfer.py:99-100,143,151)."""
    ws, bs = vgg_weights
    c, s = O.synth_image(16, 16, 0), O.synth_image(16, 16, 1)
    for n, want in ((0, 20), (19, 20), (20, 40), (50, 60)):
        r = O.run_oracle(ws, bs, c, [s], n, w_style=0.0, w_content=0.0, w_tv=20.0, w_edge=20.0)
        assert r.evals == want


def test_style_mixer_midpoint_precedence():
    """(64,64)+(64,64) -> (96,96); (64,64)+(48,80) -> (88,104)   (StyleMixer.py:31-32)."""
    a = torch.zeros(1, 2, 64, 64)
    assert O.style_mix(a, torch.zeros(1, 2, 64, 64), 0.5).shape[2:] == (96, 96)
    assert O.style_mix(a, torch.zeros(1, 2, 48, 80), 0.5).shape[2:] == (88, 104)


def test_taps_are_pre_relu(vgg_weights):
    ws, bs = vgg_weights
    f = O.VggFeatures(ws, bs, O.STYLE_LAYERS + O.CONTENT_LAYERS)(O.normalize(O.to_tensor_u8(O.synth_image(32, 32, 0)), O.VGG_MEAN, O.VGG_STD))
    assert list(f) == ["conv1_1", "conv2_1", "conv3_1", "conv4_1", "conv4_2", "conv5_1"]
    assert all(float(v.min()) < 0 for v in f.values())
    assert f["conv5_1"].shape == (1, 512, 2, 2)
    with pytest.raises(Exception):
        O.VggFeatures(ws, bs, ["conv9_9"])


def test_to_u8_truncates():
    assert O.to_u8(torch.full((1, 3, 1, 1), 0.999))[0, 0, 0] == 254
