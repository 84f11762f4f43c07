"""Host-side logic of the apply_video_process mirror (text-based-image-style-transfer_b200/app_video.py): CPU only.
The JPEG round trip is pinned to the reference's own procedure - cv2.imwrite to a .jpg file, then PIL Image.open
(app.py:790-791, 153-155)."""
import importlib
import inspect
import os

import numpy as np
import pytest

A = importlib.import_module("text-based-image-style-transfer_b200.app_video")
pkg = importlib.import_module("text-based-image-style-transfer_b200")


def test_signature_matches_reference():
    """app.py:742-746: the 27 positional parameters in order, with the reference's defaults."""
    params = list(inspect.signature(A.apply_video_process).parameters.values())
    positional = [p.name for p in params if p.kind == inspect.Parameter.POSITIONAL_OR_KEYWORD]
    assert positional == ["video_filepath", "checkbox_values", "slowmo_slider_input", "interpolation_slider_input", "input_style",
                          "text_checkbox_values", "text_box", "text_location_box", "text_style_masking_box",
                          "text_masked_transfer_edge_smoothing", "text_emoji_blur_strength", "text_emoji_step_size",
                          "text_masked_style_strength", "p_size_slider", "p_checkbox", "p_colour_dropbox", "p_colour_interpolate",
                          "p_edge_slider", "p_select_im", "p_in", "p_in_slid", "style_image_weight", "style_image1", "style_image2",
                          "color_palette_style", "d_check_box", "depth_mip_n"]
    d = {p.name: p.default for p in params}
    assert d["p_size_slider"] == 0.4 and d["p_edge_slider"] == 50 and d["depth_mip_n"] == 2 and d["p_in_slid"] == 10
    assert A.list_of_effects[3] == "Style Transfer" and A.list_of_effects[4] == "Style Mixing"
    assert (A.NUM_STEPS, A.W_STYLE, A.W_CONTENT, A.W_TV, A.W_EDGE) == (400, 5e5, 1, 2e1, 2e1)   # app.py:380-385


def test_jpeg_round_trip_equals_imwrite_then_pil_open(tmp_path):
    cv2 = pytest.importorskip("cv2")
    from PIL import Image
    from oracle import nst_oracle as O
    for seed, (h, w) in enumerate([(48, 64), (37, 53), (120, 90)]):
        bgr = np.ascontiguousarray(O.synth_image(h, w, 300 + seed)[:, :, ::-1])
        f = str(tmp_path / ("frame_%d.jpg" % seed))
        cv2.imwrite(f, bgr)                                   # app.py:791
        want = np.asarray(Image.open(f))                      # app.py:153-155
        got = A.jpeg_round_trip(bgr)
        assert got.dtype == np.uint8 and np.array_equal(got, want)
        assert not np.array_equal(got, bgr[:, :, ::-1])       # the round trip is lossy: it has to be reproduced, not skipped


def test_read_frames_and_fps(tmp_path):
    cv2 = pytest.importorskip("cv2")
    path = str(tmp_path / "clip.mp4")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"mp4v"), 12.0, (32, 24))
    assert wr.isOpened()
    rng = np.random.default_rng(0)
    for _ in range(4):
        wr.write(rng.integers(0, 256, (24, 32, 3), dtype=np.uint8))
    wr.release()
    frames, fps, n = A.read_frames(path)
    assert len(frames) == 4 and n == 4 and frames[0].shape == (24, 32, 3) and fps == pytest.approx(12.0)
    assert A.output_fps(12.0, 0, None) == 12.0                 # app.py:851
    assert A.output_fps(12.0, 3, None) == 48.0
    assert A.output_fps(12.0, 3, 0.5) == 24                    # math.floor(48 * 0.5), app.py:852-853
    assert A.output_fps(25.0, 2, 0.33) == 24


def test_stage_selection_follows_apply_image_process(tmp_path):
    from PIL import Image
    s1, s2 = str(tmp_path / "a.png"), str(tmp_path / "b.png")
    Image.new("RGB", (8, 8), (255, 0, 0)).save(s1)
    Image.new("RGB", (8, 8), (0, 255, 0)).save(s2)
    style = Image.new("RGB", (8, 8))
    assert A._stages(["Style Transfer"], None, None, None, None) is None          # app.py:468-470
    st = A._stages(["Style Transfer"], style, None, None, None)
    assert len(st) == 1 and st[0][0] == [style] and st[0][2] is False
    st = A._stages(["Style Mixing"], None, 0.3, s1, s2)
    assert len(st) == 1 and len(st[0][0]) == 2 and st[0][1] == 0.3                # app.py:487-572
    assert len(A._stages(["Style Mixing"], None, 0.3, None, s2)[0][0]) == 1       # app.py:578-580
    assert A._stages(["Style Mixing"], None, None, s1, s2) is None                # app.py:581-583
    assert len(A._stages(["Style Transfer", "Style Mixing"], style, 0.5, s1, s2)) == 2
    for effect in ("Pixel Art", "Text-Based Effects", "Convert Output to Grayscale", "Depth Based Style Transfer"):
        with pytest.raises(pkg.NstError):
            A._stages([effect, "Style Transfer"], style, None, None, None)


def test_no_cpu_path(tmp_path):
    import torch
    assert A.apply_video_process(None, ["Style Transfer"]) is None               # app.py:771-773
    with pytest.raises(pkg.NstError):
        A.apply_video_process(str(tmp_path / "x.mp4"), ["Style Transfer"], input_style=object(), device="cpu")
    if not torch.cuda.is_available():
        with pytest.raises(pkg.NstError):
            A.apply_video_process(str(tmp_path / "x.mp4"), ["Style Transfer"], input_style=object(), device="cuda")


def test_compat_names_exist():
    """north_star's tutorial-style surface (basic.py:12 imports it, commented out)."""
    m = importlib.import_module("text-based-image-style-transfer_b200.multi_style_transfer.run_style_transfer")
    sig = inspect.signature(m.run_style_transfer)
    assert list(sig.parameters) == ["cnn", "normalization_mean", "normalization_std", "content_img", "style_img", "input_img", "num_steps",
                                    "style_weight", "content_weight"]
    assert sig.parameters["num_steps"].default == 300 and sig.parameters["content_weight"].default == 1
    import torch
    assert issubclass(m.ContentLoss, torch.nn.Module) and issubclass(m.StyleLoss, torch.nn.Module)
    with pytest.raises(pkg.NstError):
        m.run_style_transfer(None, [0.5] * 3, [0.2] * 3, torch.zeros(1, 3, 8, 8), torch.zeros(1, 3, 8, 8), torch.zeros(1, 3, 8, 8))


def test_channel_attention_gate_weights_are_reproducible():
    """the gate weights must not depend on what consumed the global generator before (VGG construction on the
    first call only) - identical on every call, equal to the oracle's first-draws-after-seed-101 convention."""
    import torch
    from oracle import nst_oracle as O
    m = importlib.import_module("text-based-image-style-transfer_b200.multi_style_transfer.run_style_transfer")
    torch.manual_seed(7)
    before = torch.random.get_rng_state().clone()
    (a1, a2), = m.channel_attention_gate_weights([512])
    assert torch.equal(torch.random.get_rng_state(), before)                       # the caller's generator is untouched
    torch.randn(1000)                                                              # ... and whatever it does in between does not matter
    (b1, b2), = m.channel_attention_gate_weights([512])
    assert torch.equal(a1, b1) and torch.equal(a2, b2)
    o1, o2 = O.channel_attention_weights(512, seed=101)
    assert torch.equal(a1, o1) and torch.equal(a2, o2)
    two = m.channel_attention_gate_weights([512, 256])                             # a fresh module per content layer, one stream of draws
    assert torch.equal(two[0][0], a1) and two[1][0].shape == (128, 256)
