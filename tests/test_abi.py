"""The C-ABI shared library loads without a GPU and exports every symbol include/nst_b200.h declares; the Python
mirror keeps the reference's names and signatures and fails loudly without a CUDA device."""
import ctypes
import importlib
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header(instrumented=False):
    text = open(os.path.join(ROOT, "include", "nst_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    blocks = re.findall(r"#ifdef NST_INSTRUMENT(.*?)#endif", text, flags=re.S)
    if instrumented:
        return "\n".join(blocks)
    return re.sub(r"#ifdef NST_INSTRUMENT.*?#endif", "", text, flags=re.S)


def header_symbols(instrumented=False):
    return sorted(set(re.findall(r"\b(nst_[a-z0-9_]+)\s*\(", _header(instrumented))))


def test_library_exports_every_declared_symbol(built_libs):
    lib = ctypes.CDLL(built_libs[0])
    names = header_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), n
    lib.nst_abi_version.restype = ctypes.c_int
    assert lib.nst_abi_version() == 1


def test_product_library_has_no_instrumentation(built_libs):
    """Tuning aids (phase stamps, launch spans, wrong-result timing experiments) live in the -DNST_INSTRUMENT build only."""
    lib = ctypes.CDLL(built_libs[0])
    names = header_symbols(instrumented=True)
    assert names == ["nst_lbfgs_ctl_clocks", "nst_plan_conv_phases", "nst_plan_timeline"]
    for n in names:
        assert not hasattr(lib, n), n
    pkg = importlib.import_module("text-based-image-style-transfer_b200")
    assert set(pkg._lib.INSTR_PROTOTYPES) == set(names)
    blob = open(built_libs[0], "rb").read()
    for env in (b"NST_DBG_FLAGS", b"NST_DBG_CONV1", b"NST_CHAIN"):
        assert env not in blob, env


def test_binding_covers_the_header(built_libs):
    pkg = importlib.import_module("text-based-image-style-transfer_b200")
    assert set(pkg._lib.PROTOTYPES) == set(header_symbols())
    pkg._lib.load()
    assert ctypes.sizeof(pkg._lib.NstStatus) == 6 * 4 + 7 * 8 + 16 * 4


def test_library_contains_blackwell_tensor_core_code(built_libs):
    import subprocess
    sass = subprocess.run(["cuobjdump", "-sass", built_libs[0]], stdout=subprocess.PIPE, text=True).stdout
    assert "UTCHMMA" in sass      # tcgen05.mma
    assert "UTMALDG" in sass      # TMA tensor loads
    assert "LDTM" in sass         # tcgen05.ld
    assert "UTCHMMA.2CTA" in sass             # tcgen05.mma.cta_group::2: the convolutions run as CTA pairs
    assert "UTCBAR.2CTA.MULTICAST" in sass    # tcgen05.commit to the barriers of both CTAs of a pair
    assert "UTMALDG.4D.2CTA" in sass          # a pair CTA's patch load completing on the leader's barrier
    assert "sm_100a" in subprocess.run(["cuobjdump", "-lelf", built_libs[0]], stdout=subprocess.PIPE, text=True).stdout


def test_mirror_keeps_reference_signature(built_libs):
    m = importlib.import_module("text-based-image-style-transfer_b200.multi_style_transfer.run_style_transfer")
    sig = inspect.signature(m.run_multi_style_transfer)
    assert list(sig.parameters) == ["vgg_mean", "vgg_std", "content_img", "num_steps", "random_init", "w_style", "w_content",
                                    "w_tv", "w_edge", "style_img1", "style_img2", "style_img_weight", "print_iter",
                                    "channel_attention", "device"]  # run_style_transfer.py:27-28
    assert sig.parameters["style_img2"].default is None and sig.parameters["style_img_weight"].default == 0.5
    assert sig.parameters["print_iter"].default == 50 and sig.parameters["device"].default == "cpu"
    L = importlib.import_module("text-based-image-style-transfer_b200.multi_style_transfer.style_transfer_losses")
    for name in ("normalize", "content_loss", "gram_matrix", "style_loss", "total_variation_loss", "get_gradient_imgs",
                 "edge_loss", "to_grayscale", "seed_everything", "Vgg19", "StyleMixer"):
        assert hasattr(L, name), name
    importlib.import_module("text-based-image-style-transfer_b200.multi_style_transfer.ChannelAttention").ChannelAttention(8)


def test_no_cpu_fallback(built_libs):
    pkg = importlib.import_module("text-based-image-style-transfer_b200")
    with pytest.raises(pkg.NstError):
        pkg.run_multi_style_transfer([0.5] * 3, [0.2] * 3, None, 1, False, 1, 1, 1, 1, None, device="cpu")
    if not torch.cuda.is_available():
        with pytest.raises(pkg.NstError):
            pkg.Net([torch.zeros(64, 3, 3, 3)], [torch.zeros(64)], "cuda")
    L = importlib.import_module("text-based-image-style-transfer_b200.multi_style_transfer.style_transfer_losses")
    with pytest.raises(pkg.NstError):
        L.gram_matrix(torch.zeros(1, 64, 4, 4))


def test_mirrors_around_the_loop_keep_reference_names_and_have_no_cpu_path(built_libs):
    """text/segmentation_style_transfer.py:5, components/style_transfer_depth/*.py, app.py:800-840"""
    pkg = importlib.import_module("text-based-image-style-transfer_b200")
    seg = importlib.import_module("text-based-image-style-transfer_b200.text.segmentation_style_transfer")
    assert list(inspect.signature(seg.segmentation_style_transfer).parameters)[:4] == ["content_image", "style_image", "segmentation_mask",
                                                                                      "edge_smoothing"]
    assert inspect.signature(seg.segmentation_style_transfer).parameters["edge_smoothing"].default == 5
    D = importlib.import_module("text-based-image-style-transfer_b200.components.style_transfer_depth.style_transfer_depth")
    for name in ("get_depth_map", "style_transfer", "process_mip_layers", "style_MIP", "style_Dept", "depth_split"):
        assert callable(getattr(D.DepthStyle, name)), name
    assert list(inspect.signature(D.DepthStyle.style_MIP).parameters) == ["self", "image", "style", "n"]
    A = importlib.import_module("text-based-image-style-transfer_b200.components.style_transfer_depth.Style_a3")
    sig = inspect.signature(A.StyleA3.__init__)
    assert list(sig.parameters)[:11] == ["self", "device", "print_iter", "num_steps", "w_style", "w_content", "w_tv", "w_edge", "w_depth",
                                         "random_init", "depth_pipeline"]
    assert sig.parameters["num_steps"].default == 400 and sig.parameters["w_style"].default == 5e5          # Style_a3.py:18
    U = importlib.import_module("text-based-image-style-transfer_b200.components.style_transfer_depth.util")
    for name in ("mask_image_depth", "create_bins", "generate_mip_layers", "reconstruct_mip_image", "image_loader", "save_image", "Vgg19",
                 "normalize", "content_loss", "gram_matrix", "style_loss", "total_variation_loss", "get_gradient_imgs", "edge_loss"):
        assert hasattr(U, name), name
    assert [list(map(float, b)) for b in U.create_bins(4)] == [[0.0, 0.25], [0.25, 0.5], [0.5, 0.75], [0.75, 1.0]]
    V = importlib.import_module("text-based-image-style-transfer_b200.video")
    with pytest.raises(pkg.NstError):
        D.DepthStyle("cpu", depth_pipeline=lambda image: None)
    with pytest.raises(pkg.NstError):
        V.assemble_frames(torch.zeros((2, 4, 4, 3), dtype=torch.uint8), 1)
    with pytest.raises(pkg.NstError):
        seg.composite_tensors(torch.zeros((4, 4, 3), dtype=torch.uint8), torch.zeros((4, 4, 3), dtype=torch.uint8), torch.zeros((4, 4), dtype=torch.bool))


def test_product_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "text-based-image-style-transfer_b200")
    for dirpath, _, files in os.walk(pkg_dir):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                assert "oracle" not in open(os.path.join(dirpath, f)).read().replace("the oracle", "").replace("oracle's", ""), f


def test_pil_round_trip_helpers():
    import numpy as np
    from PIL import Image
    m = importlib.import_module("text-based-image-style-transfer_b200.multi_style_transfer.run_style_transfer")
    arr = (np.arange(5 * 7 * 3) % 256).astype(np.uint8).reshape(5, 7, 3)
    t = m.PIL_to_tensor(Image.fromarray(arr))
    assert t.shape == (1, 3, 5, 7) and t.dtype == torch.float32 and float(t.max()) <= 1.0
    assert np.array_equal(np.asarray(m.tensor_to_PIL(t[0])), arr)
    assert np.asarray(m.tensor_to_PIL(torch.full((3, 1, 1), 0.999)))[0, 0, 0] == 254   # truncation
