#!/usr/bin/env python3
"""Golden vectors for the LARGE configurations of BASELINE.json (configs[1]..[4]) written by the UNMODIFIED
reference (/root/reference/multi_style_transfer/run_style_transfer.py:27-159) on CPU.  Build container only.

Why a second script: make_golden.py pins the oracle at <= 64x64 / 60 evaluations; north_star asks for "loss curve
within 1e-2 over the run, final image >= 40 dB after N steps" on the configurations the performance is quoted on.
Each case is run TWICE with different CPU thread counts; the deviation between the two reference runs is the
reference's own self-noise and is stored beside the golden so that the GPU tests can print theirs next to
it.

Usage (each run is minutes of CPU):
    make_golden_large.py run CASE THREADS     -> /tmp/golden_large/CASE_tTHREADS.npz  (full precision)
    make_golden_large.py finalize CASE TA TB  -> tests/golden/large_CASE.npz  (TA = golden, TB = self-noise run)

Stored: the uint8 inputs (the GPU box has no /root/reference/data), the loss of every closure evaluation, the final
float image as fp16 (quantisation noise ~77 dB, far below the 40 dB bar), the final uint8 image, a few
intermediate iterates as fp16, and the self-noise (PSNR of the two runs' final images, max relative loss deviation).
"""
import os
import sys

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as G  # noqa: E402  (shim + instrumentation shared with the small cases)

O = G.O
DATA = "/root/reference/data"
TMP = "/tmp/golden_large"


def natural(path, size):
    """RGB uint8 [h, w, 3] of a shipped image resized (LANCZOS) to size = (w, h)."""
    im = Image.open(os.path.join(DATA, path)).convert("RGB")
    if im.size != size:
        im = im.resize(size, Image.LANCZOS)
    return np.asarray(im).copy()


def video_frame(path, index, size):
    import cv2
    cap = cv2.VideoCapture(os.path.join(DATA, path))
    frame = None
    for _ in range(index + 1):
        ok, frame = cap.read()
        assert ok
    cap.release()
    rgb = cv2.cvtColor(frame, cv2.COLOR_BGR2RGB)
    return np.asarray(Image.fromarray(rgb).resize(size, Image.LANCZOS)).copy()


# name: (content builder, [style builders], num_steps, style_img_weight, channel_attention, iterates to keep)
CASES = {
    # BASELINE configs[1] on the pair SURVEY A.3 found monotone for 320 evaluations
    "nat512": (lambda: natural("content_imgs/dog.jpeg", (512, 512)),
               [lambda: natural("style_imgs/starry_night.jpg", (512, 512))], 300, 0.5, False, (20, 100, 200)),
    # BASELINE configs[1] on the synthetic pair of SURVEY 8(d)
    "syn512": (lambda: O.synth_image(512, 512, 0), [lambda: O.synth_image(512, 512, 1)], 300, 0.5, False,
               (20, 100, 200)),
    # configs[1] at a shipped image's NATIVE resolution (no resampling smoothness)
    "boat512": (lambda: natural("content_imgs/boat.jpg", (512, 512)),
                [lambda: natural("style_imgs/starry_night_big.jpg", (512, 512))], 300, 0.5, False, (20, 100, 200)),
    # BASELINE configs[2]: two styles + channel attention
    "nat512_mix_ca": (lambda: natural("content_imgs/dog.jpeg", (512, 512)),
                      [lambda: natural("style_imgs/starry_night.jpg", (512, 512)),
                       lambda: natural("style_imgs/picasso.jpg", (512, 512))], 40, 0.5, True, (20, 40)),
    # BASELINE configs[3]
    "nat1024": (lambda: natural("content_imgs/face.jpg", (1024, 1024)),
                [lambda: natural("style_imgs/starry_night_big.jpg", (1024, 1024))], 40, 0.5, False, (20, 40)),
    # configs[3] / [4] on the pair that is stable at 512^2: the shipped frame / face inputs below put the
    # REFERENCE ITSELF into the overshoot regime of unit-step L-BFGS within a few evaluations (their goldens are kept to show
    # that the overshoot is reproduced; run-level tolerances can only be asked where the reference is stable against itself)
    "dog1024": (lambda: natural("content_imgs/dog.jpeg", (1024, 1024)),
                [lambda: natural("style_imgs/starry_night.jpg", (1024, 1024))], 40, 0.5, False, (20, 40)),
    "dog720p": (lambda: natural("content_imgs/dog.jpeg", (1280, 720)),
                [lambda: natural("style_imgs/starry_night.jpg", (512, 512))], 40, 0.5, False, (20, 40)),
    # BASELINE configs[4]: one 720p frame, 512x512 shared style
    "nat720p": (lambda: video_frame("content_vids/car.mp4", 10, (1280, 720)),
                [lambda: natural("style_imgs/starry_night.jpg", (512, 512))], 40, 0.5, False, (20, 40)),
}


def run(case, threads):
    G.install_shim()
    torch.set_num_threads(threads)
    cb, sbs, steps, wgt, ca, keep = CASES[case]
    content = cb()
    styles = [s() for s in sbs]
    img, losses, iterates = G.run_reference(content, styles, steps, O.APP_WEIGHTS, wgt, ca)
    os.makedirs(TMP, exist_ok=True)
    arrays = dict(content_u8=content, final_u8=img, loss_trace=losses, x_final=iterates[-1].numpy(),
                  n_evals=np.array(len(losses)), num_steps=np.array(steps), style_img_weight=np.array(wgt),
                  channel_attention=np.array(ca), threads=np.array(threads))
    for i, s in enumerate(styles):
        arrays["style%d_u8" % i] = s
    for k in keep:
        if k < len(iterates):
            arrays["x_eval%d" % k] = iterates[k].numpy()
    path = os.path.join(TMP, "%s_t%d.npz" % (case, threads))
    np.savez(path, **arrays)
    print(case, "threads", threads, "evals", len(losses), "loss", losses[0], "->", losses[-1], flush=True)


def psnr(a, b):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 10.0 * np.log10(1.0 / max(mse, 1e-30))


def finalize(case, ta, tb):
    a = np.load(os.path.join(TMP, "%s_t%d.npz" % (case, ta)))
    b = np.load(os.path.join(TMP, "%s_t%d.npz" % (case, tb)))
    out = {}
    for k in a.files:
        v = a[k]
        if k.startswith("x_"):
            v = v.astype(np.float16)
        out[k] = v
    la, lb = a["loss_trace"], b["loss_trace"]
    n = min(len(la), len(lb))
    self_dev = np.abs(la[:n] - lb[:n]) / np.abs(la[:n])
    out["self_loss_dev"] = np.array(float(np.max(self_dev)))
    out["self_dev_trace"] = self_dev
    # evaluations before the reference parts from ITSELF (two thread counts more than 1e-3 apart): only there can a run-level
    # tolerance be asked of anybody; behind it unit-step L-BFGS has overshot and the trajectory is chaotic
    bad = np.nonzero(self_dev > 1e-3)[0]
    out["stable_prefix"] = np.array(int(bad[0]) if len(bad) else n)
    out["self_psnr"] = np.array(psnr(a["x_final"], b["x_final"]))
    out["self_threads"] = np.array([ta, tb])
    out["self_n_evals"] = np.array([len(la), len(lb)])
    dec = float(np.mean(np.diff(la) < 0))
    out["frac_decreasing"] = np.array(dec)
    # keep the fixtures small: no intermediate iterates (the runs keep them under /tmp for debugging); the uint8 image only up to 512^2 (it is the truncated
    # float image); no images at all where the reference is not stable against itself (nothing can be compared with them)
    big = a["content_u8"].shape[0] * a["content_u8"].shape[1] > 512 * 512
    unstable = int(out["stable_prefix"]) < n
    for k in list(out):
        if k.startswith("x_eval") or (k == "final_u8" and (big or unstable)) or (k == "x_final" and unstable):
            del out[k]
    path = os.path.join(HERE, "large_%s.npz" % case)
    np.savez_compressed(path, **out)
    print(case, "evals", len(la), "loss", la[0], "->", la[-1], "decreasing", dec, "self-noise: loss dev",
          float(out["self_loss_dev"]), "psnr", float(out["self_psnr"]), "dB; stable prefix", int(out["stable_prefix"]), "evaluations;", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    if sys.argv[1] == "run":
        run(sys.argv[2], int(sys.argv[3]))
    else:
        finalize(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]))
