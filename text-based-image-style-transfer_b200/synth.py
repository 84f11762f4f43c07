"""Synthetic inputs named by BASELINE.json / SURVEY.md section 8(d): seeded 1/f^2 noise images (white noise makes
even the reference's own trajectory chaotic) and random-init VGG-19 weights of the
torchvision architecture (pretrained weights cannot be downloaded offline)."""
import numpy as np
import torch

VGG_MEAN = [0.485, 0.456, 0.406]   # app.py:376
VGG_STD = [0.229, 0.224, 0.225]    # app.py:377
APP_WEIGHTS = dict(w_style=5e5, w_content=1.0, w_tv=2e1, w_edge=2e1)   # app.py:380-385


def synth_image(h: int, w: int, seed: int, beta: float = 2.0) -> np.ndarray:
    """uint8 [h, w, 3]; per channel: Gaussian noise with a 1/f^beta amplitude spectrum, zero DC, unit std, mapped to
    0.5 + 0.2 z, clipped to [0, 1] and quantised."""
    rng = np.random.default_rng(seed)
    fy = np.fft.fftfreq(h)[:, None]
    fx = np.fft.fftfreq(w)[None, :]
    rad = np.sqrt(fy * fy + fx * fx)
    rad[0, 0] = 1.0
    amp = rad ** (-beta)
    amp[0, 0] = 0.0
    out = np.empty((h, w, 3), dtype=np.float64)
    for c in range(3):
        z = np.real(np.fft.ifft2(np.fft.fft2(rng.standard_normal((h, w))) * amp))
        z = (z - z.mean()) / (z.std() + 1e-12)
        out[:, :, c] = 0.5 + 0.2 * z
    return (np.clip(out, 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)


def vgg19_random_weights(seed: int = 1234, n_conv: int = 13):
    """(weights, biases) of the first n_conv convolutions of torchvision.models.vgg19(weights=None) constructed under
    torch.manual_seed(seed); the caller's RNG state is preserved."""
    import torchvision
    state = torch.random.get_rng_state()
    try:
        torch.manual_seed(seed)
        feats = torchvision.models.vgg19(weights=None).features
    finally:
        torch.random.set_rng_state(state)
    convs = [m for m in feats if isinstance(m, torch.nn.Conv2d)][:n_conv]
    return [c.weight.detach().clone() for c in convs], [c.bias.detach().clone() for c in convs]


# algorithmic work per closure evaluation
_CIN = [3, 64, 64, 128, 128, 256, 256, 256, 256, 512, 512, 512, 512]
_COUT = [64, 64, 128, 128, 256, 256, 256, 256, 512, 512, 512, 512, 512]
_LEVEL = [0, 0, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4]
_STYLE = [0, 2, 4, 8, 12]


def conv_flops(layer: int, H: int, W: int) -> float:
    h, w = H >> _LEVEL[layer], W >> _LEVEL[layer]
    return 2.0 * 9 * _CIN[layer] * _COUT[layer] * h * w


def gram_flops(layer: int, H: int, W: int) -> float:
    h, w = H >> _LEVEL[layer], W >> _LEVEL[layer]
    return 2.0 * _COUT[layer] * _COUT[layer] * h * w


def eval_flops(H: int, W: int) -> float:
    """forward + data-gradient of the 13 convolutions, Gram forward + backward of the 5 style layers."""
    return 2.0 * sum(conv_flops(i, H, W) for i in range(13)) + 2.0 * sum(gram_flops(i, H, W) for i in _STYLE)


def lbfgs_bytes(H: int, W: int, m: int) -> float:
    """two streaming passes over m stored (s, y) pairs: (4 m + 10) n floats."""
    return (4.0 * m + 10.0) * 3 * H * W * 4
