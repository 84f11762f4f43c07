"""CPU oracle for the iterative neural-style-transfer hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (text-based-image-style-transfer_b200/) never does and fails loudly without
its CUDA library.

What this is: a restatement, in plain PyTorch fp32 (or fp64) on the CPU, of the reference's algorithm
for the path - written from the reference's behaviour, not copied from it.  Each function cites the
reference lines it follows (paths relative to /root/reference/).  Third-party arithmetic the reference
relies on and that is not under /root/reference:
  * torch (requirements.txt:2 pins ==1.13.1; 2.11.0 is installed): nn.functional.conv2d / max_pool2d,
    bmm, autograd, F.interpolate, and torch.optim.LBFGS, whose algorithm is restated here as
    `LbfgsOracle` (from torch/optim/lbfgs.py:333-537) and pinned against torch.optim.LBFGS itself in
    tests/test_oracle.py;
  * torchvision (requirements.txt:6, unpinned; 0.26.0 installed): the VGG-19 "E" topology
    (torchvision/models/vgg.py:73-94) restated as `VGG_CFG`.

Parity pinning: the reference has no tests, golden vectors or fixtures of its own, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: tests/golden/make_golden.py imports
/root/reference/multi_style_transfer unmodified (three-line shim), runs it on seeded inputs and stores
losses, Gram matrices, gradients, loss traces and final images; tests/test_oracle.py checks this module
against those files.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

# torchvision vgg19().features, configuration "E", as (kind, channels); names as in helper_functions.py:52-58
VGG_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]
CONV_NAMES = [
    "conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3", "conv3_4",
    "conv4_1", "conv4_2", "conv4_3", "conv4_4", "conv5_1", "conv5_2", "conv5_3", "conv5_4",
]
CONTENT_LAYERS = ["conv4_2"]                                            # run_style_transfer.py:56
STYLE_LAYERS = ["conv1_1", "conv2_1", "conv3_1", "conv4_1", "conv5_1"]  # run_style_transfer.py:57
APP_WEIGHTS = dict(w_style=5e5, w_content=1.0, w_tv=2e1, w_edge=2e1)    # app.py:380-385
VGG_MEAN = [0.485, 0.456, 0.406]                                        # app.py:376
VGG_STD = [0.229, 0.224, 0.225]                                         # app.py:377


# ------------------------------------------------------------------------------------------------
# synthetic inputs: 1/f^2 noise images and random-init VGG-19 weights
# ------------------------------------------------------------------------------------------------
def synth_image(h: int, w: int, seed: int, beta: float = 2.0) -> np.ndarray:
    """uint8 [h, w, 3]: per-channel Gaussian noise shaped to a 1/f^beta amplitude spectrum."""
    rng = np.random.default_rng(seed)
    fy = np.fft.fftfreq(h)[:, None]
    fx = np.fft.fftfreq(w)[None, :]
    rad = np.sqrt(fy * fy + fx * fx)
    rad[0, 0] = 1.0
    amp = rad ** (-beta)
    amp[0, 0] = 0.0
    out = np.empty((h, w, 3), dtype=np.float64)
    for c in range(3):
        spec = np.fft.fft2(rng.standard_normal((h, w))) * amp
        z = np.real(np.fft.ifft2(spec))
        z = (z - z.mean()) / (z.std() + 1e-12)
        out[:, :, c] = 0.5 + 0.2 * z
    return (np.clip(out, 0.0, 1.0) * 255.0 + 0.5).astype(np.uint8)


def vgg19_random_weights(seed: int = 1234, n_conv: int = 13, dtype=torch.float32):
    """(weights, biases) of the first n_conv convolutions of torchvision's vgg19(weights=None) built under
    torch.manual_seed(seed): kaiming_normal_(fan_out, relu) weights, zero bias (torchvision/models/vgg.py:55-57).
    Draws follow torchvision's module order (all 16 convolutions are drawn; later ones are discarded), so the
    tensors equal `torchvision.models.vgg19(weights=None).features` under the same seed."""
    gen_state = torch.random.get_rng_state()
    try:
        torch.manual_seed(seed)
        ws, bs = [], []
        # torchvision first constructs every layer (default nn.Conv2d/nn.Linear init consumes the RNG), then
        # re-initialises in module order; reproduce by building the real module when torchvision is present.
        import torchvision  # noqa: F401  (installed in this image; only its constructor is used)
        m = torchvision.models.vgg19(weights=None)
        for layer in m.features:
            if isinstance(layer, torch.nn.Conv2d):
                ws.append(layer.weight.detach().to(dtype).clone())
                bs.append(layer.bias.detach().to(dtype).clone())
        return ws[:n_conv], bs[:n_conv]
    finally:
        torch.random.set_rng_state(gen_state)


# ------------------------------------------------------------------------------------------------
# loss functions
# ------------------------------------------------------------------------------------------------
def normalize(img: torch.Tensor, mean, std) -> torch.Tensor:
    """style_transfer_losses.py:9-28 - per-channel z-score on a (b, c, h, w) tensor."""
    m = torch.as_tensor(mean, dtype=img.dtype, device=img.device).reshape(1, -1, 1, 1)
    s = torch.as_tensor(std, dtype=img.dtype, device=img.device).reshape(1, -1, 1, 1)
    return (img - m) / s


def to_grayscale(img: torch.Tensor) -> torch.Tensor:
    """helper_functions.py:104-113 - channel mean, keepdim."""
    return img.mean(dim=1, keepdim=True)


def gram_matrix(x: torch.Tensor) -> torch.Tensor:
    """style_transfer_losses.py:70-95 - X X^T / (b c h w) with X = x.view(b, c, h w)."""
    b, c, h, w = x.shape
    flat = x.reshape(b, c, h * w)
    return torch.bmm(flat, flat.transpose(1, 2)) / (b * c * h * w)


def content_loss(inp: Dict[str, torch.Tensor], tgt: Dict[str, torch.Tensor], layers: Sequence[str]) -> torch.Tensor:
    """style_transfer_losses.py:31-67 - mean squared error per layer, averaged over layers."""
    total = 0.0
    for name in layers:
        total = total + F.mse_loss(inp[name], tgt[name].detach(), reduction="mean")
    return total / len(layers)


def style_mix(feat_a: torch.Tensor, feat_b: torch.Tensor, weight_b: float) -> torch.Tensor:
    """StyleMixer.py:7-38 - both maps are resized (bilinear, align_corners=True) to
    (Ha + Hb // 2, Wa + Wb // 2) - note the precedence - then blended (1 - w) * a + w * b."""
    ha, wa = feat_a.shape[2:]
    hb, wb = feat_b.shape[2:]
    size = (int(ha + hb // 2), int(wa + wb // 2))
    ra = F.interpolate(feat_a.detach(), size=size, mode="bilinear", align_corners=True)
    rb = F.interpolate(feat_b.detach(), size=size, mode="bilinear", align_corners=True)
    return (1 - weight_b) * ra + weight_b * rb


def style_targets(style_feats: List[Dict[str, torch.Tensor]], layers: Sequence[str], weight_b: float):
    """The constant half of style_loss (style_transfer_losses.py:122-135): Gram of the single style's
    features, or of the StyleMixer blend of two.  The reference recomputes this every evaluation; it
    depends on constants only."""
    out = {}
    for name in layers:
        if len(style_feats) == 1:
            out[name] = gram_matrix(style_feats[0][name])
        else:
            out[name] = gram_matrix(style_mix(style_feats[0][name], style_feats[1][name], weight_b))
    return out


def style_loss_from_targets(inp: Dict[str, torch.Tensor], targets: Dict[str, torch.Tensor], layers: Sequence[str]):
    """style_transfer_losses.py:98-146 with the target Grams precomputed; returns (loss, per-layer MSE list)."""
    total = 0.0
    per = []
    for name in layers:
        mse = F.mse_loss(gram_matrix(inp[name]), targets[name], reduction="mean")
        per.append(mse)
        total = total + mse
    return total / len(layers), per


def total_variation_loss(y: torch.Tensor) -> torch.Tensor:
    """style_transfer_losses.py:149-174 - (sum |d/dh| + sum |d/dw|) / (c h w)."""
    norm = y.shape[1] * y.shape[2] * y.shape[3]
    dh = (y[:, :, 1:, :] - y[:, :, :-1, :]).abs().sum()
    dw = (y[:, :, :, 1:] - y[:, :, :, :-1]).abs().sum()
    return (dh + dw) / norm


def get_gradient_imgs(gray: torch.Tensor) -> torch.Tensor:
    """style_transfer_losses.py:177-204 - central differences on the interior, (1, 2, H-2, W-2): x first."""
    h, w = gray.shape[2], gray.shape[3]
    dx = gray[:, :, 1:h - 1, 2:] - gray[:, :, 1:h - 1, :w - 2]
    dy = gray[:, :, 2:, 1:w - 1] - gray[:, :, :h - 2, 1:w - 1]
    return torch.cat((dx, dy), dim=1)


def edge_loss(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """style_transfer_losses.py:207-225 - mean of the two per-direction MSEs."""
    return (F.mse_loss(a[:, 0], b[:, 0]) + F.mse_loss(a[:, 1], b[:, 1])) / 2


# ------------------------------------------------------------------------------------------------
# VGG-19 feature extractor
# ------------------------------------------------------------------------------------------------
class VggFeatures:
    """helper_functions.py:44-101 - the first convolutions of VGG-19 `features`; forward returns the
    PRE-ReLU output of every requested convolution (the reference cuts its slices right after the conv,
    SURVEY quirk 1) and stops after the deepest one."""

    def __init__(self, weights: List[torch.Tensor], biases: List[torch.Tensor], layers: Sequence[str]):
        unknown = [n for n in layers if n not in CONV_NAMES]
        if unknown:
            raise Exception("Not all layers provided in content_layes and/or style_layers exist.")
        self.layers = [n for n in CONV_NAMES if n in set(layers)]
        self.last = max(CONV_NAMES.index(n) for n in self.layers)
        if self.last >= len(weights):
            raise ValueError("weights for conv index %d missing" % self.last)
        self.weights, self.biases = weights, biases

    def __call__(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        out = {}
        conv = 0
        for item in VGG_CFG:
            if item == "M":
                x = F.max_pool2d(x, kernel_size=2, stride=2)
                continue
            x = F.conv2d(x, self.weights[conv], self.biases[conv], padding=1)
            if CONV_NAMES[conv] in self.layers:
                out[CONV_NAMES[conv]] = x.clone()      # helper_functions.py:98
            if conv == self.last:
                break
            x = F.relu(x)                              # out of place: the tap above stays pre-ReLU (:72-74)
            conv += 1
        return out


def channel_attention_weights(channels: int, reduction: int = 2, seed: Optional[int] = 101):
    """ChannelAttention.py:16-17 - two bias-free nn.Linear layers with PyTorch's default init.  With
    `seed`, the global RNG is seeded first like run_style_transfer.py:52 does before the module is built
    (valid when nothing else consumed the RNG in between, which holds under the oracle's VGG shim)."""
    if seed is not None:
        torch.manual_seed(seed)
    fc1 = torch.nn.Linear(channels, channels // reduction, bias=False)
    fc2 = torch.nn.Linear(channels // reduction, channels, bias=False)
    return fc1.weight.detach().clone(), fc2.weight.detach().clone()


def channel_attention(x: torch.Tensor, w1: torch.Tensor, w2: torch.Tensor) -> torch.Tensor:
    """ChannelAttention.py:23-40 - x * sigmoid(relu(W2 relu(W1 avgpool(x))))."""
    pooled = x.mean(dim=(2, 3))
    gate = torch.sigmoid(F.relu(F.linear(F.relu(F.linear(pooled, w1)), w2)))
    return x * gate.reshape(gate.shape[0], gate.shape[1], 1, 1)


# ------------------------------------------------------------------------------------------------
# L-BFGS (torch.optim.LBFGS defaults, no line search)
# ------------------------------------------------------------------------------------------------
@dataclass
class LbfgsOracle:
    """torch/optim/lbfgs.py:333-537 restated for one flat parameter vector.  `closure()` must evaluate the
    objective at self.x (it may modify self.x in place - the reference clamps there) and return
    (loss: float, grad: Tensor)."""
    x: torch.Tensor
    lr: float = 1.0
    max_iter: int = 20
    tolerance_grad: float = 1e-7
    tolerance_change: float = 1e-9
    history_size: int = 100
    n_iter: int = 0
    func_evals: int = 0
    d: Optional[torch.Tensor] = None
    t: Optional[float] = None
    old_dirs: list = field(default_factory=list)
    old_stps: list = field(default_factory=list)
    ro: list = field(default_factory=list)
    H_diag: float = 1.0
    prev_grad: Optional[torch.Tensor] = None
    prev_loss: Optional[float] = None
    stop_reason: str = ""

    def step(self, closure):
        max_eval = self.max_iter * 5 // 4
        loss, g = closure()
        loss = float(loss)
        evals = 1
        self.func_evals += 1
        self.stop_reason = ""
        if float(g.abs().max()) <= self.tolerance_grad:
            self.stop_reason = "opt_entry"
            return loss
        first_loss = loss
        it = 0
        while it < self.max_iter:
            it += 1
            self.n_iter += 1
            if self.n_iter == 1:
                self.d = g.neg()
                self.old_dirs, self.old_stps, self.ro = [], [], []
                self.H_diag = 1.0
            else:
                y = g - self.prev_grad
                s = self.d * self.t
                ys = float(y.dot(s))
                if ys > 1e-10:
                    if len(self.old_dirs) == self.history_size:
                        self.old_dirs.pop(0)
                        self.old_stps.pop(0)
                        self.ro.pop(0)
                    self.old_dirs.append(y)
                    self.old_stps.append(s)
                    self.ro.append(1.0 / ys)
                    self.H_diag = ys / float(y.dot(y))
                k = len(self.old_dirs)
                al = [0.0] * k
                q = g.neg()
                for i in range(k - 1, -1, -1):
                    al[i] = float(self.old_stps[i].dot(q)) * self.ro[i]
                    q.add_(self.old_dirs[i], alpha=-al[i])
                r = q * self.H_diag
                for i in range(k):
                    be = float(self.old_dirs[i].dot(r)) * self.ro[i]
                    r.add_(self.old_stps[i], alpha=al[i] - be)
                self.d = r
            self.prev_grad = g.clone()
            self.prev_loss = loss
            if self.n_iter == 1:
                self.t = min(1.0, 1.0 / float(g.abs().sum())) * self.lr
            else:
                self.t = self.lr
            gtd = float(g.dot(self.d))
            if gtd > -self.tolerance_change:
                self.stop_reason = "gtd"
                break
            self.x.add_(self.d, alpha=self.t)
            if it != self.max_iter:
                loss, g = closure()
                loss = float(loss)
                evals += 1
                self.func_evals += 1
            if it == self.max_iter:
                break
            if evals >= max_eval:
                self.stop_reason = "max_eval"
                break
            if float(g.abs().max()) <= self.tolerance_grad:
                self.stop_reason = "opt"
                break
            if float((self.d * self.t).abs().max()) <= self.tolerance_change:
                self.stop_reason = "step"
                break
            if abs(loss - self.prev_loss) < self.tolerance_change:
                self.stop_reason = "loss"
                break
        return first_loss


# ------------------------------------------------------------------------------------------------
# the whole path
# ------------------------------------------------------------------------------------------------
@dataclass
class OracleResult:
    image: torch.Tensor                 # (1, 3, H, W) final clamped float image (before uint8 truncation)
    losses: List[List[float]]           # per evaluation: [total, content_w, style_w, tv_w, edge_w]
    iterates: Optional[List[torch.Tensor]] = None
    grads: Optional[List[torch.Tensor]] = None
    evals: int = 0


def to_tensor_u8(img_u8: np.ndarray, dtype=torch.float32) -> torch.Tensor:
    """transforms.ToTensor() + unsqueeze (run_style_transfer.py:5-11): HWC uint8 -> 1CHW float / 255."""
    t = torch.from_numpy(np.ascontiguousarray(img_u8)).permute(2, 0, 1).to(torch.float32).div(255)
    return t.unsqueeze(0).to(dtype)


def to_u8(img: torch.Tensor) -> np.ndarray:
    """transforms.ToPILImage() on a float tensor (run_style_transfer.py:157): mul(255).byte() truncates."""
    return img.squeeze(0).mul(255).byte().permute(1, 2, 0).contiguous().numpy()


class ClosureOracle:
    """Setup (run_style_transfer.py:52-96) and closure (:102-148) of run_multi_style_transfer."""

    def __init__(self, weights, biases, content: torch.Tensor, styles: List[torch.Tensor], *, w_style, w_content,
                 w_tv, w_edge, style_img_weight=0.5, channel_attention_on=False, mean=VGG_MEAN, std=VGG_STD,
                 content_layers=CONTENT_LAYERS, style_layers=STYLE_LAYERS, ca_seed: Optional[int] = 101,
                 emulate_reference_cost: bool = False):
        self.dtype = content.dtype
        self.device = content.device
        self.mean, self.std = mean, std
        self.w = dict(style=w_style, content=w_content, tv=w_tv, edge=w_edge)
        self.content_layers, self.style_layers = list(content_layers), list(style_layers)
        # emulate_reference_cost (timing only; values are unchanged): the reference leaves the VGG weights with
        # requires_grad=True, so backward() also computes 13 weight gradients nobody reads, and it
        # recomputes the constant style targets inside every evaluation (quirk 6)
        self.emulate_reference_cost = emulate_reference_cost
        if emulate_reference_cost:
            weights = [w.detach().clone().requires_grad_(True) for w in weights]
            biases = [b.detach().clone().requires_grad_(True) for b in biases]
        self.vgg = VggFeatures(weights, biases, self.content_layers + self.style_layers)
        normed_content = normalize(content, mean, std)
        self.edge_target = get_gradient_imgs(to_grayscale(normed_content)).detach() if w_edge > 0 else None
        with torch.no_grad():
            sfeats = [self.vgg(normalize(s, mean, std)) for s in styles]
            cfeats = self.vgg(normed_content)
            self._sfeats, self._mix_w = sfeats, style_img_weight
            self.style_t = style_targets(sfeats, self.style_layers, style_img_weight)
            if channel_attention_on:
                # run_style_transfer.py:13-25: a fresh randomly initialised gate per content layer
                cfeats = dict(cfeats)
                first = True
                for name in self.content_layers:
                    w1, w2 = channel_attention_weights(cfeats[name].shape[1], seed=ca_seed if first else None)
                    first = False
                    cfeats[name] = channel_attention(cfeats[name], w1.to(self.device, self.dtype), w2.to(self.device, self.dtype))
            self.content_t = cfeats

    def evaluate(self, x: torch.Tensor, need_grad=True):
        """x: (1,3,H,W) already clamped.  Returns dict(total, content, style, tv, edge, gram_mse, grams, grad)."""
        xv = x.detach().clone().requires_grad_(need_grad)
        normed = normalize(xv, self.mean, self.std)
        feats = self.vgg(normed)
        zero = torch.zeros((), dtype=self.dtype, device=self.device)
        c = self.w["content"] * content_loss(feats, self.content_t, self.content_layers) if self.w["content"] > 0 else zero
        per = []
        if self.w["style"] > 0:
            targets = self.style_t
            if self.emulate_reference_cost:
                targets = style_targets(self._sfeats, self.style_layers, self._mix_w)
            sl, per = style_loss_from_targets(feats, targets, self.style_layers)
            s = self.w["style"] * sl
        else:
            s = zero
        tv = self.w["tv"] * total_variation_loss(normed) if self.w["tv"] > 0 else zero
        if self.w["edge"] > 0:
            e = self.w["edge"] * edge_loss(self.edge_target, get_gradient_imgs(to_grayscale(xv)))
        else:
            e = zero
        total = torch.zeros(1, dtype=self.dtype, device=self.device) + s + c + tv + e     # run_style_transfer.py:139
        grad = None
        if need_grad:
            total.backward()
            grad = xv.grad.detach()
        return dict(total=float(total.detach()), content=float(c.detach()), style=float(s.detach()), tv=float(tv.detach()),
                    edge=float(e.detach()), gram_mse=[float(v.detach()) for v in per], grad=grad,
                    grams={k: gram_matrix(feats[k]).detach() for k in self.style_layers},
                    feats={k: v.detach() for k, v in feats.items()})


def run_oracle(weights, biases, content_u8: np.ndarray, style_u8: List[np.ndarray], num_steps: int, *,
               random_init=False, w_style, w_content, w_tv, w_edge, style_img_weight=0.5, channel_attention_on=False,
               mean=VGG_MEAN, std=VGG_STD, dtype=torch.float32, keep_iterates=False, max_evals=None,
               emulate_reference_cost=False, on_eval=None, device="cpu") -> OracleResult:
    """run_multi_style_transfer (run_style_transfer.py:27-159) on uint8 HWC arrays.  `device`: where the reference's
    torch ops run - "cpu" for every parity use; bench.py's torch_eager_gpu leg passes "cuda" to time the reference's
    own eager path (cuDNN / cuBLAS kernels picked by torch) on the B200 beside the CUDA path of this repo."""
    torch.manual_seed(101)                                                  # :52 seed_everything
    np.random.seed(101)
    weights = [w.to(device, dtype) for w in weights]
    biases = [b.to(device, dtype) for b in biases]
    content = to_tensor_u8(content_u8, dtype).to(device)                    # :59 .to(device)
    styles = [to_tensor_u8(s, dtype).to(device) for s in style_u8]
    if random_init:
        x0 = torch.randn(content.shape).to(device, dtype)                   # :84
    else:
        x0 = content.clone()                                                # :87
    co = ClosureOracle(weights, biases, content, styles, w_style=w_style, w_content=w_content, w_tv=w_tv, w_edge=w_edge,
                       style_img_weight=style_img_weight, channel_attention_on=channel_attention_on, mean=mean, std=std,
                       ca_seed=None if random_init else 101, emulate_reference_cost=emulate_reference_cost)
    x = x0.reshape(-1).clone()
    shape = content.shape
    opt = LbfgsOracle(x)
    res = OracleResult(image=None, losses=[], iterates=[] if keep_iterates else None, grads=[] if keep_iterates else None)
    counter = [0]

    class _Stop(Exception):
        pass

    stop_req = [False]

    def closure():
        if stop_req[0] or (max_evals is not None and counter[0] >= max_evals):
            raise _Stop()
        opt.x.clamp_(0, 1)                                                  # :108-109
        out = co.evaluate(opt.x.reshape(shape))
        counter[0] += 1                                                     # :143
        if on_eval is not None:
            stop_req[0] = bool(on_eval(counter[0]))     # a true return value ends the run before the next evaluation
        res.losses.append([out["total"], out["content"], out["style"], out["tv"], out["edge"]])
        if keep_iterates:
            res.iterates.append(opt.x.reshape(shape).clone())
            res.grads.append(out["grad"].clone())
        return out["total"], out["grad"].reshape(-1)

    try:
        while counter[0] <= num_steps:                                      # :100
            opt.step(closure)
    except _Stop:
        pass
    opt.x.clamp_(0, 1)                                                      # :154-155
    res.image = opt.x.reshape(shape).clone()
    res.evals = counter[0]
    return res


def psnr(a: torch.Tensor, b: torch.Tensor) -> float:
    mse = float(((a.double() - b.double()) ** 2).mean())
    return float("inf") if mse == 0 else 10.0 * math.log10(1.0 / mse)
