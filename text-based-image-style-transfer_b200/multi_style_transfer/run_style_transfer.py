"""Mirror of multi_style_transfer/run_style_transfer.py.

run_multi_style_transfer keeps the reference's 15-parameter signature, PIL in / PIL out, its two prints and
its loop semantics (run_style_transfer.py:27-159): `num_steps` counts closure evaluations, tested between
optimizer.step() calls, each of which performs up to 20 evaluations.  The work happens on the device:
one CUDA-graph launch per optimizer.step(), no host synchronisation inside it.

StyleTransferSession is the reusable form used for video frames and benchmarks: it keeps the VGG trunk,
the style Gram targets and the per-resolution plan alive across calls (the reference rebuilds all three
per call, app.py:794-798).
"""
import contextlib
import threading
from typing import List, Optional

import numpy as np
import torch

from .. import engine
from ..engine import Plan, _require_cuda
from .helper_functions import get_net, seed_everything

CONTENT_LAYERS = ['conv4_2']                                              # run_style_transfer.py:56
STYLE_LAYERS = ['conv1_1', 'conv2_1', 'conv3_1', 'conv4_1', 'conv5_1']    # run_style_transfer.py:57


def PIL_to_tensor(image):
    """transforms.ToTensor() + unsqueeze(0) (run_style_transfer.py:5-11): HWC uint8 -> (1,C,H,W) float in [0,1]."""
    arr = np.asarray(image)
    if arr.ndim == 2:
        arr = arr[:, :, None]
    t = torch.from_numpy(np.ascontiguousarray(arr)).permute(2, 0, 1).contiguous()
    if t.dtype == torch.uint8:
        t = t.to(torch.float32).div(255)
    return t.unsqueeze(0)


def tensor_to_PIL(img):
    """transforms.ToPILImage() of a (3,H,W) float tensor (run_style_transfer.py:157): mul(255).byte() truncates."""
    from PIL import Image
    arr = img.detach().cpu().mul(255).byte().permute(1, 2, 0).contiguous().numpy()
    return Image.fromarray(arr)


def _as_tensor3(v, what):
    t = torch.as_tensor(v, dtype=torch.float32).reshape(-1).cpu()
    if t.numel() != 3:
        raise ValueError("%s must have 3 entries" % what)
    return [float(x) for x in t]


def channel_attention_weights(channels, reduction_ratio=2):
    """The two bias-free nn.Linear weights ChannelAttention.__init__ draws (ChannelAttention.py:16-17), from
    the CPU generator like `ChannelAttention(c).to(device)` does."""
    fc1 = torch.nn.Linear(channels, channels // reduction_ratio, bias=False)
    fc2 = torch.nn.Linear(channels // reduction_ratio, channels, bias=False)
    return fc1.weight.detach(), fc2.weight.detach()


CA_SEED = 101   # seed_everything(101), run_style_transfer.py:52


def channel_attention_gate_weights(channel_counts, seed=CA_SEED):
    """Gate weights of channel_att_per_chosen_layers (run_style_transfer.py:13-25): one fresh ChannelAttention per content
    layer, drawn in layer order from the CPU generator as it stands right after seed_everything(101) - the reference
    reseeds at the top of every call (:52) and, on a CUDA device, draws nothing from the CPU generator before the modules
    (its random_init randn is a device draw, :84).  The draw happens inside fork_rng, so it neither depends on nor disturbs
    the caller's generator: identical inputs give identical gates on every call and on every path (session, frames, planes).
    With downloaded ImageNet weights the reference's generator has additionally been advanced by torchvision's VGG
    constructor (random init before load_state_dict); that advance is not emulated - the gate is an untrained random
    module either way and the golden vectors of tests/golden pin the first-draw convention."""
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed)
        return [channel_attention_weights(c) for c in channel_counts]


class StyleTransferSession:
    """VGG trunk + style targets + plan for one (content resolution, style set, weights) configuration."""

    def __init__(self, vgg_mean, vgg_std, content_hw, style_imgs: List[torch.Tensor], w_style, w_content, w_tv, w_edge,
                 style_img_weight=0.5, device="cuda", content_layers=CONTENT_LAYERS, style_layers=STYLE_LAYERS,
                 style_targets: Optional[dict] = None, net=None):
        self.device = _require_cuda(device)
        self.mean = _as_tensor3(vgg_mean, "vgg_mean")
        self.std = _as_tensor3(vgg_std, "vgg_std")
        self.content_layers = list(content_layers)
        self.style_layers = list(style_layers)
        self.weights = (float(w_style), float(w_content), float(w_tv), float(w_edge))
        self.net = net if net is not None else get_net(self.device)
        # high priority: the library's side stream (pixel terms, shallow Gram work) is created with the lowest priority, so
        # that when an SM frees up the block scheduler gives it to the critical path (conv chain, optimizer) first
        self.stream = torch.cuda.Stream(self.device, priority=-1)
        H, W = int(content_hw[0]), int(content_hw[1])
        self._sync_with_caller()
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            self.plan = Plan(self.net, H, W, self.content_layers + self.style_layers, self.style_layers,
                             self.content_layers, with_grad=True, mean=self.mean, std=self.std)
            self.plan.set_weights(*self.weights)
            if style_targets is None:
                style_targets = self.compute_style_targets(style_imgs, style_img_weight)
            self.style_targets = {k: v.to(self.device) for k, v in style_targets.items()}
            for name in self.style_layers:
                self.plan.set_style_target(name, self.style_targets[name])
            self.stream.synchronize()

    def _sync_with_caller(self, *tensors):
        """The session works on its own non-blocking stream; inputs were produced on the caller's current stream.  Make
        the session stream wait for everything enqueued there so far and tell the caching allocator that `tensors` are in
        use on the session stream."""
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        for t in tensors:
            if isinstance(t, torch.Tensor) and t.is_cuda:
                t.record_stream(self.stream)

    def compute_style_targets(self, style_imgs, style_img_weight):
        """Gram targets of style_loss (style_transfer_losses.py:122-135), once instead of once per evaluation."""
        if torch.cuda.current_stream(self.device) != self.stream:
            self._sync_with_caller(*style_imgs)
        plans = []
        for img in style_imgs:
            h, w = int(img.shape[2]), int(img.shape[3])
            if (h, w) == (self.plan.H, self.plan.W) and not plans:
                p = self.plan  # same resolution: reuse the main plan's buffers
            else:
                p = Plan(self.net, h, w, self.style_layers, mean=self.mean, std=self.std)
            p.features(img)
            plans.append(p)
        out = {}
        for name in self.style_layers:
            if len(plans) == 1:
                out[name] = plans[0].tap_gram(name)          # :127-129 (style_img_weight ignored)
            else:
                out[name] = engine.style_mix_gram(plans[0], plans[1], name, style_img_weight)  # :130-135
        torch.cuda.current_stream(self.device).synchronize()
        for p in plans:
            if p is not self.plan:
                p.close()
        return out

    def prepare(self, content: torch.Tensor, x0: Optional[torch.Tensor] = None, channel_attention=False,
                trace_capacity=0):
        """Content / edge targets from the content image and optimizer reset (run_style_transfer.py:71-96)."""
        self._sync_with_caller(content, x0)
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            plan = self.plan
            plan.features(content)
            gates = None
            if channel_attention:
                gates = channel_attention_gate_weights([plan.tap_shape(name)[0] for name in self.content_layers])   # :13-25
            for k, name in enumerate(self.content_layers):
                gate = None
                if gates is not None:
                    gate = plan.channel_gate(name, *gates[k])
                plan.set_content_target(name, plan, gate)
            if self.weights[3] > 0:
                plan.set_edge_target(content)
            plan.lbfgs_init(content if x0 is None else x0, trace_capacity)

    def prepare_graph(self):
        """Captures the CUDA graph of optimizer.step() now (otherwise the first step() does)."""
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            self.plan.lbfgs_prepare_graph()

    def step(self):
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            self.plan.lbfgs_step()

    def status(self):
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            return self.plan.lbfgs_status()

    def run(self, num_steps, print_iter=None):
        """The training loop (run_style_transfer.py:98-151).  Returns the number of closure evaluations."""
        calls = 0
        guard = 0
        while calls <= num_steps:
            self.step()
            st = self.status()
            if st.stop == 6:
                raise engine.NstError("non-finite loss at evaluation %d" % st.closure_calls)
            if print_iter:
                for k in range(calls + 1, st.closure_calls + 1):
                    if k % print_iter == 0:
                        print(f'Reached iteration {k}/{num_steps}')
            calls = st.closure_calls
            guard += 1
            if guard > num_steps + 2:
                break
        return calls

    def result(self) -> torch.Tensor:
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            x = self.plan.lbfgs_x()
            self.stream.synchronize()
        return x

    def trace(self, max_rows=4096) -> torch.Tensor:
        with torch.cuda.device(self.device), torch.cuda.stream(self.stream):
            return self.plan.lbfgs_trace(max_rows)

    def close(self):
        self.plan.close()


_RUN_LOCK = threading.RLock()


def run_multi_style_transfer(vgg_mean, vgg_std, content_img, num_steps, random_init, w_style, w_content, w_tv,
                             w_edge, style_img1, style_img2=None, style_img_weight=0.5, print_iter=50,
                             channel_attention=False, device="cpu"):
    """ Neural Style Transfer optimization procedure (reference: run_style_transfer.py:27-159).

    Same parameters and return value as the reference: PIL images in, the style-transferred PIL image out.
    `device` must be a CUDA device (a B200); the reference's default "cpu" raises because this implementation
    has no CPU path.
    """
    dev = _require_cuda(device)
    seed_everything(101)                                                   # :52

    content = PIL_to_tensor(content_img).to(dev)                           # :59
    styles = [PIL_to_tensor(style_img1).to(dev)]                           # :60
    if style_img2:
        styles.append(PIL_to_tensor(style_img2).to(dev))                   # :61-63
    if content.shape[1] != 3 or any(s.shape[1] != 3 for s in styles):
        raise RuntimeError("The size of tensor a (%d) must match the size of tensor b (3) at non-singleton "
                           "dimension 1" % content.shape[1])               # normalize() broadcast error of the reference

    # One optimisation at a time per process: the reference is not re-entrant either (it reseeds the global generators, :52),
    # and two chains of launches stepping side by side on one GPU need nst_plan_set_shared_gpu - callers
    # that want several images at once use video.FrameStyler(batch= / concurrent=).
    with _RUN_LOCK:
        session = StyleTransferSession(vgg_mean, vgg_std, content.shape[2:], styles, w_style, w_content, w_tv, w_edge,
                                       style_img_weight, dev)
        try:
            if random_init:
                x0 = torch.randn(content.size(), device=dev)                   # :84
            else:
                x0 = None                                                      # :87 content.clone()
            print("Channel attention enabled: " + str(channel_attention))      # :92
            evals = 20 * (int(num_steps) // 20 + 1)
            session.prepare(content, x0, channel_attention, trace_capacity=evals + 32)
            session.run(int(num_steps), print_iter)
            out = session.result()
        finally:
            session.close()
    return tensor_to_PIL(out[0])                                           # :157


# ------------------------------------------------------------------------------------------------------------------
# Tutorial-style compatibility surface named by BASELINE.json's north_star.  The reference itself only carries a
# commented-out import of it (basic.py:12: `from style_transfer.run_style_transfer import run_style_transfer`); SURVEY.md
# section 0 / 8(b) lists it as optional sugar over run_multi_style_transfer.  It maps onto the SAME device loop with the
# pixel-space terms switched off (w_tv = w_edge = 0) and the reference's layer set and loss definitions
# (style_transfer_losses.py: per-layer MSEs averaged over the layers).
# ------------------------------------------------------------------------------------------------------------------
def _net_from_cnn(cnn, device):
    """`cnn`: None (the installed VGG-19 provider), a helper_functions.Vgg19 mirror, an engine.Net, or a torchvision-style
    `vgg19().features` module (its Conv2d parameters, in order, are handed to the library)."""
    from ..engine import Net
    from .helper_functions import Vgg19
    if cnn is None:
        return get_net(device)
    if isinstance(cnn, Net):
        return cnn
    if isinstance(cnn, Vgg19):
        return cnn.net
    convs = [m for m in cnn.modules() if isinstance(m, torch.nn.Conv2d)]
    if len(convs) < 13:
        raise ValueError("cnn must contain the first 13 convolutions of VGG-19 `features` (found %d Conv2d)" % len(convs))
    return Net([c.weight.detach() for c in convs[:16]], [c.bias.detach() for c in convs[:16]], device)


class ContentLoss(torch.nn.Module):
    """Transparent layer of the tutorial contract: forward(input) stores `self.loss` = MSE(input, target) - the per-layer
    term of content_loss (style_transfer_losses.py:53-65) - and returns its input unchanged.  `self.loss` is a plain CUDA
    scalar computed by the library (no autograd graph: the optimisation loop differentiates inside the fused closure)."""

    def __init__(self, target):
        super().__init__()
        self.target = target.detach()
        self.loss = None

    def forward(self, input):
        from . import style_transfer_losses as L
        self.loss = L._mse(input, self.target)
        return input


class StyleLoss(torch.nn.Module):
    """Transparent layer of the tutorial contract: the target is the Gram matrix of `target_feature`
    (gram_matrix, style_transfer_losses.py:70-95); forward(input) stores `self.loss` = MSE(gram(input), target) - the
    per-layer term of style_loss (:138-144) - and returns its input unchanged."""

    def __init__(self, target_feature):
        super().__init__()
        from . import style_transfer_losses as L
        self.target = L.gram_matrix(target_feature).detach()
        self.loss = None

    def forward(self, input):
        from . import style_transfer_losses as L
        self.loss = L._mse(L.gram_matrix(input), self.target)
        return input


def run_style_transfer(cnn, normalization_mean, normalization_std, content_img, style_img, input_img, num_steps=300,
                       style_weight=1000000, content_weight=1):
    """run_style_transfer(cnn, norm_mean, norm_std, content_img, style_img, input_img, num_steps, style_weight,
    content_weight) -> the optimised image tensor (1,3,H,W), clamped to [0,1].

    Tensors in, tensor out (CUDA).  Equivalent to run_multi_style_transfer (run_style_transfer.py:27-159) with
    w_style = style_weight, w_content = content_weight, w_tv = w_edge = 0, a single style, no channel attention, and the
    optimisation started from `input_img` instead of the content image / randn (:83-87); `num_steps` counts closure
    evaluations between optimizer.step() calls exactly as there (:99-100).  `input_img` is updated in place, like the
    parameter the tutorial optimises."""
    dev = _require_cuda(content_img.device)
    if content_img.dim() != 4 or content_img.shape[0] != 1 or content_img.shape[1] != 3:
        raise ValueError("content_img must be a (1,3,H,W) tensor")
    if tuple(input_img.shape) != tuple(content_img.shape):
        raise ValueError("input_img must have the shape of content_img")
    session = StyleTransferSession(normalization_mean, normalization_std, content_img.shape[2:], [style_img.to(dev)],
                                   float(style_weight), float(content_weight), 0.0, 0.0, device=dev, net=_net_from_cnn(cnn, dev))
    try:
        evals = 20 * (int(num_steps) // 20 + 1)
        session.prepare(content_img, input_img.to(dev), False, trace_capacity=evals + 32)
        session.run(int(num_steps))
        out = session.result()
    finally:
        session.close()
    with torch.no_grad():
        if isinstance(input_img, torch.Tensor) and input_img.is_cuda and not input_img.requires_grad:
            input_img.copy_(out)
        elif isinstance(input_img, torch.Tensor) and input_img.is_cuda:
            input_img.data.copy_(out)
    return out
