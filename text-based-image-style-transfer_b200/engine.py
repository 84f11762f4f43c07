"""Thin object layer over the C ABI: Net (VGG-19 trunk) and Plan (per-resolution state).  torch is used
for device memory and streams only; every computation happens inside libnst_b200.so."""
import ctypes as C
from typing import Dict, List, Optional, Sequence

import torch

from . import _lib
from ._lib import NstError, NstStatus, check, f3

CONV_NAMES = [
    "conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3", "conv3_4",
    "conv4_1", "conv4_2", "conv4_3", "conv4_4", "conv5_1", "conv5_2", "conv5_3", "conv5_4",
]
_CIN = [3, 64, 64, 128, 128, 256, 256, 256, 256, 512, 512, 512, 512, 512, 512, 512]
_COUT = [64, 64, 128, 128, 256, 256, 256, 256, 512, 512, 512, 512, 512, 512, 512, 512]


def _stream_ptr(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _require_cuda(device) -> torch.device:
    device = torch.device(device)
    if device.type != "cuda":
        raise NstError("this implementation runs on a B200 only (device=%s requested); there is no CPU path" % device)
    if not torch.cuda.is_available():
        raise NstError("no CUDA device is available; there is no CPU path")
    if device.index is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return device


def mask_of(names: Sequence[str]) -> int:
    m = 0
    for n in names:
        if n not in CONV_NAMES:
            raise Exception("Not all layers provided in content_layes and/or style_layers exist.")
        m |= 1 << CONV_NAMES.index(n)
    return m


def _f32c(t: torch.Tensor, device) -> torch.Tensor:
    return t.detach().to(device=device, dtype=torch.float32).contiguous()


class Net:
    """VGG-19 `features` convolutions, repacked on the device (nst_net)."""

    def __init__(self, weights: List[torch.Tensor], biases: List[torch.Tensor], device="cuda"):
        self.lib = _lib.load()
        self.device = _require_cuda(device)
        n = len(weights)
        if n < 1 or n > _lib.NST_MAX_CONV or len(biases) != n:
            raise ValueError("expected 1..16 conv weights with matching biases")
        for i in range(n):
            if tuple(weights[i].shape) != (_COUT[i], _CIN[i], 3, 3) or tuple(biases[i].shape) != (_COUT[i],):
                raise ValueError("conv %d: unexpected parameter shape %s" % (i, tuple(weights[i].shape)))
        with torch.cuda.device(self.device):
            ws = [_f32c(w, self.device) for w in weights]
            bs = [_f32c(b, self.device) for b in biases]
            wp = (C.c_void_p * n)(*[w.data_ptr() for w in ws])
            bp = (C.c_void_p * n)(*[b.data_ptr() for b in bs])
            h = C.c_void_p()
            check(self.lib.nst_net_create(C.byref(h), wp, bp, n, _stream_ptr(self.device)))
        self.handle = h
        self.n_conv = n

    def close(self):
        if getattr(self, "handle", None):
            self.lib.nst_net_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Plan:
    """Buffers, tensor maps, targets and optimizer state for one image resolution (nst_plan)."""

    def __init__(self, net: Net, H: int, W: int, taps: Sequence[str], style_layers: Sequence[str] = (),
                 content_layers: Sequence[str] = (), with_grad: bool = False, mean=(0.0, 0.0, 0.0), std=(1.0, 1.0, 1.0)):
        self.lib = net.lib
        self.net = net
        self.device = net.device
        self.H, self.W = int(H), int(W)
        self.taps = [n for n in CONV_NAMES if n in set(taps) | set(style_layers) | set(content_layers)]
        self.style_layers = list(style_layers)
        self.content_layers = list(content_layers)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.nst_plan_create(C.byref(h), net.handle, self.H, self.W, mask_of(self.taps),
                                           mask_of(style_layers), mask_of(content_layers), 1 if with_grad else 0))
        self.handle = h
        self.set_norm(mean, std)

    # -- configuration
    def set_norm(self, mean, std):
        check(self.lib.nst_plan_set_norm(self.handle, f3(mean), f3(std)))

    def set_weights(self, w_style, w_content, w_tv, w_edge):
        check(self.lib.nst_plan_set_weights(self.handle, float(w_style), float(w_content), float(w_tv), float(w_edge)))

    def bytes(self) -> int:
        return int(self.lib.nst_plan_bytes(self.handle))

    def set_shared_gpu(self, shared: bool = True):
        """Mark a plan that steps at the same time as other plans on this GPU (own host thread / stream): its CTA-pair launches
        are then not chained by programmatic dependent launch (nst_plan_set_shared_gpu).  nst_run_frames_host does it itself."""
        check(self.lib.nst_plan_set_shared_gpu(self.handle, 1 if shared else 0))

    # -- forward
    def _img(self, x: torch.Tensor) -> torch.Tensor:
        x = _f32c(x, self.device)
        if x.dim() == 4:
            if x.shape[0] != 1:
                raise ValueError("batch size must be 1")
            x = x[0]
        if tuple(x.shape) != (3, self.H, self.W):
            raise ValueError("expected a (1,3,%d,%d) image, got %s" % (self.H, self.W, tuple(x.shape)))
        return x

    def features(self, x: torch.Tensor):
        x = self._img(x)
        with torch.cuda.device(self.device):
            check(self.lib.nst_plan_features(self.handle, _ptr(x), _stream_ptr(self.device)))
        self._keep = x

    def tap_shape(self, name: str):
        c, h, w = C.c_int(), C.c_int(), C.c_int()
        check(self.lib.nst_plan_tap_shape(self.handle, CONV_NAMES.index(name), C.byref(c), C.byref(h), C.byref(w)))
        return c.value, h.value, w.value

    def get_tap(self, name: str) -> torch.Tensor:
        c, h, w = self.tap_shape(name)
        out = torch.empty((1, c, h, w), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(self.lib.nst_plan_get_tap(self.handle, CONV_NAMES.index(name), _ptr(out), _stream_ptr(self.device)))
        return out

    def tap_gram(self, name: str) -> torch.Tensor:
        c, _, _ = self.tap_shape(name)
        out = torch.empty((1, c, c), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(self.lib.nst_plan_tap_gram(self.handle, CONV_NAMES.index(name), _ptr(out), _stream_ptr(self.device)))
        return out

    # -- targets
    def set_style_target(self, name: str, gram: torch.Tensor):
        g = _f32c(gram, self.device)
        with torch.cuda.device(self.device):
            check(self.lib.nst_plan_set_style_target(self.handle, CONV_NAMES.index(name), _ptr(g), _stream_ptr(self.device)))
            torch.cuda.current_stream(self.device).synchronize()

    def set_content_target(self, name: str, src: "Plan", gate: Optional[torch.Tensor] = None):
        g = None if gate is None else _f32c(gate, self.device)
        with torch.cuda.device(self.device):
            check(self.lib.nst_plan_set_content_target(self.handle, CONV_NAMES.index(name), src.handle, _ptr(g),
                                                       _stream_ptr(self.device)))
            torch.cuda.current_stream(self.device).synchronize()

    def channel_gate(self, name: str, w1: torch.Tensor, w2: torch.Tensor, reduction: int = 2) -> torch.Tensor:
        c, _, _ = self.tap_shape(name)
        w1 = _f32c(w1, self.device)
        w2 = _f32c(w2, self.device)
        gate = torch.empty(c, device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(self.lib.nst_plan_channel_gate(self.handle, CONV_NAMES.index(name), _ptr(w1), _ptr(w2), reduction,
                                                 _ptr(gate), _stream_ptr(self.device)))
            torch.cuda.current_stream(self.device).synchronize()
        return gate

    def set_edge_target(self, content: torch.Tensor):
        x = self._img(content)
        with torch.cuda.device(self.device):
            check(self.lib.nst_plan_set_edge_target(self.handle, _ptr(x), _stream_ptr(self.device)))
            torch.cuda.current_stream(self.device).synchronize()

    # -- closure
    def eval(self, x: torch.Tensor, need_grad: bool = True):
        x = self._img(x)
        losses = torch.empty(_lib.NST_LOSS_COUNT, device=self.device, dtype=torch.float32)
        grad = torch.empty((1, 3, self.H, self.W), device=self.device, dtype=torch.float32) if need_grad else None
        with torch.cuda.device(self.device):
            check(self.lib.nst_plan_eval(self.handle, _ptr(x), _ptr(losses), _ptr(grad), _stream_ptr(self.device)))
            torch.cuda.current_stream(self.device).synchronize()
        return losses, grad

    # -- optimizer
    def lbfgs_init(self, x0: torch.Tensor, trace_capacity: int = 0):
        x = self._img(x0)
        with torch.cuda.device(self.device):
            check(self.lib.nst_lbfgs_init(self.handle, _ptr(x), int(trace_capacity), _stream_ptr(self.device)))
            torch.cuda.current_stream(self.device).synchronize()

    def lbfgs_step(self):
        with torch.cuda.device(self.device):
            check(self.lib.nst_lbfgs_step(self.handle, _stream_ptr(self.device)))

    def lbfgs_prepare_graph(self):
        """Captures the CUDA graph of optimizer.step() now instead of inside the first lbfgs_step()."""
        with torch.cuda.device(self.device):
            check(self.lib.nst_lbfgs_prepare_graph(self.handle, _stream_ptr(self.device)))

    def lbfgs_status(self) -> NstStatus:
        st = NstStatus()
        with torch.cuda.device(self.device):
            check(self.lib.nst_lbfgs_status(self.handle, C.byref(st), _stream_ptr(self.device)))
        return st

    def lbfgs_x(self) -> torch.Tensor:
        out = torch.empty((1, 3, self.H, self.W), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            check(self.lib.nst_lbfgs_get_x(self.handle, _ptr(out), _stream_ptr(self.device)))
        return out

    def lbfgs_trace(self, max_rows: int) -> torch.Tensor:
        buf = torch.empty((max(max_rows, 1), 5), dtype=torch.float32)
        with torch.cuda.device(self.device):
            rows = check(self.lib.nst_lbfgs_trace(self.handle, C.c_void_p(buf.data_ptr()), int(max_rows),
                                                  _stream_ptr(self.device)))
        return buf[:rows].clone()

    def lbfgs_partial_step(self, n_evals: int) -> int:
        with torch.cuda.device(self.device):
            return check(self.lib.nst_lbfgs_partial_step(self.handle, int(n_evals), _stream_ptr(self.device)))

    def eval_timed(self, x: torch.Tensor):
        """One evaluation with a CUDA event after every launch -> list of (kind name, conv index, ms)."""
        x = self._img(x)
        grad = torch.empty((1, 3, self.H, self.W), device=self.device, dtype=torch.float32)
        buf = (_lib.NstLaunchTime * 128)()
        with torch.cuda.device(self.device):
            n = check(self.lib.nst_plan_eval_timed(self.handle, _ptr(x), _ptr(grad), buf, 128, _stream_ptr(self.device)))
        return [(_lib.KIND_NAMES.get(buf[i].kind, str(buf[i].kind)), buf[i].layer, buf[i].ms) for i in range(n)]

    def lbfgs_step_timed(self, grouped: bool = False):
        """One whole optimizer.step() with a CUDA event after every launch -> list of (kind name, conv index, ms); grouped:
        one event where the kind of launch changes -> list of (kind name, launches in the run, ms of the run)."""
        buf = (_lib.NstLaunchTime * 1200)()
        fn = self.lib.nst_lbfgs_step_timed_grouped if grouped else self.lib.nst_lbfgs_step_timed
        with torch.cuda.device(self.device):
            n = check(fn(self.handle, buf, 1200, _stream_ptr(self.device)))
        return [(_lib.KIND_NAMES.get(buf[i].kind, str(buf[i].kind)), buf[i].layer, buf[i].ms) for i in range(n)]

    def launches_per_step(self) -> int:
        return int(self.lib.nst_lbfgs_launches_per_step(self.handle))

    def run_frame_host(self, content_u8, out_u8, num_steps: int, ca_w1=None, ca_w2=None) -> int:
        """content_u8 / out_u8: contiguous uint8 [H,W,3] host tensors (pinned for async copies)."""
        with torch.cuda.device(self.device):
            return check(self.lib.nst_run_frame_host(self.handle, C.c_void_p(content_u8.data_ptr()),
                                                     C.c_void_p(out_u8.data_ptr()), int(num_steps),
                                                     1 if ca_w1 is not None else 0, _ptr(ca_w1), _ptr(ca_w2),
                                                     _stream_ptr(self.device)))

    def close(self):
        if getattr(self, "handle", None):
            if getattr(self, "_owned", True):
                self.lib.nst_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def _wrap(cls, net: "Net", handle, H, W, taps, style_layers, content_layers, owned):
        """A Plan object around an existing nst_plan handle (a batch head or one of its members)."""
        self = cls.__new__(cls)
        self.lib = net.lib
        self.net = net
        self.device = net.device
        self.H, self.W = int(H), int(W)
        self.taps = [n for n in CONV_NAMES if n in set(taps) | set(style_layers) | set(content_layers)]
        self.style_layers = list(style_layers)
        self.content_layers = list(content_layers)
        self.handle = handle if isinstance(handle, C.c_void_p) else C.c_void_p(handle)
        self._owned = owned
        return self


class BatchPlan:
    """Several images per launch (nst_batch_create, SURVEY 8 f row 2): `head` steps all images with ONE set of convolution /
    Gram launches per evaluation; `members[k]` is image k's single-image plan over its slice of the head's tensors (targets,
    optimizer state, status, trace, result).  The reference evaluates one image per call (run_style_transfer.py:27-159); its
    gram_matrix already divides by the batch size (style_transfer_losses.py:84-93) and apply_video_process loops over
    independent frames (app.py:784-815) - this is that loop, `batch` frames at a time."""

    def __init__(self, net: Net, H: int, W: int, style_layers: Sequence[str], content_layers: Sequence[str], batch: int,
                 mean=(0.0, 0.0, 0.0), std=(1.0, 1.0, 1.0)):
        self.lib = net.lib
        self.net = net
        self.device = net.device
        self.batch = int(batch)
        taps = list(content_layers) + list(style_layers)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            check(self.lib.nst_batch_create(C.byref(h), net.handle, int(H), int(W), mask_of([n for n in CONV_NAMES if n in set(taps)]),
                                            mask_of(style_layers), mask_of(content_layers), self.batch))
        self.head = Plan._wrap(net, h, H, W, taps, style_layers, content_layers, owned=True)
        self.members = []
        for k in range(self.batch):
            m = self.lib.nst_batch_member(h, k)
            if not m:
                raise NstError(self.lib.nst_last_error().decode())
            self.members.append(Plan._wrap(net, m, H, W, taps, style_layers, content_layers, owned=False))
        self.head.set_norm(mean, std)

    def set_weights(self, w_style, w_content, w_tv, w_edge):
        self.head.set_weights(w_style, w_content, w_tv, w_edge)   # forwarded to the members by the library

    def set_style_targets(self, targets: dict):
        for m in self.members:
            for name in m.style_layers:
                m.set_style_target(name, targets[name])

    def step(self):
        """One optimizer.step() of every member that is not frozen: one CUDA-graph launch."""
        self.head.lbfgs_step()

    def freeze(self, k: int, frozen: bool = True):
        with torch.cuda.device(self.device):
            check(self.lib.nst_lbfgs_freeze(self.members[k].handle, 1 if frozen else 0, _stream_ptr(self.device)))

    def run_batch_host(self, ins, outs, num_steps: int, ca_w1=None, ca_w2=None):
        """ins / outs: lists (<= batch) of contiguous pinned uint8 [H,W,3] host tensors; returns the evaluations each frame ran."""
        cnt = len(ins)
        pi = (C.c_void_p * cnt)(*[t.data_ptr() for t in ins])
        po = (C.c_void_p * cnt)(*[t.data_ptr() for t in outs])
        calls = (C.c_int * cnt)()
        with torch.cuda.device(self.device):
            check(self.lib.nst_run_batch_host(self.head.handle, cnt, pi, po, int(num_steps), 1 if ca_w1 is not None else 0,
                                              _ptr(ca_w1), _ptr(ca_w2), _stream_ptr(self.device), calls))
        return list(calls)

    def bytes(self) -> int:
        return self.head.bytes()

    def close(self):
        for m in self.members:
            m.handle = None
        self.members = []
        if getattr(self, "head", None) is not None:
            self.head.close()
            self.head = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def gram_chw(x: torch.Tensor) -> torch.Tensor:
    """gram_matrix of a (1, C, H, W) CUDA tensor through the tcgen05 Gram kernel."""
    lib = _lib.load()
    device = _require_cuda(x.device)
    if x.dim() != 4 or x.shape[0] != 1:
        raise NstError("gram_matrix: only batch size 1 is implemented on this path")
    xc = _f32c(x, device)
    _, c, h, w = xc.shape
    out = torch.empty((1, c, c), device=device, dtype=torch.float32)
    with torch.cuda.device(device):
        check(lib.nst_gram_chw(_ptr(xc), c, h, w, _ptr(out), _stream_ptr(device)))
    return out


def style_mix_gram(a: Plan, b: Plan, name: str, weight_b: float) -> torch.Tensor:
    c, _, _ = a.tap_shape(name)
    out = torch.empty((1, c, c), device=a.device, dtype=torch.float32)
    with torch.cuda.device(a.device):
        check(a.lib.nst_style_mix_gram(a.handle, b.handle, CONV_NAMES.index(name), float(weight_b), _ptr(out),
                                       _stream_ptr(a.device)))
    return out


def style_mix_chw(a: Plan, b: Plan, name: str, weight_b: float) -> torch.Tensor:
    c, ha, wa = a.tap_shape(name)
    _, hb, wb = b.tap_shape(name)
    out = torch.empty((1, c, ha + hb // 2, wa + wb // 2), device=a.device, dtype=torch.float32)
    with torch.cuda.device(a.device):
        check(a.lib.nst_style_mix_chw(a.handle, b.handle, CONV_NAMES.index(name), float(weight_b), _ptr(out),
                                      _stream_ptr(a.device)))
    return out
