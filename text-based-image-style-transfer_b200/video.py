"""Frame / image-pair sharding across the GPUs of one box.

The path shards across independent images only: one optimisation is a strictly sequential chain, but video frames
(apply_video_process, app.py:784-815) and batches of content/style pairs are independent - the reference processes
them one after another, rebuilding the VGG trunk and the style targets for every frame (app.py:794-798).  Here every
rank (one process per GPU, torch.distributed) takes a contiguous block of frames, the style Gram targets are computed
once on rank 0 and broadcast (5 matrices, 2.44 MB), and the finished uint8 frames are all-gathered in frame order.
No collective sits inside the optimisation step.  The JPEG round trip of app.py:790-791 and the video container I/O
(app.py:777-779, 843-859) are outside this path.
"""
from typing import Callable, Dict, List, Optional, Sequence

import torch

STYLE_CHANNELS = {"conv1_1": 64, "conv2_1": 128, "conv3_1": 256, "conv4_1": 512, "conv5_1": 512}


def shard_range(n_items: int, world: int, rank: int):
    """Contiguous block [lo, hi) of rank `rank`: blocks differ by at most one item, earlier ranks take the remainder."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


def broadcast_style_targets(targets: Optional[Dict[str, torch.Tensor]], layers: Sequence[str], device, src: int = 0):
    """Rank `src` passes its Gram targets {layer: (1,C,C)}; every rank returns the same dict.  One broadcast of the
    concatenated matrices (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return targets
    sizes = [STYLE_CHANNELS[n] for n in layers]
    flat = torch.empty(sum(c * c for c in sizes), device=device, dtype=torch.float32)
    if dist.get_rank() == src:
        flat.copy_(torch.cat([targets[n].reshape(-1).to(device=device, dtype=torch.float32) for n in layers]))
    dist.broadcast(flat, src=src)
    out, off = {}, 0
    for n, c in zip(layers, sizes):
        out[n] = flat[off:off + c * c].reshape(1, c, c).clone()
        off += c * c
    return out


def gather_frames(local: torch.Tensor, n_total: int, device) -> Optional[torch.Tensor]:
    """local: this rank's finished frames (k, H, W, 3) uint8 in frame order.  Returns all frames (n_total, H, W, 3)
    on every rank (all_gather of equal-sized, zero-padded blocks; 2.76 MB per 720p frame)."""
    dist = _dist()
    if dist is None or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    per = (n_total + world - 1) // world
    shape = tuple(local.shape[1:])
    pad = torch.zeros((per,) + shape, dtype=torch.uint8, device=device)
    pad[:local.shape[0]].copy_(local.to(device))
    out = torch.empty((world * per,) + shape, dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(out, pad)
    pieces = []
    for r in range(world):
        lo, hi = shard_range(n_total, world, r)
        pieces.append(out[r * per:r * per + (hi - lo)])
    return torch.cat(pieces, 0)


def run_sharded(frames: torch.Tensor, process_frame: Callable[[int, torch.Tensor], torch.Tensor], device) -> torch.Tensor:
    """frames: (N, H, W, 3) uint8 (host).  process_frame(index, frame_u8) -> stylised (H, W, 3) uint8 tensor.
    Every rank processes its block; returns all stylised frames in order."""
    dist = _dist()
    world = dist.get_world_size() if dist else 1
    rank = dist.get_rank() if dist else 0
    lo, hi = shard_range(frames.shape[0], world, rank)
    if hasattr(process_frame, "process_block") and hi > lo:
        # several frames at a time on this rank's GPU; with more than one rank the finished frames are also kept on the
        # device (pinned -> device, 0.1 ms per 720p frame) so that the all-gather does not start with a pageable upload
        local = process_frame.process_block(frames[lo:hi], out_device=device if world > 1 else None)
    else:
        done = [process_frame(k, frames[k]) for k in range(lo, hi)]
        local = torch.stack(done, 0) if done else torch.empty((0,) + tuple(frames.shape[1:]), dtype=torch.uint8)
    return gather_frames(local, frames.shape[0], device)


class FrameStyler:
    """Per-GPU worker for a stream of equally sized frames sharing one style: the per-frame body of
    apply_video_process (app.py:794-798 -> run_multi_style_transfer) with everything frame-independent hoisted."""

    def __init__(self, vgg_mean, vgg_std, frame_hw, style_imgs: List[torch.Tensor], w_style, w_content, w_tv, w_edge,
                 num_steps: int, style_img_weight=0.5, channel_attention=False, device="cuda", concurrent: int = 1,
                 batch: int = 1):
        from .multi_style_transfer.run_style_transfer import StyleTransferSession, STYLE_LAYERS, channel_attention_gate_weights
        from .engine import _require_cuda
        self.device = _require_cuda(device)
        dist = _dist()
        rank = dist.get_rank() if dist else 0
        targets = None
        if dist is not None and dist.get_world_size() > 1:
            if rank == 0:
                s0 = StyleTransferSession(vgg_mean, vgg_std, frame_hw, style_imgs, w_style, w_content, w_tv, w_edge,
                                          style_img_weight, self.device)
                targets = s0.style_targets
                self.session = s0
            targets = broadcast_style_targets(targets, STYLE_LAYERS, self.device)
        if not hasattr(self, "session"):
            self.session = StyleTransferSession(vgg_mean, vgg_std, frame_hw, style_imgs, w_style, w_content, w_tv, w_edge,
                                                style_img_weight, self.device, style_targets=targets)
        self.num_steps = int(num_steps)
        self.ca = None
        if channel_attention:
            # ChannelAttention is re-created (re-seeded, run_style_transfer.py:52) for every call of the reference:
            # the gate weights are the same for every frame
            (w1, w2), = channel_attention_gate_weights([512])
            self.ca = (w1.to(self.device).contiguous(), w2.to(self.device).contiguous())
        H, W = int(frame_hw[0]), int(frame_hw[1])
        self._in = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
        self._out = torch.empty((H, W, 3), dtype=torch.uint8).pin_memory()
        # `concurrent` frames at a time (process_block): one more plan / stream / CUDA graph per extra frame, same style targets.
        # Small frames are bound by per-launch latency, so the launches of one frame fill the ramps and tails of the others'
        # (256 x 256: +46 % frames/s with two, +79 % with four; 720p: +1 %) - SURVEY 8f row 2 without a batch dimension.
        self.sessions = [self.session]
        for _ in range(1, max(1, int(concurrent))):
            self.sessions.append(StyleTransferSession(vgg_mean, vgg_std, frame_hw, style_imgs, w_style, w_content, w_tv, w_edge,
                                                      style_img_weight, self.device, style_targets=self.session.style_targets))
        self._ins = [self._in] + [torch.empty((H, W, 3), dtype=torch.uint8).pin_memory() for _ in self.sessions[1:]]
        self._outs = [self._out] + [torch.empty((H, W, 3), dtype=torch.uint8).pin_memory() for _ in self.sessions[1:]]
        # `batch` frames per LAUNCH (process_block): one head plan whose tcgen05 convolution / Gram launches carry a batch
        # dimension (engine.BatchPlan, nst_batch_create) - SURVEY 8f row 2 as written.  Small frames then fill the GPU the way
        # one large frame does (four 256 x 256 frames = the tiles of one 512 x 512 frame).  Frame sizes must be multiples of 16.
        self.batch_plan = None
        if int(batch) > 1:
            from . import engine
            from .multi_style_transfer.run_style_transfer import CONTENT_LAYERS
            s0 = self.session
            with torch.cuda.device(self.device), torch.cuda.stream(s0.stream):
                self.batch_plan = engine.BatchPlan(s0.net, H, W, STYLE_LAYERS, CONTENT_LAYERS, int(batch), mean=s0.mean, std=s0.std)
                self.batch_plan.set_weights(w_style, w_content, w_tv, w_edge)
                self.batch_plan.set_style_targets(s0.style_targets)
                s0.stream.synchronize()
            while len(self._ins) < int(batch):
                self._ins.append(torch.empty((H, W, 3), dtype=torch.uint8).pin_memory())
                self._outs.append(torch.empty((H, W, 3), dtype=torch.uint8).pin_memory())

    def _run_one(self, frame_u8: torch.Tensor) -> torch.Tensor:
        """One frame through nst_run_frame_host; returns the pinned host buffer it was written to (valid until the next call)."""
        self._in.copy_(frame_u8)
        s = self.session
        with torch.cuda.device(self.device), torch.cuda.stream(s.stream):
            s.plan.run_frame_host(self._in, self._out, self.num_steps, *(self.ca or (None, None)))
        return self._out

    def __call__(self, index: int, frame_u8: torch.Tensor) -> torch.Tensor:
        return self._run_one(frame_u8).clone()

    def process_block(self, frames_u8: torch.Tensor, out_device=None) -> torch.Tensor:
        """frames_u8: (n, H, W, 3) uint8 host tensor -> the n stylised frames, `concurrent` at a time (nst_run_frames_host).
        Every frame's result is bit-identical to __call__ on that frame.  out_device: return them as a tensor on that CUDA
        device (each finished frame is copied there from the pinned host buffer it was written to) instead of a host tensor."""
        import ctypes as C
        from . import _lib
        n = int(frames_u8.shape[0])
        out = torch.empty((n,) + tuple(frames_u8.shape[1:]), dtype=torch.uint8, device=out_device if out_device is not None else "cpu")
        if self.batch_plan is not None:
            B = self.batch_plan.batch
            w1, w2 = self.ca or (None, None)
            s0 = self.session
            for lo in range(0, n, B):
                cnt = min(B, n - lo)
                for j in range(cnt):
                    self._ins[j].copy_(frames_u8[lo + j])
                with torch.cuda.device(self.device), torch.cuda.stream(s0.stream):
                    self.batch_plan.run_batch_host(self._ins[:cnt], self._outs[:cnt], self.num_steps, w1, w2)
                for j in range(cnt):
                    out[lo + j].copy_(self._outs[j])
            return out
        K = len(self.sessions)
        if K == 1:
            for k in range(n):
                out[k].copy_(self(k, frames_u8[k]) if out_device is None else self._run_one(frames_u8[k]))
            return out
        lib = _lib.load()
        w1, w2 = self.ca or (None, None)
        for lo in range(0, n, K):
            cnt = min(K, n - lo)
            for j in range(cnt):
                self._ins[j].copy_(frames_u8[lo + j])
            plans = (C.c_void_p * cnt)(*[getattr(s.plan.handle, "value", s.plan.handle) for s in self.sessions[:cnt]])
            ins = (C.c_void_p * cnt)(*[t.data_ptr() for t in self._ins[:cnt]])
            outs = (C.c_void_p * cnt)(*[t.data_ptr() for t in self._outs[:cnt]])
            streams = (C.c_void_p * cnt)(*[s.stream.cuda_stream for s in self.sessions[:cnt]])
            with torch.cuda.device(self.device):
                _lib.check(lib.nst_run_frames_host(plans, cnt, ins, outs, self.num_steps, 1 if w1 is not None else 0,
                                                   C.c_void_p(w1.data_ptr() if w1 is not None else None),
                                                   C.c_void_p(w2.data_ptr() if w2 is not None else None), streams, None))
            for j in range(cnt):
                out[lo + j].copy_(self._outs[j])
        return out

    def close(self):
        if self.batch_plan is not None:
            self.batch_plan.close()
            self.batch_plan = None
        for s in self.sessions:
            s.close()


def assemble_frames(frames_rgb: torch.Tensor, number_of_interpolations: int = 0) -> torch.Tensor:
    """The frame list apply_video_process writes to the encoder (app.py:800-806, 820-840), built on the device.

    frames_rgb: (F, H, W, 3) uint8 CUDA tensor of stylised RGB frames.  Returns ((F - 1) * (n + 1) + 1, H, W, 3) uint8 BGR
    frames: every input frame with its channels swapped (cv2.cvtColor(..., COLOR_RGB2BGR)) and, between two consecutive
    frames, n cross-dissolved ones (cv2.addWeighted(prev, 1 - alpha, frame, alpha, 0), alpha = (i + 1) / (n + 1)).
    One kernel (csrc/video.cu), bit-exact with OpenCV's 8-bit arithmetic."""
    import ctypes as C
    from . import _lib
    from .engine import _require_cuda
    dev = _require_cuda(frames_rgb.device)
    if frames_rgb.dtype != torch.uint8 or frames_rgb.dim() != 4 or frames_rgb.shape[-1] != 3 or frames_rgb.shape[0] < 1:
        raise ValueError("frames_rgb must be a non-empty (F, H, W, 3) uint8 tensor")
    n = int(number_of_interpolations or 0)
    if n < 0:
        raise ValueError("number_of_interpolations must be >= 0")
    frames_rgb = frames_rgb.contiguous()
    F, H, W, _ = frames_rgb.shape
    out = torch.empty(((F - 1) * (n + 1) + 1, H, W, 3), dtype=torch.uint8, device=dev)
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.nst_video_assemble(C.c_void_p(frames_rgb.data_ptr()), int(F), int(H), int(W), n, C.c_void_p(out.data_ptr()),
                                          C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return out


def stylise_frames(frames: torch.Tensor, styler: "FrameStyler", number_of_interpolations: int = 0) -> Optional[torch.Tensor]:
    """The compute of apply_video_process (app.py:784-842) between the decoder and the encoder: every frame through
    run_multi_style_transfer (frame-sharded over the ranks), then the BGR frame list with cross-dissolves, built on the
    device of rank 0.  frames: (N, H, W, 3) uint8 RGB (host).  Returns the ((N - 1) * (n + 1) + 1, H, W, 3) uint8 BGR
    frames as a CUDA tensor on rank 0 and None on the other ranks."""
    dist = _dist()
    rank = dist.get_rank() if dist else 0
    done = run_sharded(frames, styler, styler.device)
    if rank != 0:
        return None
    return assemble_frames(done.to(styler.device), number_of_interpolations)
