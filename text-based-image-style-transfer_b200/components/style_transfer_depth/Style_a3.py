"""Mirror of components/style_transfer_depth/Style_a3.py: StyleA3, the single-style copy of the loop with a per-call style
weight.  The loop itself is the CUDA path of multi_style_transfer (StyleTransferSession); one session (VGG plan, style Gram
targets, CUDA graph) is kept per (content size, style image) so that the n planes of a multi-plane run share it.

Kept on purpose: vgg_std[0] = 0.485 (Style_a3.py:45), w_style = 5e5 * e^(strength - 1 / strength) (:187), 420 closure
evaluations for num_steps = 400 (:108).  The depth term (w_depth = 5e4 when depth=True, :180) is built through a PIL round
trip (:142-146), so it never reaches the gradient: the iterates are those of the run without it, and only the printed loss
differs - it is not evaluated here (it needs the Depth-Anything model inside the closure)."""
import numpy as np
import torch

from ...engine import _require_cuda
from ...multi_style_transfer.run_style_transfer import StyleTransferSession
from .util import image_loader, save_image


class StyleA3:
    def __init__(self, device, print_iter=50, num_steps=400, w_style=5e5, w_content=1, w_tv=2e1, w_edge=2e1, w_depth=0,
                 random_init=False, depth_pipeline=None):
        self.device = _require_cuda(device)
        self.print_iter = print_iter
        self.num_steps = num_steps
        self.w_style = w_style
        self.w_content = w_content
        self.w_tv = w_tv
        self.w_edge = w_edge
        self.w_depth = w_depth
        self.content_layers = ['conv4_2']
        self.style_layers = ['conv1_1', 'conv2_1', 'conv3_1', 'conv4_1', 'conv5_1']
        self.vgg_mean = torch.tensor([0.485, 0.456, 0.406])
        self.vgg_std = torch.tensor([0.485, 0.224, 0.225])                       # :45, as in the reference
        self.random_init = random_init
        self.depth_pipeline = depth_pipeline                                     # :47 (a callable: image -> {"depth": PIL image})
        self._session = None
        self._session_key = None
        self.last_trace = None
        self._extra = []   # further sessions of style_transfer_planes (one per plane after the first)
        self._extra_key = None
        self.traces = []   # loss rows {total, content, style, tv, edge} of the most recent runs (at most 16)

    def _get_depth_map(self, image):
        """:50-60"""
        from PIL import Image
        if self.depth_pipeline is None:
            raise RuntimeError("StyleA3 was created without a depth_pipeline")
        return Image.fromarray(np.asarray(self.depth_pipeline(image)["depth"]))

    def _session_for(self, style_img, content_hw):
        key = (tuple(content_hw), tuple(style_img.shape), hash(style_img.cpu().numpy().tobytes()))
        if self._session is None or self._session_key != key:
            self.close()
            self._session = StyleTransferSession(self.vgg_mean, self.vgg_std, content_hw, [style_img], self.w_style, self.w_content,
                                                 self.w_tv, self.w_edge, 0.5, self.device, self.content_layers, self.style_layers)
            self._session_key = key
        else:
            with torch.cuda.device(self.device), torch.cuda.stream(self._session.stream):
                self._session.plan.set_weights(float(self.w_style), float(self.w_content), float(self.w_tv), float(self.w_edge))
                self._session.weights = (float(self.w_style), float(self.w_content), float(self.w_tv), float(self.w_edge))
        return self._session

    def _run_style_transfer(self, style, content):
        """:62-166.  Returns (optimised image tensor (1, 3, H, W), None) - the reference's second value (the edge target) is unused
        by its only caller (:191)."""
        style_img = image_loader(style, device=self.device)                      # :74-75
        content_img = image_loader(content, device=self.device)
        session = self._session_for(style_img, content_img.shape[2:])
        x0 = torch.randn(content_img.size(), device=self.device) if self.random_init else None   # :98-102
        evals = 20 * (int(self.num_steps) // 20 + 1)
        session.prepare(content_img, x0, False, trace_capacity=evals + 32)
        session.run(int(self.num_steps), None)
        self.last_trace = session.trace(evals + 32)
        self.traces = (self.traces + [self.last_trace])[-16:]
        if self.print_iter:
            t = self.last_trace.cpu()
            for k in range(self.print_iter, t.shape[0] + 1, self.print_iter):  # :153-156 (the depth term is not evaluated: 0)
                tot, c, st, tv, e = (float(v) for v in t[k - 1])
                print(f'iter {k}: Content Loss: {c:4f} | Style Loss: {st:4f} | TV Loss: {tv:4f} | Edge Loss: {e:4f} | '
                      f'Depth Loss: {0.0:4f} | Total Loss: {tot:4f} ')
        return session.result(), None

    def style_transfer(self, style, content, depth=False, strength=1):
        """:168-193"""
        self.w_depth = 5e4 if depth else 0                                        # :180-183: no effect on the iterates
        if strength < 0:
            self.w_style = 5e5
        else:
            self.w_style = 5e5 * (np.e ** (strength - 1 / strength))             # :187
        print(f"w_style: {self.w_style}")
        print(f"Style Transfer with strength {strength}")
        optim_img, _ = self._run_style_transfer(style, content)
        return save_image(optim_img)

    def style_transfer_planes(self, style, contents, strengths):
        """style_transfer (:168-193) for several equally sized content images at once - the planes of a multi-plane run
        (style_transfer_depth.py:72) - each with its own strength: one plan / stream per image, stepped in lock step
        (nst_run_frames_host), so the planes overlap on the GPU instead of queueing.  Same results as one call per image."""
        import ctypes as C
        from ... import _lib
        arrs = [np.ascontiguousarray(np.asarray(c)) for c in contents]
        if len(arrs) < 2 or self.random_init or any(a.shape != arrs[0].shape or a.ndim != 3 or a.shape[2] != 3 or a.dtype != np.uint8
                                                    for a in arrs):
            return [self.style_transfer(style, c, strength=s) for c, s in zip(contents, strengths)]
        self.w_depth = 0
        style_img = image_loader(style, device=self.device)
        H, W = arrs[0].shape[:2]
        key = ((H, W), tuple(style_img.shape), hash(style_img.cpu().numpy().tobytes()))
        first = self._session_for(style_img, (H, W))
        if getattr(self, "_extra_key", None) != key:
            self._close_extra()
            self._extra_key = key
        while len(self._extra) < len(arrs) - 1:
            self._extra.append(StyleTransferSession(self.vgg_mean, self.vgg_std, (H, W), [style_img], self.w_style, self.w_content, self.w_tv,
                                                    self.w_edge, 0.5, self.device, self.content_layers, self.style_layers,
                                                    style_targets=first.style_targets))
        sessions = [first] + self._extra[:len(arrs) - 1]
        evals = 20 * (int(self.num_steps) // 20 + 1)
        for sess, strength in zip(sessions, strengths):
            w_style = 5e5 if strength < 0 else 5e5 * (np.e ** (strength - 1 / strength))       # :184-187
            print(f"w_style: {w_style}")
            print(f"Style Transfer with strength {strength}")
            with torch.cuda.device(self.device), torch.cuda.stream(sess.stream):
                sess.plan.set_weights(float(w_style), float(self.w_content), float(self.w_tv), float(self.w_edge))
            sess.weights = (float(w_style), float(self.w_content), float(self.w_tv), float(self.w_edge))
            self.w_style = w_style
        ins = [torch.from_numpy(a).pin_memory() for a in arrs]
        outs = [torch.empty((H, W, 3), dtype=torch.uint8).pin_memory() for _ in arrs]
        n = len(arrs)
        plans = (C.c_void_p * n)(*[getattr(s.plan.handle, "value", s.plan.handle) for s in sessions])
        pin = (C.c_void_p * n)(*[t.data_ptr() for t in ins])
        pout = (C.c_void_p * n)(*[t.data_ptr() for t in outs])
        streams = (C.c_void_p * n)(*[s.stream.cuda_stream for s in sessions])
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().nst_run_frames_host(plans, n, pin, pout, int(self.num_steps), 0, None, None, streams, None))
        from PIL import Image
        for sess in sessions:
            self.last_trace = sess.trace(evals + 32)
            self.traces = (self.traces + [self.last_trace])[-16:]
        return [Image.fromarray(o.numpy().copy()) for o in outs]

    def _close_extra(self):
        for s in getattr(self, "_extra", []):
            s.close()
        self._extra = []

    def close(self):
        self._close_extra()
        if self._session is not None:
            self._session.close()
            self._session = None
