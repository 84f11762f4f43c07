"""Mirror of app.py's apply_video_process (app.py:742-864) for the effects that are the hot path: "Style Transfer"
(app.py:373-470, one style image) and "Style Mixing" (app.py:473-590, two styles blended by StyleMixer).

Same leading signature and return value (the path of the written .mp4, or None).  What changes is where the work runs:

  reference (per frame, serial)                               here
  ------------------------------------------------------------------------------------------------------------------
  cap.read() -> cv2.imwrite(frame_i.jpg)        :784-791      decode on the host; the lossy JPEG round trip is reproduced in
  Image.open(frame_i.jpg)                        :153-155      memory with the same encoder (cv2) and decoder (PIL)
  run_multi_style_transfer(... new Vgg19,        :463-467      FrameStyler: VGG trunk, style Gram targets, plan and CUDA graph
      new style features ...)                    :568-572      built once; frames sharded over the ranks (torch.distributed)
  cv2.cvtColor(RGB2BGR), cv2.addWeighted         :800-840      video.assemble_frames: one kernel on the device, bit-exact
  cv2.VideoWriter('avc1')                        :843-859      encode on the host (rank 0)

Everything else apply_image_process can chain in front of or behind the style transfer (text masks, pixel art, colour
palette, depth, grayscale: SURVEY.md section 2, out of scope) raises instead of being silently skipped.
"""
import io
import math
import os
import tempfile
from typing import Callable, List, Optional

import numpy as np
import torch

from ._lib import NstError

# app.py:86
list_of_effects = ["Convert Output to Grayscale", "Text-Based Effects", "Pixel Art", "Style Transfer", "Style Mixing",
                   "Color Palette Transfer", "Depth Based Style Transfer"]
VGG_MEAN = [0.485, 0.456, 0.406]   # app.py:376
VGG_STD = [0.229, 0.224, 0.225]    # app.py:377
# app.py:380-385 / 479-484
NUM_STEPS, W_STYLE, W_CONTENT, W_TV, W_EDGE = 400, 5e5, 1, 2e1, 2e1
# module-level switch like app.py:114 (flipped by a checkbox, app.py:1218-1228); only the Style Mixing branch reads it (:485)
CHANNEL_ATT_ENABLED = False


def jpeg_round_trip(frame_bgr: np.ndarray) -> np.ndarray:
    """What a decoded frame goes through before it reaches the style transfer (app.py:790-791 cv2.imwrite(...jpg) with
    OpenCV's default quality, then app.py:153-155 Image.open): BGR uint8 -> the RGB uint8 array PIL decodes from the JPEG.
    Done in memory: cv2.imencode runs the encoder imwrite runs, PIL decodes the same bytes it would read from the file."""
    import cv2
    from PIL import Image
    ok, buf = cv2.imencode(".jpg", frame_bgr)
    if not ok:
        raise NstError("cv2.imencode('.jpg') failed")
    return np.asarray(Image.open(io.BytesIO(buf.tobytes())).convert("RGB")).copy()


def read_frames(video_filepath: str, jpeg: bool = True):
    """(frames RGB uint8 [N,H,W,3], fps, frame count reported by the container)   app.py:777-791"""
    import cv2
    cap = cv2.VideoCapture(video_filepath)
    fps = cap.get(cv2.CAP_PROP_FPS)
    num_frames = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    frames = []
    while cap.isOpened():
        ret, frame = cap.read()
        if not ret:
            break
        frames.append(jpeg_round_trip(frame) if jpeg else cv2.cvtColor(frame, cv2.COLOR_BGR2RGB))
    cap.release()
    return frames, fps, num_frames


def output_fps(fps, number_of_interpolations, slowmo_slider_input):
    """app.py:851-853"""
    new_fps = fps if not number_of_interpolations else fps * (number_of_interpolations + 1)
    if slowmo_slider_input:
        new_fps = math.floor(new_fps * slowmo_slider_input)
    return new_fps


def _stages(checkbox_values, input_style, style_image_weight, style_image1, style_image2):
    """The run_multi_style_transfer calls apply_image_process would make per frame, in its order (:373, :473):
    list of (style PIL images, style_img_weight, channel_attention)."""
    from PIL import Image
    unsupported = [e for e in checkbox_values if e in list_of_effects and e not in (list_of_effects[3], list_of_effects[4])]
    if unsupported:
        raise NstError("apply_video_process mirror: %s is outside the style-transfer hot path this library implements "
                       "(SURVEY.md section 2); only 'Style Transfer' and 'Style Mixing' are supported" % unsupported)
    stages = []
    if list_of_effects[3] in checkbox_values:
        if not input_style:
            return None                                                     # app.py:468-470
        stages.append(([input_style], 0.5, False))                          # :463-467 (channel_attention = False, :386)
    if list_of_effects[4] in checkbox_values:
        if style_image1 and style_image2 and style_image_weight:            # :487
            stages.append(([Image.open(style_image1), Image.open(style_image2)], style_image_weight, CHANNEL_ATT_ENABLED))
        elif style_image1 and style_image_weight:                           # :575-577
            stages.append(([Image.open(style_image1)], style_image_weight, CHANNEL_ATT_ENABLED))
        elif style_image2 and style_image_weight:                           # :578-580
            stages.append(([Image.open(style_image2)], style_image_weight, CHANNEL_ATT_ENABLED))
        else:
            return None                                                     # :581-583
    return stages


def apply_video_process(video_filepath, checkbox_values, slowmo_slider_input=None, interpolation_slider_input=None, input_style=None,
                        text_checkbox_values=None, text_box=None, text_location_box=None, text_style_masking_box=None,
                        text_masked_transfer_edge_smoothing=None, text_emoji_blur_strength=None, text_emoji_step_size=None,
                        text_masked_style_strength=None,
                        p_size_slider=0.4, p_checkbox=list, p_colour_dropbox=0, p_colour_interpolate=False, p_edge_slider=50,
                        p_select_im=False, p_in=None, p_in_slid=10,
                        style_image_weight=None, style_image1=None, style_image2=None, color_palette_style=None,
                        d_check_box=None, depth_mip_n=2,
                        *, device="cuda", num_steps=NUM_STEPS, jpeg=True, concurrent=1, batch=1, output_video_filepath=None,
                        frame_sink: Optional[Callable[[np.ndarray], None]] = None):
    """Apply the selected style-transfer effects on a video and save it (reference: app.py:742-864; same positional
    parameters).  Keyword-only extras: `device` (a CUDA device), `num_steps` (the reference hard-codes 400, app.py:380),
    `jpeg` (reproduce the JPEG round trip of app.py:790-791), `concurrent` (frames in flight per GPU, bit-identical per frame), `batch` (frames per LAUNCH for small
    frames whose sides are multiples of 16 - a batch dimension inside the convolution / Gram kernels, video.FrameStyler; per frame
    equal to the one-frame path up to the order of the Gram sums), `output_video_filepath`
    (default: a fresh temporary directory like app.py:844), `frame_sink` (called on rank 0 with the final BGR frame list
    [F,H,W,3] before it is encoded).  Under torch.distributed every rank must call this; rank 0 returns the path."""
    from PIL import Image
    from . import video
    from .engine import _require_cuda
    from .multi_style_transfer.run_style_transfer import PIL_to_tensor
    if not video_filepath:
        return None                                                         # :771-773
    dev = _require_cuda(device)
    stages = _stages(list(checkbox_values or []), input_style, style_image_weight, style_image1, style_image2)
    if stages is None:
        return None
    frames, fps, num_frames = read_frames(video_filepath, jpeg)
    if len(frames) == 0:
        return None                                                         # :862-864
    dist = video._dist()
    rank = dist.get_rank() if dist else 0
    cur = torch.from_numpy(np.stack(frames, 0))
    H, W = int(cur.shape[1]), int(cur.shape[2])
    for styles, weight, ca in stages:
        style_t = [PIL_to_tensor(s).to(dev) for s in styles]
        if any(t.shape[1] != 3 for t in style_t):
            raise RuntimeError("The size of tensor a (%d) must match the size of tensor b (3) at non-singleton dimension 1"
                               % [t.shape[1] for t in style_t if t.shape[1] != 3][0])
        styler = video.FrameStyler(VGG_MEAN, VGG_STD, (H, W), style_t, W_STYLE, W_CONTENT, W_TV, W_EDGE, num_steps=num_steps,
                                   style_img_weight=weight, channel_attention=ca, device=dev, concurrent=concurrent, batch=batch)
        try:
            cur = video.run_sharded(cur, _Progress(styler, num_frames, rank), dev).cpu()
        finally:
            styler.close()
    if rank != 0:
        return None
    n_interp = int(interpolation_slider_input or 0)
    final = video.assemble_frames(cur.to(dev), n_interp).cpu().numpy()      # :800-840
    if frame_sink is not None:
        frame_sink(final)
    if output_video_filepath is None:
        output_video_filepath = os.path.join(tempfile.mkdtemp(prefix="nst_video_"), "output_video.mp4")   # :844
    write_video(output_video_filepath, final, output_fps(fps, n_interp, slowmo_slider_input))
    print(f"Current output video temporarily saved at: {output_video_filepath}")   # :861
    return output_video_filepath


class _Progress:
    """FrameStyler wrapper that prints the reference's per-frame status line (app.py:813)."""

    def __init__(self, styler, num_frames, rank):
        self.styler, self.num_frames, self.rank, self.done = styler, num_frames, rank, 0

    def process_block(self, frames_u8, out_device=None):
        out = self.styler.process_block(frames_u8, out_device=out_device)
        for _ in range(int(frames_u8.shape[0])):
            self.done += 1
            print(f"Finished processing frame: {self.done} out of {self.num_frames}")
        return out


def write_video(path, frames_bgr: np.ndarray, fps):
    """app.py:846-859: 'avc1' like the reference; OpenCV builds without an H.264 encoder fall back to 'mp4v'."""
    import cv2
    f_height, f_width = int(frames_bgr.shape[1]), int(frames_bgr.shape[2])
    out = None
    for codec in ("avc1", "mp4v"):
        out = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*codec), fps, (f_width, f_height))
        if out.isOpened():
            break
        out.release()
        out = None
    if out is None:
        raise NstError("cv2.VideoWriter could not open %s" % path)
    for frame in frames_bgr:
        out.write(np.ascontiguousarray(frame))
    out.release()
