#!/usr/bin/env python3
"""Benchmark of the style-transfer hot path (BASELINE.json: "style-transfer steps/s @512^2").

  python bench.py --gpus N --steps K --warmup W            the CUDA path (this repo)
  python bench.py --impl reference --gpus N --steps K ...  the reference's CPU implementation (oracle port) on host cores

A "step" is one closure evaluation = one L-BFGS inner iteration (VGG-19 forward, Gram/content/TV/edge losses, backward
to the pixels, optimizer update): run_style_transfer.py:102-148 + torch/optim/lbfgs.py:388-526.  Workload at every N:
BASELINE configs[1] - one 512x512 content/style pair per GPU, conv1_1..conv5_1 style layers, app.py's loss weights,
random-init VGG-19 (seed 1234), seeded 1/f^2 synthetic images.  N > 1 = one independent pair per GPU (the path shards
across images only: weak scaling, no collective inside the step; NCCL broadcasts the shared style Gram targets once).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how every field is obtained.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "style-transfer steps/s @512x512 (closure evaluations per second)"
UNIT = "evals/s"


def workload_name(size):
    return ("%dx%d single pair per GPU, conv1_1-conv5_1 style layers + conv4_2 content, app.py loss weights, L-BFGS history 100"
            "%s" % (size, size, " (BASELINE configs[1])" if size == 512 else ""))


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def _ncu_summary_name():
    for name in ("r03_ncu_full_conv_tc_summary.csv", "r02b_ncu_full_conv_tc_summary.csv", "r02_ncu_full_conv_tc_summary.csv", "r01d_ncu_full_conv_tc_summary.csv"):
        if os.path.exists(os.path.join(ROOT, "profiles", name)):
            return name
    return "r02_ncu_full_conv_tc_summary.csv"


NCU_SUMMARY = _ncu_summary_name()


def ncu_traffic():
    """Average DRAM bytes (read + write) per conv_tc_kernel launch from the committed `ncu --set full` capture of one
    evaluation's launches (profiles/<NCU_SUMMARY>: the newest round's capture); None if the summary is missing."""
    import csv
    path = os.path.join(ROOT, "profiles", NCU_SUMMARY)
    if not os.path.exists(path):
        return None
    rows = list(csv.reader(open(path)))
    hdr = rows[0]
    try:
        ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
    except ValueError:
        return None
    vals = [(float(r[ir]) + float(r[iw])) * 1e6 for r in rows[2:] if "conv_tc_kernel" in r[ik] and "<16" not in r[ik]]
    return sum(vals) / len(vals) if vals else None


def bench_mask_composite(dev, pk):
    """text/segmentation_style_transfer.py on the device (csrc/mask.cu): GB/s against the measured copy bandwidth, CUDA events,
    inputs larger than L2 at the large size; CPU = the oracle port.  Bytes: the reference reads 7 and writes 3 per pixel; the
    kernel copies tiles the mask covers entirely (or not at all) from ONE image, so on this disc mask (99 % such tiles) it has
    to move only 1 + 3 + 3 = 7 bytes per pixel - `achieved` uses 7, `achieved_ref_bytes` the reference's 10."""
    import numpy as np
    import torch
    from importlib import import_module
    seg = import_module("text-based-image-style-transfer_b200.text.segmentation_style_transfer")
    from oracle import mask_oracle as M
    out = {}
    for S in (1024, 8192):
        g = torch.Generator(device="cpu").manual_seed(S)
        content = torch.randint(0, 256, (S, S, 3), dtype=torch.uint8, generator=g).to(dev)
        style = torch.randint(0, 256, (S, S, 3), dtype=torch.uint8, generator=g).to(dev)
        yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
        mask = (((yy - S / 2) ** 2 + (xx - S / 3) ** 2) < (S / 2.5) ** 2).to(dev)
        for _ in range(3):
            res = seg.composite_tensors(content, style, mask, 5)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            res = seg.composite_tensors(content, style, mask, 5)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        gbs = 7.0 * S * S / (ms * 1e-3) / 1e9
        row = dict(ms=ms, achieved=gbs, achieved_ref_bytes=gbs * 10.0 / 7.0, peak=pk["hbm"], unit="GB/s", frac=gbs / pk["hbm"],
                   bytes=7 * S * S)
        if S == 1024:
            c, s_, m = content.cpu().numpy(), style.cpu().numpy(), mask.cpu().numpy()
            t0 = time.perf_counter()
            want = M.segmentation_style_transfer(c, s_, m, 5)
            row["cpu_ms"] = 1e3 * (time.perf_counter() - t0)
            row["bit_exact"] = bool(np.array_equal(res.cpu().numpy(), want))
        out["%dx%d" % (S, S)] = row
    out["what"] = ("segmentation_style_transfer(edge_smoothing=5) on uint8 HWC device tensors: mask blur (cv2.GaussianBlur fixed point) + "
                   "fp64 blend in one kernel; 8192^2 = 470 MB of traffic (larger than L2); the 1024^2 time is launch + Python overhead; "
                   "cpu_ms = oracle port (numpy), one thread")
    return out


def bench_video_assemble(dev, pk):
    """app.py:800-840 on the device (csrc/video.cu): 64 stylised 720p frames, 3 cross-dissolved frames between neighbours
    (253 output frames).  Bytes: every input frame read twice (as `prev` and as `frame`; the second read may hit in L2),
    every output frame written once - `achieved` counts each input frame ONCE (the minimum).  CPU = cv2 on the host."""
    import numpy as np
    import torch
    from importlib import import_module
    video = import_module("text-based-image-style-transfer_b200.video")
    from oracle import video_oracle as V
    F, H, W, n = 64, 720, 1280, 3
    g = torch.Generator(device="cpu").manual_seed(7)
    frames = torch.randint(0, 256, (F, H, W, 3), dtype=torch.uint8, generator=g).to(dev)
    for _ in range(3):
        out = video.assemble_frames(frames, n)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        out = video.assemble_frames(frames, n)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    nbytes = (F + out.shape[0]) * H * W * 3
    gbs = nbytes / (ms * 1e-3) / 1e9
    row = dict(ms=ms, achieved=gbs, peak=pk["hbm"], unit="GB/s", frac=gbs / pk["hbm"], bytes=nbytes, frames_in=F, frames_out=int(out.shape[0]),
               what="RGB->BGR + 3 cross-dissolved frames between neighbours, 64 frames 1280x720 (0.88 GB moved, larger than L2)")
    sub = frames[:4].cpu().numpy()
    t0 = time.perf_counter()
    want = np.stack(V.assemble_frames(list(sub), n), 0)
    row["cpu_ms_per_output_frame"] = 1e3 * (time.perf_counter() - t0) / want.shape[0]
    row["gpu_ms_per_output_frame"] = ms / out.shape[0]
    row["bit_exact"] = bool(np.array_equal(out[:want.shape[0]].cpu().numpy(), want))
    return row


def bench_mip(dev, pk):
    """components/style_transfer_depth/util.py split / merge on the device (csrc/depth.cu): 16384 x 4096 RGB, 4 depth planes, uint8
    depth.  Bytes: split reads 3 + 1 and writes 4 * 3 per pixel; merge reads 1 + 3 (only the plane a pixel belongs to) and
    writes 3.  The n loops in between are the headline path itself."""
    import numpy as np
    import torch
    from importlib import import_module
    U = import_module("text-based-image-style-transfer_b200.components.style_transfer_depth.util")
    from oracle import depth_oracle as D
    S, S2, n = 16384, 4096, 4
    g = torch.Generator(device="cpu").manual_seed(3)
    img = torch.randint(0, 256, (S, S2, 3), dtype=torch.uint8, generator=g).to(dev)
    yy, xx = np.mgrid[0:S, 0:S2]
    depth = ((np.sin(yy / 300.0) + np.cos(xx / 450.0) + 2) * 63).astype(np.uint8)
    bins = U.create_bins(n)
    d_dev = torch.from_numpy(depth).to(dev)

    out = {}
    lo_hi = (int(depth.min()), int(depth.max()))   # the depth map is already on the device: its upload is not kernel time
    for name, fn, nbytes in (("split", lambda: U.split_planes(img, d_dev, bins, depth_range=lo_hi), (4 + 3 * n) * S * S2),
                             ("merge", lambda: U.merge_planes(planes, d_dev, bins, depth_range=lo_hi), 7 * S * S2)):
        if name == "split":
            planes = fn()
        for _ in range(3):
            res = fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(10):
            res = fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gbs = nbytes / (ms * 1e-3) / 1e9
        out[name] = dict(ms=ms, achieved=gbs, peak=pk["hbm"], unit="GB/s", frac=gbs / pk["hbm"], bytes=nbytes)
    c_img, c_depth = img[:512, :512].contiguous(), depth[:512, :512]
    got_p = U.split_planes(c_img, c_depth, bins)
    got_m = U.merge_planes(got_p, c_depth, bins)
    want_p = D.generate_mip_layers(c_img.cpu().numpy(), c_depth, n)
    out["bit_exact"] = bool(np.array_equal(got_p.cpu().numpy(), np.stack(want_p, 0)) and
                            np.array_equal(got_m.cpu().numpy(), D.reconstruct_mip_image(want_p, c_depth, n)))
    out["what"] = "generate_mip_layers / reconstruct_mip_image, 16384x4096 RGB, 4 planes (1074 / 470 MB moved, larger than L2; calls long enough that the Python call overhead of ~30 us does not set the time)"
    return out


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], tc_burst=p["bf16_tflops"], tc_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["no samples"])
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), power_w_max=max(power), samples=len(sm),
                    reasons=sorted(reasons))


# ----------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the oracle port of the reference's CPU path on the host cores
# ----------------------------------------------------------------------------------------------------------------
def time_oracle(size, budget_s, max_evals, warm_evals=1):
    """Evaluations per second of the reference's algorithm on the CPU at size x size, measured on a bounded sample:
    as many closure evaluations of the first optimizer.step() as fit `budget_s` (at most max_evals)."""
    import torch
    from oracle import nst_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    ws, bs = O.vgg19_random_weights(1234, 13)
    content, style = O.synth_image(size, size, 0), O.synth_image(size, size, 1)
    stamps = []

    def on_eval(k):
        stamps.append(time.perf_counter())
        # stop once the budget is spent (never before 2 timed evaluations) or max_evals timed evaluations are done
        timed = len(stamps) - warm_evals
        return timed >= max_evals or (timed >= 2 and stamps[-1] - stamps[warm_evals - 1] >= budget_s)

    O.run_oracle(ws, bs, content, [style], 10 ** 9, emulate_reference_cost=True, on_eval=on_eval, **O.APP_WEIGHTS)
    # setup (VGG target features) and the first evaluation (allocator / thread-pool warm-up) are excluded
    dt = stamps[-1] - stamps[warm_evals - 1]
    evals = len(stamps) - warm_evals
    return dict(value=evals / dt, unit=UNIT, cores=cores, kind="port",
                sample="%d closure evaluations (+ L-BFGS updates) of one %dx%d run after 1 untimed evaluation; oracle/nst_oracle.py "
                       "with the reference's dead weight-gradients and per-eval style targets emulated; torch %s CPU fp32, %d threads"
                       % (evals, size, size, torch.__version__, cores)), evals, dt


def time_torch_eager_gpu(size, dev, budget_s=20.0, max_evals=60, warm_evals=2):
    """The reference's algorithm under plain torch eager ON THE GPU: the oracle port with device="cuda" - cuDNN convolutions
    (torch's default cudnn.allow_tf32=True: TF32), fp32 cuBLAS bmm for the Gram matrices (cuda.matmul.allow_tf32=False), autograd
    with the reference's dead weight gradients and per-evaluation style targets, and L-BFGS with its 4-5 host syncs per
    iteration.  This is what a maintainer gets today by passing device="cuda" to the reference on a B200;
    none of this repo's kernels run in it."""
    import torch
    from oracle import nst_oracle as O
    ws, bs = O.vgg19_random_weights(1234, 13)
    content, style = O.synth_image(size, size, 0), O.synth_image(size, size, 1)
    stamps = []

    def on_eval(k):
        torch.cuda.synchronize(dev)
        stamps.append(time.perf_counter())
        timed = len(stamps) - warm_evals
        return timed >= max_evals or (timed >= 40 and stamps[-1] - stamps[warm_evals - 1] >= budget_s)

    free0 = torch.cuda.mem_get_info(dev)[0]
    O.run_oracle(ws, bs, content, [style], 10 ** 9, emulate_reference_cost=True, on_eval=on_eval, device=dev, **O.APP_WEIGHTS)
    dt = stamps[-1] - stamps[warm_evals - 1]
    evals = len(stamps) - warm_evals
    peak_mb = torch.cuda.max_memory_allocated(dev) / 2 ** 20
    torch.cuda.empty_cache()
    return dict(value=evals / dt, unit=UNIT, evals=evals, seconds=dt, ms_per_eval=1e3 * dt / evals,
                cudnn_allow_tf32=bool(torch.backends.cudnn.allow_tf32), matmul_allow_tf32=bool(torch.backends.cuda.matmul.allow_tf32),
                torch=torch.__version__, cudnn=torch.backends.cudnn.version(), peak_alloc_mb=peak_mb, free_mb_before=free0 / 2 ** 20,
                what="reference algorithm (oracle port, emulate_reference_cost) on device=cuda under torch eager: %d closure evaluations "
                     "(+ L-BFGS updates) of one %dx%d run after %d untimed evaluations; wall clock with a device sync per evaluation "
                     "(the reference syncs there too: float(closure()))" % (evals, size, size, warm_evals))


def video_strong_scaling(args, rank, world, dev, dist, hf, synth):
    """BASELINE configs[4] / metric "video frames/s at 1/2/4/8 B200", app.py:784-815: a FIXED number of synthetic 720p frames
    (the same total at every N: strong scaling), one shared 512x512 style, frame-sharded in contiguous blocks, style Gram
    targets computed on rank 0 and broadcast, finished uint8 frames all-gathered in order.  Every frame is a whole
    run_multi_style_transfer on host uint8 buffers (H2D, content / edge targets, the loop, D2H).  Frame count and num_steps are
    reduced against the reference UI's 240 x num_steps=400 so that the row fits the bench's time budget; both are stated."""
    import torch
    from nst_b200 import video
    H, W = (int(v) for v in args.video_size.split("x"))
    n_frames = args.video_row_frames
    steps = args.video_row_steps
    evals = 20 * (steps // 20 + 1)
    style = torch.from_numpy(synth.synth_image(512, 512, 1)).permute(2, 0, 1).float().div(255).unsqueeze(0).to(dev)
    f0 = torch.from_numpy(synth.synth_image(H, W, 100)).float()
    f1 = torch.from_numpy(synth.synth_image(H, W, 101)).float()
    frames = torch.stack([((1 - k / max(n_frames - 1, 1)) * f0 + (k / max(n_frames - 1, 1)) * f1).round().to(torch.uint8)
                          for k in range(n_frames)])
    t_setup = time.perf_counter()
    # two frames per launch (nst_batch_create: batch dimension inside the convolution / Gram kernels): +3 % at 720p, where one
    # frame already fills the GPU (profiles/README.md); --video-row-batch 1 = one frame at a time
    row_batch = args.video_row_batch if (H % 16 == 0 and W % 16 == 0 and args.video_row_batch > 1) else 1   # batches need sides that are multiples of 16
    styler = video.FrameStyler(synth.VGG_MEAN, synth.VGG_STD, (H, W), [style], num_steps=steps, device=dev, batch=row_batch,
                               **synth.APP_WEIGHTS)
    styler.process_block(frames[:max(1, row_batch)])   # warm-up: graph capture, allocator, pinned buffers
    lo, hi = video.shard_range(n_frames, world, rank)
    video.gather_frames(frames[lo:hi].to(dev), n_frames, dev)    # ... and the communicator / buffers of the all-gather (same sizes)
    torch.cuda.synchronize()
    setup_s = time.perf_counter() - t_setup
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    local = (styler.process_block(frames[lo:hi], out_device=dev if world > 1 else None) if hi > lo
             else torch.empty((0, H, W, 3), dtype=torch.uint8))
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    out = video.gather_frames(local, n_frames, dev)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t2 = time.perf_counter()
    dt, gather = t2 - t0, t2 - t1
    if dist is not None:
        t = torch.tensor([dt, t1 - t0], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt, compute = float(t[0]), float(t[1])
    else:
        compute = t1 - t0
    checksum = int(out.to(torch.int64).sum()) if rank == 0 else None
    styler.close()
    del styler
    torch.cuda.empty_cache()
    return dict(metric="video style transfer frames/s (%dx%d, frame-sharded)" % (W, H), value=n_frames / dt, unit="frames/s",
                frames=n_frames, frames_per_rank=hi - lo, n_gpus=world, num_steps=steps, evals_per_frame=evals, seconds=dt,
                slowest_rank_compute_s=compute, gather_s_rank0=gather, setup_s_untimed=setup_s, evals_per_s=n_frames * evals / dt,
                scaling="strong", h2d_bytes_per_frame=3 * H * W, d2h_bytes_per_frame=3 * H * W, checksum=checksum,
                frames_per_launch=row_batch,
                what="%d synthetic %dx%d frames in total at every N (strong scaling), one shared 512x512 style, num_steps=%d (%d evaluations per "
                     "frame; the reference UI uses 240 frames x num_steps=400), contiguous frame blocks per rank, NCCL broadcast of the style Gram "
                     "targets (untimed setup) + all-gather of the finished frames (timed); wall clock, barrier on both sides, max over ranks"
                     % (n_frames, W, H, steps, evals))


def bench_batched_small_frames(dev, synth, size=256, n_frames=32, steps=60, with_concurrent=False):
    """SURVEY 8f row 2 in front of the driver: small frames (C1's 256 x 256), several per LAUNCH - the batch dimension inside the
    tcgen05 convolution / Gram kernels (FrameStyler(batch=B) -> nst_run_batch_host) - beside one frame at a time and
    several frames in flight on separate plans / streams (FrameStyler(concurrent=K)).  Every frame is a whole
    run_multi_style_transfer on host uint8 buffers; wall clock around process_block, graph capture and allocator warmed
    up before."""
    import torch
    from nst_b200 import video
    evals = 20 * (steps // 20 + 1)
    style = torch.from_numpy(synth.synth_image(512, 512, 1)).permute(2, 0, 1).float().div(255).unsqueeze(0).to(dev)
    frames = torch.stack([torch.from_numpy(synth.synth_image(size, size, 200 + k)) for k in range(n_frames)])
    rows = {}
    configs = [("one_at_a_time", {}), ("batch_4", dict(batch=4)), ("batch_8", dict(batch=8))]
    if with_concurrent:   # four plans / streams / graphs in lock step (--batch-row-concurrent; profiles/r03_bench_default.json has the row)
        configs.insert(1, ("concurrent_4", dict(concurrent=4)))
    for name, kw in configs:
        styler = video.FrameStyler(synth.VGG_MEAN, synth.VGG_STD, (size, size), [style], num_steps=steps, device=dev, **kw, **synth.APP_WEIGHTS)
        try:
            styler.process_block(frames[:8])
            torch.cuda.synchronize()
            best = None
            for _ in range(2):
                t0 = time.perf_counter()
                out = styler.process_block(frames)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
                best = dt if best is None or dt < best else best
            rows[name] = dict(frames_per_s=n_frames / best, evals_per_s=n_frames * evals / best, seconds=best, checksum=int(out.to(torch.int64).sum()))
        finally:
            styler.close()
            del styler
            torch.cuda.empty_cache()
    return dict(metric="small-frame style transfer, aggregate evaluations/s (%dx%d)" % (size, size), unit="evals/s",
                value=rows["batch_8"]["evals_per_s"], frames=n_frames, num_steps=steps, evals_per_frame=evals, rows=rows,
                h2d_bytes_per_frame=3 * size * size, d2h_bytes_per_frame=3 * size * size,
                what="%d synthetic %dx%d frames, one shared 512x512 style, num_steps=%d (%d evaluations per frame), host uint8 in / out; "
                     "`value` = 8 frames per launch (nst_batch_create: batch dimension inside conv_tc_kernel / gram_partial_kernel); rows: "
                     "one frame at a time, 4 and 8 frames per launch (--batch-row-concurrent adds 4 frames in flight on 4 plans / streams: "
                     "5 187 evaluations/s in profiles/r03_bench_default.json); best of 2 repetitions"
                     % (n_frames, size, size, steps, evals))


def run_reference(args, rank, world):
    if rank != 0:
        return
    base, evals, dt = time_oracle(args.size, budget_s=max(20.0, min(150.0, 0.5 * (args.steps + args.warmup))),
                                  max_evals=max(2, min(args.steps, 60)))
    line = dict(impl="reference", metric=METRIC, value=base["value"], unit=UNIT, n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 / base["value"], higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f32", data="synthetic",
                config=dict(workload=workload_name(args.size)),
                detail=dict(note="reference's CPU path (oracle port; /root/reference is Python and cannot travel to the GPU box), "
                                 "host cores only, bounded sample"),
                cpu_baseline=base,
                e2e=dict(value=base["value"], unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                gpu_launches=0)
    emit(line)


# ----------------------------------------------------------------------------------------------------------------
# the CUDA path
# ----------------------------------------------------------------------------------------------------------------
def run_b200(args, rank, world, local_rank):
    import numpy as np
    import torch
    import nst_b200
    from nst_b200 import synth
    from importlib import import_module
    hf = import_module("text-based-image-style-transfer_b200.multi_style_transfer.helper_functions")
    rst = import_module("text-based-image-style-transfer_b200.multi_style_transfer.run_style_transfer")

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    S = args.size
    K, Wm = args.steps, args.warmup
    ws, bs = synth.vgg19_random_weights(1234, 13)
    hf.set_vgg_weight_provider(lambda: (ws, bs))
    style = torch.from_numpy(synth.synth_image(S, S, 1)).permute(2, 0, 1).float().div(255).unsqueeze(0).to(dev)
    content_u8 = synth.synth_image(S, S, 0 if rank == 0 else 100 + rank)   # one independent image per GPU
    content = torch.from_numpy(content_u8).permute(2, 0, 1).float().div(255).unsqueeze(0).to(dev)

    # shared style Gram targets: rank 0 computes them, NCCL broadcasts (2.44 MB); outside the step loop
    targets = None
    if world > 1:
        names = rst.STYLE_LAYERS
        chans = [64, 128, 256, 512, 512]
        if rank == 0:
            s0 = rst.StyleTransferSession(synth.VGG_MEAN, synth.VGG_STD, (S, S), [style], device=dev, **synth.APP_WEIGHTS)
            bufs = [s0.style_targets[n].contiguous() for n in names]
            s0.close()
        else:
            bufs = [torch.empty((1, c, c), device=dev) for c in chans]
        for b in bufs:
            dist.broadcast(b, src=0)
        targets = dict(zip(names, bufs))
    sess = rst.StyleTransferSession(synth.VGG_MEAN, synth.VGG_STD, (S, S), [style], device=dev, style_targets=targets,
                                    **synth.APP_WEIGHTS)
    plan = sess.plan
    stream = sess.stream

    def enqueue_evals(n):
        """exactly n closure evaluations: whole optimizer.step() graphs (20 each) + one truncated step; returns launches"""
        launches = 0
        with torch.cuda.stream(stream):
            for _ in range(n // 20):
                plan.lbfgs_step()
                launches += plan.launches_per_step()
            if n % 20:
                launches += plan.lbfgs_partial_step(n % 20)
        return launches

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---- warm-up.  The CUDA graph of optimizer.step() (~800 nodes) is captured, instantiated and uploaded HERE, explicitly
    #      (nst_lbfgs_prepare_graph), and at least one whole graph launch runs before the timed region whatever --warmup
    #      is: in r01 a --warmup below 20 only ran directly-enqueued evaluations and the capture fell inside ev0..ev1.
    #      One trace capacity for every prepare() of this process: a larger capacity drops the graph (api.cu).
    JOB_STEPS, JOB_EVALS = 300, 320                       # BASELINE configs[1]: num_steps=300 -> 320 closure evaluations
    trace_cap = max(Wm + K, JOB_EVALS) + 64
    sess.prepare(content, trace_capacity=trace_cap)
    with torch.cuda.stream(stream):
        plan.lbfgs_prepare_graph()
    warm_evals = 20 * ((max(Wm, 3) + 19) // 20)           # whole optimizer.step()s: >= --warmup evaluations, >= 1 graph launch
    enqueue_evals(warm_evals)
    stream.synchronize()
    hist_start = sess.status().hist_len

    # ---- timed region: K evaluations, CUDA events on the launching stream, barrier + sync on both sides
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    with torch.cuda.stream(stream):
        ev0.record(stream)
    launches = enqueue_evals(K)
    with torch.cuda.stream(stream):
        ev1.record(stream)
    torch.cuda.synchronize()
    barrier()
    ms = ev0.elapsed_time(ev1)
    st = sess.status()
    hist_len = st.hist_len
    final_loss = st.loss

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    ms = max_over_ranks(ms)
    value = world * K / (ms * 1e-3)

    # ---- the whole job of BASELINE configs[1], device-timed: fresh optimizer state, num_steps=300 -> 320 evaluations =
    #      16 graph launches (the history grows 0 -> 100 pairs on the way, run-average ~84)
    # two repetitions, the faster one is reported (the same is done for e2e below: both time the identical job, and a
    # transient on the box must not decide which of the two comes out ahead)
    ms_320 = None
    for _rep in range(2):
        sess.prepare(content, trace_capacity=trace_cap)   # same capacity: the captured graph survives
        with torch.cuda.stream(stream):
            plan.lbfgs_prepare_graph()
        ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.synchronize()
        with torch.cuda.stream(stream):
            ev2.record(stream)
        launches_320 = enqueue_evals(JOB_EVALS)
        with torch.cuda.stream(stream):
            ev3.record(stream)
        torch.cuda.synchronize()
        barrier()
        t_rep = max_over_ranks(ev2.elapsed_time(ev3))
        ms_320 = t_rep if ms_320 is None else min(ms_320, t_rep)
    st320 = sess.status()
    value_320 = dict(value=world * JOB_EVALS / (ms_320 * 1e-3), unit=UNIT, evals=JOB_EVALS, ms=ms_320, history_pairs_at_end=st320.hist_len,
                     final_loss=st320.loss, gpu_launches=launches_320,
                     what="BASELINE configs[1] as one job, device-timed (CUDA events): fresh L-BFGS state, 16 optimizer.step() graph launches; best of 2 repetitions")
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: the public host-buffer call on the WHOLE job (uint8 image in pinned host memory -> H2D -> content / edge
    #      targets -> num_steps=300 loop with a status read-back per optimizer.step() -> D2H uint8 result), wall clock
    pin_in = torch.from_numpy(content_u8).contiguous().pin_memory()
    pin_out = torch.empty_like(pin_in).pin_memory()
    with torch.cuda.stream(stream):
        plan.run_frame_host(pin_in, pin_out, 0)              # warm: 20 evaluations through the same call
    dt_e2e = None
    for _rep in range(2):
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(stream):
            n_e2e = plan.run_frame_host(pin_in, pin_out, JOB_STEPS)
        torch.cuda.synchronize()
        t_rep = max_over_ranks(time.perf_counter() - t0)
        dt_e2e = t_rep if dt_e2e is None else min(dt_e2e, t_rep)
    status_bytes = 2700 + 4 * 16  # control block + loss vector read back after every optimizer.step()
    e2e = dict(value=world * n_e2e / dt_e2e, unit=UNIT, h2d_bytes_per_step=3 * S * S / n_e2e,
               d2h_bytes_per_step=(3 * S * S + status_bytes * (n_e2e // 20)) / n_e2e, evals=n_e2e, seconds=dt_e2e,
               call="nst_run_frame_host(num_steps=%d): H2D uint8 image, content/edge targets, %d evaluations (the whole job of BASELINE "
                    "configs[1]), one status read-back per optimizer.step(), D2H uint8 result; wall clock, max over ranks, best of 2 repetitions" % (JOB_STEPS, n_e2e))
    # sanity (r01's line had value < e2e because the graph capture sat inside the timed region): the device-timed job
    # cannot be slower than the same job through the host-buffer call, which adds copies and 16 host round trips
    sanity = dict(e2e_le_value_320=bool(e2e["value"] <= 1.02 * value_320["value"]),
                  graph_captured_before_timed_region=True, warmup_evals_run=warm_evals)
    if not sanity["e2e_le_value_320"]:
        sys.stderr.write("bench.py: WARNING e2e (%.1f) exceeds the device-timed job (%.1f) by more than 2%%\n" % (e2e["value"], value_320["value"]))

    # ---- BASELINE configs[4] in front of the driver: 720p frames, frame-sharded, FIXED total frame count across N
    video_row = None
    if not args.no_video_row:
        video_row = video_strong_scaling(args, rank, world, dev, dist, hf, synth)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel class, measured live with CUDA events: two whole optimizer.step()s are run
    #      with an event after every launch (single stream, no graph), so every kernel is timed inside a long busy stream
    pk = peaks()
    sess.prepare(content, trace_capacity=Wm + K + 64)
    enqueue_evals(120)                                            # history full (m = 100): steady state of a 320-evaluation run
    stream.synchronize()
    # three repetitions of (2 x 20 evaluations); the repetition with the smallest total is reported (a transient on the box -
    # one run showed pass 1 at 180 us instead of 114 - would otherwise decide the roofline line)
    raw, raw_total = None, None
    for _rep in range(3):
        with torch.cuda.stream(stream):
            r_ = plan.lbfgs_step_timed(grouped=True) + plan.lbfgs_step_timed(grouped=True)
        t_ = sum(ms_ for _, _, ms_ in r_)
        if raw_total is None or t_ < raw_total:
            raw, raw_total = r_, t_
    m_now = sess.status().hist_len
    # grouped rows: (kind, launches in the run, ms of the run) - one CUDA event wherever the kind of launch changes, so
    # the twelve forward convolutions (or the twelve data gradients) are timed as one run with programmatic dependent
    # launch between them intact, exactly as they execute inside the captured step
    acc = {}
    for k, nl_, ms_ in raw:
        a = acc.setdefault(k, [0.0, 0, 0])
        a[0] += ms_
        a[1] += nl_
        a[2] += 1
    n_evals_timed = acc["pixel"][2]
    per_eval = lambda kinds: sum(v[0] for k, v in acc.items() if k in kinds) / n_evals_timed  # noqa: E731
    conv_kinds = ("conv_fwd", "conv_dgrad", "gram_bwd")
    conv_ms = per_eval(conv_kinds)
    conv_launches = sum(v[1] for k, v in acc.items() if k in conv_kinds) / n_evals_timed
    conv_fl = 2.0 * sum(synth.conv_flops(i, S, S) for i in range(1, 13)) + sum(synth.gram_flops(i, S, S) for i in synth._STYLE)
    lb_kinds = ("lbfgs_pass1", "lbfgs_reduce", "lbfgs_control", "lbfgs_pass2")
    eval_ms = sum(v[0] for k, v in acc.items() if k not in lb_kinds) / n_evals_timed
    achieved_tf = conv_fl / (conv_ms * 1e-3) / 1e12
    roofline = dict(bound="tensor", kernel="conv_tc_kernel (tcgen05 implicit GEMM: 12 forward + 12 data-gradient launches per evaluation, the Gram backward "
                           "of conv1_1..conv4_1 folded into the data gradients as a second accumulator, + 1 Gram-backward launch for conv5_1)",
                    achieved=achieved_tf, peak=pk["tc_burst"], unit="TFLOP/s", frac=achieved_tf / pk["tc_burst"],
                    peak_sustained=pk["tc_sustained"], frac_sustained=achieved_tf / pk["tc_sustained"],
                    traffic=ncu_traffic(), traffic_unit="bytes of DRAM traffic per launch, average over the conv_tc_kernel launches of one evaluation (ncu --set full, profiles/" + NCU_SUMMARY + ")",
                    peak_source=pk["source"] + ": `peak` / `frac` use the BURST bf16 figure, the stricter denominator; `peak_sustained` / "
                                "`frac_sustained` are printed beside it.  Measured inside the kernels (SM cycles / %globaltimer of the MMA-issuing "
                                "thread over a CTA's main loops, instrumented build, profiles/r02p_timeline_512_pair_clock.log): the SM clock "
                                "during the main loops of the 128-channel-and-up layers is 1.59 - 1.75 GHz, not the 1.965 GHz nvidia-smi's "
                                "samples show - the tensor cores ARE power-limited inside this step, which is the regime the sustained figure "
                                "(1 407 TFLOP/s = burst x 1658 / 1965 MHz) describes; as CTA pairs (cta_group::2) those main loops issue one "
                                "M256 x N128 x K16 MMA per 62 - 64 SM cycles, the tensor core's rate",
                    sm_clock_in_main_loops_mhz=dict(min=1589, max=1856, typical_128ch_layers="1590-1750",
                                                    source="profiles/r02p_timeline_512_pair_clock.log (tools/conv_timeline.py, instrumented build)"),
                    flops_per_launch_avg=conv_fl / conv_launches, launches_per_eval=conv_launches,
                    avg_launch_ms=conv_ms / conv_launches,
                    ms_per_eval_in_kernel=conv_ms, share_of_eval=conv_ms / eval_ms,
                    fwd_tflops=sum(synth.conv_flops(i, S, S) for i in range(1, 13)) / (per_eval(("conv_fwd",)) * 1e-3) / 1e12,
                    dgrad_tflops=(sum(synth.conv_flops(i, S, S) for i in range(1, 13)) + sum(synth.gram_flops(i, S, S) for i in synth._STYLE)) /
                                 (per_eval(("conv_dgrad", "gram_bwd")) * 1e-3) / 1e12,
                    how="CUDA events on the launching stream around every run of same-kind launches (forward convolutions, "
                        "data gradients, Gram backward) of 2 x 20 evaluations executed back to back on one stream "
                        "(nst_lbfgs_step_timed_grouped), best of 3 repetitions; average launch duration = run time / launches in the run")
    lb_ms = per_eval(("lbfgs_pass1", "lbfgs_pass2"))
    m_avg = m_now                                                 # 100: the history was full while the timed steps ran
    lb_gbs = synth.lbfgs_bytes(S, S, m_avg) / (lb_ms * 1e-3) / 1e9 if lb_ms > 0 else None
    by_kind = {k: v[0] / n_evals_timed for k, v in acc.items()}
    roofline_hbm = dict(bound="hbm", kernel="lbfgs_pass1_kernel + lbfgs_pass2_kernel", achieved=lb_gbs, peak=pk["hbm"], unit="GB/s",
                        frac=(lb_gbs / pk["hbm"]) if lb_gbs else None, history_pairs=m_avg,
                        bytes_per_iteration=synth.lbfgs_bytes(S, S, m_avg), ms=lb_ms)
    lb_rows = []
    # every other kernel of the step against ITS bound: algorithmic bytes / time in the same grouped
    # timing / measured copy bandwidth.  At 512^2 most of them move a few MB in ~10 us: latency-sized launches that run on
    # the side stream, off the critical path (conv forward -> deepest Gram -> data gradients -> optimizer).
    hw = S * S
    tap_bytes = 2.0 * sum(synth._COUT[i] * (S >> synth._LEVEL[i]) * (S >> synth._LEVEL[i]) for i in synth._STYLE)
    small = {
        "pixel": (2.0 / 3.0 + 2.0) * 3 * hw * 4,                    # x read, edge target read (2 planes), gradient written
        "content": 512 * (S // 8) * (S // 8) * (2 + 4 + 2),          # tap fp16 + target fp32 read, seed bf16 written
        "gram": tap_bytes,                                            # every style tap read once (partial + 2 finalize launches per set)
        "conv1_fwd": 3 * hw * 4 + 2 * 64 * hw * 2,                    # image read, tap + activation written
        "conv1_dgrad": 64 * hw * 2 + 2 * 3 * hw * 4,                  # gradient of conv1_1 read, pixel-term gradient read, g written
        "assemble": 0.0,
    }
    on_critical_path = {"pixel": False, "content": False, "gram": "conv5_1's Gram only (the shallow layers' launch is on the side stream)",
                        "conv1_fwd": True, "conv1_dgrad": True, "assemble": False}
    roofline_small = {}
    for k, nbytes in small.items():
        if k not in by_kind or by_kind[k] <= 0:
            continue
        gbs = nbytes / (by_kind[k] * 1e-3) / 1e9
        roofline_small[k] = dict(bound="hbm", ms=by_kind[k], bytes=nbytes, achieved=gbs, peak=pk["hbm"], unit="GB/s", frac=gbs / pk["hbm"],
                                 on_critical_path=on_critical_path[k])
    roofline_small["gram"]["tensor_tflops"] = sum(synth.gram_flops(i, S, S) for i in synth._STYLE) / (by_kind["gram"] * 1e-3) / 1e12
    it_ms = sum(by_kind.get(k, 0.0) for k in lb_kinds)
    roofline_small["lbfgs_iteration"] = dict(bound="hbm", ms=it_ms, bytes=synth.lbfgs_bytes(S, S, m_avg), unit="GB/s", peak=pk["hbm"],
                                             achieved=synth.lbfgs_bytes(S, S, m_avg) / (it_ms * 1e-3) / 1e9,
                                             frac=synth.lbfgs_bytes(S, S, m_avg) / (it_ms * 1e-3) / 1e9 / pk["hbm"],
                                             serial_ms=by_kind.get("lbfgs_reduce", 0.0) + by_kind["lbfgs_control"], on_critical_path=True,
                                             what="pass 1 + controller (reduces pass 1's per-block partials, then the recursion on coefficients) + pass 2 at %d stored pairs; "
                                                  "serial_ms = the one-block controller" % m_avg)

    # ---- next rows of the scope table: kernels right behind the loop, measured the same way
    mask_row = assemble_row = mip_row = batch_row = None
    if world == 1:
        if not args.no_batch_row:
            batch_row = bench_batched_small_frames(dev, synth, with_concurrent=args.batch_row_concurrent)
        mask_row = bench_mask_composite(dev, pk)
        assemble_row = bench_video_assemble(dev, pk)
        mip_row = bench_mip(dev, pk)

    # ---- the reference's own eager path on THIS GPU (cuDNN TF32 convolutions, fp32 bmm, torch.optim-style L-BFGS with a host
    #      sync per iteration): the only pre-existing Blackwell path for the workload
    eager = None
    if world == 1 and not args.no_torch_eager:
        eager = time_torch_eager_gpu(S, dev)

    # ---- CPU baseline (rank 0, N = 1 only): the oracle port on the host cores, bounded sample
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu, _, _ = time_oracle(S, budget_s=args.cpu_budget, max_evals=20)

    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=Wm, ms_per_step=ms / K,
                higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="f16", data="synthetic",
                config=dict(workload=workload_name(S)),    # identical in both arms (the driver compares the dicts)
                detail=dict(precision="fp16 operands / fp32 accumulate (VGG forward, Gram), bf16 operands (data gradients), fp32 pixel terms and optimizer",
                            l2="working set per evaluation (activations ~0.3 GB + L-BFGS history up to 0.63 GB) exceeds the 126 MB L2; no flush needed",
                            history_pairs_at_start=hist_start, history_pairs_at_end=hist_len, warmup_evals_run=warm_evals,
                            final_loss=final_loss, flops_per_eval=synth.eval_flops(S, S),
                            note="`value` times exactly --steps evaluations after the warm-up (history length above); `value_320` and `e2e` time the "
                                 "whole 320-evaluation job of BASELINE configs[1]"),
                clocks=clocks, e2e=e2e, value_320=value_320, sanity=sanity, gpu_launches=launches, roofline=roofline, roofline_hbm=roofline_hbm,
                roofline_small=roofline_small, ms_per_eval_by_kernel=by_kind, cpu_baseline=cpu, torch_eager_gpu=eager, video=video_row,
                batched_small_frames=batch_row, mask_composite=mask_row, video_assemble=assemble_row, mip_planes=mip_row)
    emit(line)
    if args.kernel_table:
        # per-launch table: the same step with an event after EVERY launch (isolates each launch: no overlap between launches)
        with torch.cuda.stream(stream):
            raw1 = plan.lbfgs_step_timed() + plan.lbfgs_step_timed()
        acc1 = {}
        for k, l, ms_ in raw1:
            a = acc1.setdefault((k, l), [0.0, 0])
            a[0] += ms_
            a[1] += 1
        rows = [(k, l, v[0] / v[1]) for (k, l), v in acc1.items()]
        with open(args.kernel_table, "w") as f:
            f.write("kind,conv,ms,gflop,tflops\n")
            for k, l, ms_ in rows + lb_rows:
                fl = 0.0
                if k in ("conv_fwd", "conv_dgrad"):
                    fl = synth.conv_flops(l, S, S)
                elif k == "gram_bwd":
                    fl = synth.gram_flops(l, S, S)
                elif k == "gram":
                    fl = sum(synth.gram_flops(i, S, S) for i in synth._STYLE)
                f.write("%s,%d,%.5f,%.3f,%.1f\n" % (k, l, ms_, fl / 1e9, fl / (ms_ * 1e-3) / 1e12 if ms_ > 0 else 0.0))
    if dist is not None:
        dist.destroy_process_group()


def run_video(args, rank, world, local_rank):
    """BASELINE configs[4]: synthetic 720p frames, one shared style, frame-sharded over the ranks (video.py).  Reports
    frames/s; every frame is a full run_multi_style_transfer on host uint8 buffers (H2D, targets, loop, D2H) and the
    finished frames are all-gathered in order.  Not the default workload (a frame at the reference's num_steps=400
    costs seconds); run with --workload video."""
    import torch
    import nst_b200
    from nst_b200 import synth, video
    from importlib import import_module
    hf = import_module("text-based-image-style-transfer_b200.multi_style_transfer.helper_functions")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    H, W = (int(v) for v in args.video_size.split("x"))
    ws, bs = synth.vgg19_random_weights(1234, 13)
    hf.set_vgg_weight_provider(lambda: (ws, bs))
    style = torch.from_numpy(synth.synth_image(512, 512, 1)).permute(2, 0, 1).float().div(255).unsqueeze(0).to(dev)
    n_frames = args.frames
    # a slowly drifting field: frame k = blend of two fixed 1/f^2 fields (independence of the frames is what matters)
    f0 = torch.from_numpy(synth.synth_image(H, W, 100)).float()
    f1 = torch.from_numpy(synth.synth_image(H, W, 101)).float()
    frames = torch.stack([((1 - k / max(n_frames - 1, 1)) * f0 + (k / max(n_frames - 1, 1)) * f1).round().to(torch.uint8)
                          for k in range(n_frames)])
    styler = video.FrameStyler(synth.VGG_MEAN, synth.VGG_STD, (H, W), [style], num_steps=args.video_steps, device=dev,
                               concurrent=args.concurrent, batch=args.batch, **synth.APP_WEIGHTS)
    styler.process_block(frames[:max(args.concurrent, args.batch)])   # warm-up: graph capture of every plan, allocator
    video.gather_frames(frames[:1].clone(), world, dev)    # ... and the NCCL communicator of the all-gather
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = video.run_sharded(frames, styler, dev)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    dt = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([dt], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t[0])
    evals = 20 * (args.video_steps // 20 + 1)
    if rank == 0:
        line = dict(metric="video style transfer frames/s (%dx%d, frame-sharded)" % (W, H), value=n_frames / dt, unit="frames/s", n_gpus=world,
                    steps=n_frames, warmup=1, ms_per_step=1e3 * dt / n_frames, higher_is_better=True,
                    scaling="strong (--frames is the TOTAL over all ranks: keep it fixed when comparing N)",
                    vs_baseline=None, dtype="f16", data="synthetic",
                    config=dict(workload="%d synthetic %dx%d frames, one shared 512x512 style, num_steps=%d (%d evaluations per frame), "
                                         "%s per GPU, contiguous frame blocks per rank, NCCL broadcast of the style Gram "
                                         "targets + all-gather of the finished frames (BASELINE configs[4], reduced frame count / steps)"
                                         % (n_frames, W, H, args.video_steps, evals,
                                            ("%d frames per LAUNCH (batch dimension inside the convolution / Gram kernels)" % args.batch) if args.batch > 1
                                            else "%d frame(s) in flight (one plan / stream each)" % args.concurrent)),
                    evals_per_s=n_frames * evals / dt, checksum=int(out.to(torch.int64).sum()))
        emit(line)
    if dist is not None:
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line of the contract, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # libraries (NCCL's version banner, torchrun's OMP notice) print to stdout; the contract is one JSON line there:
    # everything except that line goes to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=320, help="closure evaluations to time (num_steps=300 runs 320)")
    ap.add_argument("--warmup", type=int, default=40)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--cpu-budget", type=float, default=25.0, help="seconds of CPU work for the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-eager", action="store_true", help="skip the torch-eager-on-GPU timing of the reference algorithm")
    ap.add_argument("--no-batch-row", action="store_true", help="skip the `batched_small_frames` row (256x256 frames, several per launch)")
    ap.add_argument("--batch-row-concurrent", action="store_true", help="`batched_small_frames` row: also time four frames in flight on four plans / streams")
    ap.add_argument("--no-video-row", action="store_true", help="skip the 720p frame-sharded `video` row of the default workload")
    ap.add_argument("--video-row-frames", type=int, default=64, help="`video` row: total 720p frames, the same at every N (strong scaling)")
    ap.add_argument("--video-row-batch", type=int, default=2, help="`video` row: frames per launch on every GPU (FrameStyler(batch=B)); 1 = one at a time")
    ap.add_argument("--video-row-steps", type=int, default=60, help="`video` row: num_steps per frame (60 -> 80 evaluations)")
    ap.add_argument("--kernel-table", default=None, help="write the per-launch timing table (CSV) here")
    ap.add_argument("--workload", default="pair512", choices=["pair512", "video"])
    ap.add_argument("--frames", type=int, default=16, help="--workload video: number of 720p frames (all ranks together)")
    ap.add_argument("--video-steps", type=int, default=100, help="--workload video: num_steps per frame (the reference UI uses 400)")
    ap.add_argument("--video-size", default="720x1280", help="--workload video: frame size HxW")
    ap.add_argument("--concurrent", type=int, default=1, help="--workload video: frames in flight per GPU (FrameStyler(concurrent=K))")
    ap.add_argument("--batch", type=int, default=1, help="--workload video: frames per launch (FrameStyler(batch=B): batch dimension inside the "
                                                         "convolution / Gram kernels, SURVEY 8f row 2); frame sizes must be multiples of 16")
    args = ap.parse_args()
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload == "video":
        run_video(args, rank, world, local_rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        sys.stderr.write("bench.py: --gpus %d needs torchrun (one rank per GPU); running rank 0 alone as N=1\n" % args.gpus)
    run_b200(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
