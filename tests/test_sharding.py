"""Multi-GPU host logic on the CPU: world_size-2 gloo processes exercise the frame partition, the style-target
broadcast and the ordered gather of video.py (the GPU path swaps gloo for NCCL and the fake frame processor for
FrameStyler; no collective sits inside the optimisation step)."""
import importlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def video():
    return importlib.import_module("text-based-image-style-transfer_b200.video")


def test_shard_range_covers_everything_once():
    v = video()
    for n in (0, 1, 5, 240, 241, 7):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                lo, hi = v.shard_range(n, world, r)
                assert 0 <= lo <= hi <= n
                got += list(range(lo, hi))
            assert got == list(range(n))
            sizes = [v.shard_range(n, world, r)[1] - v.shard_range(n, world, r)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    assert video().shard_range(240, 8, 3) == (90, 120)      # 240 frames divide evenly by 1/2/4/8 (BASELINE configs[4])


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        v = video()
        layers = ["conv1_1", "conv2_1", "conv3_1", "conv4_1", "conv5_1"]
        targets = None
        if rank == 0:
            g = torch.Generator().manual_seed(7)
            targets = {n: torch.randn(1, v.STYLE_CHANNELS[n], v.STYLE_CHANNELS[n], generator=g) for n in layers}
        got = v.broadcast_style_targets(targets, layers, "cpu")
        chk = float(sum(t.double().sum() for t in got.values()))
        frames = (torch.arange(n_frames * 6 * 5 * 3) % 251).to(torch.uint8).reshape(n_frames, 6, 5, 3)
        seen = []

        def process(k, f):
            seen.append(k)
            return (255 - f.to(torch.int16)).to(torch.uint8) + torch.tensor(k % 3, dtype=torch.uint8)

        out = v.run_sharded(frames, process, "cpu")

        class BlockProcessor:      # the shape of FrameStyler(concurrent=K): run_sharded hands it this rank's whole block
            blocks = []

            def __call__(self, k, f):
                raise AssertionError("run_sharded must use process_block when there is one")

            def process_block(self, fr, out_device=None):
                self.blocks.append(int(fr.shape[0]))
                return 255 - fr

        bp = BlockProcessor()
        out_b = v.run_sharded(frames, bp, "cpu")
        lo, hi = v.shard_range(n_frames, world, rank)
        assert torch.equal(out_b, 255 - frames) and bp.blocks == ([hi - lo] if hi > lo else [])
        want = torch.stack([(255 - frames[k].to(torch.int16)).to(torch.uint8) + torch.tensor(k % 3, dtype=torch.uint8)
                            for k in range(n_frames)])
        q.put((rank, chk, seen, bool(torch.equal(out, want))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [7, 8, 1])
def test_two_rank_frame_sharding_with_gloo(n_frames):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, chk0, seen0, ok0), (r1, chk1, seen1, ok1) = res
    assert chk0 == chk1                                    # every rank holds rank 0's style targets
    assert ok0 and ok1                                     # gathered frames are complete and in frame order
    assert seen0 + seen1 == list(range(n_frames))          # each frame processed exactly once, contiguous blocks


def test_single_process_paths_need_no_process_group():
    v = video()
    frames = torch.zeros(3, 4, 4, 3, dtype=torch.uint8)
    out = v.run_sharded(frames, lambda k, f: f + 1, "cpu")
    assert out.shape == frames.shape and int(out.max()) == 1
    t = {"conv1_1": torch.ones(1, 64, 64)}
    assert v.broadcast_style_targets(t, ["conv1_1"], "cpu") is t
