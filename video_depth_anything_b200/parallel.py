"""Window-sharded multi-GPU driver for `infer_video_depth` (one process per GPU, torch.distributed).

The reference has no multi-GPU inference path at all.  Its long-video driver
(video_depth_anything/video_depth.py:166-254) computes K = ceil(N/22) overlapping 32-frame windows whose model
inputs depend on input frames only (closed form in windows.window_source_indices), followed by a cheap
*sequential* scale/shift alignment (:216-252).  So the windows are sharded in contiguous blocks over the ranks with
no data-path collective; the only exchange is a gather of the raw per-window depths to one rank (NCCL send/recv
over NVLink, 34 MB per 32x518x518 fp32 window), which then runs the alignment recurrence on its GPU.

Host-side pieces (`partition_windows`, `gather_window_depths`) work on CPU tensors with the gloo backend too;
that is how tests/test_parallel_cpu.py covers the N>1 logic without GPUs.
"""
from __future__ import annotations

import os
import time
from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.distributed as dist

from .windows import INFER_LEN, num_windows


def partition_windows(n_windows: int, world: int) -> List[range]:
    """Contiguous, balanced blocks of window ids per rank (sizes differ by at most one; earlier ranks get the
    extra window).  Contiguity keeps overlapping neighbours on one GPU."""
    if world <= 0:
        raise ValueError("world size must be positive")
    base, extra = divmod(n_windows, world)
    out, start = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append(range(start, start + n))
        start += n
    return out


def _world(group) -> tuple:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def gather_window_depths(local: torch.Tensor, counts: Sequence[int], dst: int = 0, group=None) -> Optional[torch.Tensor]:
    """Gather per-rank raw window depths `[counts[rank], 32, H, W]` (fp32) to rank `dst`, in rank order (= window
    order for `partition_windows`).  Ranks are padded to the largest count so a single `gather` collective moves
    everything (NCCL: grouped ncclSend/ncclRecv over NVLink).  Returns `[sum(counts), 32, H, W]` on `dst`, None
    elsewhere."""
    rank, world = _world(group)
    if world == 1:
        return local
    assert len(counts) == world and local.shape[0] == counts[rank]
    kmax = max(counts)
    if kmax == 0:
        return local if rank == dst else None
    pad = local
    if local.shape[0] < kmax:
        pad = torch.zeros((kmax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]].copy_(local)
    pad = pad.contiguous()
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, bufs, dst=dst, group=group)
    if rank != dst:
        return None
    return torch.cat([bufs[r][:counts[r]] for r in range(world)], dim=0)


def stream_window_depths(local: Optional[torch.Tensor], counts: Sequence[int], shape_hw, dtype, device, dst: int = 0,
                         group=None):
    """Point-to-point variant of the gather for the streaming driver: every rank other than `dst` sends its raw
    window stack `[counts[rank], 32, H, W]` to `dst` (NCCL send over NVLink); on `dst` this is a generator that
    posts all receives up front and yields `(rank, stack)` in rank order (= window order), so the caller aligns
    rank r's windows while rank r+1's are still in flight.  `dst`'s own windows are not part of the stream."""
    rank, world = _world(group)
    if world == 1:
        return
    if rank != dst:
        if counts[rank] > 0:
            assert local is not None and local.shape[0] == counts[rank]
            dist.send(local.contiguous(), dst=dst, group=group)
        return
    h, w = shape_hw
    bufs, works = {}, {}
    for r in range(world):
        if r != dst and counts[r] > 0:
            bufs[r] = torch.empty(counts[r], INFER_LEN, h, w, dtype=dtype, device=device)
            works[r] = dist.irecv(bufs[r], src=r, group=group)
    for r in sorted(bufs):
        works[r].wait()
        yield r, bufs[r]


@torch.no_grad()
def infer_video_depth_sharded(model, frames: np.ndarray, target_fps, input_size: int = 518, device=None,
                              dst: int = 0, group=None):
    """Drop-in for `model.infer_video_depth(frames, target_fps, input_size, device)` under torchrun: every rank
    passes the same `frames`; rank `dst` returns `(depths float32 [N,H0,W0], target_fps)`, the others `(None,
    target_fps)`.  With one process (no process group) it is exactly `model.infer_video_depth`."""
    from .video_depth import WindowAligner
    rank, world = _world(group)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    if world == 1:
        return model.infer_video_depth(frames, target_fps, input_size=input_size, device=device)
    n, h0, w0 = frames.shape[:3]
    parts = partition_windows(num_windows(n), world)
    counts = [len(p) for p in parts]
    if dst == 0:
        # streaming form: rank 0 owns the first block of windows, so it aligns them (and streams the finished frames
        # to the host) while it computes; the other ranks' stacks arrive point-to-point and are aligned in rank order
        if rank != 0:
            t0 = time.perf_counter()
            raw = model.infer_video_depth(frames, target_fps, input_size=input_size, device=device,
                                          window_ids=list(parts[rank]), raw_only=True)
            if os.environ.get("VDA_TRACE_VIDEO") == "1":
                t1 = time.perf_counter()
                torch.cuda.synchronize(raw.device)
                print(f"video trace rank {rank}: enqueued +{(t1 - t0) * 1e3:.0f} ms, computed "
                      f"+{(time.perf_counter() - t0) * 1e3:.0f} ms", flush=True)
            for _ in stream_window_depths(raw, counts, (h0, w0), torch.float32, raw.device, dst=0, group=group):
                pass
            return None, target_fps
        dev = torch.device(device)
        trace = os.environ.get("VDA_TRACE_VIDEO") == "1"       # debug: synchronising wall-clock stamps of the phases
        stamps = [("start", time.perf_counter())]

        def stamp(name):
            if trace:
                torch.cuda.synchronize(dev)
                stamps.append((name, time.perf_counter()))

        with torch.cuda.device(dev):
            aligner = WindowAligner(n, h0, w0, dev, "identity" if model.metric else "affine")
            stamp("aligner allocated")
            model.infer_video_depth(frames, target_fps, input_size=input_size, device=dev, window_ids=list(parts[0]),
                                    aligner=aligner)
            stamp("own windows computed + aligned")
            for r, stack in stream_window_depths(None, counts, (h0, w0), torch.float32, dev, dst=0, group=group):
                stamp(f"received rank {r}")
                for k in range(stack.shape[0]):
                    aligner.push(stack[k])
                stamp(f"aligned rank {r}")
            out = aligner.result()
            stamp("drained to host")
            if trace:
                print("video trace: " + ", ".join(f"{a} +{(t - stamps[0][1]) * 1e3:.0f} ms" for a, t in stamps[1:]), flush=True)
            return out, target_fps
    raw = model.infer_video_depth(frames, target_fps, input_size=input_size, device=device,
                                  window_ids=list(parts[rank]), raw_only=True)         # [k_r,32,h0,w0] on device
    allraw = gather_window_depths(raw, counts, dst=dst, group=group)
    if rank != dst:
        return None, target_fps
    with torch.cuda.device(allraw.device):
        aligner = WindowAligner(n, h0, w0, allraw.device, "identity" if model.metric else "affine")
        for k in range(allraw.shape[0]):
            aligner.push(allraw[k])
        return aligner.result(), target_fps


__all__ = ["partition_windows", "gather_window_depths", "stream_window_depths", "infer_video_depth_sharded", "INFER_LEN"]
