#!/usr/bin/env python
"""Window latency of the other BASELINE.json configurations on one GPU (CUDA events, graph replay, L2 flushed):
configs[3] metric ViT-L 1x32x518x924 and configs[4] ViT-S 1x32x518x518 (one clip per GPU)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_depth_anything_b200 import MODEL_CONFIGS, VideoDepthAnything, synth_state_dict  # noqa: E402


def run(enc, shape, dtype=torch.bfloat16, iters=5, metric=False):
    m = VideoDepthAnything(**MODEL_CONFIGS[enc], dtype=dtype, metric=metric)
    m.load_state_dict(synth_state_dict(**MODEL_CONFIGS[enc], seed=0))
    m.to("cuda")
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(1234)).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        m.forward(x)
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        m.forward(x)
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ms = sorted(ts)[len(ts) // 2]
    frames = shape[0] * shape[1]
    print(f"{enc} {'x'.join(map(str, shape))} {str(dtype).split('.')[-1]}: {ms:.2f} ms per window, {frames / ms * 1e3:.0f} frames/s, "
          f"peak memory {torch.cuda.max_memory_allocated() / 1e9:.1f} GB", flush=True)
    del m
    torch.cuda.empty_cache()


if __name__ == "__main__":
    run("vitl", (1, 32, 3, 518, 924), metric=True)
    run("vits", (1, 32, 3, 518, 518))
    run("vits", (4, 32, 3, 518, 518))
    run("vitl", (1, 32, 3, 518, 518), dtype=torch.float16)
