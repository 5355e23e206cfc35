#!/usr/bin/env python
"""Context number (not a bench arm): the oracle port -- the reference's forward restated op for op in plain PyTorch
(oracle/vda_oracle.py, naive attention path as in the reference without xformers) -- timed on the SAME B200 under
torch.autocast fp16 (the reference's default inference mode) and in fp32, ViT-L 1x32x518x518, CUDA events.
Usage (GPU box): python tests/torch_port_gpu_timing.py > gpurun_out/torch_port_gpu.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vda_oracle as O  # noqa: E402
from video_depth_anything_b200.synth import MODEL_CONFIGS, synth_state_dict  # noqa: E402


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]


def main():
    enc = sys.argv[1] if len(sys.argv) > 1 else "vitl"
    sd = {k: v.cuda() for k, v in synth_state_dict(**MODEL_CONFIGS[enc], seed=0).items()}
    x = torch.randn(1, 32, 3, 518, 518, generator=torch.Generator().manual_seed(1234)).cuda()
    out = {}
    with torch.no_grad():
        with torch.autocast("cuda", dtype=torch.float16):
            ms = timed(lambda: O.forward(sd, x, enc), 3)
        out["fp16_autocast"] = {"ms_per_window": ms, "frames_per_s": 32e3 / ms}
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        ms = timed(lambda: O.forward(sd, x, enc), 2)
        out["fp32"] = {"ms_per_window": ms, "frames_per_s": 32e3 / ms}
    out["what"] = f"oracle port of the reference forward in plain PyTorch {torch.__version__} on {torch.cuda.get_device_name(0)}, {enc} 1x32x518x518"
    out["peak_mem_gb"] = torch.cuda.max_memory_allocated() / 1e9
    print(json.dumps(out))


if __name__ == "__main__":
    main()
