#!/usr/bin/env python
"""Stage-by-stage error of the engine against the fp32 oracle (both on the GPU) for one encoder / input shape, in bf16
and fp16: where does the end-to-end error of a configuration come from?   python tools/stage_errors.py vits 2 518 518"""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from e2e_checks import stage_report  # noqa: E402

enc = sys.argv[1] if len(sys.argv) > 1 else "vits"
T, H, W = (int(a) for a in sys.argv[2:5]) if len(sys.argv) > 4 else (2, 518, 518)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
for dt in (torch.bfloat16, torch.float16):
    fin, rows, d, ref = stage_report(enc, 0, (1, T, 3, H, W), 1234, dt, oracle_device="cuda")
    print(f"=== {enc} 1x{T}x{H}x{W} {dt}: final rel err max {fin[0]:.3e} p99.9 {fin[1]:.3e} mean {fin[2]:.3e}", flush=True)
    for k, mx, mean in rows:
        print(f"    {k:14s} range-normalised err max {mx:.3e} mean {mean:.3e}", flush=True)
