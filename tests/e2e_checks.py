"""End-to-end parity helpers: engine (libvda) forward vs the oracle restatement on the same synthetic weights /
inputs, with per-stage comparison for debugging.  Used by tests/test_forward_gpu.py and runnable as a script."""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import vda_oracle as O  # noqa: E402
from video_depth_anything_b200 import MODEL_CONFIGS, VideoDepthAnything, synth_state_dict  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def build_model(enc, seed, dtype):
    sd = synth_state_dict(**MODEL_CONFIGS[enc], seed=seed)
    m = VideoDepthAnything(**MODEL_CONFIGS[enc], dtype=dtype)
    m.load_state_dict(sd, strict=True)
    return m.to("cuda"), sd


def stage_report(enc, seed, shape, xseed, dtype, oracle_device="cpu"):
    """returns (final rel err tuple, list of (stage, rel max, rel mean))"""
    m, sd = build_model(enc, seed, dtype)
    x = torch.randn(shape, generator=torch.Generator().manual_seed(xseed))
    est, ost = {}, {}
    d = m.forward(x.cuda(), est).float().cpu()
    sdo = {k: v.to(oracle_device) for k, v in sd.items()}
    ref = O.forward(sdo, x.to(oracle_device), enc, ost).float().cpu()
    rows = []
    for k, v in ost.items():
        if k not in est:
            continue
        e = est[k]
        v = v.float().cpu()
        if isinstance(e, tuple):
            t, h, w = e
            c = v.shape[1]
            e = t.float().cpu().reshape(v.shape[0], h * w, -1)[..., :c]
            v = v.permute(0, 2, 3, 1).reshape(v.shape[0], h * w, c)
        else:
            e = e.float().cpu().reshape(v.shape)
        # intermediate activations cross zero, so use a range-normalised error (per-element relative error is
        # only meaningful for the strictly positive depth map)
        scale = v.abs().max().item() + 1e-12
        diff = (e - v).abs()
        rows.append((k, diff.max().item() / scale, diff.mean().item() / scale))
    return O.rel_err(d, ref), rows, d, ref


def golden_case(name, dtype):
    man = json.load(open(os.path.join(GOLD, "MANIFEST.json")))["cases"][name]
    m, sd = build_model(man["encoder"], man["seed"], dtype)
    x = torch.randn(man["x_shape"], generator=torch.Generator().manual_seed(man["x_seed"]))
    d = m.forward(x.cuda()).float().cpu()
    g = torch.from_numpy(np.load(os.path.join(GOLD, name + ".npz"))["depth"])
    s = man["stride"]
    return O.rel_err(d[..., ::s, ::s], g), d


def main():
    dt = {"bf16": torch.bfloat16, "fp16": torch.float16}[sys.argv[1] if len(sys.argv) > 1 else "bf16"]
    cases = [("vits", 0, (1, 8, 3, 56, 70), 1234), ("vits", 1, (1, 32, 3, 42, 42), 1235), ("vitl", 0, (1, 4, 3, 28, 42), 1236)]
    for enc, seed, shape, xs in cases:
        fin, rows, d, ref = stage_report(enc, seed, shape, xs, dt)
        print(f"=== {enc} {shape} {dt}: final rel max {fin[0]:.3e} p99.9 {fin[1]:.3e} mean {fin[2]:.3e}; "
              f"depth mean {ref.mean():.4f} engine mean {d.mean():.4f}", flush=True)
        for k, mx, mean in rows:
            print(f"    {k:14s} range-normalised err max {mx:.3e} mean {mean:.3e}", flush=True)


if __name__ == "__main__":
    main()
