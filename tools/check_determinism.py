#!/usr/bin/env python
"""Run the forward twice on the same input and report the first stage whose values differ bit-wise (debug tool)."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_depth_anything_b200 import MODEL_CONFIGS, VideoDepthAnything, synth_state_dict  # noqa: E402


def flat(st):
    out = {}
    for k, v in st.items():
        out[k] = (v[0] if isinstance(v, tuple) else v).clone()
    return out


def main():
    for enc, shape, dt in (("vits", (1, 4, 3, 42, 56), torch.bfloat16), ("vits", (1, 8, 3, 56, 70), torch.float16),
                           ("vitl", (1, 4, 3, 98, 126), torch.bfloat16), ("vits", (1, 32, 3, 518, 518), torch.bfloat16)):
        m = VideoDepthAnything(**MODEL_CONFIGS[enc], dtype=dt)
        m.load_state_dict(synth_state_dict(**MODEL_CONFIGS[enc], seed=0))
        m.to("cuda")
        x = torch.randn(*shape, generator=torch.Generator().manual_seed(3)).cuda()
        bad_runs = 0
        for rep in range(6):
            sa, sb = {}, {}
            a = m.forward(x, stages=sa)
            fa = flat(sa)
            b = m.forward(x, stages=sb)
            fb = flat(sb)
            torch.cuda.synchronize()
            if not torch.equal(a, b):
                bad_runs += 1
                for k in fa:
                    if not torch.equal(fa[k], fb[k]):
                        d = (fa[k].float() - fb[k].float()).abs()
                        print(f"  {enc} {shape} {dt} rep {rep}: first differing stage {k}: {int((d > 0).sum())} of "
                              f"{d.numel()} elements, max {d.max().item():.3e}", flush=True)
                        break
        print(f"{enc} {shape} {dt}: {bad_runs} of 6 repeat pairs differ", flush=True)


if __name__ == "__main__":
    main()
