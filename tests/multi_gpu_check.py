#!/usr/bin/env python
"""N-GPU check of the window-sharded driver (run under torchrun on a multi-GPU box):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \\
        tests/multi_gpu_check.py

Every rank computes its block of windows of a synthetic video; rank 0 gathers (NCCL), aligns and compares with the
single-GPU result of the same model: windows are independent in model compute and the kernels are batch-invariant,
so the two must be bit-identical."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from video_depth_anything_b200 import MODEL_CONFIGS, VideoDepthAnything, synth_state_dict  # noqa: E402
from video_depth_anything_b200.parallel import infer_video_depth_sharded  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    enc = os.environ.get("VDA_ENCODER", "vits")
    m = VideoDepthAnything(**MODEL_CONFIGS[enc], dtype=torch.bfloat16)
    m.load_state_dict(synth_state_dict(**MODEL_CONFIGS[enc], seed=0))
    m.to(f"cuda:{local}")
    n = int(os.environ.get("VDA_FRAMES", "100"))
    frames = np.random.default_rng(0).integers(0, 256, (n, 98, 126, 3), dtype=np.uint8)
    ref = m.infer_video_depth(frames, 24, input_size=98, device=f"cuda:{local}")[0] if rank == 0 else None
    for mode in ("two_phase", "stream"):          # shared-memory two-phase form and the rank-0 streaming form
        os.environ["VDA_SHARD_MODE"] = mode
        out, _ = infer_video_depth_sharded(m, frames, 24, input_size=98)
        if rank == 0:
            same = np.array_equal(out, ref)
            print(f"multi_gpu_check[{mode}]: world {dist.get_world_size()} frames {n} windows {-(-n // 22)}: "
                  f"{'bit-identical' if same else 'MISMATCH max abs ' + str(np.abs(out - ref).max())}", flush=True)
            assert same
        dist.barrier()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
