#!/bin/bash
# round-2 call 30 (4 GPUs): the driver's scaling command at N = 4 on the final tree
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 10 --warmup 3 > $O/c30_bench4.json 2> $O/c30_bench4.err; echo "bench4 rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/c30_bench4.json'):
    if l.startswith("{"):
        d=json.loads(l); print("window", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "video", round(d["video_frames_per_s"],1), round(d["video_seconds"],4), d["video_crc32"], "other", {k: round(v["frames_per_s"],1) for k,v in d["other_configs"].items()})
PY
