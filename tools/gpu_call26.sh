#!/bin/bash
# round-2 call 26 (8 GPUs): the driver's scaling command at N = 8 on the final tree (two-phase sharded long-video arm + window arm)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
VDA_TRACE_VIDEO=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 > $O/c26_bench8.json 2> $O/c26_bench8.err; echo "bench8 rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/c26_bench8.json'):
    if l.startswith("{"):
        d=json.loads(l); print("window", round(d["value"],1), "e2e", round(d["e2e"]["value"],1), "video", round(d["video_frames_per_s"],1), round(d["video_seconds"],4), d["video_crc32"], "other", {k: round(v["frames_per_s"],1) for k,v in d["other_configs"].items()})
PY
tail -2 $O/c26_bench8.err | cut -c1-300
