#!/bin/bash
# 8 GPUs: the driver's scaling command (no trace)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 > $O/c12_bench8.json 2> $O/c12_bench8.err; echo "bench8 rc=$?"
python - <<'PY'
import json
for l in open("gpurun_out/c12_bench8.json"):
    if l.startswith("{"):
        d=json.loads(l); print("fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1)); print("video", {k:v for k,v in d["video"].items() if k!="workload"}); print("other", d.get("other_configs"))
PY
tail -3 $O/c12_bench8.err | cut -c1-300
