#!/bin/bash
# 2-GPU trace of the two-phase tail (finer stamps)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
VDA_TRACE_VIDEO=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-other-configs --no-e2e > $O/c8_bench2.json 2> $O/c8_bench2.err; echo "bench2 rc=$?"
grep -o "video trace rank [0-9]: [^v]*" $O/c8_bench2.json | tail -4
python - <<'PY'
import json
for l in open("gpurun_out/c8_bench2.json"):
    if l.startswith("{"):
        d=json.loads(l); print("fps", round(d["value"],1), "video", {k:v for k,v in d["video"].items() if k!="workload"})
PY
