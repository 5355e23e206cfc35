#!/bin/bash
# round-2 call 7 (8 GPUs): the driver's scaling command at N=8 (window arm + 2048-frame video arm + configs[3]/[4] arms)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
nvidia-smi -L | wc -l
VDA_TRACE_VIDEO=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 6 --warmup 3 --no-other-configs --no-e2e > $O/c7_bench8.json 2> $O/c7_bench8.err; echo "bench8 rc=$?"
grep -E "video trace" $O/c7_bench8.json $O/c7_bench8.err | tail -8 | cut -c1-260
python - <<'PY'
import json
for l in open("gpurun_out/c7_bench8.json"):
    if l.startswith("{"):
        d=json.loads(l); print("fps", round(d["value"],1), "e2e", round(d["e2e"]["value"],1)); print("video", {k:v for k,v in d["video"].items() if k!="workload"}); print("other", d.get("other_configs"))
PY
tail -4 $O/c7_bench8.err | cut -c1-300
