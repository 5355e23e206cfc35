#!/bin/bash
# 2-GPU: direct D2H two-phase: bit-identity + trace
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
VDA_FRAMES=230 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > $O/c9_mgc.log 2>&1; echo "multi_gpu_check rc=$?"
grep -E "multi_gpu_check|Error|error" $O/c9_mgc.log | head
VDA_TRACE_VIDEO=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-other-configs --no-e2e > $O/c9_bench2.json 2> $O/c9_bench2.err; echo "bench2 rc=$?"
grep -o "video trace rank [0-9]: [^v]*" $O/c9_bench2.json | tail -4 | cut -c1-420
python - <<'PY'
import json
for l in open("gpurun_out/c9_bench2.json"):
    if l.startswith("{"):
        d=json.loads(l); print("fps", round(d["value"],1), "video", {k:v for k,v in d["video"].items() if k!="workload"})
PY
tail -3 $O/c9_bench2.err | cut -c1-300
