#!/bin/bash
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
VDA_TRACE_VIDEO=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 3 --warmup 3 --no-other-configs --no-e2e > $O/c13_bench8.json 2> $O/c13_bench8.err; echo "bench8 rc=$?"
python - <<'PY'
import re, json
t=open('gpurun_out/c13_bench8.json').read()
for p in re.split(r'(?=video trace rank \d:)', t)[-17:]:
    if p.startswith('video trace'): print(p.strip()[:330])
for l in t.splitlines():
    if l.startswith("{"):
        d=json.loads(l); print("video", {k:v for k,v in d["video"].items() if k!="workload"})
PY
