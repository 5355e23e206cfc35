#!/bin/bash
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
df -h /dev/shm | tail -1; ulimit -l
VDA_TRACE_VIDEO=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --no-other-configs --no-e2e > $O/c11_bench2.json 2> $O/c11_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import re, json
t=open('gpurun_out/c11_bench2.json').read()
for p in re.split(r'(?=video trace rank \d:)', t):
    if p.startswith('video trace'): print(p.strip()[:330])
for l in t.splitlines():
    if l.startswith("{"):
        d=json.loads(l); print("video", {k:v for k,v in d["video"].items() if k!="workload"})
PY
