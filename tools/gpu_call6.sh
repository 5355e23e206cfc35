#!/bin/bash
# round-2 call 6 (2 GPUs): bit-identity of the sharded driver (both exchange forms) + 2-GPU bench line with the video arm
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
nvidia-smi -L
VDA_FRAMES=230 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > $O/c6_mgc.log 2>&1; echo "multi_gpu_check rc=$?"
grep -E "multi_gpu_check|Error|error" $O/c6_mgc.log | head
VDA_TRACE_VIDEO=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 6 --warmup 3 --no-other-configs > $O/c6_bench2.json 2> $O/c6_bench2.err; echo "bench2 rc=$?"
grep -E "video trace" $O/c6_bench2.json $O/c6_bench2.err | head -8
python - <<'PY'
import json
for l in open("gpurun_out/c6_bench2.json"):
    if l.startswith("{"):
        d=json.loads(l); print("fps", round(d["value"],1), "video", d["video"])
PY
tail -5 $O/c6_bench2.err
