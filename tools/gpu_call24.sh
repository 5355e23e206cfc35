#!/bin/bash
# round-2 call 24 (2 GPUs): sharded driver with the bounded buffers: bit-identity tests + CRC of the long-video arm
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q > $O/c24_multi_gpu_tests.log 2>&1; echo "multi-gpu tests rc=$?"; tail -3 $O/c24_multi_gpu_tests.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 5 --warmup 3 --no-other-configs --no-e2e > $O/c24_bench2.json 2> $O/c24_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
l=json.loads([x for x in open("gpurun_out/c24_bench2.json") if x.startswith("{")][-1])
print("fps", round(l["value"],1), "video", round(l["video_frames_per_s"],1), round(l["video_seconds"],4), l["video_crc32"], "(expect 7cc8d955)")
PY
VDA_SHARD_MODE=stream timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 3 --warmup 3 --no-other-configs --no-e2e > $O/c24_bench2_stream.json 2> $O/c24_bench2_stream.err; echo "bench2 stream rc=$?"
python - <<'PY'
import json
l=json.loads([x for x in open("gpurun_out/c24_bench2_stream.json") if x.startswith("{")][-1])
print("stream mode: video", round(l["video_frames_per_s"],1), round(l["video_seconds"],4), l["video_crc32"], "(expect 7cc8d955)")
PY
