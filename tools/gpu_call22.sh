#!/bin/bash
# round-2 call 22: bounded device buffers of the long-video driver (upload ring, aligned-frame ring)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 600 python -m pytest tests/test_forward_gpu.py -m gpu -x -q -k "bounded or feature_reuse or infer_video or sharded or golden" > $O/c22_video_tests.log 2>&1; echo "video tests rc=$?"; tail -4 $O/c22_video_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/c22_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/c22_smoke.log
timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-other-configs --no-e2e > $O/c22_bench.json 2> $O/c22_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.loads([x for x in open("gpurun_out/c22_bench.json") if x.startswith("{")][-1])
print("fps", round(l["value"],1), "video", round(l["video_frames_per_s"],1), l["video_seconds"], l["video_crc32"], "(expect 7cc8d955)")
PY
nvidia-smi --query-gpu=memory.used --format=csv | tail -1
