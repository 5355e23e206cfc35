#!/bin/bash
# round-2 call 19: fold epilogue with the row statistics requested up front (A/B against HEAD~ build in variants/)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 200 python tests/kernel_checks.py "fold" "gemm 2740x3072x1024" > $O/c19_checks.log 2>&1; echo "checks rc=$?"; grep -E "FAIL|EXC|failing" $O/c19_checks.log | head
for v in old new old new; do
  lib=$PWD/video_depth_anything_b200/libvda.so; [[ $v == old ]] && lib=$PWD/variants/libvda_statsold.so
  echo "=== $v"; VDA_LIB=$lib timeout 120 python tools/bench_gemm.py fold 2>&1 | grep -E "fc1|qkv"
done
B="--steps 20 --warmup 4 --no-cpu-baseline --video-frames 0 --no-other-configs --no-e2e"
for v in old new old new; do
  lib=$PWD/video_depth_anything_b200/libvda.so; [[ $v == old ]] && lib=$PWD/variants/libvda_statsold.so
  VDA_LIB=$lib timeout 200 python bench.py $B 2>/dev/null | python -c "
import json,sys
l=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', 'fps', round(l['value'],1), 'p50', round(l['p50_window_latency_ms'],2), l['clocks']['sm_mhz'])"
done
