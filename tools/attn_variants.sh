#!/bin/bash
# build-and-time variants of the spatial attention kernel on the GPU box (debug defines of attention_spatial.cu)
cd "$(dirname "$0")/.."
for v in "$@"; do
  echo "=== variant: $v"
  VDA_NVCC_EXTRA="$v" python -m video_depth_anything_b200.build --force > /dev/null 2>&1 || { echo build failed; continue; }
  python tools/bench_attention.py 2>&1 | grep -E "32x1370x16|2x1370x16"
  if [[ "$v" == *VDA_SA_TIMING* ]]; then python tools/bench_attention.py timing | tail -3; fi
done
