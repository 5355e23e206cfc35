#!/bin/bash
# round-2 call 27: tail kernel with the interpolation in packed 16-bit arithmetic (variant build) vs the fp32 form
cd "$GRAFT_REPO_ROOT"
for v in base fast base fast; do
  lib=$PWD/video_depth_anything_b200/libvda.so; [[ $v == fast ]] && lib=$PWD/variants/libvda_tailfast.so
  echo "=== $v"
  VDA_LIB=$lib timeout 100 python tools/bench_gemm.py tail 2>&1 | grep "fused tail"
done
echo "=== checks fast"; VDA_LIB=$PWD/variants/libvda_tailfast.so timeout 100 python tests/kernel_checks.py "tail" 2>&1 | grep -E "ok|FAIL|EXC|failing"
echo "=== checks base"; timeout 100 python tests/kernel_checks.py "tail" 2>&1 | grep -E "ok|FAIL|EXC|failing"
