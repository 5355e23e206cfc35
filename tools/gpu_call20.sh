#!/bin/bash
# round-2 call 20: staged (smem transpose) vs direct (thread = row) LINEAR epilogue on today's kernels
cd "$GRAFT_REPO_ROOT"
for s in 1 0; do echo "=== VDA_GEMM_STAGED=$s"; VDA_GEMM_STAGED=$s timeout 150 python tools/bench_gemm.py 2>&1 | grep -v "^$"; done
