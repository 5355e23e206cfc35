#!/bin/bash
# round-2 call 18: what bounds the head's small-K GEMMs (K = 256)?  timings + ncu --set full of one launch each
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 120 python tools/bench_gemm.py smallk > $O/c18_smallk.log 2>&1; echo "smallk rc=$?"; cat $O/c18_smallk.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 6 -o $O/c18_smallk python tools/bench_gemm.py smallkprof > $O/c18_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $O/c18_ncu.log
