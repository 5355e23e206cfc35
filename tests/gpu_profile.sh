#!/bin/bash
# ncu --set full captures of the hot kernels in isolation (each command first runs once without ncu), written to
# gpurun_out/*.ncu-rep; summarise with tools/ncu_full_summary.py
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/bench_gemm.py prof > /dev/null 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -o gpurun_out/gemm_prof -f \
      python tools/bench_gemm.py prof > gpurun_out/gemm_ncu.log 2>&1
echo "gemm ncu rc=$?"
python tools/bench_attention.py prof > /dev/null 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:spatial_attention -c 1 -o gpurun_out/attn_prof -f \
      python tools/bench_attention.py prof > gpurun_out/attn_ncu.log 2>&1
echo "attn ncu rc=$?"
python tools/bench_gemm.py ln > /dev/null 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:layernorm -c 1 -o gpurun_out/ln_prof -f \
      python tools/bench_gemm.py ln > gpurun_out/ln_ncu.log 2>&1
echo "ln ncu rc=$?"
