#!/bin/bash
# first-contact script for a fresh GPU box: kernel table, then end-to-end stage report
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout -k 5 600 python tests/kernel_checks.py > gpurun_out/kernels.log 2>&1
echo "kernel_checks rc=$?" >> gpurun_out/kernels.log
timeout -k 5 600 python tests/e2e_checks.py fp16 > gpurun_out/e2e_fp16.log 2>&1
echo "e2e rc=$?" >> gpurun_out/e2e_fp16.log
timeout -k 5 600 python tests/e2e_checks.py bf16 > gpurun_out/e2e_bf16.log 2>&1
echo "e2e rc=$?" >> gpurun_out/e2e_bf16.log
tail -n 70 gpurun_out/kernels.log
tail -n 40 gpurun_out/e2e_fp16.log
