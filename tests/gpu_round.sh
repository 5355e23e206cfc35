#!/bin/bash
# standard GPU round: parity tests, smoke, bench (with per-kernel table)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout -k 5 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout -k 5 900 python bench.py --steps 5 --warmup 3 --profile-out gpurun_out/profile_vitl.json > gpurun_out/bench_vitl.log 2>&1
echo "bench rc=$?" >> gpurun_out/bench_vitl.log
tail -n 15 gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/smoke.log
tail -n 5 gpurun_out/bench_vitl.log
