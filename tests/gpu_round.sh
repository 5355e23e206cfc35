#!/bin/bash
# standard GPU round: per-kernel parity table, parity tests, smoke, bench (with per-kernel table), and -- with NCU=1 --
# the ncu launch list of one bench step (B200_PROFILING.md recipe; only after the same command exited 0 without ncu)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 600 python tests/kernel_checks.py > gpurun_out/kernels.log 2>&1
echo "kernel_checks rc=$?" >> gpurun_out/kernels.log
grep -vE '^ok ' gpurun_out/kernels.log | tail -n 20
if [ "${SKIP_PYTEST:-0}" != "1" ]; then
  timeout -k 5 1500 python -m pytest tests -q -m gpu -rA > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  grep -E 'rel err|passed|failed|FAILED|rc=' gpurun_out/pytest_gpu.log | tail -n 30
fi
timeout -k 5 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -n 3 gpurun_out/smoke.log
timeout -k 5 900 python bench.py --steps 5 --warmup 3 --profile-out gpurun_out/profile_vitl.json > gpurun_out/bench_vitl.log 2>&1
echo "bench rc=$?" >> gpurun_out/bench_vitl.log
tail -n 5 gpurun_out/bench_vitl.log
if [ "${NCU:-0}" = "1" ]; then
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --video-frames 0"
  N=$(python -c "import json;print(json.loads(open('gpurun_out/bench_vitl.log').readline())['gpu_launches']//5)" 2>/dev/null || echo 340)
  $CMD > gpurun_out/ncu_plain.log 2>&1 &&
  timeout -k 5 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $((3 * N)) -c $N --csv \
      --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
  echo "ncu launches rc=$? (N=$N)"
fi
