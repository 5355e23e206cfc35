#!/bin/bash
# round-2 final validation on one B200: full GPU test suite, smoke, the driver's bench commands (both arms), ncu launch list
# and ncu --set full of the four encoder GEMMs as the engine launches them
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r02_pytest_gpu.txt 2>&1; echo "pytest gpu rc=$?"; tail -3 $O/r02_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 --profile-out $O/r02_kernel_table.json > $O/r02_bench_vitl.json 2> $O/r02_bench_vitl.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.loads([x for x in open("gpurun_out/r02_bench_vitl.json") if x.startswith("{")][-1])
print("fps", round(l["value"],1), "p50", round(l["p50_window_latency_ms"],2), "e2e", round(l["e2e"]["value"],1), "video", l["video_frames_per_s"], l["video_crc32"], "roofline frac", round(l["roofline"]["frac"],3), "attn", round(l["roofline"]["other_kernels"]["attention_spatial"]["frac"],3), l["clocks"], "cpu", l["cpu_baseline"])
print("other", l["other_configs"])
PY
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err; echo "reference rc=$?"; cut -c1-400 $O/r02_bench_reference.json
A="--steps 2 --warmup 3 --no-cpu-baseline --no-e2e --video-frames 0 --no-other-configs"
timeout 300 python bench.py $A > $O/r02_ncu_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r02_launches.csv python bench.py $A > $O/r02_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
timeout 120 python tools/bench_gemm.py foldprof > $O/r02_gemm_plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -c 4 -o $O/gemm_prof python tools/bench_gemm.py foldprof > $O/r02_gemm_ncu.log 2>&1
echo "ncu gemm rc=$?"
timeout 120 python tools/bench_gemm.py fold > $O/r02_gemm_fold.log 2>&1; cat $O/r02_gemm_fold.log
