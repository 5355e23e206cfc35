#!/bin/bash
# round-2 final validation (second pass, tree with the handle-level C API + C host test): full GPU test suite, smoke, both
# bench arms as the driver runs them, ncu launch list of the bench command
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02d_pytest_gpu.txt 2>&1; echo "pytest gpu rc=$?"; tail -3 $O/r02d_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02d_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02d_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 --profile-out $O/r02d_kernel_table.json > $O/r02d_bench_vitl.json 2> $O/r02d_bench_vitl.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.loads([x for x in open("gpurun_out/r02d_bench_vitl.json") if x.startswith("{")][-1])
print("fps", round(l["value"],1), "p50", round(l["p50_window_latency_ms"],2), "e2e", round(l["e2e"]["value"],1), "video", l["video_frames_per_s"], l["video_crc32"], "roofline frac", round(l["roofline"]["frac"],3), "traffic", l["roofline"]["traffic"], "attn", round(l["roofline"]["other_kernels"]["attention_spatial"]["frac"],3), l["clocks"], "launches", l["gpu_launches"])
print("other", {k: round(v["frames_per_s"],1) for k,v in l["other_configs"].items()})
PY
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02d_bench_reference.json 2> $O/r02d_bench_reference.err; echo "reference rc=$?"; cut -c1-300 $O/r02d_bench_reference.json
A="--steps 2 --warmup 3 --no-cpu-baseline --no-e2e --video-frames 0 --no-other-configs"
timeout 300 python bench.py $A > $O/r02d_ncu_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/r02d_launches.csv python bench.py $A > $O/r02d_ncu_launches.log 2>&1
echo "ncu launches rc=$?"
