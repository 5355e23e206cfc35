#!/bin/bash
# round-2 call 29: re-validation after the tail-kernel change (bit-identical output): GPU tests, smoke, bench line
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02e_pytest_gpu.txt 2>&1; echo "pytest gpu rc=$?"; tail -3 $O/r02e_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02e_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02e_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 --profile-out $O/r02e_kernel_table.json > $O/r02e_bench_vitl.json 2> $O/r02e_bench_vitl.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.loads([x for x in open("gpurun_out/r02e_bench_vitl.json") if x.startswith("{")][-1])
print("fps", round(l["value"],1), "p50", round(l["p50_window_latency_ms"],2), "e2e", round(l["e2e"]["value"],1), "video", round(l["video_frames_per_s"],1), l["video_crc32"], "roofline frac", round(l["roofline"]["frac"],3), l["clocks"])
t=json.load(open("gpurun_out/r02e_kernel_table.json"))
print([ (r["kernel"], round(r["ms"],3)) for r in t["per_step"] if r["kernel"] in ("tail_fused","attention_spatial")])
print("other", {k: round(v["frames_per_s"],1) for k,v in l["other_configs"].items()})
PY
