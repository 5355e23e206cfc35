#!/bin/bash
# round-2 call 17: ViT-S per-kernel table (configs[0]/[4] geometry) with the round-2 kernels
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
B="--steps 20 --warmup 4 --no-cpu-baseline --video-frames 0 --no-other-configs --no-e2e"
timeout 300 python bench.py --encoder vits $B --profile-out $O/c17_prof_vits.json > $O/c17_bench_vits.json 2> $O/c17_bench_vits.err; echo "vits rc=$?"
python - <<'PY'
import json
l=json.loads(open("gpurun_out/c17_bench_vits.json").read().strip().splitlines()[-1])
print("vits fps", round(l["value"],1), "ms", round(l["ms_per_step"],3), l["clocks"])
t=json.load(open("gpurun_out/c17_prof_vits.json"))
for r in t["per_step"]: print(f"{r['kernel']:22s} {r['ms']:7.3f} ms x{r['launches']:4.0f} {r['tflops']:7.1f}")
for g in sorted(t["gemm_shapes"], key=lambda g: -g["ms_per_launch"]*g["launches"])[:24]:
    print(f"{g['kernel']:14s} M={g['M']:7d} N={g['N']:5d} K={g['K']:5d} x{g['launches']:3.0f} {g['ms_per_launch']*1e3:7.1f} us tot {g['ms_per_launch']*g['launches']:6.3f} ms {g['tflops']:7.1f} TF/s")
PY
