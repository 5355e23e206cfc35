#!/bin/bash
# round-2 call 5: four-stream attention (guarded: the rest of its tests only run if a tiny case passes quickly),
# validation-mode (hi|lo weights) GEMM + e2e, A/B
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
VDA_SA_KERNEL=4 timeout 60 python tests/kernel_checks.py "attn spatial 3x21x6" "attn spatial 1x128x1" "attn spatial 4x300x3" > $O/c5_attn4_small.log 2>&1
rc=$?; echo "attn4 small rc=$rc"; tail -4 $O/c5_attn4_small.log
if [[ $rc == 0 ]]; then
  VDA_SA_KERNEL=4 timeout 150 python tests/kernel_checks.py "attn spatial" "attn rescale" > $O/c5_attn4_checks.log 2>&1; rc=$?; echo "attn4 checks rc=$rc"
  grep -E "FAIL|EXC|failing" $O/c5_attn4_checks.log | head
fi
: > $O/c5_attn.log
if [[ $rc == 0 ]]; then
  for v in k2 k4 sa4poly50 sa4poly12 sa4poly0; do
    lib=$PWD/video_depth_anything_b200/libvda.so; k=4
    [[ $v == k2 ]] && k=2
    [[ $v == sa4* ]] && lib=$PWD/variants/libvda_$v.so
    echo "=== $v" >> $O/c5_attn.log
    VDA_SA_KERNEL=$k VDA_LIB=$lib timeout 90 python tools/bench_attention.py >> $O/c5_attn.log 2>&1
  done
  grep -E "===|32x1370|2x1370x16|2443" $O/c5_attn.log
fi
timeout 200 python tests/kernel_checks.py "hi|lo" "gemm 2740x3072x1024 bf16 qkv-like" "conv3x3 2x37x37" > $O/c5_split_checks.log 2>&1; echo "split checks rc=$?"
grep -E "ok|FAIL|EXC|failing" $O/c5_split_checks.log | head
timeout 400 python -m pytest tests/test_forward_gpu.py -m gpu -x -q -k "validation_mode or full_size_window" > $O/c5_fwd.log 2>&1; echo "fwd tests rc=$?"; grep -E "rel err|passed|failed|Error" $O/c5_fwd.log | head -20
if [[ $rc == 0 ]]; then
  B="--steps 20 --warmup 4 --no-cpu-baseline --video-frames 0 --no-other-configs"
  for r in 1 2; do
    timeout 200 python bench.py $B --profile-out $O/c5_prof_k2_$r.json > $O/c5_bench_k2_$r.json 2> $O/c5_bench_k2_$r.err
    VDA_SA_KERNEL=4 timeout 200 python bench.py $B --profile-out $O/c5_prof_k4_$r.json > $O/c5_bench_k4_$r.json 2> $O/c5_bench_k4_$r.err
  done
  python - <<'PY'
import json
for t in ("k2_1","k4_1","k2_2","k4_2"):
    try:
        l=json.loads(open(f"gpurun_out/c5_bench_{t}.json").read().strip().splitlines()[-1])
        print(t, "fps", round(l["value"],1), "p50", round(l["p50_window_latency_ms"],2), "e2e", round(l["e2e"]["value"],1), "attn", round(l["roofline"]["other_kernels"]["attention_spatial"]["achieved"]), l["clocks"]["sm_mhz"])
    except Exception as e: print(t, "ERR", e)
PY
fi
