#!/bin/bash
# round-2 call 4: four-stream attention kernel (checks + timings), GEMM TMA epilogue with 16-bit copy, TAE tests, A/B bench
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
VDA_SA_KERNEL=4 timeout 200 python tests/kernel_checks.py "attn spatial" "attn rescale" > $O/c4_attn4_checks.log 2>&1; echo "attn4 checks rc=$?"
grep -E "FAIL|EXC|failing" $O/c4_attn4_checks.log | head
: > $O/c4_attn.log
for v in k2 k4 sa4poly50 sa4poly12 sa4poly0; do
  lib=$PWD/video_depth_anything_b200/libvda.so; k=4
  [[ $v == k2 ]] && k=2
  [[ $v == sa4* ]] && lib=$PWD/variants/libvda_$v.so
  echo "=== $v" >> $O/c4_attn.log
  VDA_SA_KERNEL=$k VDA_LIB=$lib timeout 120 python tools/bench_attention.py >> $O/c4_attn.log 2>&1
done
grep -E "===|32x1370|2x1370x16|2443" $O/c4_attn.log
timeout 300 python tests/kernel_checks.py "gemm" > $O/c4_gemm_checks.log 2>&1; echo "gemm checks rc=$?"
grep -E "FAIL|EXC|failing" $O/c4_gemm_checks.log | head
timeout 300 python -m pytest tests/test_eval_gpu.py -m gpu -x -q > $O/c4_eval.log 2>&1; echo "eval tests rc=$?"; tail -3 $O/c4_eval.log
B="--steps 20 --warmup 4 --no-cpu-baseline --video-frames 0 --no-other-configs"
for r in 1 2; do
  VDA_LN_FOLD=0 VDA_GEMM_TMA_EPI=0 timeout 300 python bench.py $B > $O/c4_bench_base$r.json 2> $O/c4_bench_base$r.err
  timeout 300 python bench.py $B --profile-out $O/c4_prof_fold$r.json > $O/c4_bench_fold$r.json 2> $O/c4_bench_fold$r.err
  VDA_SA_KERNEL=4 timeout 300 python bench.py $B --profile-out $O/c4_prof_fold_sa4_$r.json > $O/c4_bench_foldsa4_$r.json 2> $O/c4_bench_foldsa4_$r.err
done
python - <<'PY'
import json
for t in ("base1","fold1","foldsa4_1","base2","fold2","foldsa4_2"):
    try:
        l=json.loads(open(f"gpurun_out/c4_bench_{t}.json").read().strip().splitlines()[-1])
        print(t, "fps", round(l["value"],1), "p50", round(l["p50_window_latency_ms"],2), "e2e", round(l["e2e"]["value"],1), "attn", round(l["roofline"]["other_kernels"]["attention_spatial"]["achieved"]), l["clocks"]["sm_mhz"])
    except Exception as e: print(t, "ERR", e)
PY
