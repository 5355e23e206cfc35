#!/bin/bash
# round-2 call 3: GEMM SPEC 4/5/6 checks, e2e with the LayerNorm fold + fp16 head, A/B bench, ncu of the attention kernel
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 300 python tests/kernel_checks.py "gemm" "align chain" "layernorm" > $O/c3_gemm_checks.log 2>&1; echo "gemm checks rc=$?"
grep -E "FAIL|EXC|failing|Error" $O/c3_gemm_checks.log | head -20
timeout 200 python tools/stage_errors.py vitl 2 518 518 > $O/c3_stage_vitl.log 2>&1; echo "stage vitl rc=$?"
grep -E "===|Error|error" $O/c3_stage_vitl.log | head
timeout 200 python tools/stage_errors.py vits 2 518 518 > $O/c3_stage_vits.log 2>&1; echo "stage vits rc=$?"
grep -E "===|Error|error" $O/c3_stage_vits.log | head
VDA_LN_FOLD=0 timeout 200 python tools/stage_errors.py vitl 2 518 518 > $O/c3_stage_vitl_nofold.log 2>&1; echo "stage vitl nofold rc=$?"
grep -E "===" $O/c3_stage_vitl_nofold.log | head
B="--steps 6 --warmup 3 --no-cpu-baseline --video-frames 0 --no-other-configs"
timeout 300 python bench.py $B --profile-out $O/c3_prof_fold.json > $O/c3_bench_fold.json 2> $O/c3_bench_fold.err; echo "bench fold rc=$?"
VDA_LN_FOLD=0 timeout 300 python bench.py $B --profile-out $O/c3_prof_nofold.json > $O/c3_bench_nofold.json 2> $O/c3_bench_nofold.err; echo "bench nofold rc=$?"
VDA_LN_FOLD=0 VDA_GEMM_TMA_EPI=0 timeout 300 python bench.py $B --profile-out $O/c3_prof_base.json > $O/c3_bench_base.json 2> $O/c3_bench_base.err; echo "bench base rc=$?"
python - <<'PY'
import json
for t in ("fold","nofold","base"):
    try:
        l=json.loads(open(f"gpurun_out/c3_bench_{t}.json").read().strip().splitlines()[-1])
        print(t, "fps", round(l["value"],1), "p50", round(l["p50_window_latency_ms"],2), "e2e", round(l["e2e"]["value"],1), l["clocks"])
    except Exception as e: print(t, "ERR", e)
PY
tail -3 $O/c3_bench_fold.err
timeout 120 python tools/bench_attention.py prof32 > $O/c3_attn_plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:spatial_attention -s 1 -c 1 -o $O/c3_attn python tools/bench_attention.py prof32 > $O/c3_attn_ncu.log 2>&1
echo "ncu rc=$?"
