#!/bin/bash
# round-2 call 16: attention with the softmax scale folded into q and the row maximum subtracted by the S MMA (K extension)
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 90 python tests/kernel_checks.py "attn spatial 3x21x6" "attn spatial 1x128x1" "attn spatial 4x300x3" "attn spatial 3x257x1" > $O/c16_attn_small.log 2>&1
rc=$?; echo "attn small rc=$rc"; grep -E "ok|FAIL|EXC|failing" $O/c16_attn_small.log | cut -c1-150
if [[ $rc == 0 ]]; then
  timeout 300 python tests/kernel_checks.py "attn spatial" "attn rescale" > $O/c16_attn_checks.log 2>&1; rc=$?; echo "attn checks rc=$rc"
  grep -E "FAIL|EXC|failing" $O/c16_attn_checks.log | head
  grep -E "prescaled" $O/c16_attn_checks.log | cut -c1-130 | head -30
fi
: > $O/c16_attn.log
for ps in 0 1; do
  echo "=== prescaled=$ps" >> $O/c16_attn.log
  SA_PRESCALED=$ps timeout 90 python tools/bench_attention.py >> $O/c16_attn.log 2>&1
done
grep -E "===|32x1370|2x1370x16|2443" $O/c16_attn.log
if [[ $rc == 0 ]]; then
  timeout 400 python -m pytest tests/test_cmodel_gpu.py -m gpu -x -q > $O/c16_cmodel.log 2>&1; echo "cmodel rc=$?"; tail -5 $O/c16_cmodel.log | cut -c1-300
  timeout 400 python -m pytest tests/test_forward_gpu.py -m gpu -x -q -k "full_size_window or validation_mode" > $O/c16_fwd.log 2>&1; echo "fwd rc=$?"; grep -E "rel err|passed|failed|Error" $O/c16_fwd.log | head -12
  B="--steps 20 --warmup 4 --no-cpu-baseline --video-frames 0 --no-other-configs"
  VDA_ATTN_FOLD=0 timeout 200 python bench.py $B > $O/c16_bench_nofold.json 2> $O/c16_bench_nofold.err
  timeout 200 python bench.py $B --profile-out $O/c16_prof_fold.json > $O/c16_bench_fold.json 2> $O/c16_bench_fold.err
  python - <<'PY'
import json
for t in ("nofold","fold"):
    try:
        l=json.loads(open(f"gpurun_out/c16_bench_{t}.json").read().strip().splitlines()[-1])
        print(t, "fps", round(l["value"],1), "p50", round(l["p50_window_latency_ms"],2), "e2e", round(l["e2e"]["value"],1), "attn", round(l["roofline"]["other_kernels"]["attention_spatial"]["achieved"]), l["clocks"]["sm_mhz"])
    except Exception as e: print(t, "ERR", e)
PY
fi
