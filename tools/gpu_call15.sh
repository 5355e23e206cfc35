#!/bin/bash
# round-2 call 15: S-prefetch attention variants (guarded by a 60 s leash), handle-level C API test, A/B bench
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
L=$PWD/video_depth_anything_b200/libvda.so
timeout 90 python tests/kernel_checks.py "attn spatial 3x21x6" "attn spatial 1x128x1" "attn spatial 4x300x3" "attn spatial 3x257x1" > $O/c15_attn_small.log 2>&1
rc=$?; echo "attn small rc=$rc"; tail -5 $O/c15_attn_small.log
if [[ $rc == 0 ]]; then
  timeout 200 python tests/kernel_checks.py "attn spatial" "attn rescale" > $O/c15_attn_checks.log 2>&1; rc=$?; echo "attn checks (pf1) rc=$rc"
  grep -E "FAIL|EXC|failing" $O/c15_attn_checks.log | head
  VDA_LIB=$PWD/variants/libvda_pf0.so timeout 200 python tests/kernel_checks.py "attn spatial" "attn rescale" > $O/c15_attn_checks_pf0.log 2>&1; echo "attn checks (pf0) rc=$?"
  grep -E "FAIL|EXC|failing" $O/c15_attn_checks_pf0.log | head
fi
: > $O/c15_attn.log
for v in nopf pf1 pf0 pffull; do
  lib=$PWD/variants/libvda_$v.so; [[ $v == pf1 ]] && lib=$L
  echo "=== $v" >> $O/c15_attn.log
  VDA_LIB=$lib timeout 90 python tools/bench_attention.py >> $O/c15_attn.log 2>&1
done
echo "=== timing" >> $O/c15_attn.log
VDA_LIB=$PWD/variants/libvda_pftiming.so timeout 60 python tools/bench_attention.py timing >> $O/c15_attn.log 2>&1
grep -E "===|32x1370|2x1370x16|2443|WG" $O/c15_attn.log
timeout 400 python -m pytest tests/test_cmodel_gpu.py -m gpu -x -q > $O/c15_cmodel.log 2>&1; echo "cmodel rc=$?"; tail -15 $O/c15_cmodel.log | cut -c1-300
if [[ $rc == 0 ]]; then
  B="--steps 20 --warmup 4 --no-cpu-baseline --video-frames 0 --no-other-configs"
  VDA_LIB=$PWD/variants/libvda_nopf.so timeout 200 python bench.py $B > $O/c15_bench_nopf.json 2> $O/c15_bench_nopf.err
  timeout 200 python bench.py $B --profile-out $O/c15_prof_pf1.json > $O/c15_bench_pf1.json 2> $O/c15_bench_pf1.err
  python - <<'PY'
import json
for t in ("nopf","pf1"):
    try:
        l=json.loads(open(f"gpurun_out/c15_bench_{t}.json").read().strip().splitlines()[-1])
        print(t, "fps", round(l["value"],1), "p50", round(l["p50_window_latency_ms"],2), "e2e", round(l["e2e"]["value"],1), "attn", round(l["roofline"]["other_kernels"]["attention_spatial"]["achieved"]), l["clocks"]["sm_mhz"])
    except Exception as e: print(t, "ERR", e)
PY
fi
