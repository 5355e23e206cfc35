#!/bin/bash
# round-2 call 2: attention v4 (deadlock fixed) checks + timings, other kernel checks, in-step bench new vs r1, stage errors
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
timeout 200 python tests/kernel_checks.py "attn spatial" "attn rescale" > $O/c2_attn_checks.log 2>&1; echo "attn checks rc=$?"
tail -4 $O/c2_attn_checks.log
timeout 300 python tests/kernel_checks.py "-attn spatial" "-attn rescale" > $O/c2_other_checks.log 2>&1; echo "other checks rc=$?"
grep -E "FAIL|EXC|failing" $O/c2_other_checks.log
: > $O/c2_attn.log
for v in new r1 nosplit poly7 nopoly; do
  lib=$PWD/variants/libvda_$v.so; [[ $v == new ]] && lib=$PWD/video_depth_anything_b200/libvda.so
  echo "=== $v" >> $O/c2_attn.log
  VDA_LIB=$lib timeout 120 python tools/bench_attention.py >> $O/c2_attn.log 2>&1
done
VDA_LIB=$PWD/variants/libvda_timing.so timeout 100 python tools/bench_attention.py timing >> $O/c2_attn.log 2>&1
grep -E "===|32x1370|WG" $O/c2_attn.log
timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --video-frames 0 --no-other-configs --profile-out $O/c2_prof_new.json > $O/c2_bench_new.json 2> $O/c2_bench_new.err
echo "bench new rc=$?"
VDA_LIB=$PWD/variants/libvda_r1.so timeout 300 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --video-frames 0 --no-other-configs --profile-out $O/c2_prof_r1.json > $O/c2_bench_r1.json 2> $O/c2_bench_r1.err
echo "bench r1 rc=$?"
python - <<'PY'
import json
for t in ("new","r1"):
    try:
        l=json.loads(open(f"gpurun_out/c2_bench_{t}.json").read().strip().splitlines()[-1])
        print(t, "fps", round(l["value"],1), "e2e", round(l["e2e"]["value"],1), "attn", l["roofline"]["other_kernels"].get("attention_spatial"), l["clocks"])
    except Exception as e: print(t, "ERR", e)
PY
timeout 240 python tools/stage_errors.py vits 2 518 518 > $O/c2_stage_vits.log 2>&1; echo "stage rc=$?"
grep "===" $O/c2_stage_vits.log
