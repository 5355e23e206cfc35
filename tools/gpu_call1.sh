#!/bin/bash
# round-2 call 1: new parity checks on the round-1 library, full kernel checks + attention timings of the new kernel
cd "$GRAFT_REPO_ROOT"
O=gpurun_out
nvidia-smi -L > $O/c1_gpu.txt
VDA_LIB=$PWD/variants/libvda_r1.so timeout 400 python tests/kernel_checks.py "attn rescale" "attn spatial" "groupnorm" > $O/c1_r1_checks.log 2>&1
echo "r1 checks rc=$?"
timeout 900 python tests/kernel_checks.py > $O/c1_new_checks.log 2>&1
echo "new checks rc=$?"
tail -3 $O/c1_new_checks.log
for v in r1 new nosplit poly7; do
  lib=$PWD/variants/libvda_$v.so; [[ $v == new ]] && lib=$PWD/video_depth_anything_b200/libvda.so
  echo "=== $v" >> $O/c1_attn.log
  VDA_LIB=$lib timeout 300 python tools/bench_attention.py >> $O/c1_attn.log 2>&1
done
VDA_LIB=$PWD/variants/libvda_timing.so timeout 200 python tools/bench_attention.py timing >> $O/c1_attn.log 2>&1
cat $O/c1_attn.log | grep -E "===|32x1370|2x1370|WG|err"
timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --video-frames 0 --profile-out $O/c1_prof_new.json > $O/c1_bench_new.json 2> $O/c1_bench_new.err
echo "bench new rc=$?"
VDA_LIB=$PWD/variants/libvda_r1.so timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --video-frames 0 --profile-out $O/c1_prof_r1.json > $O/c1_bench_r1.json 2> $O/c1_bench_r1.err
echo "bench r1 rc=$?"
python - <<'PY'
import json
for t in ("new","r1"):
    try:
        l=json.loads(open(f"gpurun_out/c1_bench_{t}.json").read().strip().splitlines()[-1])
        print(t, "fps", round(l["value"],1), "e2e", round(l["e2e"]["value"],1), "attn", l["roofline"]["other_kernels"].get("attention_spatial"), l["clocks"])
    except Exception as e: print(t, "ERR", e)
PY
