#!/bin/bash
# round-2 call 28: tail producer with 16 channels per unit (index math amortised; numerics unchanged) vs the previous build
cd "$GRAFT_REPO_ROOT"
for v in old new old new; do
  lib=$PWD/video_depth_anything_b200/libvda.so; [[ $v == old ]] && lib=$PWD/variants/libvda_tailold.so
  echo "=== $v"; VDA_LIB=$lib timeout 100 python tools/bench_gemm.py tail 2>&1 | grep "fused tail"
done
echo "=== checks new"; timeout 100 python tests/kernel_checks.py "tail fused" 2>&1 | grep -E "ok|FAIL|EXC|failing"
timeout 200 python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
# bit-identity of the two builds on the full-size tail problem
import ctypes, subprocess
code = '''
import os, sys, torch, zlib
sys.path.insert(0, os.getcwd())
from video_depth_anything_b200 import ops
g = torch.Generator().manual_seed(0)
for c, dt in ((128, torch.float16), (64, torch.float16), (128, torch.bfloat16)):
    n, ih, iw, oh, ow = 3, 296, 296, 518, 518
    x = torch.randn(n * ih * iw, c, generator=g).cuda().to(dt)
    w = (torch.randn(32, 9 * c, generator=g) / (9 * c) ** 0.5).cuda().to(dt)
    b = torch.randn(32, generator=g).cuda(); w2 = torch.rand(32, generator=g).cuda()
    out = torch.zeros(n, oh, ow, device="cuda")
    ops.tail_fused(x, w, b, w2, 0.1, out, n, ih, iw, oh, ow, c)
    torch.cuda.synchronize()
    print(c, dt, "%08x" % (zlib.crc32(out.cpu().numpy().tobytes()) & 0xffffffff))
'''
for lib in ("variants/libvda_tailold.so", "video_depth_anything_b200/libvda.so"):
    r = subprocess.run([sys.executable, "-c", code], env={**os.environ, "VDA_LIB": os.path.join(os.getcwd(), lib)}, capture_output=True, text=True)
    print(lib, r.stdout.strip().replace("\n", " | "), r.stderr[-300:])
PY
