#!/bin/bash
# round-2 call 25 (2 GPUs): two-phase long-video arm with / without the bounded device buffers, same box, alternating
cd "$GRAFT_REPO_ROOT"
p=29540
for r in 1 0 1 0; do
  p=$((p+1))
  VDA_VIDEO_RINGS=$r timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $p bench.py --gpus 2 --steps 3 --warmup 3 --no-other-configs --no-e2e 2>/dev/null | python -c "
import json,sys
l=json.loads([x for x in sys.stdin if x.startswith('{')][-1]); print('rings=$r window', round(l['value'],1), 'video', round(l['video_frames_per_s'],1), round(l['video_seconds'],4), l['video_crc32'])"
done
