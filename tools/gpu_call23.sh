#!/bin/bash
# round-2 call 23: long-video arm with / without the bounded device buffers, same box, alternating
cd "$GRAFT_REPO_ROOT"
for r in 1 0 1 0; do
  VDA_VIDEO_RINGS=$r timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-other-configs --no-e2e 2>/dev/null | python -c "
import json,sys
l=json.loads([x for x in sys.stdin if x.startswith('{')][-1]); print('rings=$r video', round(l['video_frames_per_s'],1), round(l['video_seconds'],4), l['video_crc32'])"
done
