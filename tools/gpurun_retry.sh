#!/bin/bash
# gpurun with retries on "transient" (nothing charged): tools/gpurun_retry.sh <timeout_s> <command...>
t=$1; shift
for i in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun --timeout $t -- "$@" 2>&1)
  echo "$out" | tail -60
  echo "$out" | grep -q "status=transient" || exit 0
  echo "[retry $i] transient, sleeping 60 s"
  sleep 60
done
