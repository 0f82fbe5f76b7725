#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout_s> [--gpus N] -- '<command>'   (retries while the pool answers "transient")
t=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$t" "$@" 2>&1)
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  break
done
