#!/bin/bash
# usage: scripts/gpu_retry.sh <timeout-seconds> '<command>'   -- retries gpurun while the pod answers busy / transient (nothing charged)
T=$1; shift
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  echo "$out" | tail -40
  if echo "$out" | grep -q "status=transient\|rc=3\|no box\|busy"; then sleep 120; continue; fi
  break
done
