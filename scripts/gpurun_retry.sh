#!/bin/bash
# usage: scripts/gpurun_retry.sh <timeout_s> '<command>'   -- retries while the pod has no free GPU slot (nothing is charged for those)
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
