#!/bin/bash
# retry a gpurun call while the pod answers "busy" (exit 3 / transient); usage: gpurun_retry.sh <timeout> <cmd...>
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"; exit 0
done
echo "gave up"; exit 3
