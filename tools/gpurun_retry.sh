#!/bin/bash
# retry a gpurun call while the pod answers "busy"; usage: [GPUS=N] gpurun_retry.sh <timeout> <cmd...>
T=$1; shift
G=${GPUS:+--gpus $GPUS}
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun $G --timeout "$T" -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|rc=3\|busy"; then sleep 45; continue; fi
  echo "$out"; exit 0
done
echo "gave up"; exit 3
