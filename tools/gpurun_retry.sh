#!/bin/bash
# retries a gpurun call while the pool answers "busy" (exit code 3 / transient): tools/gpurun_retry.sh <timeout> <command>
T=$1; shift
for attempt in 1 2 3 4 5 6 7 8 9 10 11 12; do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit 0
done
echo "$out"; exit 3
