#!/bin/bash
# usage: gpurun_retry.sh <timeout_s> [--gpus N] -- '<command>'   (retries while the pod answers busy/transient; exit code 3)
T=$1; shift
for i in $(seq 1 40); do
  OUT=$(/usr/local/graft/bin/gpurun --timeout $T "$@" 2>&1)
  if echo "$OUT" | grep -q "status=transient\|nothing was charged"; then sleep 60; continue; fi
  echo "$OUT"; exit 0
done
echo "$OUT"; echo "gave up after 40 tries"
