#!/bin/bash
# usage: scratch/gpurun_retry.sh <timeout_s> '<command>'  -- retries while the pod answers "busy / draining" (nothing charged)
T=$1; shift
for attempt in $(seq 1 20); do
  out=$(gpurun --timeout $T "$@" 2>&1)
  if echo "$out" | grep -q "status=transient\|rc=3\b"; then
    echo "[retry $attempt] busy"; sleep 90; continue
  fi
  echo "$out"; exit 0
done
echo "gave up"; exit 3
