#!/bin/bash
# usage: tools/gpurun_retry.sh <timeout-seconds> '<command>'   -- retries while the pod answers "busy / transient" (nothing charged)
T=$1; shift
for attempt in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" "$@" -- "$CMD" 2>&1)
  echo "$out" | tail -60
  if echo "$out" | grep -q "status=transient\|status=busy\|no box\|retry in a few minutes"; then
    echo "[retry $attempt] waiting 120 s"; sleep 120; continue
  fi
  break
done
