#!/bin/bash
# tools/gpurun_retry.sh TIMEOUT_S 'command' -- gpurun, retried while the pod answers "busy / transient" (nothing charged).
T=$1; shift
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout $T -- "$@" 2>&1)
  echo "$out" | tail -25
  if echo "$out" | grep -q "status=transient\|status=busy\|no box\|retry in a few minutes"; then sleep 150; continue; fi
  break
done
