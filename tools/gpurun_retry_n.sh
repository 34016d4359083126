#!/bin/bash
# tools/gpurun_retry_n.sh NGPU TIMEOUT_S 'command' -- gpurun --gpus N with retries while the pod is busy.
N=$1; T=$2; shift; shift
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --gpus $N --timeout $T -- "$@" 2>&1)
  echo "$out" | tail -25
  if echo "$out" | grep -q "status=transient\|status=busy\|no box\|retry in a few minutes\|retry later"; then sleep 200; continue; fi
  break
done
