#!/bin/bash
# usage: tools/gpu_retry.sh <log> <timeout> <command...>   -- retries gpurun while the pod answers "busy" (nothing charged)
LOG=$1; TO=$2; shift 2
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $TO -- "$@" > $LOG 2>&1
  if grep -q "status=transient" $LOG; then sleep 90; continue; fi
  break
done
