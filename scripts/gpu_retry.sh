#!/bin/bash
# usage: scripts/gpu_retry.sh <log> <gpurun args...>   -- retries while the pod answers "busy" (nothing is charged for those)
LOG=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$LOG" 2>&1
  if ! grep -q "status=transient" "$LOG"; then exit 0; fi
  sleep 45
done
exit 3
