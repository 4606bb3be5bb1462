#!/usr/bin/env bash
# gpu_retry.sh <timeout-seconds> <command...>: gpurun with retries while the pod answers "transient / busy" (nothing charged)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if grep -q "status=transient\|no box\|busy" /tmp/gpurun_last.log && [ $rc -ne 0 ]; then sleep 45; continue; fi
  if grep -q "status=transient" /tmp/gpurun_last.log; then sleep 45; continue; fi
  break
done
cat /tmp/gpurun_last.log
