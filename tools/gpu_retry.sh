#!/bin/bash
# tools/gpu_retry.sh TIMEOUT 'command' -- gpurun with retries while the pod is busy (rc 3 / transient)
T=$1; shift
for i in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 60; continue; fi
  echo "$out"; exit $rc
done
echo "gpu_retry: still busy after 20 tries"; exit 3
