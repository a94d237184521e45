#!/bin/bash
# tools/gpu.sh TIMEOUT 'command' — gpurun with retries while the pod answers busy / no slot (exit code 3 or "transient")
T=$1; shift
for i in $(seq 1 40); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$T" -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient" || [ $rc -eq 3 ]; then sleep 45; continue; fi
  echo "$out"; exit $rc
done
echo "$out"; exit 3
