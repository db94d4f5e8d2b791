#!/bin/bash
# gpurun with retries while the pod answers "busy" (exit code 3 / status=transient); usage: tools/gpurun_retry.sh <timeout_s> '<command>'
t=$1; shift
for i in $(seq 1 20); do
  out=$(gpurun --timeout "$t" -- "$@" 2>&1); rc=$?
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  echo "$out"; exit $rc
done
echo "gpurun_retry: still busy after 20 tries"; exit 3
