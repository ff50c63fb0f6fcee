#!/bin/bash
# gpu_retry.sh <timeout_s> '<command>' : retries gpurun while the pod answers "busy" (exit 3), at most 40 times.
T="$1"; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
