#!/bin/bash
# tools/gpurun_retry.sh TIMEOUT 'command' -- development helper: retries a gpurun call while the pod answers "busy" (exit 3)
t=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
