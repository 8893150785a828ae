#!/bin/bash
# tools/gpurun_retry_n.sh GPUS TIMEOUT 'command' -- like gpurun_retry.sh on N GPUs of one box
g=$1; t=$2; shift 2
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --gpus "$g" --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
