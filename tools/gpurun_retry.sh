#!/bin/bash
# usage: gpurun_retry.sh <timeout> <cmd...>  -- retries while the pod has no free slot (exit code 3)
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $T "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then cat /tmp/gpurun_last.log | tail -80; exit $rc; fi
  sleep 90
done
echo "gave up"; exit 3
