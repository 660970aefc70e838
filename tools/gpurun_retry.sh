#!/bin/bash
# usage: gpurun_retry.sh [-g N] <timeout> <cmd...>  -- retries while the pod has no free slot (exit code 3)
G=""
if [ "$1" = "-g" ]; then G="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun $G --timeout $T "$@" > /tmp/gpurun_last.log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then cat /tmp/gpurun_last.log | tail -120; exit $rc; fi
  sleep 75
done
echo "gave up"; exit 3
