#!/bin/bash
# usage: gpu_ncu_one.sh <kernel regex> <workload> <outname> [extra bench flags]
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$1" -s 3 -c 1 -f -o gpurun_out/$3 \
  python bench.py --workload $2 --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 $4 > gpurun_out/ncu_$3.log 2>&1; tail -1 gpurun_out/ncu_$3.log | cut -c1-100
