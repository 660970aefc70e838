#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fused_tiny4" -s 3 -c 1 -f -o gpurun_out/prof_poisson_tiny4 \
  python bench.py --workload poisson --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_tiny4.log 2>&1
tail -3 gpurun_out/ncu_tiny4.log
