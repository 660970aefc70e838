#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_light -s 3 -c 1 -f -o gpurun_out/prof_er_fused \
  python bench.py --workload er --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_f1.log 2>&1; tail -1 gpurun_out/ncu_f1.log | cut -c1-100
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_fused_light -s 3 -c 1 -f -o gpurun_out/prof_poisson_fused \
  python bench.py --workload poisson --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_f2.log 2>&1; tail -1 gpurun_out/ncu_f2.log | cut -c1-100
