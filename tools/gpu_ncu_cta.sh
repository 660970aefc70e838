#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_bitonic_numeric_cta" -s 3 -c 1 -f -o gpurun_out/prof_rect_cta1024 \
  python bench.py --workload rect --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 --two-phase > gpurun_out/ncu_c1.log 2>&1; tail -1 gpurun_out/ncu_c1.log | cut -c1-100
