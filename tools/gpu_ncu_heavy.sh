#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_heavy_bits|k_heavy_accum|k_heavy_rank|k_heavy_emit" -s 12 -c 4 -f -o gpurun_out/prof_rect_heavy \
  python bench.py --workload rect --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_heavy.log 2>&1; tail -2 gpurun_out/ncu_heavy.log | cut -c1-200
