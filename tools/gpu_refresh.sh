#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 2400 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -3 | tee gpurun_out/pytest.log
echo "== bench default"; timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_default.log; cut -c1-300 gpurun_out/bench_default.log
SPADA_B200_STREAMS=2 timeout 900 python bench.py --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 > gpurun_out/bench_default_2streams.log; cut -c1-300 gpurun_out/bench_default_2streams.log
echo "== ncu launch list (rect)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_rect.csv \
  python bench.py --workload rect --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_ll.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_heavy_smem_numeric|k_esc_numeric_warp|k_bitonic_numeric_cta|k_copy_rows" -s 20 -c 10 -f -o gpurun_out/prof_rect_numeric \
  python bench.py --workload rect --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_a.log 2>&1
