#!/bin/bash
# round-end measurement set: tests, bench lines (default + other configs), ncu launch list, full captures
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > gpurun_out/gpu.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== pytest"; timeout 2400 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -3 | tee gpurun_out/pytest.log
echo "== bench default"; timeout 900 python bench.py 2>&1 | tail -1 > gpurun_out/bench_default.log; cut -c1-400 gpurun_out/bench_default.log
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 > gpurun_out/bench_reference.log; cut -c1-300 gpurun_out/bench_reference.log
for w in poisson er cari; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 3 2>&1 | tail -1 > gpurun_out/bench_$w.log; cut -c1-200 gpurun_out/bench_$w.log
done
timeout 900 python bench.py --workload rmat --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 > gpurun_out/bench_rmat.log; cut -c1-200 gpurun_out/bench_rmat.log
echo "== ncu launch list (rect)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_rect.csv \
  python bench.py --workload rect --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_ll.log 2>&1
echo "== ncu full: rect sort / heavy / copy kernels, ER fused, Poisson tiny"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_heavy_smem_numeric|k_esc_numeric_warp|k_bitonic_numeric_cta|k_copy_rows" -s 20 -c 10 -f -o gpurun_out/prof_rect_numeric \
  python bench.py --workload rect --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fused_light" -s 3 -c 1 -f -o gpurun_out/prof_er_fused \
  python bench.py --workload er --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_fused_tiny4" -s 3 -c 1 -f -o gpurun_out/prof_poisson_tiny \
  python bench.py --workload poisson --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out | head -40
