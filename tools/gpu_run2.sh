#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -15 | tee gpurun_out/pytest.log
for w in rect poisson er cari; do
  echo "== bench $w"; timeout 900 python bench.py --workload $w --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_$w.log
done
echo "== ncu full captures"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_esc_numeric_warp -s 3 -c 1 -f -o gpurun_out/prof_er_numeric256 \
  python bench.py --workload er --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu1.log 2>&1; tail -2 gpurun_out/ncu1.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_esc_symbolic_warp -s 3 -c 1 -f -o gpurun_out/prof_er_symbolic256 \
  python bench.py --workload er --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_esc_numeric_warp -s 3 -c 1 -f -o gpurun_out/prof_poisson_numeric32 \
  python bench.py --workload poisson --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu3.log 2>&1; tail -2 gpurun_out/ncu3.log | cut -c1-300
ls -la gpurun_out/
