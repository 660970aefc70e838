#!/bin/bash
# first GPU pass: smoke, parity tests, three bench lines, launch list.  Every step has its own timeout.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -20 | tee gpurun_out/smoke.log
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -40 | tee gpurun_out/pytest.log
for w in rect poisson er; do
  echo "== bench $w"; timeout 900 python bench.py --workload $w --steps 10 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench_$w.log
done
echo "== ncu launches"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_rect.csv \
  python bench.py --workload rect --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
