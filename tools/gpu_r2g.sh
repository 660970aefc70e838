#!/bin/bash
mkdir -p gpurun_out
echo "== parity (new tests)"; timeout 2400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 900 -k "full_size or signed_zero or streaming or scaled or window" 2>&1 | tail -4
timeout 600 python tools/e2e_probe.py rect 2>&1 | tail -8
for w in poisson er; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 2> gpurun_out/bench_$w.err | tail -1 > gpurun_out/bench_$w.log; tail -2 gpurun_out/bench_$w.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$w.log").read())
print("$w", d["ms_per_step"], d["value"], {k:round(v,3) for k,v in d["roofline"]["launch_ms"].items()})
PY
done
echo "== ncu full: long-row kernels on rmat (one wave)"
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"k_long_chunk_sort|k_long_merge|k_long_reduce" -s 30 -c 6 -f -o gpurun_out/prof_rmat_long \
  python bench.py --workload rmat --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_rmat_long.log 2>&1
ls -la gpurun_out/prof_rmat_long.ncu-rep
