#!/bin/bash
mkdir -p gpurun_out
echo "== parity: fused kernels"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "scaled or poisson or window or each_bin or shards or empty or accelerator" 2>&1 | tail -4
timeout 600 python tools/e2e_probe.py rect 2>&1 | tail -8
for w in poisson er rect; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 2> gpurun_out/bench_$w.err | tail -1 > gpurun_out/bench_$w.log; tail -2 gpurun_out/bench_$w.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$w.log").read())
print("$w", d["ms_per_step"], d["value"], {k:round(v,3) for k,v in d["roofline"]["launch_ms"].items()})
PY
done
