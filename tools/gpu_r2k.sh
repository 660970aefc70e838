#!/bin/bash
mkdir -p gpurun_out
echo "== parity"; timeout 2400 python -m pytest tests -m gpu -x -q --timeout 900 -k "not full_size" 2>&1 | tail -4
for w in er rect poisson; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 2> gpurun_out/bench_$w.err | tail -1 > gpurun_out/bench_$w.log; tail -2 gpurun_out/bench_$w.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$w.log").read())
print("$w", d["ms_per_step"], d["value"], {k:round(v,3) for k,v in d["roofline"]["launch_ms"].items()})
PY
done
