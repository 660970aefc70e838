#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 2400 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -4 | tee gpurun_out/pytest.log
for w in ${WORKLOADS:-rect poisson er}; do
  for mode in "" "--two-phase"; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 $mode 2>&1 | tail -1 > gpurun_out/bench_$w$mode.log
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$w$mode.log").read().strip().splitlines()[-1])
print("$w $mode step %.3f ms  %.1f GFLOP/s  launches %d"%(d["ms_per_step"], d["value"], d["gpu_launches"]))
print("    "+"  ".join("%s %.3f"%(k,v) for k,v in d["roofline"]["launch_ms"].items()))
PY
  done
done
