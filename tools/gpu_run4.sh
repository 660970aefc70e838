#!/bin/bash
mkdir -p gpurun_out
echo "== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== pytest"; timeout 2400 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -8 | tee gpurun_out/pytest.log
for w in rect poisson er cari; do
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
