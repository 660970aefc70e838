#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -6 | tee gpurun_out/pytest.log
run() {  # name workload env...
  local name=$1; shift; local w=$1; shift
  env "$@" timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 > gpurun_out/s6_$name.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s6_$name.log").read().strip().splitlines()[-1])
    L=d["roofline"]["launch_ms"]
    print("$name step %.3f ms %.1f GFLOP/s | %s"%(d["ms_per_step"], d["value"], "  ".join("%s %.3f"%(k,x) for k,x in L.items() if x > 0.05)))
except Exception as e:
    print("$name FAILED", open("gpurun_out/s6_$name.log").read()[-800:])
PY
}
run poisson poisson A=1
run er er A=1
run rect rect A=1

