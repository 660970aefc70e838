#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 2400 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -4 | tee gpurun_out/pytest.log
run() { # workload, mode flags, env label
  timeout 900 python bench.py --workload $1 --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 $2 2>&1 | tail -1 > gpurun_out/b.log
  python - <<PY
import json
d=json.loads(open("gpurun_out/b.log").read().strip().splitlines()[-1])
print("$1 $2 [$3] step %.3f ms  %.1f GFLOP/s"%(d["ms_per_step"], d["value"]))
print("    "+"  ".join("%s %.3f"%(k,v) for k,v in d["roofline"]["launch_ms"].items()))
PY
}
run rect --two-phase bitonic
SPADA_B200_CTA_SORT=radix run rect --two-phase radix
run rect "" fused
run poisson --two-phase -
run poisson "" fused
run er --two-phase -
run er "" fused
