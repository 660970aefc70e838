#!/bin/bash
# parity tests, then one bench line per workload given (compact print)
mkdir -p gpurun_out
echo "== pytest"; timeout 2400 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -4 | tee gpurun_out/pytest.log
for w in "$@"; do
  st=20; [ "$w" = rmat ] && st=3
  timeout 900 python bench.py --workload $w --steps $st --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 > gpurun_out/check_$w.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/check_$w.log").read().strip().splitlines()[-1])
    L=d["roofline"]["launch_ms"]
    print("$w step %.3f ms %.1f GFLOP/s | %s"%(d["ms_per_step"], d["value"], "  ".join("%s %.3f"%(k,x) for k,x in L.items() if x > 0.1)))
except Exception as e:
    print("$w FAILED", open("gpurun_out/check_$w.log").read()[-800:])
PY
done
