#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 -k "huge or cari or skew or long_rows or each_bin or waves or rmat" 2>&1 | tail -3
for w in rmat cari; do
  st=20; [ "$w" = rmat ] && st=3
  timeout 600 python bench.py --workload $w --steps $st --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 > gpurun_out/check_$w.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/check_$w.log").read().strip().splitlines()[-1])
    L=d["roofline"]["launch_ms"]
    print("$w step %.3f ms %.1f GFLOP/s nnz_c %d | %s"%(d["ms_per_step"], d["value"], d["config"]["nnz_c"], "  ".join("%s %.3f"%(k,x) for k,x in L.items() if x > 0.05)))
except Exception as e:
    print("$w FAILED", open("gpurun_out/check_$w.log").read()[-600:])
PY
done
