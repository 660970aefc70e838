#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -5 | tee gpurun_out/pytest.log
for w in rect cari; do
  echo "== bench $w"; timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 > gpurun_out/bench_$w.log
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$w.log").read().strip().splitlines()[-1])
print("$w", d["ms_per_step"], d["value"])
for k,v in d["roofline"]["launch_ms"].items(): print("   %-18s %.3f"%(k,v))
PY
done
