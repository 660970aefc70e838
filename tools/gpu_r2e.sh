#!/bin/bash
mkdir -p gpurun_out
echo "== pytest all"; timeout 2400 python -m pytest tests -m gpu -q --timeout 900 -x 2>&1 | tail -8 | tee gpurun_out/pytest.log
echo "== bench default"; timeout 900 python bench.py 2> gpurun_out/bench_default.err | tail -1 > gpurun_out/bench_default.log; tail -3 gpurun_out/bench_default.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_default.log").read())
print("rect", d["ms_per_step"], d["value"], "e2e", d["e2e"], "cpu", d["cpu_baseline"])
print({k:round(v,3) for k,v in d["roofline"]["launch_ms"].items()})
print(d["roofline"]["kernel"], d["roofline"]["frac"], d["roofline"]["whole_path"])
PY
for w in poisson er; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_$w.err | tail -1 > gpurun_out/bench_$w.log; tail -2 gpurun_out/bench_$w.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$w.log").read())
print("$w", d["ms_per_step"], d["value"], "e2e", d["e2e"] and d["e2e"]["ms_per_step"], {k:round(v,3) for k,v in d["roofline"]["launch_ms"].items()})
PY
done
