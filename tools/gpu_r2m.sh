#!/bin/bash
mkdir -p gpurun_out
echo "== parity subset"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "each_bin or long_row or waves or cari or key_width or shards or streaming" 2>&1 | tail -3
for w in rect; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 2> gpurun_out/bench_$w.err | tail -1 > gpurun_out/bench_$w.log; tail -2 gpurun_out/bench_$w.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$w.log").read())
print("$w", d["ms_per_step"], d["value"], {k:round(v,3) for k,v in d["roofline"]["launch_ms"].items()})
PY
done
timeout 1200 python bench.py --workload rmat --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 2> gpurun_out/bench_rmat.err | tail -1 > gpurun_out/bench_rmat.log; tail -3 gpurun_out/bench_rmat.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_rmat.log").read())
print("rmat", d["ms_per_step"], d["value"], {k:round(v,3) for k,v in d["roofline"]["launch_ms"].items()})
PY
