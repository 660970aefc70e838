#!/bin/bash
mkdir -p gpurun_out
echo "== parity subset"; timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "host_to_host or streaming or each_bin or long_row" 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_rect.err | tail -1 > gpurun_out/bench_rect.log; tail -2 gpurun_out/bench_rect.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_rect.log").read())
print("rect", d["ms_per_step"], d["value"], "e2e", d["e2e"])
PY
