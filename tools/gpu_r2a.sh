#!/bin/bash
# first GPU pass of round 2: long-row path parity + timing
mkdir -p gpurun_out
echo "== long rows"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "each_bin or long_row or waves or cari or key_width" 2>&1 | tail -15
echo "== sanitizer"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -c MISMATCH gpurun_out/sanitize_memcheck.log; tail -3 gpurun_out/sanitize_memcheck.log
echo "== pytest all"; timeout 2400 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8 | tee gpurun_out/pytest.log
for w in rect cari; do
  timeout 900 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 > gpurun_out/bench_$w.log; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$w.log").read())
print("$w", d["ms_per_step"], d["value"], {k:round(v,3) for k,v in d["roofline"]["launch_ms"].items()})
PY
done
timeout 1200 python bench.py --workload rmat --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 > gpurun_out/bench_rmat.log; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_rmat.log").read())
print("rmat", d["ms_per_step"], d["value"], {k:round(v,3) for k,v in d["roofline"]["launch_ms"].items()})
PY
