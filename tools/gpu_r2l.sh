#!/bin/bash
mkdir -p gpurun_out
echo "== parity"; timeout 2400 python -m pytest tests -m gpu -x -q --timeout 900 -k "not full_size" 2>&1 | tail -4
echo "== sanitizer"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -c MISMATCH gpurun_out/sanitize_memcheck.log; tail -2 gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/sanitize_racecheck.log
for w in rect; do
  timeout 900 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_$w.err | tail -1 > gpurun_out/bench_$w.log; tail -2 gpurun_out/bench_$w.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$w.log").read())
print("$w", d["ms_per_step"], d["value"], "e2e", d["e2e"] and d["e2e"]["ms_per_step"], {k:round(v,3) for k,v in d["roofline"]["launch_ms"].items()})
PY
done
timeout 1200 python bench.py --workload rmat --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 2> gpurun_out/bench_rmat.err | tail -1 > gpurun_out/bench_rmat.log; tail -3 gpurun_out/bench_rmat.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_rmat.log").read())
print("rmat", d["ms_per_step"], d["value"], {k:round(v,3) for k,v in d["roofline"]["launch_ms"].items()})
PY
echo "== full size parity"; timeout 2400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 1200 -k "full_size" 2>&1 | tail -3
