#!/bin/bash
mkdir -p gpurun_out
echo "== long rows"; timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 600 -k "each_bin or long_row or waves or cari or key_width or skewed or mixed" 2>&1 | tail -6
echo "== sanitizer"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -c MISMATCH gpurun_out/sanitize_memcheck.log; tail -2 gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_cases.py > gpurun_out/sanitize_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/sanitize_racecheck.log
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
echo "== ncu launch list rmat"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_rmat.csv \
  python bench.py --workload rmat --steps 1 --warmup 3 --no-cpu-baseline --e2e-steps 0 > gpurun_out/ncu_ll_rmat.log 2>&1
python - <<PY
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_rmat.csv")) if len(r)>5 and r[0].isdigit()]
agg=collections.defaultdict(lambda:[0,0.0])
for r in rows:
    k=r[4].split("(")[0][:60]; v=float(r[-1].replace(",",""))
    agg[k][0]+=1; agg[k][1]+=v
for k,(n,v) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:14]: print(f"{v/1e6:10.3f} ms {n:5d} {k}")
PY
