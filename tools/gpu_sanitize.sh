#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_cases.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool"; grep -c " ok" gpurun_out/sanitize_$tool.log; grep -E "MISMATCH|ERROR SUMMARY|Race reported|Invalid|hazard" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -10
done
echo "== bench default"; timeout 900 python bench.py --no-cpu-baseline 2>&1 | tail -1 > gpurun_out/bench_default.log; cut -c1-200 gpurun_out/bench_default.log
