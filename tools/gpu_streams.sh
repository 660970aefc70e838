#!/bin/bash
# side-stream A/B: parity tests, then rect with 1 / 2 / 3 streams
mkdir -p gpurun_out
echo "== pytest"; timeout 1200 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -3 | tee gpurun_out/pytest.log
for n in 1 2 3; do
  SPADA_B200_STREAMS=$n timeout 600 python bench.py --workload rect --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 > gpurun_out/streams_$n.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/streams_$n.log").read().strip().splitlines()[-1])
    L=d["roofline"]["launch_ms"]
    print("streams=$n step %.3f ms clocks %s | %s"%(d["ms_per_step"], d["clocks"], "  ".join("%s %.3f"%(k,x) for k,x in L.items())))
except Exception as e:
    print("streams=$n FAILED", open("gpurun_out/streams_$n.log").read()[-600:])
PY
done
