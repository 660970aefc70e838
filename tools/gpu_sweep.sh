#!/bin/bash
mkdir -p gpurun_out
for g in 1 2 3 4 6 8; do
  SPADA_B200_HEAVY_CTAS_PER_SM=$g timeout 600 python bench.py --workload rect --steps 5 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 > gpurun_out/sweep_$g.log
  python - <<PY
import json
d=json.loads(open("gpurun_out/sweep_$g.log").read().strip().splitlines()[-1])
L=d["roofline"]["launch_ms"]
print("ctas/sm=$g step %.2f  bits %.3f rank %.3f emit %.3f accum %.3f"%(d["ms_per_step"],L["sym_heavy_bits"],L["sym_heavy_rank"],L["num_heavy_emit"],L["num_heavy_accum"]))
PY
done
