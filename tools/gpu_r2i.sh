#!/bin/bash
mkdir -p gpurun_out
for v in w8_c2048 w4_c2048 w4_c4096 w8_c4096; do
  lib=$PWD/spada-sim_b200/lib/libspada_b200_$v.so; [ $v = w8_c2048 ] && lib=$PWD/spada-sim_b200/lib/libspada_b200.so
  SPADA_B200_LIB=$lib SPADA_B200_TILE_PASS=1 timeout 900 python bench.py --workload rect --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2> gpurun_out/tp_$v.err | tail -1 > gpurun_out/tp_$v.log; tail -2 gpurun_out/tp_$v.err; python - <<PY
import json
d=json.loads(open("gpurun_out/tp_$v.log").read())
print("rect $v", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["roofline"]["launch_ms"].items() if "tile" in k or "copy" in k or "flop" in k})
PY
done
