#!/bin/bash
mkdir -p gpurun_out
echo "== parity"; timeout 2400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 900 -k "not full_size" 2>&1 | tail -4
for tp in 1 0; do
  SPADA_B200_TILE_PASS=$tp timeout 900 python bench.py --workload rect --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 2> gpurun_out/bench_rect_tp$tp.err | tail -1 > gpurun_out/bench_rect_tp$tp.log; tail -2 gpurun_out/bench_rect_tp$tp.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_rect_tp$tp.log").read())
print("rect tile_pass=$tp", d["ms_per_step"], d["value"], {k:round(v,3) for k,v in d["roofline"]["launch_ms"].items()})
PY
done
SPADA_B200_TILE_PASS=1 timeout 1200 python bench.py --workload rmat --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 2> gpurun_out/bench_rmat.err | tail -1 > gpurun_out/bench_rmat.log; tail -3 gpurun_out/bench_rmat.err; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_rmat.log").read())
print("rmat", d["ms_per_step"], d["value"], {k:round(v,3) for k,v in d["roofline"]["launch_ms"].items()})
PY
echo "== full size parity"; timeout 2400 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 1200 -k "full_size" 2>&1 | tail -4
