#!/bin/bash
mkdir -p gpurun_out
run() {  # name workload steps env...
  local name=$1; shift; local w=$1; shift; local st=$1; shift
  env "$@" timeout 900 python bench.py --workload $w --steps $st --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 > gpurun_out/s5_$name.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s5_$name.log").read().strip().splitlines()[-1])
    L=d["roofline"]["launch_ms"]
    print("$name step %.3f ms %.1f GFLOP/s nnz_c %d | %s"%(d["ms_per_step"], d["value"], d["config"]["nnz_c"], "  ".join("%s %.3f"%(k,x) for k,x in L.items() if x > 0.2)))
except Exception as e:
    print("$name FAILED", open("gpurun_out/s5_$name.log").read()[-800:])
PY
}
run rect rect 20 A=1
run rmat_item rmat 3 A=1
run rmat_smem2 rmat 3 SPADA_B200_HEAVY_SMEM_COLS=2097152
