#!/bin/bash
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -8 | tee gpurun_out/pytest.log
run() {  # name workload env...
  local name=$1; shift; local w=$1; shift
  env "$@" timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 > gpurun_out/s4_$name.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s4_$name.log").read().strip().splitlines()[-1])
    L=d["roofline"]["launch_ms"]
    print("$name step %.3f ms | %s"%(d["ms_per_step"], "  ".join("%s %.3f"%(k,x) for k,x in L.items() if x > 0.2)))
except Exception as e:
    print("$name FAILED", open("gpurun_out/s4_$name.log").read()[-800:])
PY
}
NS=$PWD/spada-sim_b200/lib/libspada_b200_nostream.so
for w in rect er; do
  run ${w}_pad0 $w SPADA_B200_FIBER_PAD=0
  run ${w}_pad1 $w SPADA_B200_FIBER_PAD=1
  run ${w}_pad16 $w SPADA_B200_FIBER_PAD=16
  run ${w}_pad16_nostream $w SPADA_B200_FIBER_PAD=16 SPADA_B200_LIB=$NS
done
run poisson_auto poisson A=1
run poisson_pad16 poisson SPADA_B200_FIBER_PAD=16
