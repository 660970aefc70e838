#!/bin/bash
# usage: gpu_ab.sh <variant names...>   -- A/B of library variants built by tools/build_variant.sh on rect and er
mkdir -p gpurun_out
for v in "" "$@"; do
  lib=$PWD/spada-sim_b200/lib/libspada_b200${v:+_$v}.so
  for w in ${AB_WORKLOADS:-rect er}; do
    SPADA_B200_LIB=$lib timeout 600 python bench.py --workload $w --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 > gpurun_out/ab.log
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab.log").read().strip().splitlines()[-1])
    L=d["roofline"]["launch_ms"]
    print("variant=%-6s %-5s step %.3f ms | %s"%("${v:-base}","$w",d["ms_per_step"], "  ".join("%s %.3f"%(k,x) for k,x in L.items() if x > 0.3)))
except Exception as e:
    print("variant=${v:-base} $w FAILED", open("gpurun_out/ab.log").read()[-300:])
PY
  done
done
