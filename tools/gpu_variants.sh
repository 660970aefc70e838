#!/bin/bash
mkdir -p gpurun_out
for v in "" lw2 lw8 u3; do
  lib=$PWD/spada-sim_b200/lib/libspada_b200${v:+_$v}.so
  for w in rect er poisson; do
    SPADA_B200_LIB=$lib timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 > gpurun_out/v.log
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/v.log").read().strip().splitlines()[-1])
    L=d["roofline"]["launch_ms"]
    print("variant=%-4s %-5s step %.3f ms | %s"%("${v:-base}","$w",d["ms_per_step"], "  ".join("%s %.3f"%(k,x) for k,x in L.items() if k.startswith(("sort_pass","fused")))))
except Exception as e:
    print("variant=${v:-base} $w FAILED", open("gpurun_out/v.log").read()[-300:])
PY
  done
done
