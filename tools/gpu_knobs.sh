#!/bin/bash
# A/B of library variants (build-time knobs or an older checkout) (variants built by tools/build_variant.sh): usage gpu_knobs.sh "<variant:workload:steps> ..."
mkdir -p gpurun_out
for spec in $1; do
  IFS=: read v w n <<< "$spec"; [ "$v" = new ] && v=""
  lib=$PWD/spada-sim_b200/lib/libspada_b200${v:+_$v}.so
  SPADA_B200_LIB=$lib timeout 600 python bench.py --workload $w --steps $n --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 > gpurun_out/knob_${v:-new}_$w.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/knob_${v:-new}_$w.log").read().strip().splitlines()[-1])
    agg={}
    for k,x in d["roofline"]["launch_ms"].items():
        k2="long_merge" if k.startswith("long_merge") else k.split("#")[0]
        agg[k2]=agg.get(k2,0)+x
    print("variant=%-5s %-7s step %.3f ms | %s"%("${v:-new}","$w",d["ms_per_step"], "  ".join("%s %.3f"%(k,x) for k,x in agg.items() if x > 0.25)))
except Exception as e:
    print("variant=${v:-new} $w FAILED", open("gpurun_out/knob_${v:-new}_$w.log").read()[-300:])
PY
done
