#!/bin/bash
# one-bit-short split + padded merge staging + key slot swizzle: parity first, then A/B against the previous build
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "${AB_TESTS:-key_width or top_bit or each_bin or long_row or skewed or mixed or high_compression or rmat_full or narrow_outputs}" 2>&1 | tail -5
ab() {  # variant workload steps
  lib=$PWD/spada-sim_b200/lib/libspada_b200${1:+_$1}.so
  SPADA_B200_LIB=$lib timeout 600 python bench.py --workload $2 --steps $3 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 > gpurun_out/ab_${1:-new}_$2.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_${1:-new}_$2.log").read().strip().splitlines()[-1])
    L=d["roofline"]["launch_ms"]
    agg={}
    for k,x in L.items():
        k2="long_merge" if k.startswith("long_merge") else k.split("#")[0]
        agg[k2]=agg.get(k2,0)+x
    print("variant=%-5s %-5s step %.3f ms | %s"%("${1:-new}","$2",d["ms_per_step"], "  ".join("%s %.3f"%(k,x) for k,x in agg.items() if x > 0.25)))
except Exception as e:
    print("variant=${1:-new} $2 FAILED", open("gpurun_out/ab_${1:-new}_$2.log").read()[-300:])
PY
}
[ -n "$AB_ONLY_NEW" ] || for v in base nosw; do ab "$v" rect 20; done
ab "" rect 20
[ -n "$AB_ONLY_NEW" ] || for v in base nosw; do ab "$v" rmat 6; done
ab "" rmat 6
