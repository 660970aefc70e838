#!/bin/bash
# bucket kernel A/B: parity tests, then rect with the bucketed bins off / 6..8 / 6..9
mkdir -p gpurun_out
echo "== pytest"; timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -15 | tee gpurun_out/pytest.log
for bb in 0 0x1c0 0x3c0; do
  SPADA_B200_BUCKET_BINS=$bb timeout 600 python bench.py --workload rect --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 > gpurun_out/bucket_$bb.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bucket_$bb.log").read().strip().splitlines()[-1])
    L=d["roofline"]["launch_ms"]
    print("bucket_bins=$bb step %.3f ms | %s"%(d["ms_per_step"], "  ".join("%s %.3f"%(k,x) for k,x in L.items())))
except Exception as e:
    print("bucket_bins=$bb FAILED", open("gpurun_out/bucket_$bb.log").read()[-800:])
PY
done
for tv in warp quad; do
  SPADA_B200_TINY=$tv timeout 600 python bench.py --workload poisson --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 > gpurun_out/tiny_$tv.log
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/tiny_$tv.log").read().strip().splitlines()[-1])
    L=d["roofline"]["launch_ms"]
    print("tiny=$tv step %.3f ms | %s"%(d["ms_per_step"], "  ".join("%s %.3f"%(k,x) for k,x in L.items())))
except Exception as e:
    print("tiny=$tv FAILED", open("gpurun_out/tiny_$tv.log").read()[-800:])
PY
done
