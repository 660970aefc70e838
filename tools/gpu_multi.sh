#!/bin/bash
# usage: gpu_multi.sh N [workload] [waves...]   -- bench.py under torchrun on N GPUs, one run per wave count
N=${1:-2}; W=${2:-rect}; shift; shift
WAVES=${@:-2}
mkdir -p gpurun_out
for wv in $WAVES; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 --workload $W --gather-waves $wv > gpurun_out/multi_${W}_${N}_w$wv.log 2>&1
  python - <<PY
import json
try:
    l=[x for x in open("gpurun_out/multi_${W}_${N}_w$wv.log") if x.startswith("{")][-1]
    d=json.loads(l)
    print("N=$N waves=$wv: %.3f ms/step %.1f GFLOP/s compute-only %.3f ms nnz_c %d"%(d["ms_per_step"], d["value"], d["compute_only"]["ms_per_step"], d["config"]["nnz_c"]))
except Exception as e:
    print("N=$N waves=$wv FAILED", open("gpurun_out/multi_${W}_${N}_w$wv.log").read()[-1500:])
PY
done
