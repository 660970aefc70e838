#!/bin/bash
# usage: gpu_multi.sh N [workload]
N=${1:-2}; W=${2:-rect}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 10 --warmup 3 --workload $W > gpurun_out/multi_${W}_$N.log 2>&1
tail -3 gpurun_out/multi_${W}_$N.log | cut -c1-1800
