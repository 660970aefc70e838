#!/bin/bash
# multi-GPU validation: peer-store gather vs one GPU, group API, bench at N = 2
mkdir -p gpurun_out
N=${1:-2}
echo "== group API (single process, $N devices)"
timeout 600 python tools/multi_check.py --group $N 2>&1 | tail -4
timeout 900 python tools/multi_check.py --group $N --workload rect --scale 0.125 2>&1 | tail -3
echo "== torchrun x$N: peer gather vs one GPU"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_check.py 2>&1 | grep -v "^W\|^\*" | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/multi_check.py --workload rect 2>&1 | grep -v "^W\|^\*" | tail -8
echo "== bench N=$N peer / nccl"
for g in peer nccl; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 --gather $g 2> gpurun_out/bench_n${N}_$g.err | tail -1 > gpurun_out/bench_n${N}_$g.log
tail -3 gpurun_out/bench_n${N}_$g.err | cut -c1-300
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n${N}_$g.log").read())
    print("N=$N $g", round(d["ms_per_step"],3), "ms", round(d["value"],1), "GFLOP/s", d["multi_gpu"], d["setup"])
    print({k:round(v,3) for k,v in d["roofline"]["launch_ms"].items()})
except Exception as e: print("parse failed", e)
PY
done
