#!/bin/bash
# scaling of the default workload at N GPUs of one box (peer-store gather; NCCL gather as the baseline), hardware check
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/multi_check.py > gpurun_out/multi_check_$N.log 2>&1; echo "multi_check rc=$?"
grep -cE "identical to one GPU: True" gpurun_out/multi_check_$N.log; grep -E "False|Error|Traceback" gpurun_out/multi_check_$N.log | head -5
timeout 600 python tools/multi_check.py --group $N > gpurun_out/group_check_$N.log 2>&1; echo "group rc=$?"; tail -2 gpurun_out/group_check_$N.log
for g in peer nccl; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 20 --warmup 5 --gather $g 2> gpurun_out/bench_n${N}_$g.err | tail -1 > gpurun_out/bench_n${N}_$g.log
grep -E "Error|Traceback" gpurun_out/bench_n${N}_$g.err | head -3
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n${N}_$g.log").read())
    print("N=$N $g", round(d["ms_per_step"],3), "ms", round(d["value"],1), "GFLOP/s", d["multi_gpu"], "e2e", d["e2e"] and d["e2e"]["ms_per_step"])
    print({k:round(v,3) for k,v in d["roofline"]["launch_ms"].items()})
except Exception as e: print("parse failed", e)
PY
done
