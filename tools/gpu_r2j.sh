#!/bin/bash
# multi-GPU: check + bench at N GPUs (peer gather), logs kept
mkdir -p gpurun_out
N=${1:-4}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tools/multi_check.py > gpurun_out/multi_check_$N.log 2>&1; echo "multi_check rc=$?"
grep -E "identical|Error|error|Traceback|assert" gpurun_out/multi_check_$N.log | head -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 20 --warmup 5 2> gpurun_out/bench_n${N}_peer.err | tail -1 > gpurun_out/bench_n${N}_peer.log
grep -E "Error|Traceback" gpurun_out/bench_n${N}_peer.err | head
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n${N}_peer.log").read())
    print("N=$N peer", round(d["ms_per_step"],3), "ms", round(d["value"],1), "GFLOP/s", d["multi_gpu"], d["setup"], "e2e", d["e2e"])
    print({k:round(v,3) for k,v in d["roofline"]["launch_ms"].items()})
except Exception as e: print("parse failed", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29524 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-600
