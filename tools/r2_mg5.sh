#!/bin/bash
set -u
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e ${2:-} > gpurun_out/mg${N}_bench.json 2> gpurun_out/mg${N}_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/mg${N}_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/mg${N}_bench.json").read().strip().splitlines()[-1])
    print("weak value", d["value"], "ms", d["ms_per_step"])
    s=d["sharded"]
    for m in ("p2p","dma","nccl"): print(m, round(s[m]["ms_per_step"],3), "ms", round(s[m]["value"],1), "cubes/s")
    print("parity", s.get("parity")); print("fused_y", s["p2p"].get("fused_y_pass"))
    print("config3", d.get("config3")); print("config5", d.get("config5")); print("nccl", d.get("nccl"))
except Exception as e: print("ERR", e)
PY
