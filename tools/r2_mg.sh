#!/bin/bash
# multi-GPU checks: usage tools/r2_mg.sh <ngpus>
set -u
N=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q > gpurun_out/mg${N}_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/mg${N}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e ${2:-} > gpurun_out/mg${N}_bench.json 2> gpurun_out/mg${N}_bench.err; echo "bench rc=$?"; tail -5 gpurun_out/mg${N}_bench.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/mg${N}_bench.json").read().strip().splitlines()[-1])
    print("weak value", d["value"], "ms", d["ms_per_step"])
    print("sharded", json.dumps(d["sharded"], indent=1)[:3000])
    print("config3", d.get("config3")); print("config5", d.get("config5")); print("nccl", d.get("nccl"))
except Exception as e: print("ERR", e)
PY
