#!/bin/bash
# Multi-GPU evidence with the driver's own launch line: default bench (weak batches + one sharded cube) and
# the distributed GPU tests.  usage: tools/final_ngpu.sh N
set -u
N=$1
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q 2>&1 | tail -1
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
   bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/f${N}_default.json 2> gpurun_out/f${N}_default.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/f${N}_default.json").read().strip().splitlines()[-1])
print(round(d["value"],1), d["ms_per_step"], d["e2e"]["value"], d["sharded"]["ms_per_step"], d["sharded"]["all_to_all"]["achieved_gbs"])
PY
