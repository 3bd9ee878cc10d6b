#!/bin/bash
# 2-GPU validation of the table-driven X pass on the slab-sharded path (+ configs 3 / 5 at reduced size)
set -u
timeout 900 python -m pytest tests/test_gpu_distributed.py -m gpu -q -rA > gpurun_out/r02_pytest_gpu_2gpu_distributed.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_pytest_gpu_2gpu_distributed.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --configs35 --c5-size 1024 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "bench rc=$?"; tail -5 gpurun_out/r02_bench_2gpu.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_2gpu.json").read().strip().splitlines()[-1])
    print("weak value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"])
    print("sharded", json.dumps(d["sharded"], indent=1)[:2600])
    print("config3", d.get("config3")); print("config5", json.dumps(d.get("config5"), indent=1)); print("nccl", d.get("nccl"))
except Exception as e: print("ERR", e)
PY
