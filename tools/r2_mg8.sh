#!/bin/bash
# the 8-GPU evidence run: multi-rank parity tests, then the driver's launch line for bench.py
set -u
N=${1:-8}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_distributed.py -m gpu -q -rA > gpurun_out/r02_pytest_gpu_${N}gpu_distributed.log 2>&1; echo "pytest rc=$?"; tail -18 gpurun_out/r02_pytest_gpu_${N}gpu_distributed.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo "bench rc=$?"; tail -12 gpurun_out/r02_bench_${N}gpu.err; ls -la /tmp/t21_nccl* 2>/dev/null | head -3
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_bench_${N}gpu.json").read().strip().splitlines()[-1])
    print("weak value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"])
    print("sharded", json.dumps(d["sharded"], indent=1)[:2600])
    print("config3", d.get("config3")); print("config5", d.get("config5")); print("nccl", d.get("nccl"))
except Exception as e: print("ERR", e)
PY
