#!/bin/bash
# the driver's launch line on 4 GPUs (weak batches + one sharded cube)
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r02_bench_4gpu.json 2> gpurun_out/r02_bench_4gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench_4gpu.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02_bench_4gpu.json").read().strip().splitlines()[-1])
sh=d["sharded"]
print("weak", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], "sharded", sh["p2p"]["ms_per_step"], sh["p2p"]["value"], sh["nccl"]["ms_per_step"], sh["parity"], d["nccl"])
PY
