#!/bin/bash
# Round-end multi-GPU evidence (one box, 8 GPUs): weak batches + one 1024^3 cube sharded, then the 2048^3 cross
# spectrum (BASELINE.json config 5) with both exchanges and with the n-D output left sharded.
set -u
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
run() { name=$1; shift; timeout 400 env "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err; echo "$name rc=$?"; tail -c 400 gpurun_out/$name.json; echo; }
run f${N}_weak_p2p T21_EXCHANGE=p2p $TR bench.py --gpus $N --steps 10 --warmup 3
run f${N}_x2048_nccl T21_EXCHANGE=nccl $TR bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --slab-size 2048 --slab-cross
run f${N}_x2048_p2p T21_EXCHANGE=p2p $TR bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --slab-size 2048 --slab-cross
run f${N}_nd2048_nccl T21_EXCHANGE=nccl $TR bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --slab-size 2048 --slab-cross --slab-nd
