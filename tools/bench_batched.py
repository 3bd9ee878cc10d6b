#!/usr/bin/env python
"""BASELINE.json config 2: cross_power_spectrum_1d of 64 (density, xfrac) pairs at 512^3, pair i on rank i mod P
(distributed.batched, no data-path collective), inputs resident in HBM.  Launch with torchrun; rank 0 prints one JSON
line with pairs/s (whole job, max over ranks)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import tools21cm_b200 as t2c
from tools21cm_b200 import distributed as D, synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=512)
ap.add_argument("--pairs", type=int, default=64)
ap.add_argument("--steps", type=int, default=3)
args = ap.parse_args()
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
dev = torch.device("cuda", torch.cuda.current_device())
if world > 1:
    os.environ.pop("NCCL_DEBUG", None)
    dist.init_process_group("nccl", device_id=dev)
n = args.size
xfrac = synthetic.bubble_mask_torch((n, n, n), seed=2001, device=dev).float()      # one mask, many densities
mine = D.partition(args.pairs, world, rank)
dens = {i: synthetic.gaussian_cube_torch((n, n, n), seed=3000 + i, device=dev).mul_(0.3).add_(1.0) for i in mine}


def one(i):
    return t2c.cross_power_spectrum_1d(dens[i], xfrac, kbins=15, box_dims=244 / 0.7, return_n_modes=True)


def job():
    return D.batched(one, list(range(args.pairs)))


job()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    res = job()
e1.record()
torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"workload": "cross_power_spectrum_1d, %d pairs of %d^3 fp32 cubes, pair i on rank i mod P" % (args.pairs, n),
                      "n_gpus": world, "ms_per_batch": float(ms.item()), "pairs_per_s": args.pairs / (float(ms.item()) / 1e3),
                      "includes": "the all_gather_object of the 64 small results", "n_results": len(res)}))
if world > 1:
    dist.destroy_process_group()
