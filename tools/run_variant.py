#!/usr/bin/env python
"""Run one estimator variant a few times at 1024^3 (for ncu captures).  usage: run_variant.py <name> [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tools21cm_b200 as t2c
from tools21cm_b200 import synthetic
name = sys.argv[1]; reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
n = int(os.environ.get("N", "1024")); L = 714.0
a = synthetic.toy_dt_cube_torch((n, n, n), seed=1, device="cuda")
V = {
    "ps1d": lambda: t2c.power_spectrum_1d(a, kbins=15, box_dims=L),
    "mu0": lambda: t2c.power_spectrum_mu(a, los_axis=0, mubins=5, kbins=10, box_dims=L),
    "mu1": lambda: t2c.power_spectrum_mu(a, los_axis=1, mubins=5, kbins=10, box_dims=L),
    "mu2": lambda: t2c.power_spectrum_mu(a, los_axis=2, mubins=5, kbins=10, box_dims=L),
    "mu0s": lambda: t2c.power_spectrum_mu(a, los_axis=0, mubins=8, kbins=15, box_dims=L, absolute_mus=False),
    "mu2s": lambda: t2c.power_spectrum_mu(a, los_axis=2, mubins=8, kbins=15, box_dims=L, absolute_mus=False),
    "k100": lambda: t2c.power_spectrum_1d(a, kbins=100, box_dims=L),
    "nd": lambda: t2c.power_spectrum_nd(a, box_dims=L),
    "2d": lambda: t2c.power_spectrum_2d(a, kbins=10, box_dims=L),
    "mp0": lambda: t2c.power_spectrum_multipoles(a, kbins=15, box_dims=L, los_axis=0),
}
for _ in range(reps):
    V[name]()
torch.cuda.synchronize()
print("ok", name)
