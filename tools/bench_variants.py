#!/usr/bin/env python
"""Times the other estimators of the path at one cube size (inputs resident in HBM, CUDA events,
warm-up first).  Prints one JSON object; bench.py stays the headline."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tools21cm_b200 as t2c
from tools21cm_b200 import synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=1024)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--quick", action="store_true", help="only the fp32 binned estimators")
ap.add_argument("--cross-only", action="store_true", help="only cross_power_spectrum_1d (large cubes)")
args = ap.parse_args()
n = args.size
dev = torch.device("cuda", 0)
L = 714.0
if args.cross_only:         # large cubes: plain normal deviates, no temporaries
    g = torch.Generator(device=dev).manual_seed(1)
    a = torch.randn((n, n, n), generator=g, device=dev)
    b = torch.randn((n, n, n), generator=g, device=dev)
else:
    a = synthetic.toy_dt_cube_torch((n, n, n), seed=1, device=dev)
    b = synthetic.gaussian_cube_torch((n, n, n), seed=2, device=dev)


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / args.steps


out = {"size": n, "unit": "ms per call", "dtype": "f32"}
if args.cross_only:
    out["cross_power_spectrum_1d"] = timed(lambda: t2c.cross_power_spectrum_1d(a, b, kbins=15, box_dims=L))
    print(json.dumps(out))
    sys.exit(0)
out["power_spectrum_1d"] = timed(lambda: t2c.power_spectrum_1d(a, kbins=15, box_dims=L))
out["cross_power_spectrum_1d"] = timed(lambda: t2c.cross_power_spectrum_1d(a, b, kbins=15, box_dims=L))
for los in (0, 1, 2):
    out["power_spectrum_mu_los%d_abs" % los] = timed(
        lambda: t2c.power_spectrum_mu(a, los_axis=los, mubins=5, kbins=10, box_dims=L))
out["power_spectrum_mu_los0_signed"] = timed(
    lambda: t2c.power_spectrum_mu(a, los_axis=0, mubins=8, kbins=15, box_dims=L, absolute_mus=False))
out["power_spectrum_mu_los2_signed"] = timed(
    lambda: t2c.power_spectrum_mu(a, los_axis=2, mubins=8, kbins=15, box_dims=L, absolute_mus=False))
out["power_spectrum_1d_kbins100"] = timed(lambda: t2c.power_spectrum_1d(a, kbins=100, box_dims=L))
del b
torch.cuda.empty_cache()
if not args.quick:
    a64 = a.double()
    out["power_spectrum_1d_f64"] = timed(lambda: t2c.power_spectrum_1d(a64, kbins=15, box_dims=L))
    del a64
    torch.cuda.empty_cache()
if args.quick:
    print(json.dumps(out))
    sys.exit(0)
out["power_spectrum_nd"] = timed(lambda: t2c.power_spectrum_nd(a, box_dims=L))
out["power_spectrum_2d"] = timed(lambda: t2c.power_spectrum_2d(a, kbins=10, box_dims=L))
out["power_spectrum_multipoles_los0"] = timed(lambda: t2c.power_spectrum_multipoles(a, kbins=15, box_dims=L, los_axis=0))
out["power_spectrum_multipoles_los2"] = timed(lambda: t2c.power_spectrum_multipoles(a, kbins=15, box_dims=L, los_axis=2))
if n >= 256:
    import time
    torch.cuda.synchronize(); t0 = time.perf_counter()
    t2c.integrated_bispectrum(a, Ncuts=4, kbins=15, box_dims=L, verbose=False)
    torch.cuda.synchronize()
    out["integrated_bispectrum_Ncuts4_64_subboxes"] = (time.perf_counter() - t0) * 1e3
print(json.dumps(out))
