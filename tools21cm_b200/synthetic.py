"""Seeded synthetic cubes used by the tests, the golden-vector script and bench.py.

Two families, as SURVEY.md section 8(d) defines them:

* Gaussian-random field (white noise, unit variance).
* Toy-ionised 21-cm brightness temperature: dT = 27 mK * (1 - x_i) * (1 + delta),
  delta = 0.3 * N(0,1), x_i a union of random spherical bubbles of radius
  N/32 cells with target filling fraction 0.3 (a restatement of the idea of the
  reference's ``test_case.py:34-122`` bubble model; no code shared).  Its mean
  (~19 mK) is far from zero, which exercises the mean-offset path of the FFT.

The numpy generators are the ones every parity test uses (host -> device copy);
the torch generators build the benchmark-size cubes directly in HBM.
"""
from __future__ import annotations

import numpy as np


def gaussian_cube(shape, seed=1234, dtype=np.float32):
    rng = np.random.default_rng(seed)
    return rng.standard_normal(size=tuple(shape)).astype(dtype)


def _bubble_centres(shape, seed, fill=0.3):
    n = int(round(float(np.prod(shape)) ** (1.0 / 3.0)))
    r = max(1.0, n / 32.0)
    count = max(1, int(fill * float(np.prod(shape)) / (4.0 / 3.0 * np.pi * r ** 3)))
    rng = np.random.default_rng(seed)
    centres = np.stack([rng.integers(0, s, size=count) for s in shape], axis=1)
    return centres, r


def bubble_mask(shape, seed=1236, fill=0.3):
    """x_i in {0,1}: union of spheres clipped at the box faces."""
    centres, r = _bubble_centres(shape, seed, fill)
    mask = np.zeros(shape, dtype=bool)
    ri = int(np.ceil(r))
    for c in centres:
        lo = [max(0, int(c[d]) - ri) for d in range(3)]
        hi = [min(shape[d], int(c[d]) + ri + 1) for d in range(3)]
        g = np.ogrid[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]
        d2 = (g[0] - c[0]) ** 2 + (g[1] - c[1]) ** 2 + (g[2] - c[2]) ** 2
        mask[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] |= d2 <= r * r
    return mask


def toy_dt_cube(shape, seed=1235, dtype=np.float32):
    delta = 0.3 * np.random.default_rng(seed).standard_normal(size=tuple(shape))
    xi = bubble_mask(shape, seed + 1)
    return (27.0 * (1.0 - xi) * (1.0 + delta)).astype(dtype)


def c2ray_like_pair(shape, seed=2000, dtype=np.float32):
    """(density, xfrac): density = 1 + delta, xfrac = the bubble mask."""
    delta = 0.3 * np.random.default_rng(seed).standard_normal(size=tuple(shape))
    xi = bubble_mask(shape, seed + 1)
    return (1.0 + delta).astype(dtype), xi.astype(dtype)


# ---- device-side generators for benchmark sizes -------------------------------
def gaussian_cube_torch(shape, seed=1234, device="cuda", dtype=None):
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    return torch.randn(tuple(shape), generator=g, device=device,
                       dtype=dtype or torch.float32)


def bubble_mask_torch(shape, seed=1236, fill=0.3, device="cuda"):
    import torch
    centres, r = _bubble_centres(shape, seed, fill)
    mask = torch.zeros(tuple(shape), dtype=torch.bool, device=device)
    ri = int(np.ceil(r))
    ax = torch.arange(-ri, ri + 1, device=device)
    ball = (ax[:, None, None] ** 2 + ax[None, :, None] ** 2 + ax[None, None, :] ** 2) <= r * r
    for c in centres:
        lo = [max(0, int(c[d]) - ri) for d in range(3)]
        hi = [min(shape[d], int(c[d]) + ri + 1) for d in range(3)]
        b = ball[lo[0] - int(c[0]) + ri:hi[0] - int(c[0]) + ri,
                 lo[1] - int(c[1]) + ri:hi[1] - int(c[1]) + ri,
                 lo[2] - int(c[2]) + ri:hi[2] - int(c[2]) + ri]
        mask[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]] |= b
    return mask


def toy_dt_cube_torch(shape, seed=1235, device="cuda", dtype=None):
    import torch
    dtype = dtype or torch.float32
    out = gaussian_cube_torch(shape, seed, device, dtype)
    out.mul_(0.3).add_(1.0).mul_(27.0)
    xi = bubble_mask_torch(shape, seed + 1, device=device)
    out.masked_fill_(xi, 0.0)
    return out
