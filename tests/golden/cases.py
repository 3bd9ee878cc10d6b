"""The golden-vector case matrix, shared by
``make_golden.py`` (which runs the *reference* in the build container) and by the
tests (which rebuild the same seeded inputs and compare).

Each case is a dict:  id, fn, shape, field, seed, dtype, kw (keyword arguments of
the reference function).  ``edges`` entries in kw are expanded by ``build_kwargs``.
"""
from __future__ import annotations

import numpy as np

from tools21cm_b200 import synthetic

H = 0.7


def _c(id, fn, shape, field="grf", seed=11, dtype="float32", **kw):
    return dict(id=id, fn=fn, shape=tuple(shape), field=field, seed=seed, dtype=dtype, kw=kw)


CASES = [
    # ---- config 1 of BASELINE.json and close relatives -------------------
    _c("ps1d_128_grf_L200_k15", "power_spectrum_1d", (128,) * 3, seed=1234, kbins=15, box_dims=200),
    _c("ps1d_128_grf_L100_k15", "power_spectrum_1d", (128,) * 3, seed=1234, kbins=15, box_dims=100.0),
    _c("ps1d_128_toy_L200_k15", "power_spectrum_1d", (128,) * 3, field="toy", seed=1235, kbins=15, box_dims=200),
    _c("ps1d_128_grf_mean20", "power_spectrum_1d", (128,) * 3, field="grf+20", seed=5, kbins=15, box_dims=200),
    _c("ps1d_64_grf_None_k10", "power_spectrum_1d", (64,) * 3, seed=2, kbins=10, box_dims=None),
    _c("ps1d_64_grf_37h_k100", "power_spectrum_1d", (64,) * 3, seed=3, kbins=100, box_dims=37 / H),
    _c("ps1d_64_grf_lin", "power_spectrum_1d", (64,) * 3, seed=4, kbins=15, box_dims=200, binning="linear"),
    _c("ps1d_64_grf_mixed", "power_spectrum_1d", (64,) * 3, seed=4, kbins=15, box_dims=200, binning="mixed"),
    _c("ps1d_128_grf_mixed", "power_spectrum_1d", (128,) * 3, seed=1234, kbins=15, box_dims=200, binning="mixed"),
    _c("ps1d_64_grf_edges0", "power_spectrum_1d", (64,) * 3, seed=6, kbins="edges:0,0.05,0.1,0.3,0.7,1.5", box_dims=200),
    _c("ps1d_64_f64", "power_spectrum_1d", (64,) * 3, seed=7, dtype="float64", kbins=15, box_dims=244 / H),
    _c("ps1d_16_grf", "power_spectrum_1d", (16,) * 3, seed=8, kbins=4, box_dims=10.0),
    _c("ps1d_33_odd", "power_spectrum_1d", (33,) * 3, seed=9, kbins=8, box_dims=50.0),
    _c("ps1d_33_odd_f64", "power_spectrum_1d", (33,) * 3, seed=9, dtype="float64", kbins=8, box_dims=50.0),
    _c("ps1d_noncubic", "power_spectrum_1d", (64, 32, 16), seed=10, kbins=8, box_dims=[100, 50, 25]),
    _c("ps1d_noncubic_aniso", "power_spectrum_1d", (32, 64, 128), seed=12, kbins=12, box_dims=[80.0, 100.0, 90.0]),
    _c("ps1d_boxsize_kw", "power_spectrum_1d", (32,) * 3, seed=13, kbins=6, box_dims=1.0, boxsize=120.0),
    _c("ps1d_256_grf_L200", "power_spectrum_1d", (256,) * 3, seed=1234, kbins=15, box_dims=200),
    # ---- (k, mu) ----------------------------------------------------------
    _c("psmu_256_toy_10x5", "power_spectrum_mu", (256,) * 3, field="toy", seed=1235, kbins=10, mubins=5,
       box_dims=244 / H, los_axis=0),
    _c("psmu_64_los0_abs", "power_spectrum_mu", (64,) * 3, seed=21, kbins=10, mubins=5, box_dims=200, los_axis=0),
    _c("psmu_64_los1_abs", "power_spectrum_mu", (64,) * 3, seed=21, kbins=10, mubins=8, box_dims=200, los_axis=1),
    _c("psmu_64_los2_abs", "power_spectrum_mu", (64,) * 3, seed=21, kbins=10, mubins=20, box_dims=200, los_axis=2),
    _c("psmu_64_los0_signed", "power_spectrum_mu", (64,) * 3, seed=22, kbins=10, mubins=8, box_dims=200,
       los_axis=0, absolute_mus=False),
    _c("psmu_64_los1_signed", "power_spectrum_mu", (64,) * 3, seed=22, kbins=10, mubins=5, box_dims=100.0,
       los_axis=1, absolute_mus=False),
    _c("psmu_64_los2_signed", "power_spectrum_mu", (64,) * 3, seed=22, kbins=15, mubins=8, box_dims=244 / H,
       los_axis=2, absolute_mus=False),
    _c("psmu_64_keepzero", "power_spectrum_mu", (64,) * 3, seed=23, kbins=10, mubins=5, box_dims=200,
       los_axis=0, exclude_zero_modes=False),
    _c("psmu_64_mu2", "power_spectrum_mu", (64,) * 3, seed=23, kbins=10, mubins=2, box_dims=200),
    _c("psmu_33_odd", "power_spectrum_mu", (33,) * 3, seed=24, kbins=6, mubins=5, box_dims=50.0, los_axis=2),
    _c("psmu_33_odd_signed", "power_spectrum_mu", (33,) * 3, seed=24, kbins=6, mubins=4, box_dims=50.0,
       los_axis=0, absolute_mus=False),
    _c("psmu_noncubic", "power_spectrum_mu", (64, 32, 16), seed=25, kbins=6, mubins=5, box_dims=[100, 50, 25],
       los_axis=0),
    _c("psmu_noncubic_los2_signed", "power_spectrum_mu", (64, 32, 16), seed=25, kbins=6, mubins=6,
       box_dims=[100, 50, 25], los_axis=2, absolute_mus=False),
    _c("psmu_128_tutorial", "power_spectrum_mu", (128,) * 3, field="toy", seed=26, kbins=15, mubins=8,
       box_dims=244 / H, los_axis=0, absolute_mus=False),
    _c("psmu_64_muedges", "power_spectrum_mu", (64,) * 3, seed=27, kbins="edges:0.05,0.1,0.2,0.4,0.8",
       mubins="edges:0,0.3,0.6,0.8,1.0", box_dims=200),
    # ---- cross spectra ----------------------------------------------------
    _c("xps1d_128_self", "cross_power_spectrum_1d", (128,) * 3, field="grf|same", seed=31, kbins=10, box_dims=200),
    _c("xps1d_64_pair", "cross_power_spectrum_1d", (64,) * 3, field="c2ray", seed=2000, kbins=15, box_dims=244 / H),
    _c("xps1d_128_pair", "cross_power_spectrum_1d", (128,) * 3, field="c2ray", seed=2001, kbins=15, box_dims=244 / H),
    _c("xps1d_noncubic", "cross_power_spectrum_1d", (64, 32, 16), field="grf|grf", seed=32, kbins=8,
       box_dims=[100, 50, 25]),
    _c("xpsmu_64_pair", "cross_power_spectrum_mu", (64,) * 3, field="c2ray", seed=2002, kbins=10, mubins=5,
       box_dims=200, los_axis=0),
    _c("xpsmu_64_signed", "cross_power_spectrum_mu", (64,) * 3, field="grf|grf", seed=33, kbins=10, mubins=4,
       box_dims=200, los_axis=2, absolute_mus=False),
    # ---- n-D outputs (full arrays kept only for small shapes) --------------
    _c("psnd_16", "power_spectrum_nd", (16,) * 3, seed=41, box_dims=10.0),
    _c("psnd_33", "power_spectrum_nd", (33,) * 3, seed=42, box_dims=50.0),
    _c("psnd_noncubic", "power_spectrum_nd", (32, 16, 8), field="grf+20", seed=43, box_dims=[100, 50, 25]),
    _c("psnd_32_f64", "power_spectrum_nd", (32,) * 3, seed=44, dtype="float64", box_dims=None),
    _c("xpsnd_32", "cross_power_spectrum_nd", (32,) * 3, field="c2ray", seed=2003, box_dims=200),
    _c("xpsnd_33", "cross_power_spectrum_nd", (33,) * 3, field="grf|grf", seed=45, box_dims=50.0),
    # ---- (k_perp, k_par) -----------------------------
    _c("ps2d_64_log", "power_spectrum_2d", (64,) * 3, seed=61, kbins=8, box_dims=200.0, return_modes=True),
    _c("ps2d_64_lin_nu0", "power_spectrum_2d", (64,) * 3, field="toy", seed=62, kbins=6, binning="linear",
       box_dims=244 / H, nu_axis=0, return_modes=True),
    _c("ps2d_noncubic_nu1", "power_spectrum_2d", (32, 64, 16), seed=63, kbins=5, box_dims=[80.0, 100.0, 40.0],
       nu_axis=1, return_modes=True),
    _c("ps2d_33_odd", "power_spectrum_2d", (33,) * 3, seed=64, dtype="float64", kbins=5, box_dims=50.0,
       return_modes=True),
    _c("ps2d_edges", "power_spectrum_2d", (32,) * 3, seed=65, dtype="float64",
       kbins=["edges:0,0.2,0.5,1.0,2.0", "edges:0,0.3,0.9,2.0"], box_dims=100.0, return_modes=True),
    # ---- multipoles; the reference runs only with its _get_mu call repaired and with
    # exclude_zero_modes=True (make_golden.py), so these are the cases it can pin
    _c("mpoles_64_los0", "power_spectrum_multipoles", (64,) * 3, field="toy", seed=71, kbins=8, box_dims=200.0, los_axis=0,
       exclude_zero_modes=True),
    _c("mpoles_64_los2_f64", "power_spectrum_multipoles", (64,) * 3, seed=72, dtype="float64", kbins=10, box_dims=244 / H,
       los_axis=2, exclude_zero_modes=True),
    _c("mpoles_noncubic_los1", "power_spectrum_multipoles", (32, 64, 16), seed=73, kbins=6, box_dims=[80.0, 100.0, 40.0],
       los_axis=1, exclude_zero_modes=True),
    _c("mpoles_33_odd", "power_spectrum_multipoles", (33,) * 3, seed=74, dtype="float64", kbins=5, box_dims=50.0, los_axis=0,
       exclude_zero_modes=True),
    # ---- integrated bispectrum / power-spectrum response ----------------------------
    _c("ibis_32_c2", "integrated_bispectrum", (32,) * 3, field="toy", seed=81, Ncuts=2, kbins=6, box_dims=100.0, verbose=False),
    _c("ibisx_32_c4_norm", "integrated_bispectrum_cross", (32,) * 3, field="c2ray", seed=2004, Ncuts=4, kbins=5,
       box_dims=244 / H, binning="linear", normalize=True, verbose=False),
    _c("psresp_64_c2", "power_spectrum_response", (64,) * 3, field="toy", seed=82, dtype="float64", Ncuts=2, kbins=8,
       box_dims=200.0, verbose=False),
    # ---- stand-alone binning of caller-supplied fftshifted arrays ----------
    _c("ravg_3d", "radial_average", (32,) * 3, field="positive", seed=51, box_dims=[200.0] * 3, kbins=8),
    _c("ravg_3d_lin", "radial_average", (33, 32, 16), field="positive", seed=52, box_dims=[100.0, 90.0, 40.0],
       kbins=7, binning="linear"),
    _c("ravg_2d", "radial_average", (64, 64), field="positive", seed=53, box_dims=[300.0, 300.0], kbins=10),
    _c("ravg_2d_odd", "radial_average", (45, 64), field="positive", seed=54, box_dims=[300.0, 200.0], kbins=6),
    _c("ravg_1d", "radial_average", (128,), field="positive", seed=55, box_dims=[300.0], kbins=5),
    _c("mubin_3d", "mu_binning", (32,) * 3, field="positive", seed=56, box_dims=[200.0] * 3, kbins=6, mubins=5,
       los_axis=1),
    _c("mubin_3d_signed", "mu_binning", (32, 16, 64), field="positive", seed=57, box_dims=[100.0, 50.0, 200.0],
       kbins=6, mubins=4, los_axis=2, absolute_mus=False),
]

CASE_BY_ID = {c["id"]: c for c in CASES}


def build_inputs(case):
    """Recreate the seeded input array(s) of a case (numpy, host)."""
    shape, seed, field = case["shape"], case["seed"], case["field"]
    dt = np.dtype(case["dtype"])
    if field == "grf":
        return [synthetic.gaussian_cube(shape, seed, dt)]
    if field == "grf+20":
        return [(synthetic.gaussian_cube(shape, seed, np.float64) + 20.0).astype(dt)]
    if field == "toy":
        return [synthetic.toy_dt_cube(shape, seed, dt)]
    if field == "grf|same":
        a = synthetic.gaussian_cube(shape, seed, dt)
        return [a, a.copy()]
    if field == "grf|grf":
        return [synthetic.gaussian_cube(shape, seed, dt), synthetic.gaussian_cube(shape, seed + 1000, dt)]
    if field == "c2ray":
        a, b = synthetic.c2ray_like_pair(shape, seed, dt)
        return [a, b]
    if field == "positive":  # a stand-in for an already computed fftshifted power array
        rng = np.random.default_rng(seed)
        return [rng.random(size=shape) + 0.1]
    raise ValueError(field)


def build_kwargs(case):
    kw = {}
    for key, val in case["kw"].items():
        if isinstance(val, str) and val.startswith("edges:"):
            val = np.array([float(t) for t in val[6:].split(",")])
        if isinstance(val, list) and val and isinstance(val[0], str) and val[0].startswith("edges:"):
            val = [np.array([float(t) for t in v[6:].split(",")]) for v in val]
        kw[key] = val
    return kw
