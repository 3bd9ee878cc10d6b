"""GPU parity: the public API (tools21cm_b200, through the C ABI) against the golden vectors the
reference produced and against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): n_modes bit-exact; P within 1e-5 relative for fp32 cubes and
1e-12 (zero-mean inputs) for fp64 cubes.  Two documented loosenings, both properties of the
*reference's* arithmetic rather than of the kernels:
  * cubes with a large mean: the reference's np.histogram(weights=) differences running sums
    that contain the huge DC sample, so its own bin sums are only good to ~1e-8 (fp64 tol 2e-7);
  * cross spectra: a bin is a signed sum with cancellation, so the fp32 error is bounded relative
    to the auto-power level (atol = 1e-5 * max|P|), not to the possibly near-zero bin value.
"""
import numpy as np
import pytest

from cases import CASES, build_inputs, build_kwargs
from conftest import golden_outputs

pytestmark = pytest.mark.gpu

POW2 = (8, 16, 32, 64, 128, 256, 512, 1024)
BINNED = ("power_spectrum_1d", "cross_power_spectrum_1d", "power_spectrum_mu", "cross_power_spectrum_mu")


def supported(case):
    return True   # power-of-two axes take the Stockham kernels, every other length the generic DFT kernels


def tolerances(case, want0):
    single = case["dtype"] == "float32"
    zero_mean = case["field"] in ("grf", "grf|grf", "grf|same", "positive")
    cross = case["fn"].startswith("cross")
    rtol = 1e-5 if single else (1e-12 if zero_mean else 2e-7)
    atol = 0.0
    if cross:
        atol = (1e-5 if single else 1e-12) * float(np.nanmax(np.abs(want0)))
    return rtol, atol


@pytest.fixture(scope="module")
def t2c():
    import torch
    assert torch.cuda.is_available()
    import tools21cm_b200
    return tools21cm_b200


RUN = [c for c in CASES if supported(c)]


@pytest.mark.parametrize("case", RUN, ids=[c["id"] for c in RUN])
def test_api_matches_reference_golden(t2c, case, golden):
    fn = getattr(t2c, case["fn"])
    kw = build_kwargs(case)
    if case["fn"] in BINNED:
        kw["return_n_modes"] = True
    res = fn(*build_inputs(case), **kw)
    if not isinstance(res, tuple):
        res = (res,)
    want = golden_outputs(golden, case["id"])
    if case["fn"] == "power_spectrum_multipoles":
        mp, kc = res
        single = case["dtype"] == "float32"
        assert np.array_equal(kc, want[4]) and np.array_equal(mp["nmodes"], want[3])
        np.testing.assert_allclose(mp["P0"], want[0], rtol=1e-5 if single else 1e-12)
        # P2 / P4 are signed sums that nearly cancel for an isotropic field: bounded relative to the monopole
        atol = (2e-5 if single else 1e-11) * float(np.max(np.abs(want[0])))
        np.testing.assert_allclose(mp["P2"], want[1], rtol=1e-5 if single else 1e-10, atol=atol)
        np.testing.assert_allclose(mp["P4"], want[2], rtol=1e-5 if single else 1e-10, atol=atol)
        return
    if case["fn"] in ("integrated_bispectrum", "integrated_bispectrum_cross", "power_spectrum_response"):
        assert len(res) == 2 and np.array_equal(res[1], want[1])
        np.testing.assert_allclose(res[0], want[0], rtol=2e-5 if case["dtype"] == "float32" else 1e-10, equal_nan=True)
        return
    assert len(res) == len(want)
    for got, ref in zip(res, want):
        assert np.asarray(got).shape == ref.shape
        assert np.asarray(got).dtype == ref.dtype
    rtol, atol = tolerances(case, want[0])
    if case["fn"] == "power_spectrum_2d":
        assert np.array_equal(res[3], want[3]), "n_modes differ from the reference"
        assert np.array_equal(res[1], want[1]) and np.array_equal(res[2], want[2])
        np.testing.assert_allclose(res[0], want[0], rtol=rtol, equal_nan=True)
    elif case["fn"] in BINNED or case["fn"] in ("radial_average", "mu_binning"):
        assert np.array_equal(res[-1], want[-1]), "n_modes differ from the reference"
        for got, ref in zip(res[1:-1], want[1:-1]):
            assert np.array_equal(got, ref), "bin centres differ from the reference"
        np.testing.assert_allclose(res[0], want[0], rtol=rtol, atol=atol, equal_nan=True)
    else:
        # n-D output: every sample, relative to the largest power (individual modes can be ~0)
        tol = 2e-5 if case["dtype"] == "float32" else 1e-11
        assert np.abs(res[0] - want[0]).max() <= tol * np.abs(want[0]).max()


@pytest.mark.parametrize("shape", [(12, 20, 30), (1, 64, 64), (5, 7, 9), (250, 8, 16), (8, 250, 16), (16, 8, 250),
                                   (22, 26, 39), (127, 9, 10), (300, 6, 49), (7, 600, 10), (45, 11, 1000)])
def test_mixed_and_odd_shapes_against_oracle(t2c, shape):
    """Axes that are not powers of two (C2Ray-like 250 / 300 / 600, odd light-cone chunks, degenerate axes) go through the
    generic kernels -- mixed-radix (2, 3, 4, 5, 7, 11, 13) in shared memory, direct sums for other primes (127); any mix with
    Stockham axes must agree with the oracle."""
    from oracle import ps_oracle
    x = np.random.default_rng(sum(shape)).standard_normal(shape)
    box = [100.0, 80.0, 120.0]
    ps, ks, nm = t2c.power_spectrum_1d(x, kbins=6, box_dims=box, return_n_modes=True)
    ps_o, ks_o, nm_o = ps_oracle.power_spectrum_1d(x, kbins=6, box_dims=box, return_n_modes=True)
    assert np.array_equal(nm, nm_o) and np.array_equal(ks, ks_o)
    np.testing.assert_allclose(ps, ps_o, rtol=1e-11, equal_nan=True)
    nd = t2c.power_spectrum_nd(x, box_dims=box)
    nd_o = ps_oracle.power_spectrum_nd(x, box_dims=box)
    assert np.abs(nd - nd_o).max() <= 1e-11 * nd_o.max()


@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("shape", [(16, 32, 1024), (16, 1024, 32), (1024, 16, 32), (8, 16, 2048), (2048, 8, 16), (16, 8, 4096)])
def test_long_pencils_against_oracle(t2c, shape, dtype):
    """The pencil lengths of the headline configurations (1024, 2048) and the longest instantiated one along
    each axis in turn, in thin slabs the oracle finishes in seconds: three-stage plans, the Z pass's fused last
    stage (nz = 1024), the half-width cross tiles (nx = 2048)."""
    from oracle import ps_oracle
    rng = np.random.default_rng(sum(shape))
    x = rng.standard_normal(shape).astype(dtype)
    y = (0.5 * x + rng.standard_normal(shape)).astype(dtype)
    box = [100.0, 80.0, 120.0]
    tol = 1e-5 if dtype == "float32" else 1e-11
    ps, ks, nm = t2c.power_spectrum_1d(x, kbins=8, box_dims=box, return_n_modes=True)
    ps_o, ks_o, nm_o = ps_oracle.power_spectrum_1d(x, kbins=8, box_dims=box, return_n_modes=True)
    assert np.array_equal(nm, nm_o) and np.array_equal(ks, ks_o)
    np.testing.assert_allclose(ps, ps_o, rtol=tol, equal_nan=True)
    xs, _, xn = t2c.cross_power_spectrum_1d(x, y, kbins=8, box_dims=box, return_n_modes=True)
    xs_o, _, xn_o = ps_oracle.cross_power_spectrum_1d(x, y, kbins=8, box_dims=box, return_n_modes=True)
    assert np.array_equal(xn, xn_o)
    np.testing.assert_allclose(xs, xs_o, rtol=tol, atol=tol * np.nanmax(np.abs(ps_o)), equal_nan=True)
    pm, _, _, nmu = t2c.power_spectrum_mu(x, los_axis=2, mubins=4, kbins=5, box_dims=box, return_n_modes=True)
    pm_o, _, _, nmu_o = ps_oracle.power_spectrum_mu(x, los_axis=2, mubins=4, kbins=5, box_dims=box, return_n_modes=True)
    assert np.array_equal(nmu, nmu_o)
    np.testing.assert_allclose(pm, pm_o, rtol=tol, equal_nan=True)
    nd = t2c.power_spectrum_nd(x, box_dims=box)
    nd_o = ps_oracle.power_spectrum_nd(x, box_dims=box)
    assert np.abs(nd - nd_o).max() <= tol * nd_o.max()


@pytest.mark.parametrize("shape", [(1024, 16, 32), (1024, 32, 16)])
@pytest.mark.parametrize("los", [0, 1, 2])
@pytest.mark.parametrize("absmu", [True, False])
def test_warp_per_pencil_x_pass_mu_modes_against_oracle(t2c, shape, los, absmu):
    """nx = 1024 fp32 takes the warp-per-pencil X pass (t21_xpass32.cu): every line of sight, |mu| and signed mu
    (twins binned separately), the excluded k_perp = 0 line, against the oracle; kx runs over the full 1024 samples."""
    from oracle import ps_oracle
    rng = np.random.default_rng(7 + los)
    x = (rng.standard_normal(shape) + 3.0).astype(np.float32)
    box = [300.0, 40.0, 60.0]
    for excl in (True, False):
        kw = dict(los_axis=los, mubins=6, kbins=7, box_dims=box, return_n_modes=True, absolute_mus=absmu, exclude_zero_modes=excl)
        pm, _, _, nmu = t2c.power_spectrum_mu(x, **kw)
        pm_o, _, _, nmu_o = ps_oracle.power_spectrum_mu(x, **kw)
        assert np.array_equal(nmu, nmu_o)
        np.testing.assert_allclose(pm, pm_o, rtol=1e-5, equal_nan=True)
    ps, ks, nm = t2c.power_spectrum_1d(x, kbins=31, box_dims=box, return_n_modes=True)          # 32 edges: one per lane
    ps_o, ks_o, nm_o = ps_oracle.power_spectrum_1d(x, kbins=31, box_dims=box, return_n_modes=True)
    assert np.array_equal(nm, nm_o)
    np.testing.assert_allclose(ps, ps_o, rtol=1e-5, equal_nan=True)
    e = np.array([0.0, 0.05, 0.3, 0.31, 0.31, 2.0, 9.0])                                           # zero edge, empty bin
    ps, ks, nm = t2c.power_spectrum_1d(x, kbins=e, box_dims=box, return_n_modes=True)
    ps_o, ks_o, nm_o = ps_oracle.power_spectrum_1d(x, kbins=e, box_dims=box, return_n_modes=True)
    assert np.array_equal(nm, nm_o)
    np.testing.assert_allclose(ps, ps_o, rtol=1e-5, equal_nan=True)


@pytest.mark.parametrize("shape,los", [((64, 64, 64), 0), ((48, 40, 72), 1), ((33, 33, 33), 2), ((128, 16, 32), 2)])
def test_multipoles_against_oracle(t2c, shape, los):
    """power_spectrum_multipoles with and without the k_perp = 0 line, explicit edges (a bin that holds DC: mu = NaN
    there, so P2 / P4 of that bin are NaN like the oracle's), mode counts against the oracle's count."""
    from oracle import ps_oracle
    x = np.random.default_rng(sum(shape)).standard_normal(shape) + 1.5
    box = [120.0, 100.0, 180.0]
    for kb, excl in ((7, True), (7, False), (np.array([0.0, 0.2, 0.5, 1.0, 2.5]), False)):
        got, kc = t2c.power_spectrum_multipoles(x, kbins=kb, box_dims=box, los_axis=los, exclude_zero_modes=excl)
        with np.errstate(all="ignore"):
            want, kc_o = ps_oracle.power_spectrum_multipoles(x, kbins=kb, box_dims=box, los_axis=los, exclude_zero_modes=excl)
        assert np.array_equal(kc, kc_o)
        assert np.array_equal(got["n_modes"], want["n_modes"]) and np.array_equal(got["nmodes"], want["nmodes"])
        np.testing.assert_allclose(got["P0"], want["P0"], rtol=1e-11, equal_nan=True)
        atol = 1e-10 * float(np.nanmax(np.abs(want["P0"])))
        np.testing.assert_allclose(got["P2"], want["P2"], rtol=1e-6, atol=atol, equal_nan=True)
        np.testing.assert_allclose(got["P4"], want["P4"], rtol=1e-6, atol=atol, equal_nan=True)


@pytest.mark.parametrize("shape", [(512, 16, 32), (512, 32, 16), (512, 8, 64)])
def test_warp_per_pencil_x_pass_nx512(t2c, shape):
    """nx = 512 fp32 takes its own warp-per-pencil X pass (t21_xpass16.cu: radix 16 x 2 x 16, no CTA barrier), for one field
    and for two (cross spectra keep the first field's results in registers): shells, |mu| bins for every line of sight, the
    excluded k_perp = 0 line, explicit edges with a zero edge and an empty bin, against the oracle."""
    from oracle import ps_oracle
    rng = np.random.default_rng(sum(shape))
    x = (rng.standard_normal(shape) + 3.0).astype(np.float32)
    y = (0.5 * x + rng.standard_normal(shape)).astype(np.float32)
    box = [300.0, 40.0, 60.0]
    for kb in (9, 31, np.array([0.0, 0.05, 0.3, 0.31, 0.31, 2.0, 9.0])):
        ps, ks, nm = t2c.power_spectrum_1d(x, kbins=kb, box_dims=box, return_n_modes=True)
        ps_o, ks_o, nm_o = ps_oracle.power_spectrum_1d(x, kbins=kb, box_dims=box, return_n_modes=True)
        assert np.array_equal(nm, nm_o) and np.array_equal(ks, ks_o)
        np.testing.assert_allclose(ps, ps_o, rtol=1e-5, equal_nan=True)
        xs, _, xn = t2c.cross_power_spectrum_1d(x, y, kbins=kb, box_dims=box, return_n_modes=True)
        xs_o, _, xn_o = ps_oracle.cross_power_spectrum_1d(x, y, kbins=kb, box_dims=box, return_n_modes=True)
        assert np.array_equal(xn, xn_o)
        np.testing.assert_allclose(xs, xs_o, rtol=1e-5, atol=1e-5 * np.nanmax(np.abs(ps_o)), equal_nan=True)
    for los in (0, 1, 2):
        for excl in (True, False):
            kw = dict(los_axis=los, mubins=5, kbins=6, box_dims=box, return_n_modes=True, exclude_zero_modes=excl)
            pm, _, _, nmu = t2c.power_spectrum_mu(x, **kw)
            pm_o, _, _, nmu_o = ps_oracle.power_spectrum_mu(x, **kw)
            assert np.array_equal(nmu, nmu_o)
            np.testing.assert_allclose(pm, pm_o, rtol=1e-5, equal_nan=True)
        cm, _, _, cn = t2c.cross_power_spectrum_mu(x, y, los_axis=los, mubins=4, kbins=5, box_dims=box, return_n_modes=True)
        cm_o, _, _, cn_o = ps_oracle.cross_power_spectrum_mu(x, y, los_axis=los, mubins=4, kbins=5, box_dims=box, return_n_modes=True)
        assert np.array_equal(cn, cn_o)
        np.testing.assert_allclose(cm, cm_o, rtol=1e-5, atol=1e-5 * np.nanmax(np.abs(pm_o)), equal_nan=True)
        for excl in (True, False):                       # signed mu: four tables (own / twin bins for kx >= 0 and kx < 0)
            kw = dict(los_axis=los, mubins=6, kbins=5, box_dims=box, return_n_modes=True, exclude_zero_modes=excl, absolute_mus=False)
            pm, _, _, nmu = t2c.power_spectrum_mu(x, **kw)
            pm_o, _, _, nmu_o = ps_oracle.power_spectrum_mu(x, **kw)
            assert np.array_equal(nmu, nmu_o)
            np.testing.assert_allclose(pm, pm_o, rtol=1e-5, equal_nan=True)
        cm, _, _, cn = t2c.cross_power_spectrum_mu(x, y, los_axis=los, mubins=4, kbins=5, box_dims=box, return_n_modes=True, absolute_mus=False)
        cm_o, _, _, cn_o = ps_oracle.cross_power_spectrum_mu(x, y, los_axis=los, mubins=4, kbins=5, box_dims=box, return_n_modes=True,
                                                             absolute_mus=False)
        assert np.array_equal(cn, cn_o)
        np.testing.assert_allclose(cm, cm_o, rtol=1e-5, atol=1e-5 * np.nanmax(np.abs(pm_o)), equal_nan=True)


@pytest.mark.parametrize("shape", [(1024, 16, 32), (1024, 32, 16), (512, 16, 32)])
def test_warp_per_pencil_x_pass_overflow_pencils(t2c, shape, monkeypatch):
    """Fine binnings (the reference's default is kbins = 100): pencils next to the kx axis cross more than 32 ranges, are
    marked in the segment table and classified mode by mode by the same kernel.  In these thin slabs most pencils are
    such pencils, so the share a table may carry is lifted for the test."""
    from oracle import ps_oracle
    monkeypatch.setenv("T21_SEG_MAX_OVERFLOW", "1.0")
    x = (np.random.default_rng(5).standard_normal(shape) + 2.0).astype(np.float32)
    box = [300.0, 40.0, 60.0]
    ps, ks, nm = t2c.power_spectrum_1d(x, kbins=100, box_dims=box, return_n_modes=True)
    ps_o, ks_o, nm_o = ps_oracle.power_spectrum_1d(x, kbins=100, box_dims=box, return_n_modes=True)
    assert np.array_equal(nm, nm_o) and np.array_equal(ks, ks_o)
    np.testing.assert_allclose(ps, ps_o, rtol=1e-5, equal_nan=True)
    xs, _, xn = t2c.cross_power_spectrum_1d(x, x[::-1].copy(), kbins=100, box_dims=box, return_n_modes=True)
    xs_o, _, xn_o = ps_oracle.cross_power_spectrum_1d(x, x[::-1].copy(), kbins=100, box_dims=box, return_n_modes=True)
    assert np.array_equal(xn, xn_o)
    np.testing.assert_allclose(xs, xs_o, rtol=1e-5, atol=1e-5 * np.nanmax(np.abs(ps_o)), equal_nan=True)
    for absmu in (True, False):
        kw = dict(los_axis=0, mubins=3, kbins=40, box_dims=box, return_n_modes=True, absolute_mus=absmu)
        pm, _, _, nmu = t2c.power_spectrum_mu(x, **kw)
        pm_o, _, _, nmu_o = ps_oracle.power_spectrum_mu(x, **kw)
        assert np.array_equal(nmu, nmu_o)
        np.testing.assert_allclose(pm, pm_o, rtol=1e-5, equal_nan=True)


@pytest.mark.parametrize("shape", [(1024, 16, 32), (1024, 32, 16)])
@pytest.mark.parametrize("los", [0, 1, 2])
def test_fused_multipoles_on_the_warp_per_pencil_pass(t2c, shape, los):
    """fp32 cubes with nx = 1024 take the fused route (t21_ps_multipoles: the five sums per k bin are added by the X pass
    from its registers, bins from the pencil segment table, the excluded k_perp = 0 line dropped by the table): against
    the oracle, with and without the line, and with explicit edges whose first bin holds DC (mu = NaN there)."""
    from oracle import ps_oracle
    x = (np.random.default_rng(11 + los).standard_normal(shape) + 1.5).astype(np.float32)
    box = [300.0, 40.0, 60.0]
    for kb, excl in ((7, True), (7, False), (np.array([0.0, 0.2, 0.5, 1.0, 2.5]), False)):
        got, kc = t2c.power_spectrum_multipoles(x, kbins=kb, box_dims=box, los_axis=los, exclude_zero_modes=excl)
        with np.errstate(all="ignore"):
            want, kc_o = ps_oracle.power_spectrum_multipoles(x, kbins=kb, box_dims=box, los_axis=los, exclude_zero_modes=excl)
        assert np.array_equal(kc, kc_o)
        assert np.array_equal(got["n_modes"], want["n_modes"]) and np.array_equal(got["nmodes"], want["nmodes"])
        np.testing.assert_allclose(got["P0"], want["P0"], rtol=1e-5, equal_nan=True)
        atol = 1e-5 * float(np.nanmax(np.abs(want["P0"])))
        np.testing.assert_allclose(got["P2"], want["P2"], rtol=1e-4, atol=atol, equal_nan=True)
        np.testing.assert_allclose(got["P4"], want["P4"], rtol=1e-4, atol=atol, equal_nan=True)


@pytest.mark.parametrize("shape", [(1024, 16, 32), (1024, 32, 16)])
@pytest.mark.parametrize("nu_axis", [0, 1, 2])
def test_fused_power_spectrum_2d_on_the_warp_per_pencil_pass(t2c, shape, nu_axis):
    """fp32 cubes with nx = 1024: the (k_perp, k_par) bins are index ranges along x pencils too, so power_spectrum_2d runs
    as t21_ps_binned with a table-only spec (t21_build_pencil_segments_cyl).  Against the oracle for every nu_axis; the
    mode counts come from the kernel's closed-form range counts and must equal the reference's."""
    from oracle import ps_oracle
    x = (np.random.default_rng(23 + nu_axis).standard_normal(shape) + 0.5).astype(np.float32)
    box = [300.0, 40.0, 60.0]
    for binning in ("log", "linear"):
        got = t2c.power_spectrum_2d(x, kbins=8, binning=binning, box_dims=box, return_modes=True, nu_axis=nu_axis)
        with np.errstate(all="ignore"):
            want = ps_oracle.power_spectrum_2d(x, kbins=8, binning=binning, box_dims=box, return_modes=True, nu_axis=nu_axis)
        assert np.array_equal(got[3], want[3]), "n_modes differ"
        assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2])
        np.testing.assert_allclose(got[0], want[0], rtol=1e-5, equal_nan=True)


@pytest.mark.parametrize("n,ncuts,dtype", [(64, 4, "float32"), (128, 2, "float32"), (64, 2, "float64"), (48, 3, "float64")])
def test_integrated_bispectrum_against_oracle(t2c, n, ncuts, dtype):
    """The batched windowed estimator (mask fused into the Z-pass load) against the reference's loop over masked full
    cubes; 48^3 takes the any-size route (masked cubes formed on the device)."""
    from oracle import ps_oracle
    from tools21cm_b200 import synthetic
    a = synthetic.toy_dt_cube((n,) * 3, 5, np.dtype(dtype))
    b = synthetic.gaussian_cube((n,) * 3, 6, np.dtype(dtype)) + 2.0
    tol = 2e-5 if dtype == "float32" else 1e-10
    for norm in (False, True):
        got = t2c.integrated_bispectrum_cross(a, b, Ncuts=ncuts, kbins=7, box_dims=150.0, normalize=norm, verbose=False)
        want = ps_oracle.integrated_bispectrum_cross(a.astype(np.float64), b.astype(np.float64), Ncuts=ncuts, kbins=7,
                                                     box_dims=150.0, normalize=norm, verbose=False)
        assert np.array_equal(got[1], want[1])
        np.testing.assert_allclose(got[0], want[0], rtol=tol, equal_nan=True)


@pytest.mark.parametrize("shape", [(64, 64), (45, 64), (100,)])
def test_low_dimensional_inputs_against_oracle(t2c, shape):
    from oracle import ps_oracle
    x = np.random.default_rng(len(shape)).standard_normal(shape)
    box = [300.0] * len(shape)
    nd = t2c.power_spectrum_nd(x, box_dims=box)
    nd_o = ps_oracle.power_spectrum_nd(x, box_dims=box)
    assert nd.shape == nd_o.shape and np.abs(nd - nd_o).max() <= 1e-11 * nd_o.max()
    ps, ks, nm = t2c.power_spectrum_1d(x, kbins=5, box_dims=box, return_n_modes=True)
    ps_o, ks_o, nm_o = ps_oracle.power_spectrum_1d(x, kbins=5, box_dims=box, return_n_modes=True)
    assert np.array_equal(nm, nm_o) and np.array_equal(ks, ks_o)
    np.testing.assert_allclose(ps, ps_o, rtol=1e-10, equal_nan=True)


def test_torch_cuda_input_and_determinism(t2c):
    import torch
    from tools21cm_b200 import synthetic
    x = synthetic.gaussian_cube_torch((128, 128, 128), seed=3)
    a = t2c.power_spectrum_1d(x, kbins=15, box_dims=200.0, return_n_modes=True)
    b = t2c.power_spectrum_1d(x.cpu().numpy(), kbins=15, box_dims=200.0, return_n_modes=True)
    assert np.array_equal(a[2], b[2])
    np.testing.assert_allclose(a[0], b[0], rtol=1e-12)
    nd = t2c.power_spectrum_nd(x, box_dims=200.0)
    assert isinstance(nd, torch.Tensor) and nd.is_cuda and nd.shape == x.shape


def test_reference_slope_tests(t2c):
    """The reference's own five tests (tests/test_PowerSpectrum.py:10-48), against the drop-in."""
    gauss = np.random.default_rng(0).normal(loc=0., scale=1., size=(128, 128, 128))
    kbins, mubins, box_dims = 10, 2, 200

    def slope(pp, kk):
        d = np.log10(pp * kk ** 3 / 2 / np.pi ** 2)
        return (d[kbins - 3] - d[3]) / (np.log10(kk)[kbins - 3] - np.log10(kk)[3])

    pp, kk = t2c.cross_power_spectrum_1d(gauss, gauss, kbins=kbins, box_dims=box_dims)
    assert abs(slope(pp, kk) - 3) <= 0.1
    pp, mm, kk = t2c.cross_power_spectrum_mu(gauss, gauss, kbins=kbins, mubins=mubins, box_dims=box_dims)
    assert abs(slope(pp[0, :], kk) - 3) <= 0.1
    pp, kk = t2c.power_spectrum_1d(gauss, kbins=kbins, box_dims=box_dims)
    assert abs(slope(pp, kk) - 3) <= 0.1
    dd, kk = t2c.dimensionless_ps(gauss, kbins=kbins, box_dims=box_dims)
    s = (np.log10(dd)[kbins - 3] - np.log10(dd)[3]) / (np.log10(kk)[kbins - 3] - np.log10(kk)[3])
    assert abs(s - 3) <= 0.25
    pp, mm, kk = t2c.power_spectrum_mu(gauss, kbins=kbins, mubins=mubins, box_dims=box_dims)
    assert abs(slope(pp[0, :], kk) - 3) <= 0.1


@pytest.mark.parametrize("n", [512])
def test_large_cube_properties(t2c, n):
    """Size-independent checks at sizes the oracle cannot reach quickly: Parseval through the n-D
    output, mode-count totals, linearity in amplitude, and agreement between n-D + radial_average
    and the fused estimator."""
    import torch
    from tools21cm_b200 import synthetic
    L = 300.0
    x = synthetic.toy_dt_cube_torch((n, n, n), seed=11)
    ps, ks, nm = t2c.power_spectrum_1d(x, kbins=15, box_dims=L, return_n_modes=True)
    assert nm.sum() in (n ** 3 - 1, n ** 3 - 2, n ** 3 - 7, n ** 3 - 8)  # DC, +- corner, +- 6 fundamentals
    ps2, _, nm2 = t2c.power_spectrum_1d(2.0 * x, kbins=15, box_dims=L, return_n_modes=True)
    assert np.array_equal(nm, nm2)
    np.testing.assert_allclose(ps2, 4.0 * ps, rtol=2e-6)
    nd = t2c.power_spectrum_nd(x, box_dims=L, out_dtype=torch.float64)
    scale = (L ** 3 / n ** 3) ** 2 / L ** 3
    parseval = float(nd.sum().item()) / scale / n ** 3
    direct = float((x.double() ** 2).sum().item())
    assert abs(parseval - direct) <= 1e-5 * direct
    ps3, ks3, nm3 = t2c.radial_average(nd, [L, L, L], kbins=15)
    assert np.array_equal(nm3, nm)
    np.testing.assert_allclose(ps3, ps, rtol=1e-5)


def test_window_kbins_rules_and_side_effects(t2c):
    from oracle import ps_oracle
    rng = np.random.default_rng(12)
    cube = rng.standard_normal((32, 32, 32))
    a, b = cube.copy(), cube.copy()
    got = t2c.power_spectrum_1d(a, kbins=6, box_dims=80.0, window="tukey", return_n_modes=True)
    want = ps_oracle.power_spectrum_1d(b, kbins=6, box_dims=80.0, window="tukey", return_n_modes=True)
    assert np.array_equal(a, b) and not np.array_equal(a, cube)         # the caller's array is tapered in place
    assert np.array_equal(got[2], want[2])
    np.testing.assert_allclose(got[0], want[0], rtol=1e-11)
    with pytest.raises(TypeError):                                         # a list of edges fails in the reference too
        t2c.power_spectrum_1d(cube, kbins=[0.1, 0.2, 0.4], box_dims=80.0)
    with pytest.raises(ValueError):
        t2c.power_spectrum_1d(cube, kbins=np.array([0.4, 0.2, 0.8]), box_dims=80.0)
    # mu_binning zeroes the centre sample of its input, like the reference (power_spectrum.py:414)
    p = np.abs(rng.standard_normal((16, 16, 16))) + 1.0
    t2c.mu_binning(p, los_axis=0, mubins=3, kbins=4, box_dims=[50.0] * 3)
    assert p[8, 8, 8] == 0.0
    # many bins: the per-warp / shared / global histogram variants all agree with the oracle
    for nb in (100, 1000, 3000):
        g = t2c.power_spectrum_1d(cube, kbins=nb, box_dims=80.0, return_n_modes=True)
        w = ps_oracle.power_spectrum_1d(cube, kbins=nb, box_dims=80.0, return_n_modes=True)
        assert np.array_equal(g[2], w[2])
        np.testing.assert_allclose(g[0], w[0], rtol=1e-11, equal_nan=True)
    g = t2c.power_spectrum_mu(cube, kbins=60, mubins=50, box_dims=80.0, return_n_modes=True, absolute_mus=False)
    w = ps_oracle.power_spectrum_mu(cube, kbins=60, mubins=50, box_dims=80.0, return_n_modes=True, absolute_mus=False)
    assert np.array_equal(g[3], w[3])
    np.testing.assert_allclose(g[0], w[0], rtol=1e-11, equal_nan=True)


def test_set_precision_and_run_to_run_reproducibility(t2c):
    from tools21cm_b200 import synthetic
    x = synthetic.toy_dt_cube((64, 64, 64), seed=2)
    a = t2c.power_spectrum_1d(x, kbins=15, box_dims=100.0)
    b = t2c.power_spectrum_1d(x, kbins=15, box_dims=100.0)
    assert np.array_equal(a[0], b[0], equal_nan=True), "small histograms are summed in a fixed order: results must be bit-identical"
    t2c.set_precision("float64")
    try:
        c = t2c.power_spectrum_1d(x, kbins=15, box_dims=100.0)
    finally:
        t2c.set_precision(None)
    np.testing.assert_allclose(a[0], c[0], rtol=1e-5, equal_nan=True)
    assert not np.array_equal(a[0], c[0], equal_nan=True)


def test_cross_of_a_cube_with_itself_is_its_auto_spectrum_512(t2c):
    from tools21cm_b200 import synthetic
    x = synthetic.toy_dt_cube_torch((512, 512, 512), seed=5)
    a = t2c.power_spectrum_1d(x, kbins=15, box_dims=300.0, return_n_modes=True)
    c = t2c.cross_power_spectrum_1d(x, x.clone(), kbins=15, box_dims=300.0, return_n_modes=True)
    assert np.array_equal(a[2], c[2])
    # the per-mode powers are the same fp32 numbers; the two kernels run with different numbers of CTAs, which
    # changes where a warp's partial sums are cut (fp32 partial sums of 16 modes, then float64): ~1e-11
    np.testing.assert_allclose(a[0], c[0], rtol=1e-9)
    m = t2c.power_spectrum_mu(x, kbins=10, mubins=5, box_dims=300.0, return_n_modes=True)
    assert m[3].sum() == 512 ** 3 - 512 - 1 - 0 or m[3].sum() > 0.99 * 512 ** 3


def test_2048_cube_on_one_gpu_has_no_index_overflow(t2c):
    """34 GB cube + 35 GB workspace on one 180 GB GPU: element offsets exceed 2^32 in every pass."""
    import torch
    free, _ = torch.cuda.mem_get_info()
    if free < 90 * 2 ** 30:
        pytest.skip("needs ~80 GB of free device memory")
    n, L = 2048, 714.0
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.randn((n, n, n), generator=g, device="cuda", dtype=torch.float32)
    ps, ks, nm = t2c.power_spectrum_1d(x, kbins=15, box_dims=L, return_n_modes=True)
    assert nm.sum() in (n ** 3 - 1, n ** 3 - 2, n ** 3 - 7, n ** 3 - 8)
    # white noise of unit variance: P(k) = V / N^3 in every bin, to sampling error
    expect = L ** 3 / n ** 3
    assert np.all(np.abs(ps[nm > 1000] / expect - 1.0) < 0.05)
    assert abs(ps[-1] / expect - 1.0) < 1e-3          # billions of modes in the last shells
    del x
    torch.cuda.empty_cache()
