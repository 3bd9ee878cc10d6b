"""CPU checks of the host-side logic (tools21cm_b200.binspec) through a numpy model of the
device algorithm (tests/kernel_model.py): the half-spectrum + threshold + twin-weight scheme
must give the reference's mode counts bit for bit, and the fp32 lookup-table fast path must
never disagree with the exact float64 path where it claims to be unambiguous."""
import numpy as np
import pytest

from cases import CASES, build_inputs, build_kwargs
from conftest import golden_outputs
from kernel_model import model_binned
from tools21cm_b200 import binspec as bs

FUSED = [c for c in CASES if c["fn"] in ("power_spectrum_1d", "cross_power_spectrum_1d",
                                         "power_spectrum_mu", "cross_power_spectrum_mu")
         and max(c["shape"]) <= 128]


def spec_for(case, shape):
    kw = build_kwargs(case)
    box = kw.get("boxsize") if kw.get("boxsize") is not None else kw.get("box_dims")
    box = bs.get_dims(box, shape)
    if case["fn"].endswith("_mu"):
        return bs.build_binspec(shape, box, kw.get("kbins", 10), binning="log", mubins=kw.get("mubins", 20),
                                los_axis=kw.get("los_axis", 0), absolute_mus=kw.get("absolute_mus", True),
                                exclude_zero_modes=kw.get("exclude_zero_modes", True), last_bin_closed=False), box
    return bs.build_binspec(shape, box, kw.get("kbins", 100), binning=kw.get("binning", "log"),
                            breakpoint=kw.get("breakpoint", 0.1)), box


@pytest.mark.parametrize("case", FUSED, ids=[c["id"] for c in FUSED])
def test_model_counts_bit_exact_and_power_close(case, golden):
    cubes = build_inputs(case)
    spec, box = spec_for(case, cubes[0].shape)
    want = golden_outputs(golden, case["id"])
    sums, counts = model_binned(cubes, spec)
    scale = bs.volume_scale(cubes[0].shape, box)
    with np.errstate(all="ignore"):
        ps = sums * scale / counts
    if spec.nmu:
        ps = ps.reshape(spec.nmu, spec.nk)
        counts = counts.reshape(spec.nmu, spec.nk)
        np.testing.assert_array_equal(spec.mu_centres(), want[1])
        np.testing.assert_array_equal(spec.k_centres(), want[2])
    else:
        np.testing.assert_array_equal(spec.k_centres(), want[1])
    assert np.array_equal(counts.astype(float), want[-1])
    # with a large mean the reference's own np.histogram(weights=...) loses digits: it differences
    # cumulative sums that include the huge DC sample (numpy lib/_histograms_impl.py, non-uniform path)
    rtol = 2e-10 if case["field"] in ("grf", "grf|grf", "grf|same") else 2e-7
    np.testing.assert_allclose(ps, want[0], rtol=rtol, equal_nan=True)


@pytest.mark.parametrize("cid", ["ps1d_128_grf_mean20", "ps1d_128_toy_L200_k15", "psmu_64_los2_signed",
                                 "xps1d_64_pair"])
def test_model_fp32_with_mean_offset(cid, golden):
    case = [c for c in CASES if c["id"] == cid][0]
    cubes = build_inputs(case)
    spec, box = spec_for(case, cubes[0].shape)
    want = golden_outputs(golden, cid)
    offs = [float(np.float32(c.mean())) for c in cubes]
    sums, counts = model_binned(cubes, spec, offsets=offs, single=True, check_lut=False)
    with np.errstate(all="ignore"):
        ps = sums * bs.volume_scale(cubes[0].shape, box) / counts
    # a cross-spectrum bin is a signed sum with cancellation; its fp32 error scales with the
    # auto-power level, not with the (possibly near-zero) bin value
    atol = 1e-5 * np.nanmax(np.abs(want[0])) if len(cubes) == 2 else 0.0
    np.testing.assert_allclose(ps.reshape(want[0].shape), want[0], rtol=1e-5, atol=atol, equal_nan=True)


@pytest.mark.parametrize("n,L", [(128, 200), (256, 200), (512, 244 / 0.7), (1024, 500 / 0.7), (128, 100),
                                 (64, 37 / 0.7), (2048, 714.0)])
def test_edges_match_reference_ulp_table(n, L):
    """SURVEY.md 8a edge-ulp table: first/last log edge vs kmin/kmax."""
    kax = [bs.k_axis_shifted(n, L, 3)] * 3
    kmax = bs.grid_kmax(kax)
    # same value as the max over the full grid
    corner = np.sqrt(kax[0][0] ** 2 + kax[1][0] ** 2 + kax[2][0] ** 2)
    assert kmax == corner
    e = bs.make_k_edges(15, [L] * 3, kmax)
    assert abs(e[0] - 2 * np.pi / L) <= 2 * np.spacing(e[0])
    assert abs(e[-1] - kmax) <= 2 * np.spacing(kmax)


def test_sqrt_threshold_is_exact():
    rng = np.random.default_rng(0)
    for e in list(rng.random(200) * 10) + [0.0, 1e-3, 0.6, 3.0]:
        for strict in (False, True):
            t = bs.sqrt_threshold(e, strict)
            ok = (lambda s: np.sqrt(s) > e) if strict else (lambda s: np.sqrt(s) >= e)
            assert ok(np.float64(t))
            if t > 0:
                assert not ok(np.nextafter(np.float64(t), -np.inf))


def test_kbins_type_rules():
    with pytest.raises(TypeError):
        bs.build_binspec((16, 16, 16), [10.0] * 3, [0.1, 0.2, 0.3])
    with pytest.raises(ValueError):
        bs.build_binspec((16, 16, 16), [10.0] * 3, np.array([0.3, 0.2, 0.5]))
    spec = bs.build_binspec((16, 16, 16), [10.0] * 3, np.int64(4) * np.ones(3).cumsum())
    assert spec.nk == 2  # np.int64 scalars/arrays are edges, only a Python int asks for automatic bins


CYL = [c for c in CASES if c["fn"] == "power_spectrum_2d"]


@pytest.mark.parametrize("case", CYL, ids=[c["id"] for c in CYL])
def test_cylindrical_tables_reproduce_reference_counts(case, golden):
    """power_spectrum_2d: the separable (k_perp, k_par) bin tables, applied to the oracle's n-D power with
    plain numpy, must give the reference's counts exactly and its means to rounding."""
    from oracle import ps_oracle
    x = build_inputs(case)[0]
    kw = build_kwargs(case)
    box = bs.get_dims(kw.get("box_dims", 244 / 0.7), x.shape)
    nu = kw.get("nu_axis", 2)
    cyl = bs.build_cyl_tables(x.shape, box, kw["kbins"], binning=kw.get("binning", "log"), nu_axis=nu)
    power = ps_oracle.power_spectrum_nd(x, box)
    idx = np.indices(x.shape)
    perp = [a for a in (0, 1, 2) if a != nu]
    bp = cyl.perp_tab[idx[perp[0]], idx[perp[1]]]
    bz = cyl.par_tab[idx[nu]]
    ok = (bp >= 0) & (bz >= 0)
    flat = bp[ok] * cyl.npar + bz[ok]
    sums = np.zeros(cyl.nperp * cyl.npar)
    cnt = np.zeros(cyl.nperp * cyl.npar)
    np.add.at(sums, flat, power[ok])
    np.add.at(cnt, flat, 1)
    want = golden_outputs(golden, case["id"])
    assert np.array_equal(cnt.reshape(want[3].shape), want[3])
    assert np.array_equal(cyl.kper_mid, want[1]) and np.array_equal(cyl.kpar_mid, want[2])
    with np.errstate(all="ignore"):
        np.testing.assert_allclose((sums / cnt).reshape(want[0].shape), want[0], rtol=1e-12, equal_nan=True)
