"""Host-side contract of the drop-in API that needs no GPU: signatures identical to the reference,
argument validation that happens before any device work, and the numpy helpers siblings import."""
import inspect

import numpy as np
import pytest

import tools21cm_b200 as t2c
from oracle import ps_oracle

# (name, positional parameters with defaults) exactly as in the reference power_spectrum.py
REFERENCE_SIGNATURES = {
    "power_spectrum_nd": "(input_array, box_dims=None, verbose=False, **kwargs)",
    "cross_power_spectrum_nd": "(input_array1, input_array2, box_dims, **kwargs)",
    "radial_average": "(input_array, box_dims, kbins=10, binning='log', breakpoint=0.1, **kwargs)",
    "power_spectrum_1d": "(input_array_nd, kbins=100, box_dims=None, return_n_modes=False, binning='log', "
                         "breakpoint=0.1, window=None, **kwargs)",
    "cross_power_spectrum_1d": "(input_array1_nd, input_array2_nd, kbins=100, box_dims=None, return_n_modes=False, "
                               "binning='log', breakpoint=0.1, **kwargs)",
    "power_spectrum_mu": "(input_array, los_axis=0, mubins=20, kbins=10, box_dims=None, weights=None, "
                         "exclude_zero_modes=True, return_n_modes=False, absolute_mus=True, **kwargs)",
    "cross_power_spectrum_mu": "(input_array1, input_array2, los_axis=0, mubins=20, kbins=10, box_dims=None, "
                               "weights=None, exclude_zero_modes=True, return_n_modes=False, absolute_mus=True, **kwargs)",
    "mu_binning": "(powerspectrum, los_axis=0, mubins=20, kbins=10, box_dims=None, weights=None, "
                  "exclude_zero_modes=True, binning='log', absolute_mus=True)",
    "dimensionless_ps": "(data, kbins=100, box_dims=None, binning='log', factor=10)",
    "power_spectrum_2d": "(input_array, kbins=10, binning='log', box_dims=348.5714285714286, return_modes=False, "
                         "nu_axis=2, window=None, **kwargs)",
    # position_dependent_power_spectrum.py:5,25,45
    "power_spectrum_response": "(cube, cube2=None, Ncuts=4, kbins=15, box_dims=348.5714285714286, binning='log', verbose=True)",
    "integrated_bispectrum": "(cube, Ncuts=4, kbins=15, box_dims=348.5714285714286, binning='log', normalize=False, verbose=True)",
    "integrated_bispectrum_cross": "(cube1, cube2, Ncuts=4, kbins=15, box_dims=348.5714285714286, binning='log', "
                                   "normalize=False, verbose=True)",
    # power_legendre.py:4-5
    "power_spectrum_multipoles": "(input_array, kbins=10, box_dims=None, los_axis=0, output=['P0', 'P2', 'P4', 'nmodes'], "
                                 "exclude_zero_modes=False)",
}


@pytest.mark.parametrize("name", sorted(REFERENCE_SIGNATURES))
def test_signatures_match_the_reference(name):
    assert str(inspect.signature(getattr(t2c, name))) == REFERENCE_SIGNATURES[name]
    # and the oracle restates the same interface
    assert str(inspect.signature(getattr(ps_oracle, name))) == REFERENCE_SIGNATURES[name]


def test_weights_fail_like_the_reference_before_any_device_work():
    cube = np.zeros((8, 8, 8))
    with pytest.raises(ValueError):
        t2c.power_spectrum_mu(cube, weights=np.ones((8, 8, 8)), box_dims=10.0)
    with pytest.raises(TypeError):
        t2c.power_spectrum_mu(cube, weights=2.0, box_dims=10.0)
    with pytest.raises(ValueError):
        t2c.mu_binning(cube, weights=np.ones((8, 8, 8)), box_dims=[10.0] * 3)


def test_host_helpers_match_the_oracle_grids():
    shape, box = (6, 5, 4), [30.0, 20.0, 10.0]
    kc, k = t2c._get_k(np.zeros(shape), box)
    kc_o, k_o = ps_oracle.k_magnitude(shape, box)
    assert np.array_equal(k, k_o)
    for d in range(3):
        assert np.array_equal(kc[d][tuple(slice(None) if a == d else 0 for a in range(3))], kc_o[d])
    assert t2c._get_dims(None, shape) == [244.0 / 0.7] * 3
    assert t2c._get_dims(7.0, shape) == [7.0] * 3
    e = t2c._get_kbins(5, box, k)
    assert np.array_equal(e, ps_oracle.k_edges(5, box, k.max()))
    mu = t2c._get_mu(kc, k, 1, True)
    assert np.isnan(mu[3, 2, 2]) and np.nanmax(mu) <= 1.0
    good = t2c._get_nonzero_idx((4, 4, 4), 0)
    assert good.sum() == 64 - 4 and not good[:, 2, 2].any()
    assert np.array_equal(t2c.get_kbins(5, box, array=np.zeros(shape)), e)


def test_apply_window_mutates_like_the_reference():
    a = np.ones((2, 2, 8))
    b = a.copy()
    out = t2c.apply_window(a, "blackmanharris")
    ps_oracle.apply_window(b, "blackmanharris")
    assert out is a and np.array_equal(a, b) and not np.allclose(a, 1.0)
    with pytest.raises(AttributeError):          # an ndarray window fails at window.lower() in the reference too
        t2c.apply_window(a, np.ones(8))
