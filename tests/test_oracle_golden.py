"""Pin the CPU oracle (oracle/ps_oracle.py) against outputs of the reference itself
(tests/golden/golden.npz, produced by tests/golden/make_golden.py in the build
container).  Mode counts must be identical; powers agree to 1e-13 relative (the
oracle calls the same scipy.fftpack / numpy.histogram routines, so in practice
they are bit-identical)."""
import numpy as np
import pytest

from cases import CASES, build_inputs, build_kwargs
from conftest import golden_outputs
from oracle import ps_oracle

BINNED = ("power_spectrum_1d", "cross_power_spectrum_1d", "power_spectrum_mu", "cross_power_spectrum_mu")


@pytest.mark.parametrize("case", CASES, ids=[c["id"] for c in CASES])
def test_oracle_matches_reference_golden(case, golden):
    fn = getattr(ps_oracle, case["fn"])
    kw = build_kwargs(case)
    if case["fn"] in BINNED:
        kw["return_n_modes"] = True
    res = fn(*build_inputs(case), **kw)
    if not isinstance(res, tuple):
        res = (res,)
    if case["fn"] == "power_spectrum_multipoles":      # (dict, k) -> P0, P2, P4, nmodes, k as make_golden.py stores them
        res = (res[0]["P0"], res[0]["P2"], res[0]["P4"], res[0]["nmodes"], res[1])
    want = golden_outputs(golden, case["id"])
    assert len(res) == len(want)
    for got, ref in zip(res, want):
        got = np.asarray(got)
        assert got.shape == ref.shape
        assert got.dtype == ref.dtype
    if case["fn"] in BINNED or case["fn"] in ("radial_average", "mu_binning"):
        assert np.array_equal(res[-1], want[-1]), "n_modes differ"
    for got, ref in zip(res, want):
        np.testing.assert_allclose(got, ref, rtol=1e-13, atol=0, equal_nan=True)
