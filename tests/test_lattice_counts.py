"""The exact lattice mode counter (tests/lattice.py) against the reference's golden n_modes and the oracle, so
that it can stand in for the oracle at 1024^3 / 2048^3 in the GPU tests (tests/test_gpu_fullsize.py)."""
import numpy as np
import pytest

import lattice
from cases import CASE_BY_ID
from conftest import golden_outputs
from oracle import ps_oracle

GOLDEN_1D = ["ps1d_128_grf_L200_k15", "ps1d_128_grf_L100_k15", "ps1d_64_grf_None_k10", "ps1d_64_grf_37h_k100",
             "ps1d_64_grf_lin", "ps1d_33_odd", "ps1d_noncubic", "ps1d_noncubic_aniso", "ps1d_256_grf_L200", "ps1d_16_grf"]


@pytest.mark.parametrize("cid", GOLDEN_1D)
def test_lattice_counts_equal_the_reference_golden_counts(cid, golden):
    case = CASE_BY_ID[cid]
    kw = case["kw"]
    shape = case["shape"]
    box = kw["box_dims"]
    box = [244.0 / 0.7] * 3 if box is None else ([box] * 3 if not hasattr(box, "__iter__") else list(box))
    edges = lattice.reference_edges(shape, box, kw["kbins"], kw.get("binning", "log"))
    want = golden_outputs(golden, cid)
    got = lattice.shell_counts(shape, box, edges)
    assert got.dtype == want[2].dtype
    assert np.array_equal(got, want[2])
    # the centres the reference returns are the mid-points of exactly these edges
    assert np.array_equal(edges[:-1] + (edges[1:] - edges[:-1]) / 2.0, want[1])


@pytest.mark.parametrize("shape,box", [((48, 48, 48), [37 / 0.7] * 3), ((50, 50, 50), [100.0] * 3), ((21, 30, 64), [60.0, 80.0, 100.0])])
def test_lattice_counts_equal_the_oracle_on_adversarial_boxes(shape, box):
    """first / last edges land +-1 ulp from kmin / kmax depending on (N, L) (SURVEY.md 8a edge table)"""
    x = np.zeros(shape)
    for nb, binning in ((7, "log"), (15, "linear")):
        _, _, nm = ps_oracle.power_spectrum_1d(x, kbins=nb, box_dims=box, return_n_modes=True, binning=binning)
        edges = lattice.reference_edges(shape, box, nb, binning)
        assert np.array_equal(lattice.shell_counts(shape, box, edges), nm)


def test_headline_totals():
    """1024^3, box 500/0.7, 15 log bins (bench.py's workload): the first edge sits one ulp above kmin and the last one
    ulp below kmax, so the 6 fundamental modes, the corner and DC are in no bin: N^3 - 8."""
    shape, box = (1024,) * 3, [500 / 0.7] * 3
    c = lattice.shell_counts(shape, box, lattice.reference_edges(shape, box, 15))
    assert c.sum() == 1024 ** 3 - 8
