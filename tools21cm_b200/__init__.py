"""tools21cm_b200 -- the tools21cm power-spectrum estimator on B200 (sm_100a).

Swap ``import tools21cm as t2c`` for ``import tools21cm_b200 as t2c`` on the power-spectrum
path: the functions below keep the reference's signatures and return conventions
(src/tools21cm/power_spectrum.py, exported through ``from .power_spectrum import *`` at
src/tools21cm/__init__.py:37).  Importing this package never touches CUDA; the first
estimator call loads tools21cm_b200/lib/libt21ps.so and fails loudly if it, or a CUDA
device, is missing.
"""
from .power_spectrum import (  # noqa: F401
    power_spectrum_nd, cross_power_spectrum_nd, radial_average, power_spectrum_1d, cross_power_spectrum_1d,
    power_spectrum_mu, cross_power_spectrum_mu, mu_binning, power_spectrum_2d, dimensionless_ps, anisotropy_ratio_r_mu,
    get_k, get_kbins, apply_window, set_precision,
    _get_k, _get_mu, _get_kbins, _get_dims, _get_nonzero_idx,
)

from .power_legendre import power_spectrum_multipoles  # noqa: F401  (src/tools21cm/__init__.py:38)
from .position_dependent_power_spectrum import (  # noqa: F401  (src/tools21cm/__init__.py: from .position_dependent_power_spectrum import *)
    power_spectrum_response, integrated_bispectrum, integrated_bispectrum_cross,
)

from .inputs import XfracFile, DensityFile, calc_dt, mean_dt, subtract_mean_signal  # noqa: F401

__version__ = "0.2.0"
