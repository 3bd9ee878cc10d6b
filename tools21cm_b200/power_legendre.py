"""Drop-in for ``tools21cm/power_legendre.py``: the power spectrum expanded in the first three even Legendre
polynomials of mu = k_los / k (reference ``power_legendre.py:4-90``).

The reference function cannot run as shipped: it calls ``_get_mu`` with three arguments (``:47``; the function needs
four, ``power_spectrum.py:481``), and with ``exclude_zero_modes=False`` it multiplies a boolean mask in place by a
float array (``:51,79``: numpy refuses the cast).  What it computes once those two lines run is unambiguous, and that is
what is done here: per k bin (half-open, ``:78``)

    P_l = sum(P * L_l) / sum(L_l ** 2),   L_0 = 1,  L_2 = (3 mu^2 - 1)/2,  L_4 = 4.375 (mu^2 - 0.115587)(mu^2 - 0.741556)

(``:55-61,81-86``; only even powers of mu, so the sign convention of mu does not matter).  ``multipoles['nmodes']``
keeps the reference's value -- ``len(np.nonzero(idx))`` (``:80``) is the number of array dimensions, 3.0 in every bin,
not a mode count -- and the actual number of modes per bin is added under ``'n_modes'``.

On the device: the Z/Y/X passes write the power of the computed half spectrum in the transform's precision
(``t21_ps_half``) and one streaming kernel accumulates the five sums per bin (``t21_bin_multipoles_half``); the
fftshifted float64 cube, the k / mu grids and the per-bin boolean masks of the reference are never built.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from . import binspec as _bs
from . import power_spectrum as _ps

__all__ = ["power_spectrum_multipoles"]


def power_spectrum_multipoles(input_array, kbins=10, box_dims=None, los_axis=0, output=['P0', 'P2', 'P4', 'nmodes'],
                              exclude_zero_modes=False):
    """Returns ``(multipoles, k)``: a dict with the requested ``'P0'``, ``'P2'``, ``'P4'`` arrays (plus ``'nmodes'`` and
    ``'n_modes'``, see the module docstring) and the mid-points of the k bins."""
    assert (los_axis >= 0 and los_axis <= len(input_array.shape))
    box_dims = _ps._get_dims(box_dims, input_array.shape)
    x, _ = _ps._stage(input_array)
    torch = _ps._torch()
    shape = tuple(x.shape)
    if len(shape) != 3 or los_axis > 2:
        raise ValueError("power_spectrum_multipoles needs a 3-D array and a line of sight along one of its axes")
    dev = _ps._device(x.device.index)
    # the excluded line only exists when both perpendicular axes are even (power_spectrum.py:567-573 compares an
    # integer index with N/2 as a float)
    perp = [a for a in (0, 1, 2) if a != los_axis]
    excl = bool(exclude_zero_modes) and all(shape[a] % 2 == 0 for a in perp)
    key = ("multipoles", _ps._freeze(list(box_dims)), _ps._freeze(kbins), int(los_axis), excl)

    def make_spec():
        sp = _bs.build_binspec(shape, box_dims, kbins, binning="log", los_axis=los_axis, last_bin_closed=False,
                               shells_with_los=True)
        sp.exclude_zero = excl        # shells otherwise ignore both; the segment table of the fused pass drops the
        sp.los_axis = int(los_axis)   # k_perp = 0 line of this axis
        return sp

    code = _ps._dtype_code(x)
    fused = code == _lib.T21_F32 and shape[0] == 1024
    spec, cspec, _owner = dev.spec((shape, key), make_spec, segments=fused)
    nk = spec.nk
    scale = _bs.volume_scale(shape, box_dims)
    with torch.cuda.device(x.device):
        lib = _lib.load()
        sums = torch.empty(nk * 5, dtype=torch.float64, device=x.device)
        counts = torch.empty(nk, dtype=torch.int64, device=x.device)
        if fused and cspec.seg_tab and nk <= 64:
            # one pipeline: Z / Y passes, then the warp-per-pencil X pass adds the five sums per bin from its registers
            plan = dev.plan(shape, code)
            ws = dev.workspace(lib.t21_workspace_bytes(plan, 1, 5 * nk))
            rc = lib.t21_ps_multipoles(plan, x.data_ptr(), 1, C.byref(cspec), int(los_axis), sums.data_ptr(), counts.data_ptr(),
                                       ws.data_ptr(), ws.numel(), _ps._stream_ptr(torch, x.device.index))
            _lib.check(rc, "t21_ps_multipoles")
            s = sums.cpu().numpy().reshape(nk, 5)
            s[:, :3] *= scale                      # the three sums that carry the power
        else:
            power = _ps._fused_half([x], scale, x.dtype)
            ws = dev.workspace(lib.t21_bin_multipoles_workspace_bytes(nk))
            rc = lib.t21_bin_multipoles_half(power.data_ptr(), _ps._dtype_code(power), int(power.shape[2]), C.byref(cspec), int(los_axis), int(excl),
                                             sums.data_ptr(), counts.data_ptr(), ws.data_ptr(), ws.numel(),
                                             _ps._stream_ptr(torch, x.device.index))
            _lib.check(rc, "t21_bin_multipoles_half")
            s = sums.cpu().numpy().reshape(nk, 5)
        n = counts.cpu().numpy().astype("float")
    multipoles = {}
    with np.errstate(all="ignore"):
        if 'P0' in output:
            multipoles['P0'] = s[:, 0] / n
        if 'P2' in output:
            multipoles['P2'] = s[:, 1] / s[:, 3]
        if 'P4' in output:
            multipoles['P4'] = s[:, 2] / s[:, 4]
    multipoles['nmodes'] = np.full(nk, float(len(shape)))      # the reference's len(np.nonzero(idx)) (:80)
    multipoles['n_modes'] = n
    return multipoles, spec.k_centres()
