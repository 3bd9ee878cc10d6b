"""Drop-in for ``tools21cm/position_dependent_power_spectrum.py``: the integrated bispectrum / power-spectrum response
of Giri et al. 2019 (reference ``position_dependent_power_spectrum.py:5-88``).

The reference loops over the ``Ncuts**3`` sub-boxes and, for each, multiplies the FULL cube by a cubical top-hat
(``_W_L`` ``:90-103``) and takes ``power_spectrum_1d`` of that full-size masked cube (``:65-77``): ``Ncuts**3`` complete
FFT + binning passes plus three full-size temporaries each.  Here the whole loop is ONE call into the library
(``t21_ps_binned_windows``): the top-hat is fused into the first FFT pass's load -- no masked cube is ever formed, only
the sub-box's samples are read -- the launches of all sub-boxes are queued back to back, and the sub-box means of the
second cube are one blocked reduction on the device.  Same signatures, defaults and return tuples as the reference.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from . import binspec as _bs
from . import power_spectrum as _ps

__all__ = ["power_spectrum_response", "integrated_bispectrum", "integrated_bispectrum_cross"]


def power_spectrum_response(cube, cube2=None, Ncuts=4, kbins=15, box_dims=244 / .7, binning='log', verbose=True):
    """Response of the small-scale power to large-scale fluctuations (reference :5-22)."""
    return integrated_bispectrum_cross(cube, cube if cube2 is None else cube2, Ncuts=Ncuts, kbins=kbins,
                                       box_dims=box_dims, binning=binning, normalize=True, verbose=verbose)


def integrated_bispectrum(cube, Ncuts=4, kbins=15, box_dims=244 / .7, binning='log', normalize=False, verbose=True):
    """Integrated bispectrum of one field (reference :25-43)."""
    return integrated_bispectrum_cross(cube, cube, Ncuts=Ncuts, kbins=kbins, box_dims=box_dims, binning=binning,
                                       normalize=normalize, verbose=verbose)


def _W_L(array, rL, L):
    """Cubical heaviside filter (reference :90-103), on the host: kept for callers that import it."""
    assert array.ndim == np.array(rL).size
    out = np.zeros(array.shape)
    if np.array(L).size == 1:
        L = [L for i in range(array.ndim)]
    xl = [int(rL[0] - L[0] / 2), int(rL[0] + L[0] / 2)]
    yl = [int(rL[1] - L[1] / 2), int(rL[1] + L[1] / 2)]
    zl = [int(rL[2] - L[2] / 2), int(rL[2] + L[2] / 2)]
    out[xl[0]:xl[1], yl[0]:yl[1], zl[0]:zl[1]] = 1
    return out


def _window_spectra(x, windows, kbins, box_dims, binning):
    """power_spectrum_1d of x * W for every window [x0, x1, y0, y1, z0, z1]: (pk [nwin, nk], k centres)."""
    torch = _ps._torch()
    shape = tuple(x.shape)
    box = _ps._get_dims(box_dims, shape)
    scale = _bs.volume_scale(shape, box)
    pow2 = all(n >= 8 and n <= 4096 and (n & (n - 1)) == 0 for n in shape)
    if not pow2:
        # other sizes: the same estimator, one masked cube at a time (formed on the device)
        pks, ks = [], None
        for w in windows:
            m = torch.zeros_like(x)
            m[w[0]:w[1], w[2]:w[3], w[4]:w[5]] = x[w[0]:w[1], w[2]:w[3], w[4]:w[5]]
            pk, ks = _ps.power_spectrum_1d(m, kbins=kbins, box_dims=box_dims, binning=binning)
            pks.append(pk)
        return np.array(pks), ks
    dev = _ps._device(x.device.index)
    with torch.cuda.device(x.device):
        lib = _lib.load()
        code = _ps._dtype_code(x)
        plan = dev.plan(shape, code)
        key = ("1d", _ps._freeze(list(box)), _ps._freeze(kbins), binning, 0.1)
        spec, cspec, _owner = dev.spec((shape, key), lambda: _bs.build_binspec(shape, box, kbins, binning=binning), segments=True)
        nb, nw = spec.nbins, len(windows)
        ws = dev.workspace(lib.t21_workspace_bytes(plan, 1, nb))
        sums = torch.empty((nw, nb), dtype=torch.float64, device=x.device)
        counts = torch.empty(nb, dtype=torch.int64, device=x.device)
        warr = (C.c_int * (6 * nw))(*[int(v) for w in windows for v in w])
        rc = lib.t21_ps_binned_windows(plan, x.data_ptr(), warr, nw, C.byref(cspec), sums.data_ptr(), counts.data_ptr(),
                                       ws.data_ptr(), ws.numel(), _ps._stream_ptr(torch, x.device.index))
        _lib.check(rc, "t21_ps_binned_windows")
        s = sums.cpu().numpy()
        n = counts.cpu().numpy().astype("float")
    with np.errstate(all="ignore"):
        return s * scale / n[None, :], spec.k_centres()


def integrated_bispectrum_cross(cube1, cube2, Ncuts=4, kbins=15, box_dims=244 / .7, binning='log', normalize=False,
                                verbose=True):
    """Integrated bispectrum of cube1's local power with cube2's local mean (reference :45-88).
    Returns (B_k, k) or, with ``normalize``, (B_k / P_k / sigma^2, k)."""
    assert cube1.shape == cube2.shape
    assert cube1.shape[0] % Ncuts == 0 and cube1.shape[1] % Ncuts == 0 and cube1.shape[2] % Ncuts == 0
    Lx, Ly, Lz = cube1.shape[0] / Ncuts, cube1.shape[1] / Ncuts, cube1.shape[2] / Ncuts
    rLs = [[Lx / 2. + i * Lx, Ly / 2. + j * Ly, Lz / 2. + k * Lz] for i in range(Ncuts) for j in range(Ncuts) for k in range(Ncuts)]
    windows = [[int(r[0] - Lx / 2), int(r[0] + Lx / 2), int(r[1] - Ly / 2), int(r[1] + Ly / 2), int(r[2] - Lz / 2), int(r[2] + Lz / 2)]
               for r in rLs]                                                  # the slices _W_L sets to one (:98-102)
    n_box = Ncuts ** 3
    V_L = (Lx * Ly * Lz)
    x1, _ = _ps._stage(cube1)
    x2 = x1 if cube2 is cube1 else _ps._stage(cube2)[0]
    torch = _ps._torch()
    pks, ks = _window_spectra(x1, windows, kbins, box_dims, binning)
    # d_mean = (cube2 * W).sum() / V_L for every sub-box (:70): one blocked float64 reduction
    blocks = x2.reshape(Ncuts, int(Lx), Ncuts, int(Ly), Ncuts, int(Lz)).sum(dim=(1, 3, 5), dtype=torch.float64)
    d_mean = (blocks.reshape(-1) / V_L).cpu().numpy()
    B_k = np.zeros(pks.shape[1], dtype=np.float64)
    P_k = np.zeros(pks.shape[1], dtype=np.float64)
    sig2 = 0
    for i in range(n_box):                                                    # the reference's accumulation order (:71-73)
        B_k += pks[i] * d_mean[i]
        P_k += pks[i]
        sig2 += (d_mean[i]) ** 2
    B_k = B_k / n_box
    P_k = P_k / n_box
    sig2 = sig2 / n_box
    if verbose:
        print('The long wavelength mode is %.3f/cMpc' % (2 * np.pi / (box_dims / Ncuts)))
    if normalize:
        with np.errstate(all="ignore"):
            return B_k / P_k / sig2, ks
    return B_k, ks
