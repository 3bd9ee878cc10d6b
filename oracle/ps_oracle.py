"""CPU oracle for the tools21cm power-spectrum path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy/scipy *restatement* of the algorithm in the reference
``src/tools21cm/power_spectrum.py`` (tools21cm v2.4.0).  It is the checker the
CUDA path is compared against; it is never imported by the product package
``tools21cm_b200``.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.

Parity pin: the reference holds no golden vectors for this path (its tests are
unseeded slope checks, ``tests/test_PowerSpectrum.py:10-48``).  This oracle is
therefore pinned against outputs of the reference itself, generated in the
build container by ``tests/golden/make_golden.py`` and committed under
``tests/golden/*.npz`` (``tests/test_oracle_golden.py`` checks them bit-for-bit
for mode counts and to 1e-13 for powers).

The third-party arithmetic the reference leans on is kept identical on purpose:
``scipy.fftpack.fftn/fftshift`` (reference ``power_spectrum.py:45,85-86``),
``numpy.histogram`` (``:128,131``) and ``numpy.mean`` (``:430``).  Neither numpy
nor scipy is pinned by the reference (``requirements.txt:2-3``); the versions in
this image are numpy 2.3 / scipy 1.18.

Every function cites the reference lines it follows.  Arithmetic order matters
for bit-exact mode counts (see ``k_axis`` and ``k_magnitude``).
"""
from __future__ import annotations

import numpy as np
from scipy import fftpack

# reference conv.py:10-11 with const.py:37 (h = 0.7): default comoving box in Mpc
DEFAULT_BOX_MPC = 244.0 / 0.7


# --------------------------------------------------------------------------
# geometry helpers
# --------------------------------------------------------------------------
def box_lengths(box_dims, shape):
    """Per-axis box lengths.  Reference ``_get_dims`` power_spectrum.py:522-534:
    None -> default box on every axis, scalar -> repeated, iterable -> as is."""
    if box_dims is None:
        return [DEFAULT_BOX_MPC] * len(shape)
    if not hasattr(box_dims, "__iter__"):
        return [box_dims] * len(shape)
    return box_dims


def axis_centre(n, ndim):
    """Index of the k=0 sample on the fftshifted grid as the reference defines it.

    3-D (power_spectrum.py:471-472): n/2 for even n, (n-1)/2 for odd n.
    1-D / 2-D (power_spectrum.py:458,463): (n-1)/2 for every n (a reference quirk
    that mis-centres even-length 1-D/2-D maps; reproduced deliberately)."""
    if ndim == 3:
        return n / 2 if n % 2 == 0 else (n - 1) / 2
    return (n - 1) / 2.0


def k_axis(n, length, ndim):
    """k component along one axis, float64, in fftshift index order.
    Operation order of power_spectrum.py:473-475: ``2.*np.pi * (x-c) / L``
    i.e. (2*pi*(x-c))/L with x int32 and c float64."""
    x = np.arange(n, dtype="int32")
    return 2.0 * np.pi * (x - axis_centre(n, ndim)) / length


def k_components(shape, box_dims):
    """Per-axis 1-D k tables (the reference builds full grids with np.indices,
    power_spectrum.py:462,470; broadcasting these tables gives the same values)."""
    nd = len(shape)
    return [k_axis(shape[d], box_dims[d], nd) for d in range(nd)]


def k_magnitude(shape, box_dims):
    """|k| on the full fftshifted grid.  power_spectrum.py:477 sums squares left
    to right: sqrt((kx**2 + ky**2) + kz**2).  1-D returns the signed kx itself
    (power_spectrum.py:459-460)."""
    kc = k_components(shape, box_dims)
    nd = len(shape)
    if nd == 1:
        return kc, kc[0]
    if nd == 2:
        kx, ky = kc[0][:, None], kc[1][None, :]
        return kc, np.sqrt(kx ** 2 + ky ** 2)
    kx, ky, kz = kc[0][:, None, None], kc[1][None, :, None], kc[2][None, None, :]
    return kc, np.sqrt(kx ** 2 + ky ** 2 + kz ** 2)


def k_edges(kbins, box_dims, kmax, binning="log", breakpoint=0.1):
    """Bin edges.  Reference ``_get_kbins`` power_spectrum.py:505-519.  Only a
    Python ``int`` asks for automatic edges; anything else is returned as given.
    kmin = 2*pi/min(L); kmax = max |k| on the grid."""
    if isinstance(kbins, int):
        kmin = 2.0 * np.pi / min(box_dims)
        if binning == "linear":
            return np.linspace(kmin, kmax, kbins + 1)
        if binning == "log":
            return 10 ** np.linspace(np.log10(kmin), np.log10(kmax), kbins + 1)
        lin = np.linspace(kmin, kmax, kbins + 1)
        log = 10 ** np.linspace(np.log10(kmin), np.log10(kmax), kbins + 1)
        return np.hstack((lin[lin < breakpoint], log[log > breakpoint]))
    return kbins


def volume_scale(shape, box_dims):
    """(V/prod N)^2 / V.  power_spectrum.py:50-52."""
    boxvol = np.prod(box_dims)
    pixelsize = boxvol / np.prod(shape)
    return pixelsize ** 2 / boxvol


# --------------------------------------------------------------------------
# n-D spectra
# --------------------------------------------------------------------------
def power_spectrum_nd(input_array, box_dims=None, verbose=False, **kwargs):
    """|FFT|^2 on the fftshifted grid, float64.  power_spectrum.py:24-54."""
    if kwargs.get("boxsize") is not None:
        box_dims = kwargs.get("boxsize")
    box_dims = box_lengths(box_dims, input_array.shape)
    ft = fftpack.fftshift(fftpack.fftn(input_array.astype("float64")))
    ps = np.abs(ft) ** 2
    ps *= volume_scale(input_array.shape, box_dims)
    return ps


def cross_power_spectrum_nd(input_array1, input_array2, box_dims, **kwargs):
    """Re(F1 conj F2) on the fftshifted grid.  power_spectrum.py:57-95."""
    assert input_array1.shape == input_array2.shape
    if kwargs.get("boxsize") is not None:
        box_dims = kwargs.get("boxsize")
    box_dims = box_lengths(box_dims, input_array1.shape)
    f1 = fftpack.fftshift(fftpack.fftn(input_array1.astype("float64")))
    f2 = fftpack.fftshift(fftpack.fftn(input_array2.astype("float64")))
    ps = np.real(f1) * np.real(f2) + np.imag(f1) * np.imag(f2)
    ps *= volume_scale(input_array1.shape, box_dims)
    return ps


# --------------------------------------------------------------------------
# binning
# --------------------------------------------------------------------------
def radial_average(input_array, box_dims, kbins=10, binning="log", breakpoint=0.1, **kwargs):
    """Shell average by two histograms.  power_spectrum.py:98-134.  Bins are
    half-open except the last, which is closed (np.histogram semantics); the
    returned centres are arithmetic mid-points; counts come back as float64."""
    if kwargs.get("boxsize") is not None:
        box_dims = kwargs.get("boxsize")
    _, k = k_magnitude(input_array.shape, box_dims)
    edges = k_edges(kbins, box_dims, k.max(), binning=binning, breakpoint=breakpoint)
    half = (edges[1:] - edges[:-1]) / 2.0
    total = np.histogram(k.flatten(), bins=edges, weights=input_array.flatten())[0]
    n_modes = np.histogram(k.flatten(), bins=edges)[0].astype("float")
    with np.errstate(all="ignore"):  # reference __init__.py:94-95 silences FP warnings
        total = total / n_modes
    return total, edges[:-1] + half, n_modes


def mu_of(kc, k, los_axis, absolute_mus):
    """cos(theta) to the line of sight; NaN where k < 0.001.  power_spectrum.py:481-502."""
    if los_axis not in (0, 1, 2):
        raise Exception("Your space is not %d-dimensional!" % los_axis)
    shape = [1, 1, 1]
    shape[los_axis] = -1
    klos = kc[los_axis].reshape(shape) * np.ones_like(k)
    with np.errstate(all="ignore"):
        mu = np.abs(klos / k) if absolute_mus else klos / np.abs(k)
    mu[np.where(k < 0.001)] = np.nan
    return mu


def kperp_nonzero(shape, los_axis):
    """Mask of modes off the k_perp = 0 line.  power_spectrum.py:563-575.  The
    comparison is against N/2 as a float, so it never matches for odd N."""
    idx = np.indices(shape)
    perp = [a for a in (0, 1, 2) if a != los_axis] if los_axis in (0, 1) else [0, 1]
    on_line = (idx[perp[0]] == shape[perp[0]] / 2) * (idx[perp[1]] == shape[perp[1]] / 2)
    return np.invert(on_line)


def mu_binning(powerspectrum, los_axis=0, mubins=20, kbins=10, box_dims=None, weights=None,
               exclude_zero_modes=True, binning="log", absolute_mus=True):
    """(k, mu) averages; every bin half-open in both k and mu.  power_spectrum.py:380-435.
    Like the reference this zeroes the centre sample of its input in place (:414)."""
    if weights is not None:
        raise ValueError("weights are not usable in the reference either (power_spectrum.py:386,433)")
    assert len(powerspectrum.shape) == 3
    kc, k = k_magnitude(powerspectrum.shape, box_dims)
    mu = mu_of(kc, k, los_axis, absolute_mus)
    edges = k_edges(kbins, box_dims, k.max(), binning=binning)
    half_k = (edges[1:] - edges[:-1]) / 2.0
    nk = len(edges) - 1
    good = kperp_nonzero(powerspectrum.shape, los_axis) if exclude_zero_modes \
        else np.ones(powerspectrum.shape, dtype=bool)
    if isinstance(mubins, int):
        mubins = np.linspace(0.0 if absolute_mus else -1.0, 1.0, mubins + 1)
    half_mu = (mubins[1:] - mubins[:-1]) / 2.0
    nmu = len(mubins) - 1
    powerspectrum[tuple((np.array(powerspectrum.shape) / 2).astype(int))] = 0.0
    out = np.zeros((nmu, nk))
    n_modes = np.zeros((nmu, nk))
    with np.errstate(all="ignore"):
        for ik in range(nk):
            in_k = (k >= edges[ik]) & (k < edges[ik + 1]) & good
            for im in range(nmu):
                sel = (mu >= mubins[im]) & (mu < mubins[im + 1]) & in_k
                vals = powerspectrum[sel]
                out[im, ik] = np.mean(vals) if vals.size else np.nan
                n_modes[im, ik] = vals.size
    return out, mubins[:-1] + half_mu, edges[:-1] + half_k, n_modes


# --------------------------------------------------------------------------
# user-facing estimators
# --------------------------------------------------------------------------
def apply_window(input_array, window):
    """Taper along the last axis, in place.  power_spectrum.py:12-21."""
    if window is None:
        return input_array
    from scipy.signal import windows
    name = window.lower()
    if name == "blackmanharris":
        input_array *= windows.blackmanharris(input_array.shape[-1])[None, None, :]
    elif name == "tukey":
        input_array *= windows.tukey(input_array.shape[-1])[None, None, :]
    else:
        input_array *= window
    return input_array


def power_spectrum_1d(input_array_nd, kbins=100, box_dims=None, return_n_modes=False,
                      binning="log", breakpoint=0.1, window=None, **kwargs):
    """power_spectrum.py:137-174."""
    input_array_nd = apply_window(input_array_nd, window)
    if kwargs.get("boxsize") is not None:
        box_dims = kwargs.get("boxsize")
    box_dims = box_lengths(box_dims, input_array_nd.shape)
    p = power_spectrum_nd(input_array_nd, box_dims=box_dims)
    ps, centres, n_modes = radial_average(p, kbins=kbins, box_dims=box_dims,
                                          binning=binning, breakpoint=breakpoint)
    return (ps, centres, n_modes) if return_n_modes else (ps, centres)


def cross_power_spectrum_1d(input_array1_nd, input_array2_nd, kbins=100, box_dims=None,
                            return_n_modes=False, binning="log", breakpoint=0.1, **kwargs):
    """power_spectrum.py:257-289."""
    if kwargs.get("boxsize") is not None:
        box_dims = kwargs.get("boxsize")
    box_dims = box_lengths(box_dims, input_array1_nd.shape)
    p = cross_power_spectrum_nd(input_array1_nd, input_array2_nd, box_dims=box_dims)
    ps, centres, n_modes = radial_average(p, kbins=kbins, box_dims=box_dims,
                                          binning=binning, breakpoint=breakpoint)
    return (ps, centres, n_modes) if return_n_modes else (ps, centres)


def power_spectrum_mu(input_array, los_axis=0, mubins=20, kbins=10, box_dims=None, weights=None,
                      exclude_zero_modes=True, return_n_modes=False, absolute_mus=True, **kwargs):
    """power_spectrum.py:292-332 (``binning`` is not forwarded: always 'log')."""
    if kwargs.get("boxsize") is not None:
        box_dims = kwargs.get("boxsize")
    box_dims = box_lengths(box_dims, input_array.shape)
    p = power_spectrum_nd(input_array, box_dims=box_dims)
    ps, mu_c, k_c, n_modes = mu_binning(p, los_axis, mubins, kbins, box_dims, weights,
                                        exclude_zero_modes, absolute_mus=absolute_mus)
    return (ps, mu_c, k_c, n_modes) if return_n_modes else (ps, mu_c, k_c)


def cross_power_spectrum_mu(input_array1, input_array2, los_axis=0, mubins=20, kbins=10,
                            box_dims=None, weights=None, exclude_zero_modes=True,
                            return_n_modes=False, absolute_mus=True, **kwargs):
    """power_spectrum.py:335-377."""
    if kwargs.get("boxsize") is not None:
        box_dims = kwargs.get("boxsize")
    box_dims = box_lengths(box_dims, input_array1.shape)
    p = cross_power_spectrum_nd(input_array1, input_array2, box_dims=box_dims)
    ps, mu_c, k_c, n_modes = mu_binning(p, los_axis, mubins, kbins, box_dims, weights,
                                        exclude_zero_modes, absolute_mus=absolute_mus)
    return (ps, mu_c, k_c, n_modes) if return_n_modes else (ps, mu_c, k_c)


def power_spectrum_2d(input_array, kbins=10, binning="log", box_dims=244 / .7, return_modes=False, nu_axis=2,
                      window=None, **kwargs):
    """(k_perp, k_par) means with scipy.stats.binned_statistic_2d.  power_spectrum.py:177-254."""
    from scipy import stats
    input_array = apply_window(input_array, window)
    if type(kbins) == list:
        binning = None
    elif np.array(kbins).size == 1:
        kbins = [kbins, kbins]
    elif not isinstance(kbins[0], int):
        binning = None
    if kwargs.get("boxsize") is not None:
        box_dims = kwargs.get("boxsize")
    box_dims = box_lengths(box_dims, input_array.shape)
    power = power_spectrum_nd(input_array, box_dims)
    kc = k_components(input_array.shape, box_dims)
    grids = np.meshgrid(*kc, indexing="ij")
    perp = [a for a in (0, 1, 2) if a != nu_axis]
    kz = np.abs(grids[nu_axis])
    kp = np.sqrt(grids[perp[0]] ** 2 + grids[perp[1]] ** 2)
    with np.errstate(all="ignore"):
        if binning is None:
            kper, kpar = np.array(kbins[0]), np.array(kbins[1])
        elif binning == "log":
            kper = np.linspace(np.log10(kp[kp != 0].min()), np.log10(kp.max()), kbins[0] + 1)
            kpar = np.linspace(np.log10(kz[kz != 0].min()), np.log10(kz.max()), kbins[1] + 1)
            kp, kz = np.log10(kp), np.log10(kz)
        elif binning == "linear":
            kper = np.linspace(kp[kp != 0].min(), kp.max(), kbins[0] + 1)
            kpar = np.linspace(kz[kz != 0].min(), kz.max(), kbins[1] + 1)
        kp, kz, power = kp.flatten(), kz.flatten(), power.flatten()
        res = stats.binned_statistic_2d(x=kp, y=kz, values=power, statistic="mean", bins=[kper, kpar])
        if binning == "log":
            kper_mid = np.power(10, 0.5 * (kper[:-1] + kper[1:]))
            kpar_mid = np.power(10, 0.5 * (kpar[:-1] + kpar[1:]))
        else:
            kper_mid = (kper[:-1] + kper[1:]) / 2.0
            kpar_mid = (kpar[:-1] + kpar[1:]) / 2.0
        if return_modes:
            cnt = stats.binned_statistic_2d(x=kp, y=kz, values=None, statistic="count", bins=[kper, kpar])
            return res.statistic, kper_mid, kpar_mid, cnt.statistic
    return res.statistic, kper_mid, kpar_mid


def dimensionless_ps(data, kbins=100, box_dims=None, binning="log", factor=10):
    """Delta^2(k) = P k^3 / 2 pi^2 interpolated from a finer binning.  power_spectrum.py:536-561."""
    from scipy.interpolate import interp1d
    pk, ks = power_spectrum_1d(data, kbins=kbins * factor, box_dims=box_dims, binning=binning)
    f = interp1d(ks, pk * ks ** 3 / 2 / np.pi ** 2)
    if binning == "log":
        knew = 10 ** np.linspace(np.log10(ks[1]), np.log10(ks[-1]), kbins)
    else:
        knew = np.linspace(ks[1], ks[-1], kbins)
    return f(knew), knew


def power_spectrum_multipoles(input_array, kbins=10, box_dims=None, los_axis=0,
                              output=['P0', 'P2', 'P4', 'nmodes'], exclude_zero_modes=False):
    """Reference ``power_legendre.py:4-90`` with the two lines that stop it from running repaired and nothing else:
    ``_get_mu`` gets its fourth argument (``:47`` passes three; only even powers of mu are used, so
    ``absolute_mus`` is immaterial) and the all-pass mask of ``exclude_zero_modes=False`` is boolean (``:51`` builds a
    float array, which ``idx *= good_idx`` at ``:79`` cannot cast).  ``nmodes`` stays ``len(np.nonzero(idx))``
    (``:80``): the number of dimensions, as the reference returns it; the mode count is added as ``n_modes``."""
    assert (los_axis >= 0 and los_axis <= len(input_array.shape))
    box_dims = box_lengths(box_dims, input_array.shape)
    ps = power_spectrum_nd(input_array, box_dims)
    kc, k = k_magnitude(input_array.shape, box_dims)
    edges = k_edges(kbins, box_dims, k.max())                       # binning='log' default of _get_kbins (:45)
    dk = (edges[1:] - edges[:-1]) / 2.
    mu = mu_of(kc, k, los_axis, True)
    good_idx = kperp_nonzero(ps.shape, los_axis) if exclude_zero_modes else np.ones(ps.shape, dtype=bool)
    P2 = 0.5 * (3. * mu ** 2 - 1.)
    P4 = 4.375 * (mu ** 2 - 0.115587) * (mu ** 2 - 0.741556)
    n_kbins = len(edges) - 1
    multipoles = {key: np.zeros(n_kbins) for key in ('P0', 'P2', 'P4') if key in output}
    nmodes = np.zeros(n_kbins)
    count = np.zeros(n_kbins)
    for i in range(n_kbins):
        idx = (k >= edges[i]) & (k < edges[i + 1])
        idx = idx * good_idx
        nmodes[i] = len(np.nonzero(idx))
        count[i] = np.count_nonzero(idx)
        if 'P0' in output:
            multipoles['P0'][i] = np.sum(ps[idx]) / np.sum(np.ones_like(mu)[idx] ** 2)
        if 'P2' in output:
            multipoles['P2'][i] = np.sum(ps[idx] * P2[idx]) / np.sum(P2[idx] ** 2)
        if 'P4' in output:
            multipoles['P4'][i] = np.sum(ps[idx] * P4[idx]) / np.sum(P4[idx] ** 2)
    multipoles['nmodes'] = nmodes
    multipoles['n_modes'] = count
    return multipoles, edges[:-1] + dk


def integrated_bispectrum_cross(cube1, cube2, Ncuts=4, kbins=15, box_dims=244 / .7, binning='log', normalize=False,
                                verbose=True):
    """Reference ``position_dependent_power_spectrum.py:45-88`` (and ``_W_L`` ``:90-103``), restated line for line:
    every sub-box masks the FULL cube and goes through power_spectrum_1d."""
    assert cube1.shape == cube2.shape
    assert cube1.shape[0] % Ncuts == 0 and cube1.shape[1] % Ncuts == 0 and cube1.shape[2] % Ncuts == 0
    Lx, Ly, Lz = cube1.shape[0] / Ncuts, cube1.shape[1] / Ncuts, cube1.shape[2] / Ncuts
    rLs = [[Lx / 2. + i * Lx, Ly / 2. + j * Ly, Lz / 2. + k * Lz] for i in range(Ncuts) for j in range(Ncuts) for k in range(Ncuts)]
    B_k = np.zeros(kbins, dtype=np.float64)
    P_k = np.zeros(kbins, dtype=np.float64)
    sig2 = 0
    n_box = Ncuts ** 3
    V_L = (Lx * Ly * Lz)
    for i in range(n_box):
        w = np.zeros(cube1.shape)
        r = rLs[i]
        w[int(r[0] - Lx / 2):int(r[0] + Lx / 2), int(r[1] - Ly / 2):int(r[1] + Ly / 2), int(r[2] - Lz / 2):int(r[2] + Lz / 2)] = 1
        c1 = cube1 * w
        c2 = cube2 * w
        pk, ks = power_spectrum_1d(c1, kbins=kbins, box_dims=box_dims, binning=binning)
        d_mean = c2.sum(dtype=np.float64) / V_L
        B_k += pk * d_mean
        P_k += pk
        sig2 += (d_mean) ** 2
    B_k = B_k / n_box
    P_k = P_k / n_box
    sig2 = sig2 / n_box
    if normalize:
        with np.errstate(all="ignore"):
            return B_k / P_k / sig2, ks
    return B_k, ks


def integrated_bispectrum(cube, Ncuts=4, kbins=15, box_dims=244 / .7, binning='log', normalize=False, verbose=True):
    """Reference ``position_dependent_power_spectrum.py:25-43``."""
    return integrated_bispectrum_cross(cube, cube, Ncuts=Ncuts, kbins=kbins, box_dims=box_dims, binning=binning,
                                       normalize=normalize, verbose=verbose)


def power_spectrum_response(cube, cube2=None, Ncuts=4, kbins=15, box_dims=244 / .7, binning='log', verbose=True):
    """Reference ``position_dependent_power_spectrum.py:5-22``."""
    return integrated_bispectrum_cross(cube, cube if cube2 is None else cube2, Ncuts=Ncuts, kbins=kbins,
                                       box_dims=box_dims, binning=binning, normalize=True, verbose=verbose)


# --------------------------------------------------------------------------
# input side: C2Ray-format readers, calc_dt, subtract_mean_signal
# --------------------------------------------------------------------------
def read_xfrac(filename, old_format=False, neutral=False, binary_format=False):
    """Reference ``xfrac_file.py:34-73``: header of 3 (binary) or 6 (Fortran unformatted: record markers around the
    three mesh sizes, then one more marker) int32, float64 (float32 for old files) payload in Fortran order."""
    with open(filename, 'rb') as f:
        if binary_format:
            mesh = np.fromfile(f, count=3, dtype='int32')
        else:
            mesh = np.fromfile(f, count=6, dtype='int32')[1:4]
        xi = np.fromfile(f, dtype='float32' if old_format else 'float64')
    xi = xi.reshape(tuple(int(m) for m in mesh), order='F')
    return 1.0 - xi if neutral else xi


def read_density(filename):
    """Reference ``density_file.py:41-57``: three int32 mesh sizes, float32 payload in Fortran order; returns
    (raw_density, cgs_density) with cgs = raw * rho_crit_0 * (mesh / nbox_fine)**3 * OmegaB (float32 array times a
    Python float: float32)."""
    with open(filename, 'rb') as f:
        mesh = np.fromfile(f, count=3, dtype='int32')
        raw = np.fromfile(f, dtype='float32')
    raw = raw.reshape(tuple(int(m) for m in mesh), order='F')
    return raw, raw * (_RHO_CRIT_0 * (float(mesh[0]) / 8000.0) ** 3 * _OMEGA_B)


# const.py:37-111 defaults
_H = 0.7
_OMEGA0 = 0.27
_OMEGA_B = 0.044
_LAM = 1.0 - _OMEGA0
_H0CGS = 100.0 * _H * 1e5 / (1e6 * 3.086e18)
_RHO_CRIT_0 = 3.0 * _H0CGS * _H0CGS / (8.0 * np.pi * 6.6732e-8)
_LAMBDA0 = 3.0e5 * 1.0e5 / (1.42e3 * 1.0e6)
_NUM_H_0 = (1 - 0.2486) * _OMEGA_B * _RHO_CRIT_0 / 1.672661e-24
_MEANDT = 3.0 * _LAMBDA0 ** 3 / (32. * np.pi) * 2.85e-15 * 0.068 * _NUM_H_0 / (_H0CGS / _H) * 1000.


def mean_dt(z):
    """Reference ``temperature.py:200-212``."""
    Ez = np.sqrt(_OMEGA0 * (1.0 + z) ** 3 + _LAM + (1.0 - _OMEGA0 - _LAM) * (1.0 + z) ** 2)
    return _MEANDT / _H * (1.0 + z) ** 2 / Ez


def calc_dt(xi, rho, z):
    """Reference ``temperature.py:51-80`` and ``_dt`` ``:215-222`` on arrays."""
    xi = xi.astype('float64')
    rho = rho.astype('float64')
    rho_mean = _RHO_CRIT_0 * _OMEGA_B if rho.min() >= 0 else 1
    return mean_dt(z) * (1.0 - xi) * rho / rho_mean


def subtract_mean_signal(signal, los_axis=2):
    """Reference ``statistics.py:62-85``."""
    out = signal.copy()
    for i in range(signal.shape[los_axis]):
        if los_axis in [0, -3]:
            out[i, :, :] -= signal[i, :, :].mean()
        if los_axis in [1, -2]:
            out[:, i, :] -= signal[:, i, :].mean()
        if los_axis in [2, -1]:
            out[:, :, i] -= signal[:, :, i].mean()
    return out
