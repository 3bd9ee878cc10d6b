"""Host-side (numpy, float64) construction of everything the binning kernels need.

The reference decides bin membership with float64 comparisons on full-size
k / mu grids (``power_spectrum.py:121-131`` for shells, ``:391-431`` for (k, mu)).
The device recomputes k and mu per mode from per-axis tables instead of
materialising those grids.  To keep the mode counts *bit-identical* the tables and
edges are produced here with exactly the reference's float64 expressions, and the
comparisons are moved to a domain where the device needs no sqrt:

    k = sqrt(s),  s = (kx^2 + ky^2) + kz^2      (sum order of power_spectrum.py:477)
    e <= sqrt(s)  <=>  s >= thr(e),   thr(e) = min{ s : sqrt_rn(s) >= e }

because IEEE sqrt is monotone.  ``thr`` is found by stepping ulps around e*e.

A second, approximate structure makes the common case cheap: a lookup table over
log2(s) cells (and one over mu cells) that gives the bin directly unless the cell
is within one cell of an edge, in which case the device runs the exact float64
sequence.  The device evaluates the table index in fp32; one cell is ~1e3 times
wider than its fp32 evaluation error, so unflagged cells are unambiguous.

Nothing here runs on the hot path per mode: it is O(sum of axis lengths + bins).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

DEFAULT_BOX_MPC = 244.0 / 0.7  # reference conv.py:10-11 (boxsize / const.h)

LUT_OUTSIDE = 0xFFFF   # cell lies outside every bin
LUT_EXACT = 0xFFFE     # cell touches an edge: run the exact float64 path
K_LUT_CELLS = 8192
MU_LUT_CELLS = 4096
MU_NAN_K = 0.001       # reference power_spectrum.py:501: mu = NaN where k < 0.001


# --------------------------------------------------------------------------
# reference-equivalent scalar helpers
# --------------------------------------------------------------------------
def get_dims(box_dims, ashape):
    """Reference ``_get_dims`` power_spectrum.py:522-534."""
    if box_dims is None:
        return [DEFAULT_BOX_MPC] * len(ashape)
    if not hasattr(box_dims, "__iter__"):
        return [box_dims] * len(ashape)
    return box_dims


def axis_centre(n, ndim):
    """power_spectrum.py:471-472 (3-D) and :458,463 (1-D/2-D quirk)."""
    if ndim == 3:
        return n / 2 if n % 2 == 0 else (n - 1) / 2
    return (n - 1) / 2.0


def k_axis_shifted(n, length, ndim):
    """float64 k along one axis in fftshift index order; (2*pi*(x-c))/L as in :473-475."""
    x = np.arange(n, dtype="int32")
    return 2.0 * np.pi * (x - axis_centre(n, ndim)) / length


def to_fft_order(a):
    """fftshift index i = (m + n//2) % n  ->  table indexed by the FFT index m."""
    n = a.shape[0]
    return a[(np.arange(n) + n // 2) % n]


def grid_kmax(kaxes):
    """max |k| over the grid == the expression at the per-axis maxima (rounding is monotone)."""
    sq = [float(np.max(k ** 2)) for k in kaxes]
    if len(sq) == 1:
        return float(np.max(kaxes[0]))
    tot = sq[0]
    for v in sq[1:]:
        tot = tot + v
    return float(np.sqrt(tot))


def make_k_edges(kbins, box_dims, kmax, binning="log", breakpoint=0.1):
    """Reference ``_get_kbins`` power_spectrum.py:505-519."""
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


def make_mu_edges(mubins, absolute_mus):
    """power_spectrum.py:406-409."""
    if isinstance(mubins, int):
        return np.linspace(0.0 if absolute_mus else -1.0, 1.0, mubins + 1)
    return mubins


def volume_scale(shape, box_dims):
    """power_spectrum.py:50-52."""
    boxvol = np.prod(box_dims)
    pixelsize = boxvol / np.prod(shape)
    return float(pixelsize ** 2 / boxvol)


# --------------------------------------------------------------------------
# exact thresholds in the s = k^2 domain
# --------------------------------------------------------------------------
def sqrt_threshold(edge, strict=False):
    """Smallest float64 s >= 0 with sqrt(s) >= edge (or > edge when strict)."""
    e = float(edge)
    if np.isnan(e):
        return np.inf

    def ok(s):
        r = np.sqrt(np.float64(s))
        return (r > e) if strict else (r >= e)

    if e < 0 or (e == 0 and not strict):
        return 0.0
    if np.isinf(e):
        return np.inf
    s = np.float64(e) * np.float64(e)
    while s > 0 and ok(s):
        s = np.nextafter(s, -np.inf)
    while not ok(s):
        s = np.nextafter(s, np.inf)
    return float(s)


def linear_threshold(edge, strict=False):
    """1-D arrays bin the signed k itself (power_spectrum.py:459-460): no sqrt."""
    e = float(edge)
    return float(np.nextafter(e, np.inf)) if strict else e


# --------------------------------------------------------------------------
# the spec handed to the device
# --------------------------------------------------------------------------
@dataclass
class BinSpec:
    shape: tuple                 # logical 3-D shape (1-D/2-D inputs are padded with 1s at the front)
    nk: int
    nmu: int                     # 0 -> shells only
    k_edges: np.ndarray          # float64 [nk+1] as the reference would build them
    mu_edges: np.ndarray         # float64 [nmu+1] or empty
    thr: np.ndarray              # float64 [nk+1], s-domain thresholds: bin b <=> thr[b] <= s < thr[b+1]
    tab_s: list                  # 3 x float64 [n_d]: per-axis addends of s (k_d^2, or k for 1-D)
    tab_los: np.ndarray          # float64 [n_los]: signed k along the line of sight (mu only)
    los_axis: int                # axis index into the padded 3-D shape, -1 when nmu == 0
    absolute_mus: bool
    exclude_zero: bool           # drop the k_perp = 0 line (only effective if both perp axes are even)
    zero_idx: tuple              # per-axis index of the k=0 sample in the tables' index order
    nyq_idx: tuple               # per-axis index of the self-conjugate Nyquist sample, -1 if none
    s_nan: float                 # mu is NaN (mode dropped) where s < s_nan  [thr of k >= 0.001]
    k_lut: np.ndarray            # uint16 [K_LUT_CELLS]
    k_lut_u0: float
    k_lut_inv: float
    mu_lut: np.ndarray           # uint16 [MU_LUT_CELLS] or empty
    k_lo_f: np.ndarray           # float32 [nk+2]: fp32 s^ in [k_lo_f[b], k_hi_f[b]) => exact s is in k bin b
    k_hi_f: np.ndarray           #   entries nk / nk+1 are the "below all bins" / "above all bins" pseudo bins
    mu_lo_f: np.ndarray          # float32 [nmu]: same for the fp32 mu^ estimate
    mu_hi_f: np.ndarray
    order: str                   # 'fft' (fused FFT kernels) or 'shifted' (caller-supplied arrays)
    extra: dict = field(default_factory=dict)

    @property
    def nbins(self):
        return self.nk * max(self.nmu, 1)

    def k_centres(self):
        e = self.k_edges
        return e[:-1] + (e[1:] - e[:-1]) / 2.0      # power_spectrum.py:126,134

    def mu_centres(self):
        e = self.mu_edges
        return e[:-1] + (e[1:] - e[:-1]) / 2.0      # power_spectrum.py:410,435


def _build_k_lut(thr, smin, smax, extra_flags=()):
    """Cells over u = log2(s).  Entry = bin index, LUT_OUTSIDE, or LUT_EXACT."""
    cells = K_LUT_CELLS
    u0 = np.log2(smin) - 0.5
    u1 = np.log2(smax) + 0.5
    inv = cells / (u1 - u0)
    nb = len(thr) - 1
    centres = u0 + (np.arange(cells) + 0.5) / inv
    with np.errstate(divide="ignore"):
        tlog = np.where(thr > 0, np.log2(np.maximum(thr, 1e-300)), -np.inf)
    b = np.searchsorted(tlog, centres, side="right") - 1
    lut = np.where((b >= 0) & (b < nb), b, LUT_OUTSIDE).astype(np.int64)
    for v in extra_flags:            # (k, mu) only: s < s_nan has mu = NaN and is in no bin
        if v > 0:
            lut[centres < np.log2(v)] = LUT_OUTSIDE
    marks = list(tlog[np.isfinite(tlog)]) + [np.log2(v) for v in extra_flags if v > 0]
    for t in marks:
        c = int(np.floor((t - u0) * inv))
        for cc in (c - 1, c, c + 1):
            if 0 <= cc < cells:
                lut[cc] = LUT_EXACT
    lut[0] = LUT_EXACT            # s == 0, negatives and NaN clamp here
    lut[cells - 1] = LUT_EXACT    # overflow clamps here
    return lut.astype(np.uint16), float(u0), float(inv)


def _build_mu_lut(mu_edges, absolute_mus):
    """Cells over mu in [-1, 1].  Half-open bins lo <= mu < hi (power_spectrum.py:428)."""
    cells = MU_LUT_CELLS
    nb = len(mu_edges) - 1
    centres = -1.0 + (np.arange(cells) + 0.5) * (2.0 / cells)
    b = np.searchsorted(mu_edges, centres, side="right") - 1
    lut = np.where((b >= 0) & (b < nb), b, LUT_OUTSIDE).astype(np.int64)
    for t in mu_edges:
        c = int(np.floor((t + 1.0) * cells / 2.0))
        for cc in (c - 1, c, c + 1):
            if 0 <= cc < cells:
                lut[cc] = LUT_EXACT
    lut[0] = LUT_EXACT
    lut[cells - 1] = LUT_EXACT
    return lut.astype(np.uint16)


S_GUARD = 1.0e-6    # relative guard on s^: >= 5x the fp32 error of (sx + sy) + sz from rounded tables
MU_GUARD = 4.0e-6   # absolute guard on mu^ = k_los * rsqrt(s^) (|mu| <= 1)


def _guard_intervals(thr, mu_edges, s_nan):
    """fp32 intervals strictly inside each bin, for the device's no-lookup fast path."""
    nk = len(thr) - 1
    lo = np.empty(nk + 2, dtype=np.float64)
    hi = np.empty(nk + 2, dtype=np.float64)
    lo[:nk] = np.maximum(thr[:-1], s_nan) * (1.0 + S_GUARD)
    hi[:nk] = thr[1:] * (1.0 - S_GUARD)
    lo[nk], hi[nk] = -np.inf, thr[0] * (1.0 - S_GUARD)           # below every bin
    lo[nk + 1], hi[nk + 1] = thr[-1] * (1.0 + S_GUARD), np.inf   # above every bin
    lo_f = np.nextafter(lo.astype(np.float32), np.float32(np.inf))      # round inwards
    hi_f = np.nextafter(hi.astype(np.float32), np.float32(-np.inf))
    lo_f[nk] = -np.inf
    hi_f[nk + 1] = np.inf
    if thr[0] <= 0:
        hi_f[nk] = -np.inf                                               # nothing lies below
    if len(mu_edges) >= 2:
        mlo = np.nextafter((mu_edges[:-1] + MU_GUARD).astype(np.float32), np.float32(np.inf))
        mhi = np.nextafter((mu_edges[1:] - MU_GUARD).astype(np.float32), np.float32(-np.inf))
    else:
        mlo = np.zeros(1, dtype=np.float32)
        mhi = np.zeros(1, dtype=np.float32)
    return lo_f, hi_f, mlo, mhi


def build_binspec(shape, box_dims, kbins, *, binning="log", breakpoint=0.1, mubins=None, los_axis=0,
                  absolute_mus=True, exclude_zero_modes=True, last_bin_closed=True, order="fft", shells_with_los=False):
    """Assemble a BinSpec for an array of ``shape`` (1-, 2- or 3-D).

    ``last_bin_closed`` is True for ``radial_average`` (np.histogram closes the last bin)
    and False for ``mu_binning`` (``k < kmax`` strictly, power_spectrum.py:424).
    ``shells_with_los``: shell bins (``mubins`` None) that still carry the signed k along ``los_axis`` and the
    mu = NaN threshold -- what ``power_spectrum_multipoles`` weights its modes with (power_legendre.py:47,55-61).
    """
    ndim = len(shape)
    if ndim not in (1, 2, 3):
        raise ValueError("only 1-, 2- and 3-dimensional arrays are supported")
    box_dims = list(box_dims)
    kaxes = [k_axis_shifted(shape[d], box_dims[d], ndim) for d in range(ndim)]
    kmax = grid_kmax(kaxes)
    edges = make_k_edges(kbins, box_dims, kmax, binning=binning, breakpoint=breakpoint)
    if not isinstance(edges, np.ndarray):
        raise TypeError("kbins must be an int or a numpy array of bin edges")  # reference fails at :126 too
    edges = edges.astype(np.float64)
    if edges.ndim != 1 or edges.size < 2:
        raise ValueError("need at least two bin edges")
    if np.any(np.diff(edges) < 0):
        raise ValueError("`bins` must increase monotonically, when an array")  # np.histogram's message
    nk = edges.size - 1

    if ndim == 1:
        tabs = [kaxes[0].copy()]
        thr = np.array([linear_threshold(e) for e in edges[:-1]]
                       + [linear_threshold(edges[-1], strict=last_bin_closed)])
    else:
        tabs = [k ** 2 for k in kaxes]
        thr = np.array([sqrt_threshold(e) for e in edges[:-1]]
                       + [sqrt_threshold(edges[-1], strict=last_bin_closed)])

    if order == "fft":
        if ndim != 3:
            raise ValueError("FFT-order tables are only used by the fused 3-D path")
        tabs = [to_fft_order(t) for t in tabs]
        kaxes_o = [to_fft_order(k) for k in kaxes]
        zero_idx = tuple(0 for _ in shape)
        nyq_idx = tuple(n // 2 if n % 2 == 0 else -1 for n in shape)
    else:
        kaxes_o = kaxes
        zero_idx = tuple(int(np.argmin(np.abs(k))) for k in kaxes)
        nyq_idx = tuple(0 if n % 2 == 0 else -1 for n in shape)

    nmu = 0
    mu_edges = np.zeros(0)
    mu_lut = np.zeros(0, dtype=np.uint16)
    tab_los = np.zeros(1)
    los3 = -1
    s_nan = 0.0
    extra_flags = ()
    if mubins is not None:
        if ndim != 3:
            raise AssertionError("mu binning needs a 3-D array")  # reference assert :389
        if los_axis not in (0, 1, 2):
            raise Exception("Your space is not %d-dimensional!" % los_axis)  # :496
        mu_edges = np.asarray(make_mu_edges(mubins, absolute_mus), dtype=np.float64)
        if mu_edges.ndim != 1 or mu_edges.size < 2:
            raise ValueError("need at least two mu edges")
        nmu = mu_edges.size - 1
        mu_lut = _build_mu_lut(mu_edges, absolute_mus)
        tab_los = kaxes_o[los_axis].copy()
        los3 = los_axis
        s_nan = sqrt_threshold(MU_NAN_K)
        extra_flags = (s_nan,)

    elif shells_with_los:
        if ndim != 3 or los_axis not in (0, 1, 2):
            raise ValueError("the line of sight must be one of the three axes of a 3-D array")
        tab_los = kaxes_o[los_axis].copy()
        s_nan = sqrt_threshold(MU_NAN_K)

    # pad to three axes (front) so one kernel signature serves 1-, 2- and 3-D inputs
    pad = 3 - ndim
    shape3 = (1,) * pad + tuple(int(n) for n in shape)
    tabs3 = [np.zeros(1)] * pad + [np.ascontiguousarray(t, dtype=np.float64) for t in tabs]
    zero3 = (0,) * pad + zero_idx
    nyq3 = (-1,) * pad + nyq_idx

    nz = [t[t != 0] for t in tabs3]
    if ndim == 1:
        pos = np.abs(tabs3[2][tabs3[2] != 0])
        smin, smax = float(pos.min()), float(pos.max())
    else:
        smin = min(float(np.min(t)) for t in nz if t.size)
        smax = sum(float(np.max(t)) for t in tabs3)
    k_lut, u0, inv = _build_k_lut(thr, smin, max(smax, smin * 2), extra_flags)

    # the perpendicular-line exclusion of power_spectrum.py:567-573 compares an integer
    # index with N/2 as a float: it can only ever match when both perpendicular N are even
    excl = False
    if nmu and exclude_zero_modes:
        perp = [a for a in (0, 1, 2) if a != los_axis]
        excl = all(shape[a] % 2 == 0 for a in perp)

    k_lo_f, k_hi_f, mu_lo_f, mu_hi_f = _guard_intervals(thr, mu_edges, s_nan)
    return BinSpec(shape=shape3, nk=nk, nmu=nmu, k_edges=edges, mu_edges=mu_edges, thr=thr,
                   tab_s=tabs3, tab_los=np.ascontiguousarray(tab_los, dtype=np.float64), los_axis=los3,
                   absolute_mus=bool(absolute_mus), exclude_zero=excl, zero_idx=zero3, nyq_idx=nyq3,
                   s_nan=float(s_nan), k_lut=k_lut, k_lut_u0=u0, k_lut_inv=inv, mu_lut=mu_lut,
                   k_lo_f=k_lo_f, k_hi_f=k_hi_f, mu_lo_f=mu_lo_f, mu_hi_f=mu_hi_f, order=order)


# --------------------------------------------------------------------------
# (k_perp, k_par) tables for power_spectrum_2d
# --------------------------------------------------------------------------
@dataclass
class CylSpec:
    nperp: int
    npar: int
    perp_tab: np.ndarray      # int32 [n_a, n_b]: perpendicular bin of every index pair across the line of sight, -1 = none
    par_tab: np.ndarray       # int32 [n_nu]: parallel bin of every index along the line of sight, -1 = none
    kper_mid: np.ndarray
    kpar_mid: np.ndarray
    nu_axis: int

    def counts(self):
        """Modes per (k_perp, k_par) bin, float64 like scipy's 'count' statistic: the bins are separable, so the count is
        (index pairs whose k_perp falls in the perp bin) x (indices whose |k_par| falls in the par bin), exactly."""
        if getattr(self, "_counts", None) is None:
            nperp = np.bincount(self.perp_tab[self.perp_tab >= 0].ravel(), minlength=self.nperp).astype(np.float64)
            npar = np.bincount(self.par_tab[self.par_tab >= 0].ravel(), minlength=self.npar).astype(np.float64)
            self._counts = nperp[:, None] * npar[None, :]
        return self._counts.copy()


def build_cyl_tables(shape, box_dims, kbins, binning="log", nu_axis=2, order="shifted"):
    """Bin tables of the reference's power_spectrum_2d (power_spectrum.py:209-254).

    A mode's bin is separable, so the reference's digitisation (scipy.stats.binned_statistic_2d, i.e.
    np.digitize plus scipy's right-edge rule) is run HERE, with the same scipy routine, on the n_a*n_b distinct
    k_perp values and the n_nu distinct k_par values instead of on all N^3 modes.  The device only looks the
    two tables up, so counts are identical to the reference by construction, whatever the bins."""
    from scipy import stats
    if len(shape) != 3:
        raise ValueError("power_spectrum_2d needs a 3-D array")
    if type(kbins) == list:                      # :211-216, including the reference's dispatch quirks
        binning = None
    elif np.array(kbins).size == 1:
        kbins = [kbins, kbins]
    elif not isinstance(kbins[0], int):
        binning = None
    kax = [k_axis_shifted(shape[d], box_dims[d], 3) for d in range(3)]
    xy = [0, 1, 2]
    xy.remove(nu_axis)
    kz = np.abs(kax[nu_axis])
    kp = np.sqrt(kax[xy[0]][:, None] ** 2 + kax[xy[1]][None, :] ** 2)
    with np.errstate(all="ignore"):
        if binning is None:
            kper = np.array(kbins[0])
            kpar = np.array(kbins[1])
        elif binning == "log":
            kper = np.linspace(np.log10(kp[kp != 0].min()), np.log10(kp.max()), kbins[0] + 1)
            kpar = np.linspace(np.log10(kz[kz != 0].min()), np.log10(kz.max()), kbins[1] + 1)
            kp, kz = np.log10(kp), np.log10(kz)
        elif binning == "linear":
            kper = np.linspace(kp[kp != 0].min(), kp.max(), kbins[0] + 1)
            kpar = np.linspace(kz[kz != 0].min(), kz.max(), kbins[1] + 1)
        else:
            raise UnboundLocalError("cannot access local variable 'kper'")     # what the reference does
        # the same digitisation scipy applies per dimension inside binned_statistic_2d
        rp = stats.binned_statistic(kp.ravel(), None, statistic="count", bins=kper)
        rz = stats.binned_statistic(kz.ravel(), None, statistic="count", bins=kpar)
    nperp, npar = len(rp.bin_edges) - 1, len(rz.bin_edges) - 1
    bp = rp.binnumber.astype(np.int64) - 1
    bz = rz.binnumber.astype(np.int64) - 1
    perp_tab = np.where((bp >= 0) & (bp < nperp), bp, -1).astype(np.int32).reshape(kp.shape)
    par_tab = np.where((bz >= 0) & (bz < npar), bz, -1).astype(np.int32)
    if binning == "log":
        kper_mid = np.power(10, 0.5 * (kper[:-1] + kper[1:]))
        kpar_mid = np.power(10, 0.5 * (kpar[:-1] + kpar[1:]))
    else:
        kper_mid = (kper[:-1] + kper[1:]) / 2.0
        kpar_mid = (kpar[:-1] + kpar[1:]) / 2.0
    if order == "fft":        # tables indexed by FFT index m: fftshift index i = (m + n//2) % n along every axis
        ia = (np.arange(perp_tab.shape[0]) + perp_tab.shape[0] // 2) % perp_tab.shape[0]
        ib = (np.arange(perp_tab.shape[1]) + perp_tab.shape[1] // 2) % perp_tab.shape[1]
        perp_tab = perp_tab[np.ix_(ia, ib)]
        par_tab = to_fft_order(par_tab)
    return CylSpec(nperp=nperp, npar=npar, perp_tab=np.ascontiguousarray(perp_tab),
                   par_tab=np.ascontiguousarray(par_tab), kper_mid=kper_mid, kpar_mid=kpar_mid, nu_axis=nu_axis)
