"""numpy model of what the CUDA kernels compute (test infrastructure).

It follows the device algorithm step by step -- half spectrum from a real-to-complex
transform of (x - offset), analytic DC repair, per-axis tables, s-domain thresholds,
Hermitian-twin weights, and the fp32 lookup-table fast path -- so that the host-side
tables built by ``tools21cm_b200.binspec`` can be validated on a CPU box against the
oracle / golden vectors: if this model reproduces the reference's mode counts
bit for bit, a kernel that implements the same steps does too.
"""
import numpy as np

from tools21cm_b200 import binspec as bs


def half_spectrum(x, offset=0.0, single=False):
    """rfftn of (x - offset) with the DC sample repaired analytically."""
    if single:
        import scipy.fft
        y = (x.astype(np.float32) - np.float32(offset))
        f = scipy.fft.rfftn(y).astype(np.complex128)
    else:
        f = np.fft.rfftn(x.astype(np.float64) - offset)
    f[0, 0, 0] += offset * x.size
    return f


def lut_cells_k(spec, mx, my, mz):
    t = [a.astype(np.float32) for a in spec.tab_s]
    s_hat = (t[0][mx] + t[1][my]) + t[2][mz]
    with np.errstate(all="ignore"):
        u = np.log2(s_hat).astype(np.float32)
        cf = u * np.float32(spec.k_lut_inv) + np.float32(-spec.k_lut_u0 * spec.k_lut_inv)
    cf = np.where(np.isnan(cf), 0.0, cf)
    c = np.clip(np.floor(cf), 0, bs.K_LUT_CELLS - 1).astype(np.int64)
    return c, s_hat


def lut_cells_mu(spec, klos_f32, s_hat, twin_sign=1.0):
    with np.errstate(all="ignore"):
        mu_hat = (klos_f32 / np.sqrt(s_hat)).astype(np.float32) * np.float32(twin_sign)
    if spec.absolute_mus:
        mu_hat = np.abs(mu_hat)
    cf = (mu_hat + np.float32(1.0)) * np.float32(bs.MU_LUT_CELLS / 2.0)
    cf = np.where(np.isnan(cf), 0.0, cf)
    return np.clip(np.floor(cf), 0, bs.MU_LUT_CELLS - 1).astype(np.int64)


def bin_half_spectrum(power, spec, check_lut=True, iy0=0):
    """power: float64 [nx, ny_local, nzh] on the FFT-order half grid (rows iy0 .. iy0+ny_local of the
    global cube; the whole cube by default).  Returns (sums, counts)."""
    nx, ny, nz = spec.shape
    nzh = nz // 2 + 1
    nyl = power.shape[1]
    assert power.shape == (nx, nyl, nzh) and iy0 + nyl <= ny
    mx, my, mz = np.meshgrid(np.arange(nx), np.arange(iy0, iy0 + nyl), np.arange(nzh), indexing="ij")
    s = (spec.tab_s[0][mx] + spec.tab_s[1][my]) + spec.tab_s[2][mz]
    kb = np.searchsorted(spec.thr, s, side="right") - 1
    k_ok = (kb >= 0) & (kb < spec.nk)
    twin = (mz != 0) & (mz != spec.nyq_idx[2])          # conjugate partner is not stored
    nb = spec.nbins
    sums = np.zeros(nb)
    counts = np.zeros(nb, dtype=np.int64)

    if check_lut:
        cells, s_hat = lut_cells_k(spec, mx, my, mz)
        ent = spec.k_lut[cells].astype(np.int64)
        fast = ent != bs.LUT_EXACT
        fast_bin = np.where(ent == bs.LUT_OUTSIDE, -1, ent)
        exact_bin = np.where(k_ok, kb, -1)
        assert np.array_equal(fast_bin[fast], exact_bin[fast]), "k lookup table disagrees with exact path"
        lut_cells_k.last_exact_fraction = 1.0 - fast.mean()

    if spec.nmu == 0:
        w = np.where(twin, 2, 1)
        np.add.at(sums, kb[k_ok], (power * w)[k_ok])
        np.add.at(counts, kb[k_ok], w[k_ok])
        return sums, counts

    los = spec.los_axis
    midx = (mx, my, mz)[los]
    klos = spec.tab_los[midx]
    with np.errstate(all="ignore"):
        k = np.sqrt(s)
        mu = klos / k
    alive = k_ok & (s >= spec.s_nan)
    if spec.exclude_zero:
        perp = [a for a in (0, 1, 2) if a != los]
        on_line = ((mx, my, mz)[perp[0]] == spec.zero_idx[perp[0]]) & \
                  ((mx, my, mz)[perp[1]] == spec.zero_idx[perp[1]])
        alive &= ~on_line
    # the stored mode and, where it is not stored, its conjugate twin
    is_nyq_los = midx == spec.nyq_idx[los]
    mu_twin = np.where(is_nyq_los, mu, -mu)
    for which, (mval, present) in enumerate(((mu, np.ones_like(twin)), (mu_twin, twin))):
        m = np.abs(mval) if spec.absolute_mus else mval
        mb = np.searchsorted(spec.mu_edges, m, side="right") - 1
        ok = alive & present & (mb >= 0) & (mb < spec.nmu) & ~np.isnan(m)
        if check_lut:
            sign = np.where(is_nyq_los, 1.0, -1.0) if which == 1 else 1.0
            mc = lut_cells_mu(spec, spec.tab_los.astype(np.float32)[midx], s_hat, sign)
            ment = spec.mu_lut[mc].astype(np.int64)
            mfast = (ment != bs.LUT_EXACT) & fast & alive & present
            want = np.where(ok, mb, -1)
            got = np.where(ment == bs.LUT_OUTSIDE, -1, ment)
            # fast path is only trusted where the k cell is unflagged too
            got = np.where(fast_bin >= 0, got, -1)
            assert np.array_equal(got[mfast], want[mfast]), "mu lookup table disagrees with exact path"
        flat = mb * spec.nk + kb
        np.add.at(sums, flat[ok], power[ok])
        np.add.at(counts, flat[ok], 1)
    return sums, counts


def model_binned(cubes, spec, offsets=None, single=False, check_lut=True):
    offsets = offsets or [0.0] * len(cubes)
    f = [half_spectrum(c, o, single) for c, o in zip(cubes, offsets)]
    if len(f) == 1:
        power = f[0].real ** 2 + f[0].imag ** 2
    else:
        power = f[0].real * f[1].real + f[0].imag * f[1].imag
    return bin_half_spectrum(power, spec, check_lut)
