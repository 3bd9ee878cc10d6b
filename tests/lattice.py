"""Exact mode counts per spherical bin from the integer lattice, without any FFT or N^3 array.

``radial_average`` (reference power_spectrum.py:98-134) histograms k = sqrt(kx^2 + ky^2 + kz^2) (left-to-right
float64 sum, :477) with np.histogram: bin b holds edges[b] <= k < edges[b+1], the last bin is closed.  k is a
non-decreasing function of |f_z| for fixed (f_x, f_y) because every float64 operation involved is monotonic, so
the number of k below an edge along a z row is found by bisection with the reference's own float64 expressions;
x <-> y is an exact symmetry of the first addition.  Cost O(N^2 log N) instead of O(N^3): 1024^3 and 2048^3 take
seconds.  This is a CHECKER (tests only); it shares no code with tools21cm_b200/binspec.py.
"""
import numpy as np


def _axis_k(n, length):
    """|k| for f = 0 .. n//2 with the reference's expression 2*pi*(x - c)/L (:471-475), and the multiplicity of
    each |f| (both signs, except f = 0 and the Nyquist sample of an even axis)."""
    c = n / 2 if n % 2 == 0 else (n - 1) / 2
    x = np.arange(n, dtype="int32")
    k = 2.0 * np.pi * (x - c) / length
    f = np.rint(np.abs(x - c)).astype(np.int64)
    nf = int(f.max()) + 1
    kabs = np.zeros(nf)
    mult = np.zeros(nf, dtype=np.int64)
    for i in range(n):
        kabs[f[i]] = abs(k[i])
        mult[f[i]] += 1
    return kabs, mult


def reference_edges(shape, box_dims, kbins, binning="log"):
    """_get_kbins (:505-519) for an int kbins, with k.max() taken at the grid corner."""
    ks = [_axis_k(n, L)[0] for n, L in zip(shape, box_dims)]
    kmax = np.sqrt((ks[0][-1] ** 2 + ks[1][-1] ** 2) + ks[2][-1] ** 2)
    kmin = 2.0 * np.pi / min(box_dims)
    if binning == "linear":
        return np.linspace(kmin, kmax, kbins + 1)
    return 10 ** np.linspace(np.log10(kmin), np.log10(kmax), kbins + 1)


def shell_counts(shape, box_dims, edges):
    """n_modes of radial_average for a 3-D array of `shape`: exact integers, returned as float64 like the reference."""
    (kx, mx), (ky, my), (kz, mz) = [_axis_k(n, L) for n, L in zip(shape, box_dims)]
    a = (kx[:, None] ** 2 + ky[None, :] ** 2).ravel()             # kx**2 + ky**2, first addition of :477
    w = (mx[:, None] * my[None, :]).ravel()
    kz2 = kz ** 2
    cum = np.concatenate([[0], np.cumsum(mz)])                     # signed z samples with |f_z| < j
    nz = kz2.size

    def below(edge, strict):
        """per (fx, fy): number of signed f_z with k < edge (strict) or k <= edge"""
        lo = np.zeros(a.size, dtype=np.int64)
        hi = np.full(a.size, nz, dtype=np.int64)
        for _ in range(int(np.ceil(np.log2(nz + 1))) + 1):
            mid = (lo + hi) // 2
            k = np.sqrt(a + kz2[np.minimum(mid, nz - 1)])
            inside = (k < edge) if strict else (k <= edge)
            go = lo < hi
            lo = np.where(go & inside, mid + 1, lo)
            hi = np.where(go & ~inside, mid, hi)
        return cum[lo]

    edges = np.asarray(edges, dtype=np.float64)
    nb = edges.size - 1
    tot = np.empty(nb + 1, dtype=np.int64)
    for e in range(nb + 1):
        # modes with k < edges[e]; the last edge is inclusive (np.histogram closes the last bin)
        tot[e] = int(np.sum(w * below(edges[e], strict=(e < nb))))
    return np.diff(tot).astype(np.float64)
