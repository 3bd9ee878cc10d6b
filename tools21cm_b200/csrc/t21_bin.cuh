// Per-mode bin classification and run-length accumulation.
//
// classify():  (ix, iy, iz) -> bin of the stored mode and of its Hermitian twin.
//   fast path: fp32 s^ = sum of per-axis table entries, cell = f(log2 s^) (and f(mu^)),
//              table entry gives the bin unless flagged T21_LUT_EXACT;
//   exact path: the reference's float64 sequence -- s = (kx^2 + ky^2) + kz^2 with
//              non-contracted adds (power_spectrum.py:477), thresholds in the s
//              domain (binspec.py), k = sqrt(s), mu = k_los / k (power_spectrum.py:498),
//              half-open comparisons (:424-428) -- so mode counts are bit-identical.
//
// Accumulation: see WarpBins (X pass: warp-uniform bin cache, one histogram update per
//   eviction) and RunAcc (per-thread run-length accumulator for slowly varying bins).
#pragma once
#include "t21_common.cuh"
#include "../../include/t21ps.h"

namespace t21 {

T21_HD float fast_log2(float x) {
#if T21_DEVICE_CODE
    return __log2f(x);
#else
    return log2f(x);
#endif
}
T21_HD float fast_rsqrt(float x) {
#if T21_DEVICE_CODE
    return rsqrtf(x);
#else
    return 1.0f / sqrtf(x);
#endif
}
// floor to int with NaN -> 0 and saturation, then clamp to [0, hi]
T21_HD int cell_of(float f, int hi) {
#if T21_DEVICE_CODE
    int c = __float2int_rd(f);
#else
    int c = (f != f) ? 0 : (f <= -2.0e9f ? INT32_MIN : (f >= 2.0e9f ? INT32_MAX : (int)floorf(f)));
#endif
    return c < 0 ? 0 : (c > hi ? hi : c);
}

// largest b with a[b] <= v, or -1; a has n entries, non-decreasing
T21_HD int upper_index(const double* __restrict__ a, int n, double v) {
    int lo = 0, hi = n;  // first index with a[i] > v
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ldg(a + mid) <= v) lo = mid + 1; else hi = mid;
    }
    return lo - 1;
}

T21_HD int kbin_exact(const t21_binspec& B, double s) {
    const int b = upper_index(B.thr, B.nk + 1, s);
    return (b >= 0 && b < B.nk) ? b : -1;
}
T21_HD int mubin_exact(const t21_binspec& B, double mu) {
    if (!(mu == mu)) return -1;
    const int b = upper_index(B.mu_edges, B.nmu + 1, mu);
    return (b >= 0 && b < B.nmu) ? b : -1;
}

struct ModeBins {
    int own;      // flat bin of the stored mode, -1 if none
    int own_w;    // its weight (2 when the twin falls in the same bin by symmetry)
    int twin;     // flat bin of the conjugate twin when it must be binned separately, else -1
};

// has_twin: the conjugate partner of this mode is not itself stored (0 < iz < Nyquist on the
// halved axis for the fused kernels; always false for caller-supplied full arrays).
T21_HD ModeBins classify(const t21_binspec& B, int ix, int iy, int iz, bool has_twin) {
    ModeBins r;
    r.own = -1; r.own_w = 0; r.twin = -1;
    const float sh = (ldg(B.tab_sf[0] + ix) + ldg(B.tab_sf[1] + iy)) + ldg(B.tab_sf[2] + iz);
    const int kc = cell_of(fmaf(fast_log2(sh), B.k_lut_inv, B.k_lut_c0), B.k_cells - 1);
    const unsigned ke = ldg(B.k_lut + kc);

    if (B.nmu == 0) {
        int kb;
        if (ke == T21_LUT_EXACT) {
            const double s = dadd_rn(dadd_rn(ldg(B.tab_s[0] + ix), ldg(B.tab_s[1] + iy)), ldg(B.tab_s[2] + iz));
            kb = kbin_exact(B, s);
        } else {
            kb = (ke == T21_LUT_OUTSIDE) ? -1 : (int)ke;
        }
        r.own = kb;
        r.own_w = has_twin ? 2 : 1;
        return r;
    }

    // ---- (k, mu) ----
    const int los = B.los_axis;
    const int il = los == 0 ? ix : (los == 1 ? iy : iz);
    if (B.exclude_zero) {
        const int pa = los == 0 ? iy : ix;
        const int pb = los == 2 ? iy : iz;
        const int za = los == 0 ? B.zero_idx[1] : B.zero_idx[0];
        const int zb = los == 2 ? B.zero_idx[1] : B.zero_idx[2];
        if (pa == za && pb == zb) return r;
    }
    const bool abs_mu = B.absolute_mus != 0;
    const bool twin_sep = has_twin && !abs_mu;           // signed mu: twin lands in the mirrored bin
    const bool los_nyq = (il == B.nyq_idx[los]);         // Nyquist k_los does not change sign under conjugation
    const float muh = ldg(B.tab_losf + il) * fast_rsqrt(sh);
    const float m_own = abs_mu ? fabsf(muh) : muh;
    const float half_cells = 0.5f * (float)B.mu_cells;
    const unsigned me = ldg(B.mu_lut + cell_of((m_own + 1.0f) * half_cells, B.mu_cells - 1));
    unsigned mt = T21_LUT_OUTSIDE;
    if (twin_sep) {
        const float m_tw = los_nyq ? muh : -muh;
        mt = ldg(B.mu_lut + cell_of((m_tw + 1.0f) * half_cells, B.mu_cells - 1));
    }
    if (ke == T21_LUT_EXACT || me == T21_LUT_EXACT || mt == T21_LUT_EXACT) {
        const double s = dadd_rn(dadd_rn(ldg(B.tab_s[0] + ix), ldg(B.tab_s[1] + iy)), ldg(B.tab_s[2] + iz));
        if (s < B.s_nan) return r;                        // mu is NaN there (power_spectrum.py:501)
        const int kb = kbin_exact(B, s);
        if (kb < 0) return r;
        const double k = sqrt(s);
        const double mu = ldg(B.tab_los + il) / k;
        const int mb = mubin_exact(B, abs_mu ? fabs(mu) : mu);
        if (mb >= 0) { r.own = mb * B.nk + kb; r.own_w = (has_twin && abs_mu) ? 2 : 1; }
        if (twin_sep) {
            const int mbt = mubin_exact(B, los_nyq ? mu : -mu);
            if (mbt >= 0) r.twin = mbt * B.nk + kb;
        }
        return r;
    }
    if (ke == T21_LUT_OUTSIDE) return r;
    const int kb = (int)ke;
    if (me != T21_LUT_OUTSIDE) { r.own = (int)me * B.nk + kb; r.own_w = (has_twin && abs_mu) ? 2 : 1; }
    if (twin_sep && mt != T21_LUT_OUTSIDE) r.twin = (int)mt * B.nk + kb;
    return r;
}

// ---- pencil segment table (warp-per-pencil X pass) ------------------------------------------------------------------
// Along the x pencil (iy, iz) the bin of a mode changes only a few times: k grows with |kx| and mu is monotonic in it.
// The table lists the ranges: entry j = (p_j << 16) | (bin_j + 1), modes p_j <= |kx index| < p_{j+1} fall in bin_j.
// It is built by running classify() -- the per-mode classification every other kernel uses -- over |kx| = 0 .. nx/2 of
// the pencil and cutting wherever its answer changes: no monotonicity is assumed, only counted on for the table to be
// short.  The mirror mode nx - f must agree (shells and |mu| bins are even in kx; signed mu is refused by the caller):
// a pencil where it does not, or with more than `stride` ranges, reports stride + 1 and the table is not used.
constexpr unsigned kSegSentinel = 0xFFFF0000u;
// modes of an x pencil of nx samples (H = nx / 2) with a <= |kx index| < b, 0 <= a < b <= H + 1, in the kx >= 0 half
// (LO), the kx < 0 half (HI), or both.  Even nx: the Nyquist sample |kx| = H is single and counted with the kx < 0 half.
T21_HD unsigned segment_modes(bool lo, bool hi, unsigned a, unsigned b, unsigned H, bool even = true) {
    const unsigned lo_end = even ? H : H + 1u;                      // kx >= 0 half: |kx| in [0, lo_end)
    const unsigned nlo = (b < lo_end ? b : lo_end) - (a < lo_end ? a : lo_end);
    const unsigned nhi = b - (a > 1u ? a : 1u);                      // kx < 0 half: |kx| in [1, H]
    return (lo ? nlo : 0u) + (hi ? nhi : 0u);
}
// bin_of(ix) -> flat bin or -1, for ix = f (and ix = nx - f when the mirror is asked to agree).  A pencil that needs more
// than `stride` ranges, or whose mirror disagrees, gets the overflow mark in entry 0 and reports stride + 1: the
// warp-per-pencil X pass classifies such a pencil mode by mode; the other kernels only take tables without any.
constexpr unsigned kSegOverflow = 0xFFFE0000u;
template <class BinOf>
T21_HD int emit_pencil_segments(int nx, bool check_mirror, bool negative_half, BinOf&& bin_of, unsigned* __restrict__ out, int stride) {
    int s = 0, prev = -2;
    bool bad = false;
    for (int f = 0; f <= nx / 2; ++f) {
        const int b = bin_of(negative_half ? (nx - f) % nx : f);
        if (check_mirror && f > 0 && f < nx - f && bin_of(nx - f) != b) bad = true;
        if (b != prev) {
            if (s < stride) out[s] = ((unsigned)f << 16) | (unsigned)(b + 1);
            ++s;
            prev = b;
        }
    }
    for (int j = s; j < stride; ++j) out[j] = kSegSentinel;
    if (bad || s > stride) {
        out[0] = kSegOverflow;
        for (int j = 1; j < stride; ++j) out[j] = kSegSentinel;
        return stride + 1;
    }
    return s;
}

T21_HD int build_pencil_segments(const t21_binspec& B, int iy, int iz, int which, unsigned* __restrict__ out, int stride) {
    const int nz = B.n[2];
    const bool has_twin = (iz != 0) && (iz != nz - iz);
    return emit_pencil_segments(B.n[0], which == 0 || which == 5, which == 2 || which == 4, [&](int ix) {
        const ModeBins mb = classify(B, ix, iy, iz, has_twin);
        int b = (which == 3 || which == 4) ? mb.twin : mb.own;
        if (which == 5 && B.exclude_zero) {
            const int los = B.los_axis;
            const int pa = los == 0 ? iy : ix, pb = los == 2 ? iy : iz;
            const int za = los == 0 ? B.zero_idx[1] : B.zero_idx[0], zb = los == 2 ? B.zero_idx[1] : B.zero_idx[2];
            if (pa == za && pb == zb) b = -1;
        }
        return b;
    }, out, stride);
}

// (k_perp, k_par) bins of power_spectrum_2d (power_spectrum.py:177-254) from the separable host-built tables in FFT
// index order (binspec.build_cyl_tables): flat bin = perp * npar + par, -1 if either part is outside its bins
T21_HD int build_pencil_segments_cyl(int n0, int n1, int n2, int nu_axis, const int* __restrict__ perp_tab,
                                     const int* __restrict__ par_tab, int npar, int iy, int iz, unsigned* __restrict__ out, int stride) {
    return emit_pencil_segments(n0, true, false, [&](int ix) {
        const int pe = nu_axis == 2 ? ldg(perp_tab + (long long)ix * n1 + iy)
                                    : ldg(perp_tab + (long long)(nu_axis == 1 ? ix : iy) * n2 + iz);
        const int pa = ldg(par_tab + (nu_axis == 2 ? iz : (nu_axis == 1 ? iy : ix)));
        return (pe < 0 || pa < 0) ? -1 : pe * npar + pa;
    }, out, stride);
}

#ifdef __CUDACC__
// One warp (or a team of `tw` warps, this one being number `sub`) walks the ranges of a pencil whose powers sit in shared
// memory, pp[0 .. nx-1] in FFT order: lanes and team warps interleave over |kx|, every warp books its partial sum of a
// range and warp 0 of the team the range's closed-form mode count.  LO / HI: the kx >= 0 / kx < 0 half.
template <typename T, bool LO, bool HI, class Hist>
__device__ __forceinline__ void walk_pencil_ranges(const T* __restrict__ pp, unsigned nx, unsigned ent, int lane, int sub, int tw,
                                                   unsigned wt, Hist& hist) {
    const unsigned pos = ent >> 16, binp1 = ent & 0xFFFFu;
    const unsigned S = (unsigned)__popc(__ballot_sync(0xffffffffu, pos != 0xFFFFu));
    const unsigned H = nx / 2u;
    const bool even = (nx & 1u) == 0u;
    const unsigned lo_end = even ? H : H + 1u;
    for (unsigned j = 0; j < S; ++j) {                              // warp-uniform
        const int bin = (int)__shfl_sync(0xffffffffu, binp1, j) - 1;
        const unsigned a = __shfl_sync(0xffffffffu, pos, j);
        const unsigned b = j + 1u < S ? __shfl_sync(0xffffffffu, pos, (j + 1u) & 31u) : H + 1u;
        if (bin < 0) continue;
        T sum = (T)0;
        for (unsigned f = a + (unsigned)lane + 32u * (unsigned)sub; f < b; f += 32u * (unsigned)tw) {
            if (LO && f < lo_end) sum += pp[f];
            if (HI && f >= 1u) sum += pp[nx - f];
        }
        double t = (double)sum;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (lane == 0) hist.add(bin, t * (double)wt, sub == 0 ? segment_modes(LO, HI, a, b, H, even) * wt : 0u);
    }
}
#endif

// ---- accumulators ---------------------------------------------------------------------
// RunAcc: per-thread run-length accumulator.  Used where one thread walks modes whose bin
// changes slowly (t21_bin_only: consecutive z samples) and by the CPU replay of the X pass.
struct RunAcc {
    int bin;
    unsigned cnt;
    double sum;
};
T21_HD void run_init(RunAcc& a) { a.bin = -1; a.cnt = 0; a.sum = 0.0; }

// Hist: provides add(bin, sum, cnt)
template <class Hist>
T21_HD void run_flush(RunAcc& a, Hist& h) {
    if (a.cnt) h.add(a.bin, a.sum, a.cnt);
    a.cnt = 0; a.sum = 0.0;
}
template <class Hist>
T21_HD void run_add(RunAcc& a, Hist& h, int bin, double p, int w) {
    if (bin < 0) return;
    if (bin != a.bin) { run_flush(a, h); a.bin = bin; }
    a.sum += (w == 2) ? (p + p) : p;
    a.cnt += (unsigned)w;
}

// WarpBins: what the X pass uses on the device.  Its epilogue walks a power tile so that the
// 32 lanes of a warp hold neighbouring modes (adjacent kz x adjacent kx) and consecutive steps
// move kx by a few cells: the lanes nearly always fall in one bin, or in two adjacent bins
// while the warp crosses an edge.  So a warp keeps TWO warp-uniform bin entries, each with the
// fp32 interval of s^ (and mu^) that provably lies inside that bin (binspec.py guard bands)
// and a per-lane partial sum:
//   fast step : every lane's (s^, mu^) is inside one of the two intervals -> add to the
//               lane-private partial sum.  No table lookup, no shuffle, no atomic.
//   slow step : some lane is outside both -> classify() all lanes exactly, then serve each
//               distinct bin: hit, or evict the least recently used entry (shuffle reduction,
//               one histogram update by lane 0) and install the new bin's intervals.
// "Below / above every bin" are cached as pseudo bins nb, nb+1 so they do not force slow steps.
// Every lane of the warp must make the same sequence of calls.
template <bool MU> struct WarpBinsMu { float mlo[2], mhi[2]; };   // warp-uniform guard interval of mu^
template <> struct WarpBinsMu<false> {};

template <bool MU>
struct WarpBins : WarpBinsMu<MU> {
#if T21_DEVICE_CODE
    int bin[2];              // warp-uniform flat bin id (or pseudo bin), -1 = empty
    float slo[2], shi[2];    // warp-uniform guard interval of s^
    int lru;                 // warp-uniform: entry to evict next
    unsigned cnt[2];         // per lane
    double sum[2];           // per lane
#else
    RunAcc run;
#endif
};

template <bool MU>
T21_HD void warpbins_init(WarpBins<MU>& a) {
#if T21_DEVICE_CODE
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        a.bin[c] = -1; a.cnt[c] = 0; a.sum[c] = 0.0;
        a.slo[c] = 1.f; a.shi[c] = 0.f;                                   // empty interval
        if constexpr (MU) { a.mlo[c] = 1.f; a.mhi[c] = 0.f; }
    }
    a.lru = 0;
#else
    run_init(a.run);
#endif
}

#if T21_DEVICE_CODE
template <bool MU, class Hist>
__device__ __forceinline__ void warpbins_evict(WarpBins<MU>& a, Hist& h, int c, int nbins) {
    // c is a compile-time constant at every call site, so a.*[c] stay in registers
    double s = a.sum[c];
    unsigned n = a.cnt[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        n += __shfl_xor_sync(0xffffffffu, n, o);
    }
    if ((threadIdx.x & 31) == 0 && n && a.bin[c] < nbins) h.add(a.bin[c], s, n);   // pseudo bins are dropped
    a.sum[c] = 0.0;
    a.cnt[c] = 0;
}

template <bool MU>
__device__ __forceinline__ void warpbins_install(WarpBins<MU>& a, const t21_binspec& B, int c, int ref) {
    const int nb = B.nk * (MU ? B.nmu : 1);
    a.bin[c] = ref;
    int kb = ref, mb = -1;
    if (ref >= nb) kb = B.nk + (ref - nb);            // pseudo bins: any mu
    else if (MU) { mb = ref / B.nk; kb = ref - mb * B.nk; }
    a.slo[c] = ldg(B.k_lo_f + kb);
    a.shi[c] = ldg(B.k_hi_f + kb);
    if constexpr (MU) {
        if (mb >= 0) { a.mlo[c] = ldg(B.mu_lo_f + mb); a.mhi[c] = ldg(B.mu_hi_f + mb); }
        else { a.mlo[c] = -2.f; a.mhi[c] = 2.f; }
    }
}

// fast step; returns false (and adds nothing) if any lane needs the slow step.
// MU: whether mu^ takes part.  valid = this lane has a mode (false for padding lanes).
template <bool MU>
__device__ __forceinline__ bool warpbins_fast(WarpBins<MU>& a, float sh, float muh, double p, unsigned w, bool valid) {
    bool in0 = sh >= a.slo[0] && sh < a.shi[0];
    bool in1 = sh >= a.slo[1] && sh < a.shi[1];
    if constexpr (MU) {
        in0 = in0 && muh >= a.mlo[0] && muh < a.mhi[0];
        in1 = in1 && muh >= a.mlo[1] && muh < a.mhi[1];
    }
    if (!__all_sync(0xffffffffu, !valid || in0 || in1)) return false;
    if (valid) {
        if (in0) { a.sum[0] += p; a.cnt[0] += w; }
        else { a.sum[1] += p; a.cnt[1] += w; }
    }
    return true;
}

// slow step: `bin` is this lane's exact flat bin (-1: in no bin and not cacheable)
template <bool MU, class Hist>
__device__ __forceinline__ void warpbins_slow(WarpBins<MU>& a, Hist& h, const t21_binspec& B, int nbins, int bin,
                                              double p, unsigned w, bool valid) {
    unsigned todo = __ballot_sync(0xffffffffu, valid && bin >= 0);
    while (todo) {                                   // warp-uniform loop: one trip per distinct bin
        const int ref = __shfl_sync(0xffffffffu, bin, __ffs(todo) - 1);
        const bool mine = valid && bin == ref;
        todo &= ~__ballot_sync(0xffffffffu, mine);
        if (a.bin[0] == ref) {
            if (mine) { a.sum[0] += p; a.cnt[0] += w; }
            a.lru = 1;
        } else if (a.bin[1] == ref) {
            if (mine) { a.sum[1] += p; a.cnt[1] += w; }
            a.lru = 0;
        } else if (a.lru == 0) {
            if (a.bin[0] >= 0) warpbins_evict(a, h, 0, nbins);
            warpbins_install(a, B, 0, ref);
            if (mine) { a.sum[0] = p; a.cnt[0] = w; }
            a.lru = 1;
        } else {
            if (a.bin[1] >= 0) warpbins_evict(a, h, 1, nbins);
            warpbins_install(a, B, 1, ref);
            if (mine) { a.sum[1] = p; a.cnt[1] = w; }
            a.lru = 0;
        }
    }
}
#endif

// WarpBinCache: the same two-entry warp-uniform cache keyed by the exact bin id alone (no guard intervals),
// for kernels whose lanes read consecutive elements of a caller's array and get their bin from tables.
struct WarpBinCache {
    int bin[2];              // warp-uniform, -1 = empty
    int lru;
    unsigned cnt[2];         // per lane
    double sum[2];           // per lane
};
#ifdef __CUDACC__
__device__ __forceinline__ void wbc_init(WarpBinCache& a) {
    a.bin[0] = a.bin[1] = -1; a.lru = 0;
    a.cnt[0] = a.cnt[1] = 0u; a.sum[0] = a.sum[1] = 0.0;
}
template <class Hist>
__device__ __forceinline__ void wbc_evict(WarpBinCache& a, Hist& h, int c) {
    double s = a.sum[c];
    unsigned n = a.cnt[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        n += __shfl_xor_sync(0xffffffffu, n, o);
    }
    if ((threadIdx.x & 31) == 0 && n) h.add(a.bin[c], s, n);
    a.sum[c] = 0.0;
    a.cnt[c] = 0u;
}
// every lane of the warp calls this together; bin < 0 = this lane has nothing to add
template <class Hist>
__device__ __forceinline__ void wbc_add(WarpBinCache& a, Hist& h, int bin, double v, unsigned w) {
    const bool in0 = bin == a.bin[0], in1 = bin == a.bin[1];
    if (__all_sync(0xffffffffu, bin < 0 || in0 || in1)) {
        if (bin >= 0) {
            if (in0) { a.sum[0] += v; a.cnt[0] += w; }
            else { a.sum[1] += v; a.cnt[1] += w; }
        }
        return;
    }
    unsigned todo = __ballot_sync(0xffffffffu, bin >= 0);
    while (todo) {                                   // warp-uniform loop: one trip per distinct bin
        const int ref = __shfl_sync(0xffffffffu, bin, __ffs(todo) - 1);
        const bool mine = bin == ref;
        todo &= ~__ballot_sync(0xffffffffu, mine);
        if (a.bin[0] == ref) {
            if (mine) { a.sum[0] += v; a.cnt[0] += w; }
            a.lru = 1;
        } else if (a.bin[1] == ref) {
            if (mine) { a.sum[1] += v; a.cnt[1] += w; }
            a.lru = 0;
        } else if (a.lru == 0) {
            if (a.bin[0] >= 0) wbc_evict(a, h, 0);
            a.bin[0] = ref;
            if (mine) { a.sum[0] = v; a.cnt[0] = w; }
            a.lru = 1;
        } else {
            if (a.bin[1] >= 0) wbc_evict(a, h, 1);
            a.bin[1] = ref;
            if (mine) { a.sum[1] = v; a.cnt[1] = w; }
            a.lru = 0;
        }
    }
}
template <class Hist>
__device__ __forceinline__ void wbc_flush(WarpBinCache& a, Hist& h) {
    if (a.bin[0] >= 0) wbc_evict(a, h, 0);
    if (a.bin[1] >= 0) wbc_evict(a, h, 1);
    a.bin[0] = a.bin[1] = -1;
}
#endif

template <bool MU, class Hist>
T21_HD void warpbins_flush(WarpBins<MU>& a, Hist& h, int nbins) {
#if T21_DEVICE_CODE
    if (a.bin[0] >= 0) { warpbins_evict(a, h, 0, nbins); a.bin[0] = -1; a.slo[0] = 1.f; a.shi[0] = 0.f; }
    if (a.bin[1] >= 0) { warpbins_evict(a, h, 1, nbins); a.bin[1] = -1; a.slo[1] = 1.f; a.shi[1] = 0.f; }
#else
    (void)nbins;
    run_flush(a.run, h);
#endif
}

// ThreadBins: a per-thread run-length accumulator.  Used by CTAs smaller than a
// warp (pencils shorter than 32 samples, where speed is irrelevant) and by the CPU replay.
struct ThreadBins {
    RunAcc run[1];
};

// per-thread accumulation into either accumulator type (the WarpBins overload is only reached by
// the CPU replay, where WarpBins degenerates to a run accumulator)
template <class Hist>
T21_HD void acc_add_thread(ThreadBins& a, Hist& h, int bin, double p, int w) { run_add(a.run[0], h, bin, p, w); }
template <bool MU, class Hist>
T21_HD void acc_add_thread(WarpBins<MU>& a, Hist& h, int bin, double p, int w) {
#if !T21_DEVICE_CODE
    run_add(a.run, h, bin, p, w);
#endif
}
template <bool MU> T21_HD void acc_init(WarpBins<MU>& a) { warpbins_init(a); }
T21_HD void acc_init(ThreadBins& a) { run_init(a.run[0]); }
template <bool MU, class Hist>
T21_HD void acc_flush(WarpBins<MU>& a, Hist& h, int nbins) { warpbins_flush(a, h, nbins); }
template <class Hist>
T21_HD void acc_flush(ThreadBins& a, Hist& h, int) { run_flush(a.run[0], h); }

// CTA histogram in shared memory, updated with atomics (device) / plain arrays (host replay)
struct SmemHist {
    double* sum;
    unsigned* cnt;
    T21_HD void add(int bin, double s, unsigned c) {
#if T21_DEVICE_CODE
        atomicAdd(sum + bin, s);
        atomicAdd(cnt + bin, c);
#else
        sum[bin] += s; cnt[bin] += c;
#endif
    }
};
// One histogram per warp in shared memory: only lane 0 of the owning warp ever writes it, in
// program order, so no atomics are needed and the result is bit-reproducible run to run.
struct WarpPrivHist {
    double* sum;      // already offset to this warp's copy
    unsigned* cnt;
    T21_HD void add(int bin, double s, unsigned c) { sum[bin] += s; cnt[bin] += c; }
};
// direct global accumulation, used when the histogram does not fit in shared memory
struct GlobalHist {
    double* sum;
    unsigned long long* cnt;
    T21_HD void add(int bin, double s, unsigned c) {
#if T21_DEVICE_CODE
        atomicAdd(sum + bin, s);
        atomicAdd(cnt + bin, (unsigned long long)c);
#else
        sum[bin] += s; cnt[bin] += c;
#endif
    }
};

}  // namespace t21
