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

// WarpBins: what the X pass uses on the device.  In the X pass the 32 lanes of a warp hold
// neighbouring modes (16 adjacent kz x 2 adjacent kx for the same register slot), which almost
// always share a bin, while the 16 slots of one thread are spread over all kx and do not.
// So the bin ids are kept WARP-UNIFORM in a small cache, each lane keeps a private partial
// sum per cache entry, and the shared histogram is touched only when an entry is evicted:
// one shuffle reduction and one add by lane 0 ("warp-aggregated atomics").  Lanes whose bin
// differs from the leader's are served in a second trip of the loop; nobody ever issues a
// per-lane atomic.  Every lane of the warp must call add() the same number of times.
constexpr int kWarpBinEntries = 4;
struct WarpBins {
#if T21_DEVICE_CODE
    int bin[kWarpBinEntries];        // warp-uniform, -1 = empty
    unsigned cnt[kWarpBinEntries];   // per lane
    double sum[kWarpBinEntries];     // per lane
    int victim;                      // warp-uniform round-robin pointer
#else
    RunAcc run;
#endif
};

T21_HD void warpbins_init(WarpBins& a) {
#if T21_DEVICE_CODE
#pragma unroll
    for (int c = 0; c < kWarpBinEntries; ++c) { a.bin[c] = -1; a.cnt[c] = 0; a.sum[c] = 0.0; }
    a.victim = 0;
#else
    run_init(a.run);
#endif
}

#if T21_DEVICE_CODE
template <class Hist>
__device__ __forceinline__ void warpbins_evict(WarpBins& a, Hist& h, int c) {
    // c is a compile-time constant at every call site (unrolled), so a.*[c] stay in registers
    double s = a.sum[c];
    unsigned n = a.cnt[c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        n += __shfl_xor_sync(0xffffffffu, n, o);
    }
    if ((threadIdx.x & 31) == 0 && n) h.add(a.bin[c], s, n);
    a.sum[c] = 0.0;
    a.cnt[c] = 0;
}
#endif

// valid: this lane has a mode to add (false for padding lanes); bin < 0 means "in no bin"
template <class Hist>
T21_HD void warpbins_add(WarpBins& a, Hist& h, int bin, double p, int w, bool valid) {
#if T21_DEVICE_CODE
    unsigned todo = __ballot_sync(0xffffffffu, valid && bin >= 0);
    while (todo) {                                   // warp-uniform loop: one trip per distinct bin
        const int leader = __ffs(todo) - 1;
        const int ref = __shfl_sync(0xffffffffu, bin, leader);
        const bool mine = valid && bin == ref;
        todo &= ~__ballot_sync(0xffffffffu, mine);
        bool hit = false;
#pragma unroll
        for (int c = 0; c < kWarpBinEntries; ++c) {
            if (!hit && a.bin[c] == ref) {           // warp-uniform branch
                hit = true;
                if (mine) { a.sum[c] += (w == 2) ? (p + p) : p; a.cnt[c] += (unsigned)w; }
            }
        }
        if (!hit) {
#pragma unroll
            for (int c = 0; c < kWarpBinEntries; ++c) {
                if (c == a.victim) {                 // warp-uniform branch
                    if (a.bin[c] >= 0) warpbins_evict(a, h, c);
                    a.bin[c] = ref;
                    if (mine) { a.sum[c] = (w == 2) ? (p + p) : p; a.cnt[c] = (unsigned)w; }
                }
            }
            a.victim = (a.victim + 1 == kWarpBinEntries) ? 0 : a.victim + 1;
        }
    }
#else
    if (valid) run_add(a.run, h, bin, p, w);
#endif
}

template <class Hist>
T21_HD void warpbins_flush(WarpBins& a, Hist& h) {
#if T21_DEVICE_CODE
#pragma unroll
    for (int c = 0; c < kWarpBinEntries; ++c)
        if (a.bin[c] >= 0) { warpbins_evict(a, h, c); a.bin[c] = -1; }
#else
    run_flush(a.run, h);
#endif
}

// ThreadBins: per-thread run-length accumulators (own / twin).  Used by CTAs smaller than a
// warp (pencils shorter than 32 samples, where speed is irrelevant) and by the CPU replay.
struct ThreadBins {
    RunAcc run[2];
};

// one interface over both accumulators; `which` = 0 for the stored mode, 1 for its twin
T21_HD void acc_init(WarpBins& a) { warpbins_init(a); }
T21_HD void acc_init(ThreadBins& a) { run_init(a.run[0]); run_init(a.run[1]); }
template <class Hist>
T21_HD void acc_add(WarpBins& a, Hist& h, int, int bin, double p, int w, bool valid) {
    warpbins_add(a, h, bin, p, w, valid);
}
template <class Hist>
T21_HD void acc_add(ThreadBins& a, Hist& h, int which, int bin, double p, int w, bool valid) {
    if (valid) run_add(a.run[which], h, bin, p, w);
}
template <class Hist>
T21_HD void acc_flush(WarpBins& a, Hist& h) { warpbins_flush(a, h); }
template <class Hist>
T21_HD void acc_flush(ThreadBins& a, Hist& h) { run_flush(a.run[0], h); run_flush(a.run[1], h); }

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
