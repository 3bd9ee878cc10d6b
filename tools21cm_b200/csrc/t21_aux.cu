// Small kernels around the three passes: mean-offset estimate, histogram reduction,
// and element-wise binning of caller-supplied arrays (radial_average / mu_binning).
#include <string.h>
#include "t21_launch.cuh"

namespace t21 {

// ---- mean offset --------------------------------------------------------------------
// c ~ mean(x) from a strided sample (<= 16384 values, odd stride so it walks every axis).
// The FFT of x - c differs from the FFT of x only in the DC sample, which the X pass
// repairs analytically, so c only needs to be close to the mean: it is there to keep
// the fp32 partial sums small, not to define the result.
template <typename T>
__global__ void __launch_bounds__(1024) mean_estimate_kernel(const T* __restrict__ x, long long n, T* out) {
    constexpr int kPer = 16;                               // independent loads in flight per thread
    const long long want = n < 1024 * kPer ? n : 1024 * kPer;
    const long long stride = (n / want) | 1;
    const long long count = (n - 1) / stride + 1;
    T v[kPer];
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
        const long long i = threadIdx.x + 1024LL * u;
        v[u] = i < count ? x[i * stride] : (T)0;
    }
    double s = 0.0;
#pragma unroll
    for (int u = 0; u < kPer; ++u) s += (double)v[u];
    for (long long i = threadIdx.x + 1024LL * kPer; i < count; i += 1024) s += (double)x[i * stride];
    __shared__ double red[32];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = red[threadIdx.x];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) *out = (T)(s / (double)count);
    }
}

int launch_mean_estimate(int dtype, const void* x, long long n, void* offset_out, cudaStream_t st) {
    if (dtype == T21_F32) mean_estimate_kernel<float><<<1, 1024, 0, st>>>((const float*)x, n, (float*)offset_out);
    else mean_estimate_kernel<double><<<1, 1024, 0, st>>>((const double*)x, n, (double*)offset_out);
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- deterministic inter-CTA reduction ------------------------------------------------
// one warp per bin: lane l adds partials l, l+32, ... in order, then a fixed shuffle tree joins the 32
// lane sums -- the order never depends on timing, so the result is reproducible
__global__ void __launch_bounds__(128)
reduce_partials_kernel(const double* __restrict__ part_sum, const unsigned* __restrict__ part_cnt,
                       int nparts, int nbins, double* __restrict__ sums, long long* __restrict__ counts) {
    const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (b >= nbins) return;
    double s = 0.0;
    long long c = 0;
    for (int p = lane; p < nparts; p += 32) {
        s += part_sum[(size_t)p * nbins + b];
        c += part_cnt[(size_t)p * nbins + b];
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        c += __shfl_down_sync(0xffffffffu, c, o);
    }
    if (lane == 0) { sums[b] = s; counts[b] = c; }
}

int launch_reduce_partials(const double* part_sum, const unsigned* part_cnt, int nparts, int nbins,
                           double* sums_out, long long* counts_out, cudaStream_t st) {
    reduce_partials_kernel<<<(nbins + 3) / 4, 128, 0, st>>>(part_sum, part_cnt, nparts, nbins, sums_out, counts_out);
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    return 0;
}

__global__ void zero_hist_kernel(double* sums, long long* counts, int nbins) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < nbins) { sums[b] = 0.0; counts[b] = 0; }
}
int launch_zero_hist(double* sums, long long* counts, int nbins, cudaStream_t st) {
    zero_hist_kernel<<<(nbins + 127) / 128, 128, 0, st>>>(sums, counts, nbins);
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- element-wise binning of a full (fftshifted) array ---------------------------------
// A warp reads kUnroll runs of 32 consecutive samples (coalesced); each lane finds its sample's bin and adds
// it to the warp's two-entry bin cache (WarpBinCache): neighbouring samples share a bin most of the time, so
// the shared-memory histogram is touched only when the warp moves on to another bin.
constexpr int kBinThreads = 256;
constexpr int kUnroll = 4;

// (ix, iy, iz) of element o0 + d, given those of o0 (0 <= d < 2^20): 32-bit divisions only
__device__ __forceinline__ void advance_coords(int ix0, int iy0, int iz0, int d, int n1, int n2, int& ix, int& iy, int& iz) {
    const unsigned tz = (unsigned)iz0 + (unsigned)d;
    const unsigned qz = tz / (unsigned)n2;
    iz = (int)(tz - qz * (unsigned)n2);
    const unsigned ty = (unsigned)iy0 + qz;
    const unsigned qy = ty / (unsigned)n1;
    iy = (int)(ty - qy * (unsigned)n1);
    ix = ix0 + (int)qy;
}

template <typename T, int HK, class BinOf>
__device__ __forceinline__ void bin_array(const T* __restrict__ p, int n0, int n1, int n2, int nbins, int warp_private,
                                          double* __restrict__ part_sum, unsigned* __restrict__ part_cnt, double* gsum,
                                          unsigned long long* gcnt, BinOf&& bin_of) {
    // HK == 0: CTA histogram in shared memory -- one copy per warp when that fits (only lane 0 of the owning warp
    // writes it, in program order; warps, then CTAs, are summed in a fixed order => bit-reproducible run to run),
    // otherwise one copy updated with atomics
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NW = kBinThreads / 32;
    const int copies = (HK == 0 && warp_private) ? NW : 1;
    double* hs = reinterpret_cast<double*>(smem_raw);
    unsigned* hc = reinterpret_cast<unsigned*>(hs + (size_t)copies * nbins);
    if (HK == 0) {
        for (int b = threadIdx.x; b < copies * nbins; b += kBinThreads) { hs[b] = 0.0; hc[b] = 0u; }
        __syncthreads();
    }
    const long long total = (long long)n0 * n1 * n2;
    const int lane = threadIdx.x & 31;
    const long long span = 32LL * kUnroll;                                   // samples per warp per trip
    const long long warp0 = (long long)blockIdx.x * (kBinThreads / 32) + (threadIdx.x >> 5);
    const long long nwarps = (long long)gridDim.x * (kBinThreads / 32);
    WarpBinCache acc;
    wbc_init(acc);
    SmemHist sh{hs, hc};
    WarpPrivHist wh{hs + (size_t)(threadIdx.x >> 5) * nbins, hc + (size_t)(threadIdx.x >> 5) * nbins};
    GlobalHist gh{gsum, gcnt};
    for (long long o0 = warp0 * span; o0 < total; o0 += nwarps * span) {     // warp-uniform trip count
        const long long t = o0 / n2;                                         // one 64-bit division per 128 samples
        const int iz0 = (int)(o0 - t * n2);
        const int ix0 = (int)(t / n1), iy0 = (int)(t - (long long)ix0 * n1);
        T v[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long o = o0 + u * 32 + lane;
            v[u] = o < total ? p[o] : (T)0;
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long o = o0 + u * 32 + lane;
            int ix, iy, iz, w = 1;
            advance_coords(ix0, iy0, iz0, u * 32 + lane, n1, n2, ix, iy, iz);
            const int bin = o < total ? bin_of(ix, iy, iz, w) : -1;
            if (HK == 0 && warp_private) wbc_add(acc, wh, bin, (double)v[u], (unsigned)w);
            else if (HK == 0) wbc_add(acc, sh, bin, (double)v[u], (unsigned)w);
            else wbc_add(acc, gh, bin, (double)v[u], (unsigned)w);
        }
    }
    if (HK == 0) {
        if (warp_private) wbc_flush(acc, wh); else wbc_flush(acc, sh);
        __syncthreads();
        for (int b = threadIdx.x; b < nbins; b += kBinThreads) {
            double sacc = 0.0;
            unsigned cacc = 0u;
            for (int w = 0; w < copies; ++w) { sacc += hs[(size_t)w * nbins + b]; cacc += hc[(size_t)w * nbins + b]; }
            part_sum[(size_t)blockIdx.x * nbins + b] = sacc;
            part_cnt[(size_t)blockIdx.x * nbins + b] = cacc;
        }
    } else {
        wbc_flush(acc, gh);
    }
}

template <typename T, int HK>
__global__ void __launch_bounds__(kBinThreads)
bin_only_kernel(const T* __restrict__ p, const __grid_constant__ t21_binspec B, int nbins, int warp_private,
                double* __restrict__ part_sum, unsigned* __restrict__ part_cnt, double* gsum,
                unsigned long long* gcnt) {
    bin_array<T, HK>(p, B.n[0], B.n[1], B.n[2], nbins, warp_private, part_sum, part_cnt, gsum, gcnt,
                     [&](int ix, int iy, int iz, int& w) {
                         const ModeBins mb = classify(B, ix, iy, iz, false);
                         w = mb.own_w;
                         return mb.own;
                     });
}

// ---- (k_perp, k_par) binning of a full fftshifted 3-D array (power_spectrum_2d) -----------------------
// The bin of a mode is separable: the perpendicular bin depends on the two indices across the line of
// sight, the parallel bin on the index along it.  Both are host-built tables (binspec.build_cyl_tables
// runs the reference's own numpy/scipy digitisation on the n_a*n_b + n_nu distinct values), so the
// kernel only looks them up: mode counts are identical to the reference by construction.
template <typename T, int HK>
__global__ void __launch_bounds__(kBinThreads)
bin_cyl_kernel(const T* __restrict__ p, int n0, int n1, int n2, int nu_axis, const int* __restrict__ perp_tab,
               const int* __restrict__ par_tab, int npar, int nbins, int warp_private, double* __restrict__ part_sum,
               unsigned* __restrict__ part_cnt, double* gsum, unsigned long long* gcnt) {
    const int nb_perp_inner = nu_axis == 2 ? n1 : n2;      // length of the second perpendicular axis
    bin_array<T, HK>(p, n0, n1, n2, nbins, warp_private, part_sum, part_cnt, gsum, gcnt,
                     [&](int ix, int iy, int iz, int& w) {
                         const int inu = nu_axis == 0 ? ix : (nu_axis == 1 ? iy : iz);
                         const int ia = nu_axis == 0 ? iy : ix;
                         const int ib = nu_axis == 2 ? iy : iz;
                         const int bp = ldg(perp_tab + (long long)ia * nb_perp_inner + ib);
                         const int bz = ldg(par_tab + inu);
                         w = 1;
                         return (bp >= 0 && bz >= 0) ? bp * npar + bz : -1;
                     });
}

int launch_bin_cyl(const void* p, int dtype, int n0, int n1, int n2, int nu_axis, const int* perp_tab,
                   const int* par_tab, int nperp, int npar, double* part_sum, unsigned* part_cnt, int max_grid,
                   double* gsum, long long* gcnt, int* grid_used, cudaStream_t st) {
    const int nbins = nperp * npar;
    const bool global = nbins > kSmemHistMaxBins;
    const long long total = (long long)n0 * n1 * n2;
    const long long per_cta = (long long)kBinThreads * kUnroll;
    long long g = (total + per_cta - 1) / per_cta;
    const long long cap = (long long)device_sm_count() * 8;
    if (g > cap) g = cap;
    if (!global && g > max_grid) g = max_grid;
    if (g < 1) g = 1;
    // per-warp histogram copies (deterministic sums) while they fit the default 48 KB of dynamic shared memory
    const int wp = (!global && (size_t)(kBinThreads / 32) * nbins * 12 <= 48 * 1024) ? 1 : 0;
    const size_t smem = global ? 0 : (size_t)(wp ? kBinThreads / 32 : 1) * nbins * (sizeof(double) + sizeof(unsigned));
    unsigned long long* gc = reinterpret_cast<unsigned long long*>(gcnt);
    if (dtype == T21_F32) {
        if (global) bin_cyl_kernel<float, 1><<<(int)g, kBinThreads, smem, st>>>((const float*)p, n0, n1, n2, nu_axis, perp_tab, par_tab, npar, nbins, 0, nullptr, nullptr, gsum, gc);
        else bin_cyl_kernel<float, 0><<<(int)g, kBinThreads, smem, st>>>((const float*)p, n0, n1, n2, nu_axis, perp_tab, par_tab, npar, nbins, wp, part_sum, part_cnt, nullptr, nullptr);
    } else {
        if (global) bin_cyl_kernel<double, 1><<<(int)g, kBinThreads, smem, st>>>((const double*)p, n0, n1, n2, nu_axis, perp_tab, par_tab, npar, nbins, 0, nullptr, nullptr, gsum, gc);
        else bin_cyl_kernel<double, 0><<<(int)g, kBinThreads, smem, st>>>((const double*)p, n0, n1, n2, nu_axis, perp_tab, par_tab, npar, nbins, wp, part_sum, part_cnt, nullptr, nullptr);
    }
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    *grid_used = (int)g;
    return 0;
}

int launch_bin_only(const void* p, int dtype, const t21_binspec& spec, double* part_sum, unsigned* part_cnt,
                    int max_grid, double* gsum, long long* gcnt, int* grid_used, cudaStream_t st) {
    const int nbins = spec.nk * (spec.nmu > 0 ? spec.nmu : 1);
    const bool global = nbins > kSmemHistMaxBins;   // then gsum/gcnt are the pre-zeroed final outputs
    const long long total = (long long)spec.n[0] * spec.n[1] * spec.n[2];
    const long long per_cta = (long long)kBinThreads * kUnroll;
    long long g = (total + per_cta - 1) / per_cta;
    const long long cap = (long long)device_sm_count() * 8;
    if (g > cap) g = cap;
    if (!global && g > max_grid) g = max_grid;
    if (g < 1) g = 1;
    // per-warp histogram copies (deterministic sums) while they fit the default 48 KB of dynamic shared memory
    const int wp = (!global && (size_t)(kBinThreads / 32) * nbins * 12 <= 48 * 1024) ? 1 : 0;
    const size_t smem = global ? 0 : (size_t)(wp ? kBinThreads / 32 : 1) * nbins * (sizeof(double) + sizeof(unsigned));
    unsigned long long* gc = reinterpret_cast<unsigned long long*>(gcnt);
    if (dtype == T21_F32) {
        if (global) bin_only_kernel<float, 1><<<(int)g, kBinThreads, smem, st>>>((const float*)p, spec, nbins, 0, nullptr, nullptr, gsum, gc);
        else bin_only_kernel<float, 0><<<(int)g, kBinThreads, smem, st>>>((const float*)p, spec, nbins, wp, part_sum, part_cnt, nullptr, nullptr);
    } else {
        if (global) bin_only_kernel<double, 1><<<(int)g, kBinThreads, smem, st>>>((const double*)p, spec, nbins, 0, nullptr, nullptr, gsum, gc);
        else bin_only_kernel<double, 0><<<(int)g, kBinThreads, smem, st>>>((const double*)p, spec, nbins, wp, part_sum, part_cnt, nullptr, nullptr);
    }
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    *grid_used = (int)g;
    return 0;
}

// ---- row-wise binning of a half-spectrum power array (t21_ps_half) ------------------------------------------------
// p[n0][n1][pitch] holds nzh = nz/2+1 valid columns per (x, y) row, FFT order.  One WARP walks one row at a time, lane l
// reading columns l, l+32, ... (coalesced), in blocks of consecutive rows.  Along a row only part of the bin index
// changes, and a lane always meets the same columns, so every lane adds into its OWN column of a small shared-memory
// table acc[bin part][lane] -- no atomics, no votes, no cache to thrash (the element-wise kernels above keep a
// two-entry warp cache, which a row that crosses ten bins evicts ten times: 4.0 ms for 2.2 GB at 1024^3).  After a
// block of rows the touched table rows are summed over the lanes (shuffles, float64) into the warp's own histogram;
// warps and CTAs are then summed in a fixed order, so the result is bit-reproducible.
constexpr int kRowWarps = 4;                      // warps per CTA of the row kernels
constexpr int kRowThreads = kRowWarps * 32;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned warp_sum(unsigned v) { return __reduce_add_sync(0xffffffffu, v); }

// (k_perp, k_par) bins (power_spectrum.py:177-254).  The bin is separable: `zpart` changes along the row, `rpart` is
// fixed per row.  nu_axis = 2: zpart = par_tab[iz], rpart = perp_tab[ix][iy]; nu_axis = 1: zpart = perp_tab[ix][iz],
// rpart = par_tab[iy]; nu_axis = 0: zpart = perp_tab[iy][iz], rpart = par_tab[ix].  Mode counts are not formed here:
// for separable bins they are (index pairs in the perp bin) x (indices in the par bin), exact, on the host.
template <typename T, typename A>      // A: type of the lane-owned partial sums (float for float32 cubes, double for float64)
__global__ void __launch_bounds__(kRowThreads)
cyl_rows_kernel(const T* __restrict__ p, int n0, int n1, int nz, int pitch, int nu_axis, const int* __restrict__ perp_tab,
                const int* __restrict__ par_tab, int nperp, int npar, int rows_per_block, double* __restrict__ part_sum) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nzh = nz / 2 + 1, nbins = nperp * npar;
    const int nzp = nu_axis == 2 ? npar : nperp;                       // values of the part that changes along z
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* hist = reinterpret_cast<double*>(smem_raw) + (size_t)warp * nbins;                        // [nbins]
    A* acc = reinterpret_cast<A*>(smem_raw + (size_t)kRowWarps * nbins * sizeof(double)) + (size_t)warp * nzp * 32;
    for (int b = lane; b < nbins; b += 32) hist[b] = 0.0;
    for (int b = lane; b < nzp * 32; b += 32) acc[b] = (A)0;
    __syncwarp();
    const long long nrows = (long long)n0 * n1;
    const long long nblocks = (nrows + rows_per_block - 1) / rows_per_block;
    const long long gw = (long long)blockIdx.x * kRowWarps + warp, nw = (long long)gridDim.x * kRowWarps;
    int cur_r = -1;                 // row part of the rows accumulated in acc so far
    unsigned long long touched = 0ull;   // z parts this lane has added to (nzp <= 64)
    auto flush = [&]() {
        const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)touched), hi = __reduce_or_sync(0xffffffffu, (unsigned)(touched >> 32));
        unsigned long long t = ((unsigned long long)hi << 32) | lo;
        while (t) {
            const int zb = __ffsll((long long)t) - 1;
            t &= t - 1;
            const double v = warp_sum((double)acc[zb * 32 + lane]);
            acc[zb * 32 + lane] = (A)0;
            if (lane == 0) hist[nu_axis == 2 ? cur_r * npar + zb : zb * npar + cur_r] += v;
        }
        touched = 0ull;
    };
    for (long long blk = gw; blk < nblocks; blk += nw) {
        const long long r0 = blk * rows_per_block;
        const long long r1 = r0 + rows_per_block < nrows ? r0 + rows_per_block : nrows;
        for (long long r = r0; r < r1; ++r) {
            const int ix = (int)(r / n1), iy = (int)(r - (long long)ix * n1);
            const int rp = nu_axis == 2 ? ldg(perp_tab + (long long)ix * n1 + iy) : ldg(par_tab + (nu_axis == 1 ? iy : ix));
            if (rp < 0) continue;                                      // the whole row is outside the bins
            if (rp != cur_r) { if (cur_r >= 0) flush(); cur_r = rp; }
            const T* row = p + r * pitch;
            const int* ztab = nu_axis == 2 ? par_tab : perp_tab + (long long)(nu_axis == 1 ? ix : iy) * nz;
            for (int iz = lane; iz < nzh; iz += 32) {
                const int zb = ldg(ztab + iz);
                if (zb < 0) continue;
                const A w = (iz != 0 && iz != nz - iz) ? (A)2 : (A)1;      // the mode stands for its Hermitian twin too
                acc[zb * 32 + lane] += (A)row[iz] * w;
                touched |= 1ull << zb;
            }
        }
        if (cur_r >= 0) flush();
        cur_r = -1;
    }
    __syncthreads();
    double* h0 = reinterpret_cast<double*>(smem_raw);
    for (int b = threadIdx.x; b < nbins; b += kRowThreads) {
        double sacc = 0.0;
        for (int w = 0; w < kRowWarps; ++w) sacc += h0[(size_t)w * nbins + b];
        part_sum[(size_t)blockIdx.x * nbins + b] = sacc;
    }
}

// one warp per output: lane l adds parts l, l+32, ... in order, then a fixed shuffle tree (deterministic)
__global__ void reduce_sums_kernel(const double* __restrict__ part, int nparts, int n, double* __restrict__ out) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    double s = 0.0;
    for (int q = lane; q < nparts; q += 32) s += part[(size_t)q * n + i];
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if (lane == 0) out[i] = s;
}

int launch_cyl_rows(const void* p, int dtype, int n0, int n1, int nz, int pitch, int nu_axis, const int* perp_tab, const int* par_tab,
                    int nperp, int npar, double* part_sum, int max_grid, double* sums_out, cudaStream_t st) {
    const int nbins = nperp * npar, nzp = nu_axis == 2 ? npar : nperp;
    if (nzp > 64) { set_error("at most 64 bins along the row direction (got %d)", nzp); return T21_EINVAL; }
    const size_t smem = (size_t)kRowWarps * (nbins * sizeof(double) + (size_t)nzp * 32 * (dtype == T21_F32 ? sizeof(float) : sizeof(double)));
    if (smem > 200 * 1024) { set_error("too many (k_perp, k_par) bins (%d) for the row kernel", nbins); return T21_EINVAL; }
    const int rows_per_block = 16;
    long long g = ((long long)n0 * n1 + (long long)rows_per_block * kRowWarps - 1) / ((long long)rows_per_block * kRowWarps);
    const long long cap = (long long)device_sm_count() * 8;
    if (g > cap) g = cap;
    if (g > max_grid) g = max_grid;
    if (g < 1) g = 1;
    if (dtype == T21_F32) {
        auto kern = cyl_rows_kernel<float, float>;
        if (int rc = ensure_smem(kern, smem)) return rc;
        kern<<<(int)g, kRowThreads, smem, st>>>((const float*)p, n0, n1, nz, pitch, nu_axis, perp_tab, par_tab, nperp, npar, rows_per_block, part_sum);
    } else {
        auto kern = cyl_rows_kernel<double, double>;
        if (int rc = ensure_smem(kern, smem)) return rc;
        kern<<<(int)g, kRowThreads, smem, st>>>((const double*)p, n0, n1, nz, pitch, nu_axis, perp_tab, par_tab, nperp, npar, rows_per_block, part_sum);
    }
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    reduce_sums_kernel<<<(nbins + 3) / 4, 128, 0, st>>>(part_sum, (int)g, nbins, sums_out);
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- multipole moments of a half-spectrum power array (power_legendre.py:4-90) ----------------------------------
// Per k bin (half-open, `(k >= kmin) & (k < kmax)` :78): sum P, sum P L2, sum P L4, sum L2^2, sum L4^2 and the mode
// count, with L2 = (3 mu^2 - 1)/2 and L4 = 4.375 (mu^2 - 0.115587)(mu^2 - 0.741556) as the reference writes them
// (:57-60), mu = k_los / k.  Even in mu, so a stored mode stands for its Hermitian twin with weight 2.  The k bin is
// decided exactly (classify()); for float32 cubes mu^2 comes from the fp32 tables (1e-7 relative: it is a weight,
// not a decision), for float64 cubes from the reference's own float64 sequence.  Modes with k < 0.001 have mu = NaN
// in the reference (:501): their L2 / L4 terms are NaN here too.  Same row-wise scheme as cyl_rows_kernel with
// lane-owned tables acc[k bin][5][lane] (A = float for float32 cubes, double for float64) and cnt[k bin][lane].
constexpr int kMom = 5;
// first iz in [0, nzh] with ((sx + sy) + sz[iz]) >= thr: s only grows along the stored half of a z row (rounding is
// monotonic), so the k bins of a row are index ranges; bisection with the reference's own float64 sum (:477)
__device__ __forceinline__ int row_boundary(const double* __restrict__ tz, int nzh, double sxy, double thr) {
    int lo = 0, hi = nzh;
    while (__any_sync(0xffffffffu, lo < hi)) {
        const int mid = (lo + hi) >> 1;
        const bool ge = dadd_rn(sxy, ldg(tz + (mid < nzh ? mid : nzh - 1))) >= thr;
        if (lo < hi) { if (ge) hi = mid; else lo = mid + 1; }
    }
    return lo;
}

// BOUND: nk <= 31, one lane per bin edge finds where the edge falls along the row (row_boundary); a step of 32 columns
// no edge falls into takes its bin from two votes, the others compare per lane.  Otherwise classify() per element.
template <typename T, typename A, bool BOUND>
__global__ void __launch_bounds__(kRowThreads)
multipole_rows_kernel(const T* __restrict__ p, const __grid_constant__ t21_binspec B, int pitch, int los, int exclude_zero,
                      int rows_per_block, double* __restrict__ part_sum, unsigned* __restrict__ part_cnt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nk = B.nk, n0 = B.n[0], n1 = B.n[1], nz = B.n[2], nzh = nz / 2 + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // layout: hist[warps][nk][5] f64 | hcnt[warps][nk] u64 | acc[warps][nk][5][32] A | cnt[warps][nk][32] u32
    double* hist0 = reinterpret_cast<double*>(smem_raw);
    unsigned long long* hcnt0 = reinterpret_cast<unsigned long long*>(hist0 + (size_t)kRowWarps * nk * kMom);
    A* acc0 = reinterpret_cast<A*>(hcnt0 + (size_t)kRowWarps * nk);
    unsigned* cnt0 = reinterpret_cast<unsigned*>(acc0 + (size_t)kRowWarps * nk * kMom * 32);
    double* hist = hist0 + (size_t)warp * nk * kMom;
    unsigned long long* hcnt = hcnt0 + (size_t)warp * nk;
    A* acc = acc0 + (size_t)warp * nk * kMom * 32;
    unsigned* cnt = cnt0 + (size_t)warp * nk * 32;
    for (int b = lane; b < nk * kMom; b += 32) hist[b] = 0.0;
    for (int b = lane; b < nk; b += 32) hcnt[b] = 0ull;
    for (int b = lane; b < nk * kMom * 32; b += 32) acc[b] = (A)0;
    for (int b = lane; b < nk * 32; b += 32) cnt[b] = 0u;
    __syncwarp();
    const long long nrows = (long long)n0 * n1;
    const long long nblocks = (nrows + rows_per_block - 1) / rows_per_block;
    const long long gw = (long long)blockIdx.x * kRowWarps + warp, nw = (long long)gridDim.x * kRowWarps;
    const float s_nan_f = (float)B.s_nan;
    unsigned long long touched = 0ull;                                  // nk <= 64
    for (long long blk = gw; blk < nblocks; blk += nw) {
        const long long r0 = blk * rows_per_block;
        const long long r1 = r0 + rows_per_block < nrows ? r0 + rows_per_block : nrows;
        for (long long r = r0; r < r1; ++r) {
            const int ix = (int)(r / n1), iy = (int)(r - (long long)ix * n1);
            const T* row = p + r * pitch;
            int m = 1 << 30, nle = 0;
            if (BOUND) {
                const double sxy = dadd_rn(ldg(B.tab_s[0] + ix), ldg(B.tab_s[1] + iy));
                const int mm = row_boundary(B.tab_s[2], nzh, sxy, ldg(B.thr + (lane <= nk ? lane : nk)));
                if (lane <= nk) m = mm;
                nle = __popc(__ballot_sync(0xffffffffu, m <= 0));
            }
            constexpr int UN = 6;                                       // row values in flight per lane (the kernel is
            for (int izc = 0; izc < nzh; izc += 32 * UN) {              // latency bound with one load at a time: 6.2 ms)
                T vals[UN];
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const int izl = izc + 32 * u + lane;
                    vals[u] = izl < nzh ? row[izl] : (T)0;
                }
#pragma unroll
                for (int u = 0; u < UN; ++u) {
                    const int iz0 = izc + 32 * u;
                    if (iz0 >= nzh) break;                              // warp-uniform
                    const int iz = iz0 + lane;
                    int kb_b = -1;
                    if (BOUND) {                                        // (every lane takes part in the votes and shuffles)
                        const int nhi = __popc(__ballot_sync(0xffffffffu, m <= iz0 + 31));
                        int ne = nle;
                        for (int e = nle; e < nhi; ++e) ne += iz >= __shfl_sync(0xffffffffu, m, e) ? 1 : 0;
                        nle = nhi;
                        kb_b = (ne >= 1 && ne <= nk) ? ne - 1 : -1;
                    }
                    if (iz >= nzh) continue;
                    if (exclude_zero) {                                 // the k_perp = 0 line (power_spectrum.py:563-575)
                        const int pa = los == 0 ? iy : ix, pb = los == 2 ? iy : iz;
                        const int za = los == 0 ? B.zero_idx[1] : B.zero_idx[0], zb = los == 2 ? B.zero_idx[1] : B.zero_idx[2];
                        if (pa == za && pb == zb) continue;
                    }
                    const bool has_twin = (iz != 0) && (iz != nz - iz);
                    ModeBins mb;
                    if (BOUND) { mb.own = kb_b; mb.own_w = has_twin ? 2 : 1; mb.twin = -1; }
                    else mb = classify(B, ix, iy, iz, has_twin);
                    const int kb = mb.own;
                    if (kb < 0) continue;
                    const int il = los == 0 ? ix : (los == 1 ? iy : iz);
                    A mu2;
                    if (sizeof(A) == 8) {                               // the reference's own sequence (:498)
                        const double sd = dadd_rn(dadd_rn(ldg(B.tab_s[0] + ix), ldg(B.tab_s[1] + iy)), ldg(B.tab_s[2] + iz));
                        const double mu = ldg(B.tab_los + il) / sqrt(sd);
                        mu2 = (A)(sd >= B.s_nan ? mu * mu : nan(""));
                    } else {
                        const float sh = (ldg(B.tab_sf[0] + ix) + ldg(B.tab_sf[1] + iy)) + ldg(B.tab_sf[2] + iz);
                        const float kl = ldg(B.tab_losf + il);
                        mu2 = (A)(sh >= s_nan_f ? kl * kl / sh : nanf(""));
                    }
                    const A w = (A)mb.own_w, pv = (A)vals[u] * w;
                    const A l2 = (A)0.5 * ((A)3.0 * mu2 - (A)1.0);
                    const A l4 = (A)4.375 * (mu2 - (A)0.115587) * (mu2 - (A)0.741556);
                    A* a = acc + (size_t)kb * kMom * 32 + lane;
                    a[0] += pv; a[32] += pv * l2; a[64] += pv * l4; a[96] += w * l2 * l2; a[128] += w * l4 * l4;
                    cnt[kb * 32 + lane] += (unsigned)mb.own_w;
                    touched |= 1ull << kb;
                }
            }
        }
        const unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)touched), hi = __reduce_or_sync(0xffffffffu, (unsigned)(touched >> 32));
        unsigned long long t = ((unsigned long long)hi << 32) | lo;
        while (t) {
            const int kb = __ffsll((long long)t) - 1;
            t &= t - 1;
#pragma unroll
            for (int q = 0; q < kMom; ++q) {
                const double v = warp_sum((double)acc[((size_t)kb * kMom + q) * 32 + lane]);
                acc[((size_t)kb * kMom + q) * 32 + lane] = (A)0;
                if (lane == 0) hist[kb * kMom + q] += v;
            }
            const unsigned c = warp_sum(cnt[kb * 32 + lane]);
            cnt[kb * 32 + lane] = 0u;
            if (lane == 0) hcnt[kb] += c;
        }
        touched = 0ull;
    }
    __syncthreads();
    for (int b = threadIdx.x; b < nk * kMom; b += kRowThreads) {
        double sacc = 0.0;
        for (int w = 0; w < kRowWarps; ++w) sacc += hist0[(size_t)w * nk * kMom + b];
        part_sum[(size_t)blockIdx.x * nk * kMom + b] = sacc;
    }
    for (int b = threadIdx.x; b < nk; b += kRowThreads) {
        unsigned long long cacc = 0ull;
        for (int w = 0; w < kRowWarps; ++w) cacc += hcnt0[(size_t)w * nk + b];
        part_cnt[(size_t)blockIdx.x * nk + b] = (unsigned)cacc;     // < 2^32 modes per CTA (checked by the launcher)
    }
}

// sums_out[nk][5], counts_out[nk]; part buffers sized for max_grid CTAs
int launch_bin_multipoles(const void* p, int dtype, const t21_binspec& spec, int pitch, int los, int exclude_zero,
                          double* part_sum, unsigned* part_cnt, int max_grid, int* grid_used, cudaStream_t st) {
    const int nk = spec.nk;
    if (nk > 64) { set_error("at most 64 k bins for the multipole kernel (got %d)", nk); return T21_EINVAL; }
    const size_t asz = dtype == T21_F32 ? sizeof(float) : sizeof(double);
    const size_t smem = (size_t)kRowWarps * ((size_t)nk * kMom * 8 + (size_t)nk * 8 + (size_t)nk * kMom * 32 * asz + (size_t)nk * 32 * 4);
    if (smem > 220 * 1024) { set_error("too many k bins (%d) for the multipole kernel", nk); return T21_EINVAL; }
    const int rows_per_block = 8;
    const long long nrows = (long long)spec.n[0] * spec.n[1];
    long long g = (nrows + (long long)rows_per_block * kRowWarps - 1) / ((long long)rows_per_block * kRowWarps);
    const long long cap = (long long)device_sm_count() * 4;
    if (g > cap) g = cap;
    if (g > max_grid) g = max_grid;
    if (g < 1) g = 1;
    if ((double)nrows * spec.n[2] / (double)g > 4.0e9) { set_error("cube too large for 32-bit per-CTA mode counts"); return T21_EINVAL; }
#define T21_MP_LAUNCH(TT, BB)                                                                                   \
    {                                                                                                           \
        auto kern = multipole_rows_kernel<TT, TT, BB>;                                                          \
        if (int rc = ensure_smem(kern, smem)) return rc;                                                        \
        kern<<<(int)g, kRowThreads, smem, st>>>((const TT*)p, spec, pitch, los, exclude_zero, rows_per_block, part_sum, part_cnt); \
    }
    const bool bound = nk <= 31;
    if (dtype == T21_F32) { if (bound) T21_MP_LAUNCH(float, true) else T21_MP_LAUNCH(float, false) }
    else { if (bound) T21_MP_LAUNCH(double, true) else T21_MP_LAUNCH(double, false) }
#undef T21_MP_LAUNCH
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    *grid_used = (int)g;
    return 0;
}

// sums over CTAs of the [grid][nk][5] partials, in order (deterministic)
__global__ void reduce_moments_kernel(const double* __restrict__ part_sum, const unsigned* __restrict__ part_cnt, int nparts,
                                      int nk, double* __restrict__ sums, long long* __restrict__ counts) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i < nk * kMom) {
        double s = 0.0;
        for (int q = lane; q < nparts; q += 32) s += part_sum[(size_t)q * nk * kMom + i];
        for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
        if (lane == 0) sums[i] = s;
    }
    if (i < nk) {
        long long c = 0;
        for (int q = lane; q < nparts; q += 32) c += part_cnt[(size_t)q * nk + i];
        for (int o = 16; o > 0; o >>= 1) c += __shfl_down_sync(0xffffffffu, c, o);
        if (lane == 0) counts[i] = c;
    }
}
int launch_reduce_moments(const double* part_sum, const unsigned* part_cnt, int nparts, int nk, double* sums_out,
                          long long* counts_out, cudaStream_t st) {
    reduce_moments_kernel<<<(nk * kMom + 3) / 4, 128, 0, st>>>(part_sum, part_cnt, nparts, nk, sums_out, counts_out);
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- pencil segment table (t21_build_pencil_segments) ------------------------------------------------------------------
// one thread per pencil of the computed half spectrum; see build_pencil_segments() in t21_bin.cuh
__global__ void __launch_bounds__(128)
pencil_segments_kernel(const __grid_constant__ t21_binspec B, int which, unsigned* __restrict__ out, int stride, int* __restrict__ max_seg) {
    const int nzh = B.n[2] / 2 + 1;
    const long long np = (long long)B.n[1] * nzh;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int s = 0;
    if (i < np) s = build_pencil_segments(B, (int)(i / nzh), (int)(i % nzh), which, out + i * stride, stride);
    const unsigned novf = __popc(__ballot_sync(0xffffffffu, s > stride));
    for (int o = 16; o > 0; o >>= 1) { const int t = __shfl_xor_sync(0xffffffffu, s, o); s = t > s ? t : s; }
    if ((threadIdx.x & 31) == 0 && s > 0) { atomicMax(max_seg, s); if (novf) atomicAdd(max_seg + 1, (int)novf); }
}

__global__ void __launch_bounds__(128)
pencil_segments_cyl_kernel(int n0, int n1, int n2, int nu_axis, const int* __restrict__ perp_tab, const int* __restrict__ par_tab,
                           int npar, unsigned* __restrict__ out, int stride, int* __restrict__ max_seg) {
    const int nzh = n2 / 2 + 1;
    const long long np = (long long)n1 * nzh;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int s = 0;
    if (i < np) s = build_pencil_segments_cyl(n0, n1, n2, nu_axis, perp_tab, par_tab, npar, (int)(i / nzh), (int)(i % nzh), out + i * stride, stride);
    const unsigned novf = __popc(__ballot_sync(0xffffffffu, s > stride));
    for (int o = 16; o > 0; o >>= 1) { const int t = __shfl_xor_sync(0xffffffffu, s, o); s = t > s ? t : s; }
    if ((threadIdx.x & 31) == 0 && s > 0) { atomicMax(max_seg, s); if (novf) atomicAdd(max_seg + 1, (int)novf); }
}

int launch_pencil_segments_cyl(int n0, int n1, int n2, int nu_axis, const int* perp_tab, const int* par_tab, int npar,
                               unsigned* out, int stride, int* max_seg_dev, cudaStream_t st) {
    const long long np = (long long)n1 * (n2 / 2 + 1);
    pencil_segments_cyl_kernel<<<(unsigned)((np + 127) / 128), 128, 0, st>>>(n0, n1, n2, nu_axis, perp_tab, par_tab, npar, out, stride, max_seg_dev);
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_pencil_segments(const t21_binspec& spec, int which, unsigned* out, int stride, int* max_seg_dev, cudaStream_t st) {
    const long long np = (long long)spec.n[1] * (spec.n[2] / 2 + 1);
    pencil_segments_kernel<<<(unsigned)((np + 127) / 128), 128, 0, st>>>(spec, which, out, stride, max_seg_dev);
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace t21
