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

}  // namespace t21
