// __global__ wrappers that run the phase-structured pass bodies, and their launchers.
#pragma once
#include <stdlib.h>
#include "t21_internal.h"
#include "t21_passes.cuh"

namespace t21 {

template <class K, typename T, int PH = 0>
__device__ __forceinline__ void dev_phases(const typename K::Params& p, long long cta, int tid,
                                           typename K::Regs& r, cx<T>* smem) {
    K::template phase<PH>(p, cta, tid, r, smem);
    if constexpr (PH + 1 < K::NPHASE) {
        __syncthreads();
        dev_phases<K, T, PH + 1>(p, cta, tid, r, smem);
    }
}
template <class K, typename T, class Hist, int PH = 0>
__device__ __forceinline__ void dev_phases_hist(const typename K::Params& p, long long tile, int tid,
                                                typename K::Regs& r, cx<T>* smem, Hist& h) {
    K::template phase<PH>(p, tile, tid, r, smem, h);
    if constexpr (PH + 1 < K::NPHASE) {
        __syncthreads();
        dev_phases_hist<K, T, Hist, PH + 1>(p, tile, tid, r, smem, h);
    }
}

// one CTA per tile (Z and Y passes)
template <class K, typename T>
__global__ void __launch_bounds__(K::THREADS, K::MIN_CTAS) tile_kernel(const __grid_constant__ typename K::Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    typename K::Regs r;
    dev_phases<K, T>(p, (long long)blockIdx.x, (int)threadIdx.x, r, reinterpret_cast<cx<T>*>(smem_raw));
}

#ifndef T21_X_MIN_CTAS
#define T21_X_MIN_CTAS 2
#endif
// cross spectra keep the first field's 16 results per thread while the second is transformed: at 64 registers
// that spills ~700 bytes per thread, so they are compiled for one CTA per SM (1024^3 cross: 11.6 -> 10.7 ms)
#ifndef T21_X_NF2_WIDE
#define T21_X_NF2_WIDE 1
#endif
// CTAs per SM the X pass is compiled for (i.e. its register budget): 64 registers per thread for fp32
// pencils of up to 512 threads, 128 for fp64 (16 complex doubles are 64 registers by themselves)
template <typename T>
constexpr int x_min_ctas(int threads, int e, int nf) {
    if (threads * e * (T21_X_NF2_WIDE ? nf : 1) > 512 * 16) return 1;
    const int by_threads = 1024 / threads < 1 ? 1 : 1024 / threads;      // 1024 resident threads per SM
    const int want = sizeof(T) == 4 ? by_threads : by_threads / 2;
    return want < 1 ? 1 : (want > 2 * T21_X_MIN_CTAS ? 2 * T21_X_MIN_CTAS : want);
}

struct NoHist {
    __device__ void add(int, double, unsigned) {}
};

// persistent X pass: each CTA walks tiles blockIdx.x, +gridDim.x, ... and keeps its histogram
// in shared memory until the end (HK = 0), accumulates into global memory (HK = 1), or has no
// histogram at all (HK = 2, n-D store).
template <class K, typename T, int HK>
__global__ void __launch_bounds__(K::THREADS, (x_min_ctas<T>(K::THREADS, K::E, K::kNF)))
xpass_kernel(const __grid_constant__ typename K::Params p, int nbins, int warp_private,
             double* __restrict__ part_sum, unsigned* __restrict__ part_cnt, double* gsum, unsigned long long* gcnt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<T>* ex = reinterpret_cast<cx<T>*>(smem_raw);
    const int tid = threadIdx.x;
    typename K::Regs r;
#pragma unroll
    for (int c = 0; c < K::NACC; ++c) acc_init(r.acc[c]);
    if constexpr (HK == 0) {
        // shared-memory histogram: one copy per warp when that fits (plain adds by lane 0, summed
        // in warp order at the end => bit-reproducible), otherwise one copy updated with atomics
        constexpr int NW = (K::THREADS + 31) / 32;
        const int copies = warp_private ? NW : 1;
        double* hs = reinterpret_cast<double*>(smem_raw + K::SMEM_BYTES);
        unsigned* hc = reinterpret_cast<unsigned*>(hs + (size_t)copies * nbins);
        for (int b = tid; b < copies * nbins; b += K::THREADS) { hs[b] = 0.0; hc[b] = 0u; }
        __syncthreads();
        if (warp_private) {
            WarpPrivHist hist{hs + (size_t)(tid >> 5) * nbins, hc + (size_t)(tid >> 5) * nbins};
            for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x)
                { dev_phases_hist<K, T, WarpPrivHist>(p, tile, tid, r, ex, hist); __syncthreads(); }
#pragma unroll
            for (int c = 0; c < K::NACC; ++c) acc_flush(r.acc[c], hist, nbins);
        } else {
            SmemHist hist{hs, hc};
            for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x)
                { dev_phases_hist<K, T, SmemHist>(p, tile, tid, r, ex, hist); __syncthreads(); }
#pragma unroll
            for (int c = 0; c < K::NACC; ++c) acc_flush(r.acc[c], hist, nbins);
        }
        __syncthreads();
        for (int b = tid; b < nbins; b += K::THREADS) {
            double s = 0.0;
            unsigned c = 0;
            for (int w = 0; w < copies; ++w) { s += hs[(size_t)w * nbins + b]; c += hc[(size_t)w * nbins + b]; }
            part_sum[(size_t)blockIdx.x * nbins + b] = s;
            part_cnt[(size_t)blockIdx.x * nbins + b] = c;
        }
    } else if constexpr (HK == 1) {
        GlobalHist hist{gsum, gcnt};
        for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x)
            { dev_phases_hist<K, T, GlobalHist>(p, tile, tid, r, ex, hist); __syncthreads(); }
#pragma unroll
        for (int c = 0; c < K::NACC; ++c) acc_flush(r.acc[c], hist, nbins);
    } else {
        NoHist hist;
        for (long long tile = blockIdx.x; tile < p.ntiles; tile += gridDim.x)
            { dev_phases_hist<K, T, NoHist>(p, tile, tid, r, ex, hist); __syncthreads(); }
    }
}

template <class Kern>
inline int ensure_smem(Kern kern, size_t bytes) {
    if (bytes > 48 * 1024) {
        T21_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    }
    return 0;
}

template <class K, typename T>
inline int launch_tiles(const typename K::Params& p, long long nctas, cudaStream_t st) {
    if (nctas <= 0) return 0;
    if (nctas > 2147483647LL) { set_error("grid too large"); return T21_EINVAL; }
    size_t smem = K::SMEM_BYTES;
    auto kern = tile_kernel<K, T>;
    if (int rc = ensure_smem(kern, smem)) return rc;
    kern<<<(unsigned)nctas, K::THREADS, smem, st>>>(p);
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    return 0;
}

// T21_X_SEGMENTS=0: the CTA-per-tile X pass classifies per mode even when the spec carries a segment table (A/B runs)
inline bool x_use_segments() {
    static const bool on = [] { const char* e = getenv("T21_X_SEGMENTS"); return !(e && e[0] == '0'); }();
    return on;
}

template <class K, typename T>
inline int launch_x(XArgs& a, cudaStream_t st) {
    typename K::Params p;
    memset(&p, 0, sizeof(p));
    p.w[0] = (const cx<T>*)a.w[0];
    p.w[1] = (const cx<T>*)a.w[1];
    p.ny = a.ny; p.nz = a.nz; p.nzh = a.nzh; p.pitch = a.pitch; p.iy0 = a.iy0;
    p.xs_log2 = a.xs_log2 >= 0 ? a.xs_log2 : ilog2_rt(a.nx);
    constexpr int WZ = K::THREADS / K::S::TP;
    p.nchunks = (a.nzh + WZ - 1) / WZ;
    constexpr int PARTS = WZ < WLay<T>::LW ? WLay<T>::LW / WZ : 1;     // see XPass::phase
    p.ntiles = (long long)a.ny * ((p.nchunks + PARTS - 1) / PARTS) * PARTS;
    p.tw = (const cx<T>*)a.tw;
    p.offset[0] = (const T*)a.offset[0];
    p.offset[1] = (const T*)a.offset[1];
    p.ntot = a.ntot;
    if (a.bins) p.bins = *a.bins;
    if constexpr (K::kMode == MODE_BIN) {
        // bins as index ranges from the spec's pencil segment table, when it carries one this kernel can use
        const bool is_signed = a.bins && a.bins->nmu > 0 && !a.bins->absolute_mus;
        if (K::kSegOk && a.bins && a.bins->seg_tab && a.bins->seg_overflow == 0 && a.bins->seg_ntab == (is_signed ? 4 : 1) && x_use_segments()) {
            p.seg = a.bins->seg_tab;
            p.seg_stride = a.bins->seg_stride;
            p.seg_ntab = a.bins->seg_ntab;
            p.seg_pencils = (long long)a.bins->n[1] * a.nzh;
        }
    }
    p.nd_out = a.nd_out; p.nd_f64 = a.nd_f64; p.nd_half = a.nd_half; p.scale = a.scale;

    const bool is_bin = a.bins != nullptr;
    const int hk = !is_bin ? 2 : (a.hist_global ? 1 : 0);
    size_t smem = K::SMEM_BYTES;
    // per-warp histogram copies when they are small (and the CTA is made of whole warps)
    constexpr int NW = (K::THREADS + 31) / 32;
    const int warp_private = (hk == 0 && K::THREADS % 32 == 0 &&
                              (size_t)NW * a.nbins * (sizeof(double) + sizeof(unsigned)) <= 16 * 1024) ? 1 : 0;
    if (hk == 0) smem += (size_t)(warp_private ? NW : 1) * a.nbins * (sizeof(double) + sizeof(unsigned));
    int occ = 0, grid = 0;
#define T21_LAUNCH_HK(HKV)                                                                              \
    {                                                                                                   \
        auto kern = xpass_kernel<K, T, HKV>;                                                            \
        if (int rc = ensure_smem(kern, smem)) return rc;                                                \
        T21_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, K::THREADS, smem));       \
        if (occ < 1) { set_error("x pass does not fit on an SM (smem %zu)", smem); return T21_EINVAL; } \
        long long g = (long long)occ * device_sm_count();                                               \
        if (g > p.ntiles) g = p.ntiles;                                                                 \
        if (HKV == 0 && g > a.max_grid) g = a.max_grid;                                                 \
        grid = (int)g;                                                                                  \
        p.tile_step = grid;                                                                             \
        kern<<<grid, K::THREADS, smem, st>>>(p, a.nbins, warp_private, a.part_sum, a.part_cnt, a.gsum, a.gcnt);        \
    }
    if constexpr (K::kMode == MODE_BIN) {
        if (hk == 0) T21_LAUNCH_HK(0) else if (hk == 1) T21_LAUNCH_HK(1)
        else { set_error("binning pass called without a bin spec"); return T21_EINVAL; }
    } else {
        if (hk != 2) { set_error("n-D pass called with a bin spec"); return T21_EINVAL; }
        T21_LAUNCH_HK(2)
    }
#undef T21_LAUNCH_HK
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    a.grid_used = grid;
    return 0;
}

// switch over the instantiated pencil lengths
#define T21_FOR_FAST_N(n, MACRO)          \
    switch (n) {                          \
        case 8: MACRO(8); break;          \
        case 16: MACRO(16); break;        \
        case 32: MACRO(32); break;        \
        case 64: MACRO(64); break;        \
        case 128: MACRO(128); break;      \
        case 256: MACRO(256); break;      \
        case 512: MACRO(512); break;      \
        case 1024: MACRO(1024); break;    \
        case 2048: MACRO(2048); break;    \
        case 4096: MACRO(4096); break;    \
        default: set_error("length %d is not instantiated", n); return T21_EINVAL; \
    }

}  // namespace t21
