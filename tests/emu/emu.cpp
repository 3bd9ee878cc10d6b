// CPU replay of the CUDA pass bodies (test infrastructure, never linked into the product).
//
// The kernels in tools21cm_b200/csrc are written as barrier-separated phases over
// host/device templates.  This file runs the same phases thread by thread with a plain
// array standing in for shared memory, so the index arithmetic, twiddle tables, real-FFT
// untangling, Hermitian weights and binning can be checked in a container without a GPU.
// Build: g++ -O2 -std=c++17 -shared -fPIC -o tests/emu/libt21emu.so tests/emu/emu.cpp
#include <cstring>
#include <vector>

#include "../../tools21cm_b200/csrc/t21_passes.cuh"
#include "../../tools21cm_b200/csrc/t21_tables.h"

using namespace t21;

// ---- generic phase runner -----------------------------------------------------------
template <class K, typename T, int PH = 0>
static void run_simple(const typename K::Params& p, long long cta, std::vector<typename K::Regs>& regs, cx<T>* smem) {
    for (int tid = 0; tid < K::THREADS; ++tid) K::template phase<PH>(p, cta, tid, regs[tid], smem);
    if constexpr (PH + 1 < K::NPHASE) run_simple<K, T, PH + 1>(p, cta, regs, smem);
}
template <class K, typename T, class Hist, int PH = 0>
static void run_hist(const typename K::Params& p, long long tile, std::vector<typename K::Regs>& regs, cx<T>* smem, Hist& h) {
    for (int tid = 0; tid < K::THREADS; ++tid) K::template phase<PH>(p, tile, tid, regs[tid], smem, h);
    if constexpr (PH + 1 < K::NPHASE) run_hist<K, T, Hist, PH + 1>(p, tile, regs, smem, h);
}

static int round_up(int a, int b) { return (a + b - 1) / b * b; }

// ---- passes with run-time sizes ------------------------------------------------------
template <int NZ, typename T>
static void zpass(const T* in, cx<T>* out, long long nrows, int pitch, const T* offset, int ny) {
    using K = ZPass<NZ, T>;
    auto tw = build_stage_twiddles<T>(NZ / 2, emax_z(NZ / 2));
    auto twr = build_roots<T>(NZ, NZ / 2 + 1);
    typename K::Params p{in, out, nrows, pitch, ny, (nrows / ny) * ny * WLay<T>::LW, offset, tw.data(), twr.data()};
    std::vector<cx<T>> smem(K::SMEM_BYTES / sizeof(cx<T>) + 1);
    long long nctas = (nrows + K::ROWS - 1) / K::ROWS;
    for (long long cta = 0; cta < nctas; ++cta) {
        std::vector<typename K::Regs> regs(K::THREADS);
        run_simple<K, T>(p, cta, regs, smem.data());
    }
}
template <int NY, typename T, int WZ>
static void ypass(cx<T>* w, long long nouter, int nzh, int pitch) {
    using K = YPass<NY, T, WZ>;
    auto tw = build_stage_twiddles<T>(NY, emax_y(NY));
    int nchunks = (nzh + WZ - 1) / WZ;
    typename K::Params p;
    std::memset(&p, 0, sizeof(p));
    p.w = w; p.nzh = nzh; p.pitch = pitch; p.nchunks = nchunks; p.cstride = nouter * NY * WLay<T>::LW; p.tw = tw.data(); p.nouter = nouter;
    std::vector<cx<T>> smem(K::SMEM_BYTES / sizeof(cx<T>) + 1);
    for (long long cta = 0; cta < nouter * nchunks; ++cta) {
        std::vector<typename K::Regs> regs(K::THREADS);
        run_simple<K, T>(p, cta, regs, smem.data());
    }
}
template <int NX, typename T, int WZ, int NF, int MODE, int MU>
static void xpass_mu(const cx<T>* w0, const cx<T>* w1, int ny, int nz, int pitch, const T* off0, const T* off1,
                  const t21_binspec* spec, double* sums, long long* counts, void* nd_out, int nd_f64, double scale,
                  int nctas_emulated) {
    using K = XPass<NX, T, WZ, NF, MODE, MU>;
    auto tw = build_stage_twiddles<T>(NX, emax_x(NX));
    typename K::Params p;
    std::memset(&p, 0, sizeof(p));
    p.w[0] = w0; p.w[1] = w1;
    p.ny = ny; p.nz = nz; p.nzh = nz / 2 + 1; p.pitch = pitch;
    p.nchunks = (p.nzh + WZ - 1) / WZ;
    constexpr int PARTS = WZ < WLay<T>::LW ? WLay<T>::LW / WZ : 1;
    p.ntiles = (long long)ny * ((p.nchunks + PARTS - 1) / PARTS) * PARTS;
    p.tile_step = nctas_emulated;
    p.xs_log2 = ilog2_rt(NX);
    p.iy0 = 0;
    p.tw = tw.data();
    p.offset[0] = off0; p.offset[1] = off1;
    p.ntot = (double)NX * ny * nz;
    if (spec) p.bins = *spec;
    p.nd_out = nd_out; p.nd_f64 = nd_f64; p.nd_half = 0; p.scale = scale;
    std::vector<cx<T>> smem(K::SMEM_BYTES / sizeof(cx<T>) + 1);
    const int nb = spec ? spec->nk * (spec->nmu > 0 ? spec->nmu : 1) : 1;
    // persistent CTAs: each keeps its run accumulators and histogram across its tiles
    for (int cta = 0; cta < nctas_emulated; ++cta) {
        std::vector<double> hs(nb, 0.0);
        std::vector<unsigned> hc(nb, 0u);
        SmemHist hist{hs.data(), hc.data()};
        std::vector<typename K::Regs> regs(K::THREADS);
        for (auto& r : regs) for (int c = 0; c < K::NACC; ++c) acc_init(r.acc[c]);
        for (long long tile = cta; tile < p.ntiles; tile += nctas_emulated)
            run_hist<K, T, SmemHist>(p, tile, regs, smem.data(), hist);
        for (auto& r : regs) for (int c = 0; c < K::NACC; ++c) acc_flush(r.acc[c], hist, nb);
        if (MODE == MODE_BIN)
            for (int b = 0; b < nb; ++b) { sums[b] += hs[b]; counts[b] += hc[b]; }
    }
}

template <int NX, typename T, int WZ, int NF, int MODE>
static void xpass(const cx<T>* w0, const cx<T>* w1, int ny, int nz, int pitch, const T* off0, const T* off1,
                  const t21_binspec* spec, double* sums, long long* counts, void* nd_out, int nd_f64, double scale,
                  int nctas_emulated) {
    const int mu = (MODE == MODE_BIN && spec && spec->nmu > 0) ? (spec->absolute_mus ? 1 : 2) : 0;
    if (mu == 0) xpass_mu<NX, T, WZ, NF, MODE, 0>(w0, w1, ny, nz, pitch, off0, off1, spec, sums, counts, nd_out, nd_f64, scale, nctas_emulated);
    else if (mu == 1) xpass_mu<NX, T, WZ, NF, MODE, 1>(w0, w1, ny, nz, pitch, off0, off1, spec, sums, counts, nd_out, nd_f64, scale, nctas_emulated);
    else xpass_mu<NX, T, WZ, NF, MODE, 2>(w0, w1, ny, nz, pitch, off0, off1, spec, sums, counts, nd_out, nd_f64, scale, nctas_emulated);
}

#define DISPATCH_N(n, MACRO)                 \
    switch (n) {                             \
        case 8: MACRO(8); break;             \
        case 16: MACRO(16); break;           \
        case 32: MACRO(32); break;           \
        case 64: MACRO(64); break;           \
        case 128: MACRO(128); break;         \
        case 256: MACRO(256); break;         \
        case 1024: MACRO(1024); break;       \
        default: return -1;                  \
    }

template <typename T>
static int emu_all(int nx, int ny, int nz, const T* in0, const T* in1, double off0d, double off1d, int upto,
                   cx<T>* half_out, const t21_binspec* spec, double* sums, long long* counts, void* nd_out,
                   int nd_f64, double scale) {
    const int nzh = nz / 2 + 1;
    const int pitch = round_up(nzh, 16);
    const int nf = in1 ? 2 : 1;
    std::vector<cx<T>> w[2];
    T offs[2] = {(T)off0d, (T)off1d};
    const T* ins[2] = {in0, in1};
    for (int f = 0; f < nf; ++f) {
        w[f].assign((size_t)nx * ny * pitch, cmake<T>(0, 0));
#define ZP(N) zpass<N, T>(ins[f], w[f].data(), (long long)nx * ny, pitch, &offs[f], ny)
        DISPATCH_N(nz, ZP)
#undef ZP
        if (upto >= 2) {
            constexpr int WZ = sizeof(T) == 4 ? 16 : 8;
#define YP(N) ypass<N, T, (N <= 512 ? WZ : WZ / 2)>(w[f].data(), nx, nzh, pitch)
            DISPATCH_N(ny, YP)
#undef YP
        }
    }
    if (half_out) {       // hand the chunk-major array back as plain [nx][ny][pitch] rows
        const WLay<T> lay = make_wlay<T>(nx, ny, pitch, 62);
        for (int x = 0; x < nx; ++x)
            for (int y = 0; y < ny; ++y)
                for (int k = 0; k < pitch; ++k)
                    half_out[((size_t)x * ny + y) * pitch + k] = w[0][(size_t)(((long long)(k >> WLay<T>::LWS) * nx + x) * ny + y) * WLay<T>::LW + (k & (WLay<T>::LW - 1))];
        (void)lay;
    }
    if (upto < 3) return 0;
    constexpr int WZ = sizeof(T) == 4 ? 16 : 8;
    const int nctas = 3;
    if (spec) {
        if (nf == 1) {
#define XP(N) xpass<N, T, (N <= 512 ? WZ : WZ / 2), 1, MODE_BIN>(w[0].data(), nullptr, ny, nz, pitch, &offs[0], nullptr, spec, sums, counts, nullptr, 0, 0.0, nctas)
            DISPATCH_N(nx, XP)
#undef XP
        } else {
#define XP(N) xpass<N, T, (N <= 512 ? WZ : WZ / 2), 2, MODE_BIN>(w[0].data(), w[1].data(), ny, nz, pitch, &offs[0], &offs[1], spec, sums, counts, nullptr, 0, 0.0, nctas)
            DISPATCH_N(nx, XP)
#undef XP
        }
    } else {
        if (nf == 1) {
#define XP(N) xpass<N, T, (N <= 512 ? WZ : WZ / 2), 1, MODE_ND>(w[0].data(), nullptr, ny, nz, pitch, &offs[0], nullptr, nullptr, nullptr, nullptr, nd_out, nd_f64, scale, nctas)
            DISPATCH_N(nx, XP)
#undef XP
        } else {
#define XP(N) xpass<N, T, (N <= 512 ? WZ : WZ / 2), 2, MODE_ND>(w[0].data(), w[1].data(), ny, nz, pitch, &offs[0], &offs[1], nullptr, nullptr, nullptr, nd_out, nd_f64, scale, nctas)
            DISPATCH_N(nx, XP)
#undef XP
        }
    }
    return 0;
}

extern "C" {

int emu_pitch(int nz) { return round_up(nz / 2 + 1, 16); }

// upto: 1 = z pass, 2 = z+y passes (half_out receives [nx][ny][pitch] complex), 3 = full
int emu_run(int dtype, int nx, int ny, int nz, const void* in0, const void* in1, double off0, double off1, int upto,
            void* half_out, const t21_binspec* spec, double* sums, long long* counts, void* nd_out, int nd_f64,
            double scale) {
    if (dtype == T21_F32)
        return emu_all<float>(nx, ny, nz, (const float*)in0, (const float*)in1, off0, off1, upto, (cx<float>*)half_out,
                              spec, sums, counts, nd_out, nd_f64, scale);
    return emu_all<double>(nx, ny, nz, (const double*)in0, (const double*)in1, off0, off1, upto,
                           (cx<double>*)half_out, spec, sums, counts, nd_out, nd_f64, scale);
}

// element-wise binning of a caller-supplied array (the body of the bin-only kernel)
int emu_bin_only(const double* p, const t21_binspec* spec, double* sums, long long* counts) {
    const int nb = spec->nk * (spec->nmu > 0 ? spec->nmu : 1);
    std::vector<double> hs(nb, 0.0);
    std::vector<unsigned> hc(nb, 0u);
    SmemHist hist{hs.data(), hc.data()};
    RunAcc acc; run_init(acc);
    long long o = 0;
    for (int ix = 0; ix < spec->n[0]; ++ix)
        for (int iy = 0; iy < spec->n[1]; ++iy)
            for (int iz = 0; iz < spec->n[2]; ++iz, ++o) {
                ModeBins mb = classify(*spec, ix, iy, iz, false);
                run_add(acc, hist, mb.own, p[o], mb.own_w);
            }
    run_flush(acc, hist);
    for (int b = 0; b < nb; ++b) { sums[b] = hs[b]; counts[b] = hc[b]; }
    return 0;
}

// The epilogue of the warp-per-pencil X pass as plain loops: bins are index ranges along x pencils, taken from the
// segment tables build_pencil_segments() makes (one table for shells / |mu| bins, four for signed mu).  phalf is the
// power of the computed half spectrum, [nx][ny][nz/2+1].  Returns 1 if a pencil needs more than 32 ranges,
// 2 if the closed-form mode count of a range (segment_modes, what the kernel books) disagrees with counting them.
int emu_bin_segments(const double* phalf, const t21_binspec* spec, double* sums, long long* counts) {
    const t21_binspec& B = *spec;
    const int nx = B.n[0], ny = B.n[1], nz = B.n[2], nzh = nz / 2 + 1, H = nx / 2;
    const int nb = B.nk * (B.nmu > 0 ? B.nmu : 1);
    const bool is_signed = B.nmu > 0 && !B.absolute_mus;
    const int ntab = is_signed ? 4 : 1;
    for (int b = 0; b < nb; ++b) { sums[b] = 0.0; counts[b] = 0; }
    unsigned ent[32];
    for (int iy = 0; iy < ny; ++iy)
        for (int iz = 0; iz < nzh; ++iz) {
            const bool has_twin = (iz != 0) && (iz != nz - iz);
            for (int t = 0; t < ntab; ++t) {
                const int S = build_pencil_segments(B, iy, iz, is_signed ? t + 1 : 0, ent, 32);
                if (S > 32) return 1;
                const bool lo = ntab == 1 || t == 0 || t == 2, hi = ntab == 1 || t == 1 || t == 3;
                const unsigned wt = (ntab == 1 && has_twin) ? 2u : 1u;
                for (int j = 0; j < S; ++j) {
                    const unsigned a = ent[j] >> 16, b = j + 1 < S ? ent[j + 1] >> 16 : (unsigned)H + 1u;
                    const int bin = (int)(ent[j] & 0xFFFFu) - 1;
                    if (bin < 0) continue;
                    double sum = 0.0;
                    unsigned cnt = 0;
                    for (unsigned f = a; f < b; ++f) {
                        if (lo && f < (unsigned)(nx % 2 == 0 ? H : H + 1)) { sum += phalf[((size_t)f * ny + iy) * nzh + iz]; ++cnt; }
                        if (hi && f >= 1) { sum += phalf[((size_t)(nx - f) * ny + iy) * nzh + iz]; ++cnt; }
                    }
                    if (cnt != segment_modes(lo, hi, a, b, (unsigned)H, nx % 2 == 0)) return 2;
                    sums[bin] += sum * wt;
                    counts[bin] += (long long)cnt * wt;
                }
            }
        }
    return 0;
}
}
