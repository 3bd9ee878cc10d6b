// The three passes of the R2C 3-D transform as barrier-separated phases.
//
// Data layout (DESIGN.md "HBM layout"):
//   input   x[nx][ny][nz]            real, C order (z contiguous)
//   half    W[nx][ny][pitch]         complex, kz = 0..nz/2 valid, pitch = round_up(nz/2+1, 16)
//
//   ZPass : rows along z.  Packs a real row into nz/2 complex values, FFTs it and untangles
//           the result into the nz/2+1 non-redundant frequencies (the classic real-FFT trick),
//           subtracting a mean offset on load.              reads 4*nz, writes 8*(nz/2+1) B/row
//   YPass : pencils along y at fixed x, in place.  A CTA owns all ny samples of WZ adjacent
//           kz columns (WZ*8 B contiguous per row => full 32 B sectors).     reads+writes W
//   XPass : pencils along x at fixed ky, fused with |F|^2 / Re(F1 conj F2) and the binning
//           (or with the fftshifted n-D store); W is read once and never written back.
//
// Every pass body is `phase<PH>()` for PH = 0..NPHASE-1 with a CTA barrier between
// consecutive phases, so tests/emu can replay it on the CPU.
#pragma once
#include "t21_fft.cuh"
#include "t21_bin.cuh"

namespace t21 {

// ======================================================================================
// Z pass
// ======================================================================================
template <int NZ, typename T>
struct ZPass {
    static_assert(is_pow2(NZ) && NZ >= 4, "nz");
    static constexpr int M = NZ / 2;
    using P = Plan<M>;
    static constexpr int E = P::E, TP = P::T, NS = P::NS;
    static constexpr int THREADS = TP >= 256 ? TP : 256;
    static constexpr int ROWS = THREADS / TP;
    static constexpr int ROWPITCH = M + (M >> 4) + 1;          // padded: i -> i + (i >> 4)
    static constexpr int SMEM_ELEMS = ROWS * ROWPITCH;
    static constexpr int NPHASE = 2 * NS;

    struct Params {
        const T* in;          // [nrows][NZ]
        cx<T>* out;           // [nrows][pitch]
        long long nrows;
        int pitch;
        const T* offset;      // device scalar subtracted from every sample, or null
        const cx<T>* tw;      // Plan<M> stage twiddles
        const cx<T>* twr;     // exp(-2 pi i k / NZ), k = 0..M
    };
    struct Regs {
        cx<T> v[E];
    };

    static T21_HD int pad(int i) { return i + (i >> 4); }

    template <int PH>
    static T21_HD void phase(const Params& p, long long cta, int tid, Regs& r, cx<T>* smem) {
        const int j = tid % TP;
        const long long row = cta * ROWS + tid / TP;
        const bool active = row < p.nrows;
        cx<T>* srow = smem + (tid / TP) * ROWPITCH;
        auto st_smem = [&](int i, cx<T> v, int) { srow[pad(i)] = v; };
        auto ld_smem = [&](int i) { return srow[pad(i)]; };

        if constexpr (PH == 0) {
            const T off = p.offset ? ldg(p.offset) : (T)0;
            const cx<T>* src = reinterpret_cast<const cx<T>*>(p.in + (active ? row : 0) * NZ);
            stage_load<M, 0>(r.v, j, [&](int i) {
                cx<T> v = active ? src[i] : cmake<T>(0, 0);
                return cmake<T>(v.x - off, v.y - off);
            });
            stage_compute<M, 0>(r.v, j, p.tw, st_smem);
        } else if constexpr (PH < NPHASE - 1) {
            constexpr int S = (PH + 1) / 2;
            if constexpr (PH % 2 == 1) stage_load<M, S>(r.v, j, ld_smem);
            else stage_compute<M, S>(r.v, j, p.tw, st_smem);
        } else {
            // untangle: X[k] = (Z[k] + conj Z[M-k])/2 - i/2 * w^k * (Z[k] - conj Z[M-k])
            if (!active) return;
            cx<T>* dst = p.out + row * p.pitch;
#pragma unroll
            for (int i = 0; i < E; ++i) {
                const int k = j + i * TP;
                const cx<T> zk = ld_smem(k), zm = ld_smem((M - k) & (M - 1));
                const cx<T> w = ldg(p.twr + k);
                const cx<T> a = cmake<T>(zk.x + zm.x, zk.y - zm.y);
                const cx<T> d = cmake<T>(zk.x - zm.x, zk.y + zm.y);
                const cx<T> t = cmul(w, d);
                dst[k] = cmake<T>((T)0.5 * (a.x + t.y), (T)0.5 * (a.y - t.x));
                if (k == 0) dst[M] = cmake<T>(zk.x - zk.y, (T)0);
            }
        }
    }
};

// ======================================================================================
// strided pencils: shared pieces of the Y and X passes
// ======================================================================================
template <int N, typename T, int WZ>
struct Strided {
    using P = Plan<N>;
    static constexpr int E = P::E, TP = P::T, NS = P::NS;
    static constexpr int THREADS = WZ * TP;
    static constexpr int SMEM_ELEMS = (NS > 1) ? N * WZ : 1;
    static constexpr int FFT_PHASES = 2 * NS - 1;     // phases of one pencil transform

    // phase PH of one transform.  ld_g(i): load element i from global; final(i, v): consume
    // output element i (natural frequency order).
    template <int PH, class LoadG, class Final>
    static T21_HD void fft_phase(cx<T>* v, int j, int zl, const cx<T>* tw, cx<T>* smem, LoadG&& ld_g, Final&& fin) {
        auto st_smem = [&](int i, cx<T> val, int) { smem[i * WZ + zl] = val; };
        auto ld_smem = [&](int i) { return smem[i * WZ + zl]; };
        if constexpr (PH == 0) {
            stage_load<N, 0>(v, j, ld_g);
            if constexpr (NS == 1) stage_compute<N, 0>(v, j, tw, fin);
            else stage_compute<N, 0>(v, j, tw, st_smem);
        } else {
            constexpr int S = (PH + 1) / 2;
            if constexpr (PH % 2 == 1) {
                stage_load<N, S>(v, j, ld_smem);
            } else {
                if constexpr (S == NS - 1) stage_compute<N, S>(v, j, tw, fin);
                else stage_compute<N, S>(v, j, tw, st_smem);
            }
        }
    }
};

// ======================================================================================
// Y pass (in place)
// ======================================================================================
template <int NY, typename T, int WZ>
struct YPass {
    using S = Strided<NY, T, WZ>;
    static constexpr int THREADS = S::THREADS;
    static constexpr int SMEM_ELEMS = S::SMEM_ELEMS;
    static constexpr int NPHASE = S::FFT_PHASES;

    struct Params {
        cx<T>* w;             // [nouter][NY][pitch]
        int nzh;              // valid columns
        int pitch;
        int nchunks;          // ceil(nzh / WZ)
        const cx<T>* tw;
    };
    struct Regs {
        cx<T> v[S::E];
    };

    template <int PH>
    static T21_HD void phase(const Params& p, long long cta, int tid, Regs& r, cx<T>* smem) {
        const int zl = tid % WZ, j = tid / WZ;
        const int chunk = (int)(cta % p.nchunks);
        const long long outer = cta / p.nchunks;
        const int z = chunk * WZ + zl;
        const bool active = z < p.nzh;
        cx<T>* base = p.w + outer * (long long)NY * p.pitch + (active ? z : 0);
        const long long stride = p.pitch;
        S::template fft_phase<PH>(
            r.v, j, zl, p.tw, smem,
            [&](int i) { return active ? base[i * stride] : cmake<T>(0, 0); },
            [&](int i, cx<T> val, int) { if (active) base[i * stride] = val; });
    }
};

// ======================================================================================
// X pass fused with power + binning (MODE_BIN) or the fftshifted n-D store (MODE_ND)
// ======================================================================================
enum { MODE_BIN = 0, MODE_ND = 1 };

template <int NX, typename T, int WZ, int NF, int MODE>
struct XPass {
    using S = Strided<NX, T, WZ>;
    using P = Plan<NX>;
    static constexpr int E = S::E;
    static constexpr int THREADS = S::THREADS;
    static constexpr int SMEM_ELEMS = S::SMEM_ELEMS;
    static constexpr int NPHASE = NF * S::FFT_PHASES;   // the last phase of the last field holds the epilogue
    static constexpr int kMode = MODE;
    // whole warps use the warp-uniform bin cache; CTAs smaller than a warp keep per-thread runs
    template <bool B, class A, class C> struct Pick { using type = A; };
    template <class A, class C> struct Pick<false, A, C> { using type = C; };
    using Acc = typename Pick<(S::THREADS % 32 == 0), WarpBins, ThreadBins>::type;

    struct Params {
        const cx<T>* w[2];    // [NX][ny][pitch] per field
        int ny, nz, nzh, pitch, nchunks;
        long long ntiles;     // ny * nchunks
        const cx<T>* tw;
        const T* offset[2];   // device scalars (mean offsets) or null
        double ntot;          // nx*ny*nz, for the analytic DC repair
        t21_binspec bins;     // MODE_BIN
        void* nd_out;         // MODE_ND: [NX][ny][nz] fftshifted
        int nd_f64;           // MODE_ND: output dtype
        double scale;         // MODE_ND
    };
    struct Regs {
        cx<T> v[NF][E];
        Acc acc;
    };

    // one output element of the last stage: both fields are final for index ix
    // Called by every thread of the CTA for every slot (warp-collective inside); `active` is
    // false for the padding lanes of a ragged last chunk.
    template <class Hist>
    static T21_HD void consume(const Params& p, Regs& r, Hist& hist, int ix, int iy, int iz, bool active,
                               cx<T> a, cx<T> b) {
        double pw;
        if ((ix | iy | iz) == 0) {
            // DC of (x - c) plus c*N: the only sample the offset changes
            const double c0 = p.offset[0] ? (double)ldg(p.offset[0]) : 0.0;
            const double re0 = (double)a.x + c0 * p.ntot;
            if (NF == 1) {
                pw = re0 * re0 + (double)a.y * (double)a.y;
            } else {
                const double c1 = p.offset[1] ? (double)ldg(p.offset[1]) : 0.0;
                const double re1 = (double)b.x + c1 * p.ntot;
                pw = re0 * re1 + (double)a.y * (double)b.y;
            }
        } else {
            pw = (NF == 1) ? (double)(a.x * a.x + a.y * a.y) : (double)(a.x * b.x + a.y * b.y);
        }
        const bool has_twin = (iz != 0) && (iz != p.nz - iz);
        if constexpr (MODE == MODE_BIN) {
            ModeBins mb;
            mb.own = -1; mb.own_w = 0; mb.twin = -1;
            if (active) mb = classify(p.bins, ix, iy, iz, has_twin);
            acc_add(r.acc, hist, 0, mb.own, pw, mb.own_w, active);
            if (p.bins.nmu > 0 && !p.bins.absolute_mus)      // signed mu: the twin goes to the mirrored bin
                acc_add(r.acc, hist, 1, mb.twin, pw, 1, active);
        } else {
            if (!active) return;
            const double val = pw * p.scale;
            const int sx = ix + NX / 2 - (ix + NX / 2 >= NX ? NX : 0);
            const int hy = p.ny / 2, hz = p.nz / 2;
            const int sy = iy + hy - (iy + hy >= p.ny ? p.ny : 0);
            const int sz = iz + hz - (iz + hz >= p.nz ? p.nz : 0);
            const long long o = ((long long)sx * p.ny + sy) * p.nz + sz;
            if (p.nd_f64) ((double*)p.nd_out)[o] = val; else ((float*)p.nd_out)[o] = (float)val;
            if (has_twin) {
                const int tx = ix ? NX - ix : 0, ty = iy ? p.ny - iy : 0, tz = p.nz - iz;
                const int ux = tx + NX / 2 - (tx + NX / 2 >= NX ? NX : 0);
                const int uy = ty + hy - (ty + hy >= p.ny ? p.ny : 0);
                const int uz = tz + hz - (tz + hz >= p.nz ? p.nz : 0);
                const long long ot = ((long long)ux * p.ny + uy) * p.nz + uz;
                if (p.nd_f64) ((double*)p.nd_out)[ot] = val; else ((float*)p.nd_out)[ot] = (float)val;
            }
        }
    }

    template <int PH, class Hist>
    static T21_HD void phase(const Params& p, long long tile, int tid, Regs& r, cx<T>* smem, Hist& hist) {
        constexpr int F = PH / S::FFT_PHASES;          // field this phase works on
        constexpr int FPH = PH % S::FFT_PHASES;
        const int zl = tid % WZ, j = tid / WZ;
        const int chunk = (int)(tile % p.nchunks);
        const int iy = (int)(tile / p.nchunks);
        const int iz = chunk * WZ + zl;
        const bool active = iz < p.nzh;
        const cx<T>* base = p.w[F] + (long long)iy * p.pitch + (active ? iz : 0);
        const long long stride = (long long)p.ny * p.pitch;
        S::template fft_phase<FPH>(
            r.v[F], j, zl, p.tw, smem,
            [&](int i) { return active ? base[i * stride] : cmake<T>(0, 0); },
            [&](int i, cx<T> val, int slot) {
                // last stage of field F: element i is final and still sits in register `slot`.
                // Field 0 of a cross spectrum just stays there; the last field consumes both.
                if constexpr (F == NF - 1)
                    consume(p, r, hist, i, iy, iz, active, (NF == 1) ? val : r.v[0][slot], val);
            });
    }
};

}  // namespace t21
