// The three passes of the R2C 3-D transform as barrier-separated phases.
//
// Data layout (DESIGN.md "HBM layout"):
//   input   x[nx][ny][nz]            real, C order (z contiguous)
//   half    W[nch][nx][ny][LW]       complex, "chunk major": the kz axis is cut into 64-byte pieces of
//                                    LW = 8 (fp32) / 4 (fp64) columns and the piece index is the SLOWEST
//                                    axis.  A Y-pass tile (all y, one piece, fixed x) is then one contiguous
//                                    block, an X-pass tile reads 64-byte pieces ny*64 B apart inside one
//                                    piece-plane, and the Z pass writes LW columns of its 8 consecutive rows
//                                    as one 512-byte run.  kz = 0..nz/2 valid, nch = pitch / LW,
//                                    pitch = round_up(nz/2+1, 16).
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
#include "t21_config.h"

namespace t21 {

// element offset of (x, y, kz) in a chunk-major half-spectrum array of nxa x planes and nya y rows.
// Arrays received from an NCCL all-to-all are a sequence of x slabs, each chunk-major on its own:
// xs_log2 = log2(planes per slab), slab = elements per slab (for a plain array one slab holds everything).
template <typename T>
struct WLay {
    static constexpr int LW = 64 / (int)sizeof(cx<T>);
    static constexpr int LWS = ilog2(LW);
    long long nya;        // y rows in this array
    long long slab;       // elements per x slab
    int xs_log2;          // log2(x planes per slab)
    T21_HD long long off(long long x, long long y, int kz) const {
        const long long nxs = 1LL << xs_log2;
        return (x >> xs_log2) * slab + (((((long long)(kz >> LWS) << xs_log2) + (x & (nxs - 1))) * nya + y) << LWS) +
               (kz & (LW - 1));
    }
};
template <typename T>
T21_HD WLay<T> make_wlay(long long nxa, long long nya, int pitch, int xs_log2) {
    WLay<T> w;
    w.nya = nya;
    w.xs_log2 = xs_log2;
    w.slab = (long long)pitch * nya << xs_log2;
    (void)nxa;
    return w;
}
T21_HD int ilog2_rt(long long n) { int l = 0; while ((1LL << (l + 1)) <= n) ++l; return l; }

// ======================================================================================
// Z pass
// ======================================================================================
template <int NZ, typename T>
struct ZPass {
    static_assert(is_pow2(NZ) && NZ >= 4, "nz");
    static constexpr int M = NZ / 2;
    using P = Plan<M, emax_z(M)>;
    static constexpr int E = P::E, TP = P::T, NS = P::NS;
    static constexpr int THREADS = TP >= 256 ? TP : 256;
    static constexpr int ROWS = THREADS / TP;
    // padded: i -> i + (i >> 4).  Rows are 64 (mod 128) bytes apart: the untangle phase reads 64-byte runs
    // of neighbouring rows in one wavefront, and they must fall on different halves of the banks.
    static constexpr int kRun = 64 / (int)sizeof(cx<T>);
    static constexpr int ROWPITCH = (M + (M >> 4) + 2 * kRun) / (2 * kRun) * (2 * kRun) + kRun;
    static constexpr int SMEM_ELEMS = ROWS * ROWPITCH;
    static constexpr int SMEM_BYTES = SMEM_ELEMS * (int)sizeof(cx<T>);
    // fp32: 64 registers is what it needs anyway.  fp64: uncapped it takes 146 registers -- ONE 256-thread CTA per SM;
    // capped at 128 two fit (1024^3 fp64 Z pass: see DESIGN.md section 4)
    static constexpr int MIN_CTAS = THREADS <= 256 ? (sizeof(T) == 4 ? 4 : 2) : 1;
    // When the last stage is a radix-2 one, it is fused with the untangle phase: outputs k, M-k, M/2-k, M/2+k
    // need exactly the four stage inputs a[k], a[k+M/2], a[M/2-k], a[M-k], so one thread reads those four values
    // and writes four outputs -- one shared-memory exchange (16 stores, 16 loads, two barriers) less.
    static constexpr int kQ = M / 4;
    static constexpr int kGroups = THREADS / ((64 / (int)sizeof(cx<T>)) * ROWS);
    static constexpr bool kFuseLast = NS >= 2 && P::radix(NS - 1) == 2 && kQ % (64 / (int)sizeof(cx<T>)) == 0 &&
                                      THREADS % ((64 / (int)sizeof(cx<T>)) * ROWS) == 0 && kGroups >= 2 &&
                                      (kQ / (64 / (int)sizeof(cx<T>))) % kGroups == 0;
    // ... and with a 16*16*2 plan the middle stage runs in place (Plan::kInPlace1): stage 0 | stage 1 | fused last
    static constexpr bool kInPlace = kFuseLast && P::kInPlace1 && T21_INPLACE_STAGE1 != 0;
    static constexpr int NPHASE = kInPlace ? 3 : (kFuseLast ? 2 * NS - 2 : 2 * NS);
    // where the in-place middle stage left element e of its (Stockham-ordered) output
    static T21_HD int where(int e) {
        if constexpr (kInPlace) return 16 * (e >> 8) + (e & 15) + TP * ((e >> 4) & 15);
        else return e;
    }

    struct Params {
        const T* in;          // [nrows][NZ]
        cx<T>* out;           // chunk-major half spectrum of nrows / ny x planes
        long long nrows;
        int pitch;
        int ny;               // rows are (x, y) pairs
        long long cstride;    // elements between 64-byte pieces of `out`: (x planes) * ny * LW
        const T* offset;      // device scalar subtracted from every sample, or null
        const cx<T>* tw;      // Plan<M> stage twiddles
        const cx<T>* twr;     // exp(-2 pi i k / NZ), k = 0..M
        // Cubical window fused into the load (position_dependent_power_spectrum.py:65-68,90-103: cube * W_L): samples
        // outside [win_lo, win_hi) along any axis are read as zero -- and not read at all.  win_on = 0: no window.
        int win_on;
        int win_lo[3], win_hi[3];
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
            if (p.win_on) {
                // (rows fit 32 bits with a window: the launcher checks; one division per thread, only on this path)
                const unsigned xr = (unsigned)row / (unsigned)p.ny, yr = (unsigned)row - xr * (unsigned)p.ny;
                const bool row_in = active && (int)xr >= p.win_lo[0] && (int)xr < p.win_hi[0] && (int)yr >= p.win_lo[1] &&
                                    (int)yr < p.win_hi[1];
                const int z0 = p.win_lo[2], z1 = p.win_hi[2];
                stage_load<P, 0>(r.v, j, [&](int i) {
                    const int z = 2 * i;
                    cx<T> v = cmake<T>(0, 0);
                    if (row_in && z + 1 >= z0 && z < z1) {
                        v = src[i];
                        if (z < z0) v.x = (T)0;
                        if (z + 1 >= z1) v.y = (T)0;
                    }
                    return v;
                });
            } else {
                stage_load<P, 0>(r.v, j, [&](int i) {
                    cx<T> v = active ? src[i] : cmake<T>(0, 0);
                    return cmake<T>(v.x - off, v.y - off);
                });
            }
            stage_compute<P, 0>(r.v, j, p.tw, st_smem);
        } else if constexpr (kInPlace && PH == 1) {
            stage_load<P, 1>(r.v, j, ld_smem);
            stage_compute<P, 1, true>(r.v, j, p.tw, st_smem);
        } else if constexpr (PH < NPHASE - 1) {
            constexpr int S = (PH + 1) / 2;
            if constexpr (PH % 2 == 1) stage_load<P, S>(r.v, j, ld_smem);
            else stage_compute<P, S>(r.v, j, p.tw, st_smem);
        } else if constexpr (kFuseLast) {
            // last radix-2 stage: Z[i] = a[i] + w2^i a[i+M/2], Z[i+M/2] = a[i] - w2^i a[i+M/2] (w2 = exp(-2 pi i / M)),
            // then the untangle of the pairs (k, M-k) and (M/2-k, M/2+k); a[] is what stage NS-2 left in smem.
            constexpr int LW = WLay<T>::LW, LWS = WLay<T>::LWS;
            constexpr int M2 = M / 2;
            const long long cstride = p.cstride;
            const cx<T>* tw2 = p.tw + P::tw_off(NS - 1);
            const int l = tid % LW, rsub = (tid / LW) % ROWS, g = tid / (LW * ROWS);
            const long long rw = cta * ROWS + rsub;
            const cx<T>* ar = smem + rsub * ROWPITCH;
            cx<T>* d0 = p.out + (rw << LWS);          // row (x, y) sits at x * ny + y = rw
            auto put = [&](int k, cx<T> v) { d0[(long long)(k >> LWS) * cstride + (k & (LW - 1))] = v; };
            // X[k] and X[M-k] from Z[k] and Z[M-k]
            auto untangle = [&](int k, cx<T> zk, cx<T> zm) {
                const cx<T> w = ldg(p.twr + k);
                const cx<T> a = cmake<T>(zk.x + zm.x, zk.y - zm.y);       // Z[k] + conj Z[M-k]
                const cx<T> d = cmake<T>(zk.x - zm.x, zk.y + zm.y);       // Z[k] - conj Z[M-k]
                const cx<T> t = cmul(w, d);
                put(k, cmake<T>((T)0.5 * (a.x + t.y), (T)0.5 * (a.y - t.x)));
                put(M - k, cmake<T>((T)0.5 * (a.x - t.y), (T)0.5 * (-a.y - t.x)));   // conj(a)/2 - i/2 conj(t)
            };
            if (rw < p.nrows) {
#pragma unroll
                for (int it = 0; it < (kQ / LW) / kGroups; ++it) {
                    const int k = (g + it * kGroups) * LW + l;            // 0 <= k < M/4
                    const cx<T> a0 = ar[pad(where(k))], a1 = ar[pad(where(k + M2))];
                    const cx<T> u = k ? cmul(ldg(tw2 + k), a1) : a1;
                    const cx<T> zk = cadd(a0, u), zk2 = csub(a0, u);      // Z[k], Z[k+M/2]
                    if (k == 0) {
                        untangle(0, zk, zk);                              // X[0], X[M] (M-0 wraps to 0)
                        untangle(M2, zk2, zk2);                           // the self-paired middle sample
                    } else {
                        const cx<T> b0 = ar[pad(where(M2 - k))], b1 = ar[pad(where(M - k))];
                        const cx<T> v = cmul(ldg(tw2 + (M2 - k)), b1);
                        untangle(k, zk, csub(b0, v));                     // Z[M-k] = a[M/2-k] - w2^(M/2-k) a[M-k]
                        untangle(M2 - k, cadd(b0, v), zk2);               // pair (M/2-k, M/2+k)
                    }
                }
                if (g == 1 && l == 0) {                                   // k = M/4 pairs with 3M/4
                    const cx<T> a0 = ar[pad(where(kQ))], a1 = ar[pad(where(kQ + M2))];
                    const cx<T> u = cmul(ldg(tw2 + kQ), a1);
                    untangle(kQ, cadd(a0, u), csub(a0, u));
                }
            }
        } else {
            // untangle: X[k] = (Z[k] + conj Z[M-k])/2 - i/2 * w^k * (Z[k] - conj Z[M-k]),  w = exp(-2 pi i / NZ).
            // k and M-k need the same two inputs and w^(M-k) = -conj(w^k), so one thread produces both.
            constexpr int LW = WLay<T>::LW, LWS = WLay<T>::LWS;
            const long long cstride = p.cstride;
            // Work is re-dealt for the stores: a warp takes LW consecutive k of 32/LW consecutive rows, so that
            // one store instruction writes one contiguous run (the rows of a CTA are consecutive in y and a
            // 64-byte piece of row y+1 follows that of row y in the chunk-major layout).
            constexpr int HALF = M / 2;
            constexpr bool kDeal = (HALF % LW == 0) && (THREADS % (LW * ROWS) == 0) && (HALF / LW) % (THREADS / (LW * ROWS)) == 0;
            auto pair_at = [&](int rsub, int k) {
                const long long rw = cta * ROWS + rsub;
                if (rw >= p.nrows) return;
                const cx<T>* zr = smem + rsub * ROWPITCH;
                const cx<T> zk = zr[pad(k)], zm = zr[pad((M - k) & (M - 1))];
                const cx<T> w = ldg(p.twr + k);
                const cx<T> a = cmake<T>(zk.x + zm.x, zk.y - zm.y);       // Z[k] + conj Z[M-k]
                const cx<T> d = cmake<T>(zk.x - zm.x, zk.y + zm.y);       // Z[k] - conj Z[M-k]
                const cx<T> t = cmul(w, d);
                cx<T>* d0 = p.out + (rw << LWS);      // row (x, y) sits at x * ny + y = rw
                d0[(long long)(k >> LWS) * cstride + (k & (LW - 1))] = cmake<T>((T)0.5 * (a.x + t.y), (T)0.5 * (a.y - t.x));
                // partner: conj(a)/2 - i/2 * conj(t)
                const int km = M - k;
                d0[(long long)(km >> LWS) * cstride + (km & (LW - 1))] = cmake<T>((T)0.5 * (a.x - t.y), (T)0.5 * (-a.y - t.x));
            };
            if constexpr (kDeal) {
                constexpr int GROUPS = THREADS / (LW * ROWS);          // k pieces handled per step by the CTA
                const int l = tid % LW, rsub = (tid / LW) % ROWS, g = tid / (LW * ROWS);
#pragma unroll
                for (int it = 0; it < (HALF / LW) / GROUPS; ++it) pair_at(rsub, (g + it * GROUPS) * LW + l);
                if (g == 0 && l == 0) pair_at(rsub, HALF);             // self-paired middle sample of every row
            } else {
                if (!active) return;
#pragma unroll
                for (int i = 0; i < (E + 1) / 2; ++i) {
                    const int k = j + i * TP;
                    if (k < HALF || HALF == 0) pair_at(tid / TP, k);
                }
                if (j == 0 && M >= 2) pair_at(tid / TP, HALF);
            }
        }
    }
};

// ======================================================================================
// strided pencils: shared pieces of the Y and X passes
// ======================================================================================
// MERGE_LAST: the consumer of the last stage's outputs does not touch the exchange buffer (the Y pass stores to
// global memory), so the last stage's loads and its compute need no barrier between them either.
template <int N, typename T, int WZ, int EMAX, bool MERGE_LAST = false>
struct Strided {
    using P = Plan<N, EMAX>;
    static constexpr int E = P::E, TP = P::T, NS = P::NS;
    static constexpr int THREADS = WZ * TP;
    static constexpr int SMEM_ELEMS = (NS > 1) ? N * WZ : 1;
    // 16*16*r plans run their middle stage in place (Plan::kInPlace1): phases are then
    //   0: stage 0 (global loads, stores) | 1: stage 1 (loads, stores to the same places) | 2: stage 2 loads |
    //   3: stage 2 compute + consumer        (2 and 3 are one phase with MERGE_LAST)
    static constexpr bool kInPlace = P::kInPlace1 && T21_INPLACE_STAGE1 != 0;
    static constexpr int FFT_PHASES = kInPlace ? (MERGE_LAST ? 3 : 4) : 2 * NS - 1;     // phases of one pencil transform

    // phase PH of one transform.  ld_g(i): load element i from global; final(i, v): consume
    // output element i (natural frequency order).
    // (Measured and dropped: padding the exchange buffer by one row in 16 removes the 2-way bank conflict of
    // the first-stage stores -- Y 1.565 -> 1.557 ms, but X 2.08 -> 2.22 ms from the extra address registers.)
    // loaded = true: the caller has filled v with the first stage's inputs already (ld_g is not used)
    template <int PH, class LoadG, class Final>
    static T21_HD void fft_phase(cx<T>* v, int j, int zl, const cx<T>* tw, cx<T>* smem, LoadG&& ld_g, Final&& fin,
                                 bool loaded = false) {
        auto st_smem = [&](int i, cx<T> val, int) { smem[i * WZ + zl] = val; };
        auto ld_smem = [&](int i) { return smem[i * WZ + zl]; };
        if constexpr (PH == 0) {
            if (!loaded) stage_load<P, 0>(v, j, ld_g);
            if constexpr (NS == 1) stage_compute<P, 0>(v, j, tw, fin);
            else stage_compute<P, 0>(v, j, tw, st_smem);
        } else if constexpr (kInPlace) {
            if constexpr (PH == 1) {
                stage_load<P, 1>(v, j, ld_smem);
                stage_compute<P, 1, true>(v, j, tw, st_smem);
            } else {
                if constexpr (PH == 2) stage_load<P, 2, true>(v, j, ld_smem);
                if constexpr (PH == 3 || MERGE_LAST) stage_compute<P, 2>(v, j, tw, fin);
            }
        } else {
            constexpr int S = (PH + 1) / 2;
            if constexpr (PH % 2 == 1) {
                stage_load<P, S>(v, j, ld_smem);
            } else {
                if constexpr (S == NS - 1) stage_compute<P, S>(v, j, tw, fin);
                else stage_compute<P, S>(v, j, tw, st_smem);
            }
        }
    }
};

// ======================================================================================
// Y pass (in place)
// ======================================================================================
// SEND: 0 in place; 1 store into the packed all-to-all send buffer of a slab-sharded cube;
//       2 store straight into the y-slab buffers of the PEER GPUs over NVLink (the all-to-all is fused
//         into this pass: no send buffer, no collective, the transfer overlaps the transform tile by tile).
// Columns past nz/2 inside the last chunk are scratch (pitch >= nchunks*WZ): they are transformed
// like the rest and ignored by the X-pass epilogue, so no access here needs a bounds predicate.
constexpr int kMaxPeers = 16;

template <int NY, typename T, int WZ, int SEND = 0>
struct YPass {
    using S = Strided<NY, T, WZ, emax_y(NY), true>;
    static constexpr int THREADS = S::THREADS;
    static constexpr int SMEM_BYTES = S::SMEM_ELEMS * (int)sizeof(cx<T>);
    static constexpr int NPHASE = S::FFT_PHASES;
    // aim at 64 registers per thread (1024 resident threads per SM) as far as shared memory allows: the
    // pass is latency bound and wants every CTA it can get
    static constexpr int kByRegs = THREADS >= 1024 ? 1 : (1024 / THREADS > 8 ? 8 : 1024 / THREADS);
    static constexpr int kBySmem = SMEM_BYTES > 0 ? (200 * 1024) / (SMEM_BYTES + 1024) : 8;
    static constexpr int MIN_CTAS = S::E > 16 ? 1 : (kByRegs < kBySmem ? kByRegs : (kBySmem < 1 ? 1 : kBySmem));

    struct Params {
        cx<T>* w;             // chunk-major, nouter x planes of NY rows
        int nzh;              // valid columns
        int pitch;
        int nchunks;          // ceil(nzh / WZ)
        long long cstride;    // elements between 64-byte pieces: nouter * NY * LW
        const cx<T>* tw;
        // slab-sharded cubes: write the result straight into the all-to-all send buffer (P blocks, block s =
        // the chunk-major [nch][nouter][NY/P][LW] part for rank s) instead of in place; null otherwise
        cx<T>* send;
        int nyl_log2;         // log2(NY / P)
        long long nouter;     // x planes in this slab
        // SEND == 2: peer[s] = rank s's chunk-major y-slab buffer [nch][nx][NY/P][LW] (peer-mapped); this
        // slab's planes start at global x = xg0
        cx<T>* peer[SEND == 2 ? kMaxPeers : 1];
        long long xg0;
        long long peer_cstride;   // nx * (NY/P) * LW
    };
    struct Regs {
        cx<T> v[S::E];
    };

    template <int PH>
    static T21_HD void phase(const Params& p, long long cta, int tid, Regs& r, cx<T>* smem) {
        const int zl = tid % WZ, j = tid / WZ;
        // tile counts fit 31 bits (launch_tiles checks): 32-bit division, the 64-bit one costs ~0.2 ms a pass
        const unsigned outer32 = (unsigned)cta / (unsigned)p.nchunks;
        const int chunk = (int)((unsigned)cta - outer32 * (unsigned)p.nchunks);
        const long long outer = outer32;
        const int z = chunk * WZ + zl;
        // pencil (x = outer, column z): y rows are LW elements apart inside one contiguous block
        constexpr int LW = WLay<T>::LW, LWS = WLay<T>::LWS;
        cx<T>* base = p.w + (long long)(z >> LWS) * p.cstride + ((outer * NY) << LWS) + (z & (LW - 1));
        const long long stride = LW;
        using P = typename S::P;
        // rows are visited in loop order, so both the first-stage loads and the last-stage stores walk
        // a cursor instead of multiplying an index by the run-time row stride
        constexpr int R0 = P::radix(0), RL = P::radix(P::NS - 1), NSL = P::ns(P::NS - 1);
        RowCursor<cx<T>> lc(base + (long long)j * stride, (long long)(NY / R0) * stride, (long long)P::T * stride, R0);
        RowCursor<cx<T>> sc(base + (long long)j * stride, (long long)NSL * stride, (long long)P::T * stride, RL);
        S::template fft_phase<PH>(
            r.v, j, zl, p.tw, smem,
#ifdef T21_DBG_NOLOAD
            [&](int i) { return cmake<T>((T)i, (T)zl); },
#else
            [&](int) { return *lc.next(); },
#endif
            [&](int i, cx<T> val, int) {
                if constexpr (SEND == 1) {
                    // block of rank i >> nyl_log2, chunk-major [nch][nouter][NY/P][LW] inside the block
                    const long long blk = ((long long)p.pitch * p.nouter) << p.nyl_log2;
                    const long long o = (((((long long)(z >> LWS) * p.nouter + outer) << p.nyl_log2) +
                                          (i & ((1 << p.nyl_log2) - 1))) << LWS) + (z & (LW - 1));
                    p.send[(long long)(i >> p.nyl_log2) * blk + o] = val;
                } else if constexpr (SEND == 2) {
                    // the same element, written into rank (i >> nyl_log2)'s chunk-major [nch][nx][NY/P][LW]
                    const long long o = (long long)(z >> LWS) * p.peer_cstride +
                                        ((((p.xg0 + outer) << p.nyl_log2) + (i & ((1 << p.nyl_log2) - 1))) << LWS) +
                                        (z & (LW - 1));
                    p.peer[i >> p.nyl_log2][o] = val;
                } else {
                    *sc.next() = val;
                }
            });
    }
};

// ======================================================================================
// X pass fused with power + binning (MODE_BIN) or the fftshifted n-D store (MODE_ND)
// ======================================================================================
// After the last butterfly stage every thread holds 16 outputs spread over all kx.  It turns
// them into power, parks the power tile P[kx][kz] in the (now free) exchange buffer, and after
// one more barrier the CTA walks the tile in a different order: each warp sweeps a contiguous
// range of kx, 32/WZ rows per step, lanes along kz.  In that order the bin of a warp's lanes
// is almost always uniform and changes only a few times per sweep, so the rolled epilogue
// loop (small code) feeds the warp-uniform bin cache with near-perfect hit rate.
enum { MODE_BIN = 0, MODE_ND = 1 };

// MU (MODE_BIN only): 0 shells, 1 |mu| bins, 2 signed mu bins (Hermitian twins binned separately).
// It is a compile-time parameter so that the accumulator state that lives across tiles is as
// small as the mode allows (shells: 6 lane registers; signed mu: two accumulators with mu intervals).
template <int NX, typename T, int WZ, int NF, int MODE, int MU = 0>
struct XPass {
    using S = Strided<NX, T, WZ, emax_x(NX)>;
    using P = typename S::P;
    static constexpr int E = S::E;
    static constexpr int THREADS = S::THREADS;
    static constexpr int TILE = NX * WZ;                       // modes per tile
    static constexpr int EX_BYTES = (S::NS > 1) ? TILE * (int)sizeof(cx<T>) : 0;
    static constexpr int P_BYTES = TILE * (int)sizeof(T);
    static constexpr int BUF_BYTES = ((EX_BYTES > P_BYTES ? EX_BYTES : P_BYTES) + 15) / 16 * 16;
    static constexpr int SMEM_BYTES = BUF_BYTES + 16;          // + the DC power as a double
    static constexpr int NPHASE = NF * S::FFT_PHASES + 1;      // transforms, then the epilogue
    static constexpr int kMode = MODE;
    static constexpr int kNF = NF;
    static constexpr int kMu = MU;
    static constexpr int NACC = (MU == 2) ? 2 : 1;
    static constexpr int RPI = (THREADS >= 32) ? 32 / WZ : S::TP;   // kx rows a warp covers per step
    // whole warps use the warp-uniform bin cache; CTAs smaller than a warp keep per-thread runs
    template <bool B, class A, class C> struct Pick { using type = A; };
    template <class A, class C> struct Pick<false, A, C> { using type = C; };
    using Acc = typename Pick<(THREADS % 32 == 0), WarpBins<(MU != 0)>, ThreadBins>::type;

    struct Params {
        const cx<T>* w[2];    // chunk-major per field: NX planes of ny rows (possibly as a sequence of x slabs)
        int ny, nz, nzh, pitch, nchunks;
        long long ntiles;     // ny * nchunks
        long long tile_step;  // tiles between two consecutive tiles of one CTA (the grid size)
        int xs_log2;          // log2(x planes per slab): log2(NX) for a plain array, log2(NX/P) after an NCCL exchange
        int iy0;              // global ky index of local row 0 (slab-sharded cubes; 0 otherwise)
        const cx<T>* tw;
        const T* offset[2];   // device scalars (mean offsets) or null
        double ntot;          // nx*ny*nz, for the analytic DC repair
        t21_binspec bins;     // MODE_BIN
        void* nd_out;         // MODE_ND: [NX][ny][nz] fftshifted, or [NX][ny][nz/2+1] in FFT order (nd_half)
        int nd_f64;           // MODE_ND: output dtype
        int nd_half;          // MODE_ND: > 0 = store only the computed half spectrum, unshifted, rows of nd_half elements
                              // (a multiple of 8 keeps every 32-byte run of a tile inside one DRAM sector)
        double scale;         // MODE_ND
        // MODE_BIN with a pencil segment table (t21_build_pencil_segments): bins are index ranges along the pencil
        const unsigned* seg;  // null: classify per mode (warp-uniform bin cache)
        int seg_stride, seg_ntab;
        long long seg_pencils;
    };
    // pencil-major, padded power tile of the table-driven epilogue: pencil zl at zl * PP (PP = 4 mod 32: the stores
    // of a warp -- 8 pencils x 4 consecutive kx -- fall on 32 different banks, a warp reading one pencil is contiguous)
    static constexpr int PP = NX + 4;
    static constexpr bool kSegOk = (S::NS > 1) && (THREADS % 32 == 0) && (NX % 2 == 0);
    struct Regs {
        cx<T> v[NF][E];
        Acc acc[NACC];        // [0] stored modes, [1] Hermitian twins when they bin separately (signed mu)
    };

    static T21_HD T* ptile(cx<T>* smem) { return reinterpret_cast<T*>(smem); }
    static T21_HD double* dc_slot(cx<T>* smem) {
        return reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(smem) + BUF_BYTES);
    }

    // tile -> (chunk of WZ columns, local ky row).  (32-bit division: ntiles < 2^31.)  ky runs fastest: CTAs
    // working side by side read neighbouring 64-byte pieces.  A CTA narrower than a 64-byte piece takes the
    // piece's parts as consecutive tiles, so that they are read from DRAM together.
    static T21_HD void tile_coords(const Params& p, long long tile, int& chunk, int& iy_loc) {
        constexpr int PARTS = WZ < WLay<T>::LW ? WLay<T>::LW / WZ : 1;
        const unsigned t2 = (unsigned)tile / PARTS;
        const int crel = (int)(t2 / (unsigned)p.ny);
        iy_loc = (int)(t2 - (unsigned)crel * (unsigned)p.ny);
        chunk = crel * PARTS + (int)((unsigned)tile % PARTS);
    }

    // the global loads of one tile's pencils (field F) into v, in the order stage 0 consumes them
    template <int F>
    static T21_HD void load_tile(const Params& p, long long tile, int tid, cx<T>* v) {
        const int zl = tid % WZ, j = tid / WZ;
        int chunk, iy_loc;
        tile_coords(p, tile, chunk, iy_loc);
        const int iz = chunk * WZ + zl;
        // element i (= x) of this pencil sits at WLay::off(i, iy_loc, iz) = off(0, iy_loc, iz) + xi(i) * xstr
        // with xi(i) = i + (i >> xs_log2) * D: plane i of a sequence of x slabs (D = 0 for a plain array);
        // scratch columns of the last chunk are read too
        constexpr int LW = WLay<T>::LW, LWS = WLay<T>::LWS;
        const long long xstr = (long long)p.ny << LWS;
        const cx<T>* base = p.w[F] + (((((long long)(iz >> LWS) << p.xs_log2) * p.ny) + iy_loc) << LWS) + (iz & (LW - 1));
#ifdef T21_DBG_NOLOAD
        stage_load<P, 0>(v, j, [&](int i) { return cmake<T>((T)i, (T)zl); });
#else
        if ((NX >> p.xs_log2) <= 1) {
            // plain array: rows are linear in x, walked by a cursor
            constexpr int R0 = P::radix(0);
            RowCursor<const cx<T>> lc(base + (long long)j * xstr, (long long)(NX / R0) * xstr, (long long)P::T * xstr, R0);
            stage_load<P, 0>(v, j, [&](int) { return *lc.next(); });
        } else {
            const long long D = ((long long)(p.pitch >> LWS) - 1) << p.xs_log2;
            stage_load<P, 0>(v, j, [&](int i) { return base[(i + (long long)(i >> p.xs_log2) * D) * xstr]; });
        }
#endif
    }

    // (Measured and dropped: issuing the NEXT tile's loads before the epilogue keeps 32 more registers live
    // through the sweep -- ptxas spills seven of the loaded values right behind the loads, X 2.08 -> 2.53 ms;
    // prefetch.global.L2 of the next tile instead: 2.31 ms.)
    template <int PH, class Hist>
    static T21_HD void phase(const Params& p, long long tile, int tid, Regs& r, cx<T>* smem, Hist& hist) {
        const int zl = tid % WZ, j = tid / WZ;
        // ky fastest: CTAs running together read neighbouring 64-byte pieces of the same piece-plane
        int chunk, iy_loc;                                 // iy_loc: row inside this rank's y slab
        tile_coords(p, tile, chunk, iy_loc);
        const int iy = iy_loc + p.iy0;                     // global ky index (tables, DC, exclusions)
        const int iz = chunk * WZ + zl;
        const bool active = iz < p.nzh;
        if constexpr (PH < NF * S::FFT_PHASES) {
            constexpr int F = PH / S::FFT_PHASES;          // field this phase works on
            constexpr int FPH = PH % S::FFT_PHASES;
            if constexpr (FPH == 0) {
                load_tile<F>(p, tile, tid, r.v[F]);
            }
            auto fin = [&](int i, cx<T> b, int slot) {
                    // last stage of field F: element i (= kx) is final and sits in register `slot`.
                    // Field 0 of a cross spectrum just stays there; the last field forms the power.
                    if constexpr (F == NF - 1) {
                        const cx<T> a = (NF == 1) ? b : r.v[0][slot];
                        T pw = a.x * b.x + a.y * b.y;
                        if (i == 0 && iy == 0 && iz == 0) {
                            // DC of (x - c) plus c*N: the only sample the mean offset changes
                            const double c0 = p.offset[0] ? (double)ldg(p.offset[0]) : 0.0;
                            const double c1 = (NF == 2 && p.offset[1]) ? (double)ldg(p.offset[1]) : c0;
                            const double ra = (double)a.x + c0 * p.ntot, rb = (double)b.x + c1 * p.ntot;
                            const double dc = ra * rb + (double)a.y * (double)b.y;
                            if constexpr (MODE == MODE_BIN) {
                                // the tile keeps 0 for this sample (so it is counted like any other mode);
                                // its power goes to the histogram here, once, with count 0
                                pw = (T)0;
                                int b0;
                                if (kSegOk && p.seg) b0 = (int)(ldg(p.seg) & 0xFFFFu) - 1;     // range 0 of pencil (0, 0)
                                else b0 = classify(p.bins, 0, 0, 0, false).own;
                                if (b0 >= 0) hist.add(b0, dc, 0u);
                            } else {
                                *dc_slot(smem) = dc;
                            }
                        }
                        if (MODE == MODE_BIN && kSegOk && p.seg) ptile(smem)[zl * PP + i] = pw;
                        else ptile(smem)[i * WZ + zl] = pw;
                    }
                };
            S::template fft_phase<FPH>(r.v[F], j, zl, p.tw, smem, [&](int) { return cmake<T>((T)0, (T)0); }, fin, true);
        } else {
            epilogue(p, tid, iy, iy_loc, iz, zl, active, r, smem, hist);
        }
    }

    // exact bin of one mode for accumulator `which`; -1 = none.  k below / above every bin maps to
    // the cacheable pseudo bins nb / nb+1 (decided on the guard intervals, which are conservative).
    static T21_HD int exact_bin(const Params& p, int ix, int iy, int iz, bool has_twin, int which, float sh,
                                int& weight) {
        const ModeBins mb = classify(p.bins, ix, iy, iz, has_twin);
        int bin = which == 0 ? mb.own : mb.twin;
        weight = which == 0 ? mb.own_w : 1;
        if (bin < 0) {
            const int nb = p.bins.nk * (p.bins.nmu > 0 ? p.bins.nmu : 1);
            weight = 1;
            if (sh < ldg(p.bins.k_hi_f + p.bins.nk)) bin = nb;
            else if (sh >= ldg(p.bins.k_lo_f + p.bins.nk + 1)) bin = nb + 1;
        }
        return bin;
    }

    template <class Hist>
    static T21_HD void epilogue(const Params& p, int tid, int iy, int iy_loc, int iz, int zl, bool active, Regs& r,
                                cx<T>* smem, Hist& hist) {
        if constexpr (MODE == MODE_BIN) {
            epilogue_bin(p, tid, iy, iz, zl, active, r, smem, hist);
        } else {
            epilogue_nd(p, tid, iy, iy_loc, iz, zl, active, smem);
        }
    }

#if T21_DEVICE_CODE
    // Table-driven epilogue: warp teams walk the pencils of the (pencil-major) power tile range by range.  A team of
    // TW warps shares a pencil (lanes and team warps interleave over |kx|), each warp books its partial sum of a range,
    // the first one also the range's closed-form mode count.  LO / HI: the kx >= 0 / kx < 0 half (signed mu walks them
    // separately, with the own and the twin tables).
    template <bool LO, bool HI, class Hist>
    static __device__ __forceinline__ void walk_tile_pencil(const T* __restrict__ pp, unsigned ent, int lane, int sub, int tw,
                                                            unsigned wt, Hist& hist) {
        walk_pencil_ranges<T, LO, HI>(pp, (unsigned)NX, ent, lane, sub, tw, wt, hist);
    }

    template <class Hist>
    static __device__ __forceinline__ void epilogue_bin_segments(const Params& p, int tid, int iy, int iz0, cx<T>* smem, Hist& hist) {
        constexpr int NWARP = THREADS / 32;
        constexpr int TW = NWARP > WZ ? NWARP / WZ : 1;            // warps per pencil
        const int lane = tid & 31, w = tid >> 5;
        const int sub = w / WZ;                                     // (TW == 1: w < WZ or the loop below strides)
        for (int pz = NWARP > WZ ? w % WZ : w; pz < WZ; pz += (NWARP > WZ ? WZ : NWARP)) {
            const int iz = iz0 + pz;
            if (iz >= p.nzh) continue;
            const bool has_twin = (iz != 0) && (iz != p.nz - iz);
            const T* pp = ptile(smem) + pz * PP;
            const unsigned* e0 = p.seg + ((size_t)iy * p.nzh + iz) * p.seg_stride + lane;
            const size_t tstep = (size_t)p.seg_pencils * p.seg_stride;
            const bool on = lane < p.seg_stride;
            if (MU != 2) {
                walk_tile_pencil<true, true>(pp, on ? ldg(e0) : kSegSentinel, lane, NWARP > WZ ? sub : 0, TW, has_twin ? 2u : 1u, hist);
            } else {
                walk_tile_pencil<true, false>(pp, on ? ldg(e0) : kSegSentinel, lane, NWARP > WZ ? sub : 0, TW, 1u, hist);
                walk_tile_pencil<false, true>(pp, on ? ldg(e0 + tstep) : kSegSentinel, lane, NWARP > WZ ? sub : 0, TW, 1u, hist);
                if (has_twin) {
                    walk_tile_pencil<true, false>(pp, on ? ldg(e0 + 2 * tstep) : kSegSentinel, lane, NWARP > WZ ? sub : 0, TW, 1u, hist);
                    walk_tile_pencil<false, true>(pp, on ? ldg(e0 + 3 * tstep) : kSegSentinel, lane, NWARP > WZ ? sub : 0, TW, 1u, hist);
                }
            }
        }
    }
#endif

    template <class Hist>
    static T21_HD void epilogue_bin(const Params& p, int tid, int iy, int iz, int zl, bool active, Regs& r,
                                    cx<T>* smem, Hist& hist) {
#if T21_DEVICE_CODE
        if constexpr (kSegOk) {
            if (p.seg) { epilogue_bin_segments(p, tid, iy, iz - zl, smem, hist); return; }
        }
#endif
        const t21_binspec& B = p.bins;
        const int q = tid / WZ;                                   // row id of this thread, 0..TP-1
        const int ix0 = (q / RPI) * (E * RPI) + (q % RPI);
        const bool has_twin = (iz != 0) && (iz != p.nz - iz);
        const int nb = B.nk * (MU ? B.nmu : 1);
        const int los = MU ? B.los_axis : -1;
        const int zx = B.zero_idx[0], nyx = B.nyq_idx[0];
        // k_perp = 0 line (power_spectrum.py:563-575): the part of the test that is fixed for this thread
        bool line_fixed = false;
        if (MU && B.exclude_zero)
            line_fixed = los == 0 ? (iy == B.zero_idx[1] && iz == B.zero_idx[2])
                                  : (los == 1 ? iz == B.zero_idx[2] : iy == B.zero_idx[1]);
        const int iz_c = active ? iz : 0;
        const float c_yz = ldg(B.tab_sf[1] + iy) + ldg(B.tab_sf[2] + iz_c);
        const unsigned wt = (has_twin && MU != 2) ? 2u : 1u;
        const T wf = (T)wt;
        float klos_c = 0.f;
        bool nyq_c = false;
        if (MU && los != 0) {
            const int il = los == 1 ? iy : iz_c;
            klos_c = ldg(B.tab_losf + il);
            nyq_c = il == B.nyq_idx[los];
        }
#if T21_DEVICE_CODE
        if constexpr (THREADS % 32 == 0 && E % 2 == 0) {
            // kx and -kx have the same kx^2, hence the same s, |mu| and (unless the line of sight is x)
            // the same signed mu: rows ix and NX-ix share ONE classification.  The warp sweeps the lower
            // half of kx and adds both powers; row 0 pairs with itself, the Nyquist row is done at the end.
            if (MU != 2 || los != 0) {
                constexpr int EH = E / 2;
                const int r0 = (q / RPI) * (EH * RPI) + (q % RPI);
                const T* pt2 = ptile(smem) + zl;
                auto step = [&](int ix, const T* plo_p, const T* phi_p, const float* sx_p, bool lane_on, bool self) {
                    const T plo = *plo_p, phi = *phi_p;
                    const T psum = self ? plo : plo + phi;
                    const double pw = (double)(psum * wf);
                    const unsigned wn = self ? wt : 2u * wt;
                    bool valid = lane_on && active;
                    if (MU) valid = valid && !(line_fixed && (los == 0 || ix == zx));
                    const float sh = ldg(sx_p) + c_yz;
                    float muh = 0.f;
                    if (MU) muh = (los == 0 ? ldg(B.tab_losf + ix) : klos_c) * fast_rsqrt(sh);
                    const float m_own = MU == 1 ? fabsf(muh) : muh;
                    if (!warpbins_fast<(MU != 0)>(r.acc[0], sh, m_own, pw, wn, valid)) {
                        int w1 = 1;
                        const int bin = valid ? exact_bin(p, ix, iy, iz, has_twin, 0, sh, w1) : -1;
                        const unsigned wx = self ? (unsigned)w1 : 2u * (unsigned)w1;
                        warpbins_slow(r.acc[0], hist, B, nb, bin, (w1 == (int)wt) ? pw : (double)psum, wx, valid);
                    }
                    if constexpr (MU == 2) {                  // los != 0 here: the twin's mu is the same for ix and NX-ix
                        const bool tv = valid && has_twin;
                        const float m_tw = nyq_c ? muh : -muh;
                        if (!warpbins_fast<true>(r.acc[1], sh, m_tw, (double)psum, self ? 1u : 2u, tv)) {
                            int w1 = 1;
                            const int bin = tv ? exact_bin(p, ix, iy, iz, has_twin, 1, sh, w1) : -1;
                            warpbins_slow(r.acc[1], hist, B, nb, bin, (double)psum, self ? 1u : 2u, tv);
                        }
                    }
                };
                // loop-carried row index, made opaque so it is advanced, not recomputed from the thread id
                // (an integer, not a pointer: the tile accesses must stay shared-memory loads)
                int ix = r0;
                asm volatile("" : "+r"(ix));
                if constexpr (MU == 0 && WZ <= 32) {
                    // Shells, every column of the tile a stored mode (CTA-uniform): s^ grows with kz and with kx
                    // over the rows a warp covers in one step (fp32 addition is monotonic), so the step's smallest
                    // and largest s^ sit in two known lanes.  If both are inside one cached interval, so is every
                    // lane: the whole step is two loads and an add per lane, with no per-lane test and no vote.
                    if ((iz - zl) + WZ <= p.nzh) {
                        WarpBins<false>& acc = r.acc[0];
                        const float c_lo = __shfl_sync(0xffffffffu, c_yz, 0), c_hi = __shfl_sync(0xffffffffu, c_yz, WZ - 1);
                        const unsigned w2 = 2u * wt;
                        const int lane_row = q % RPI;
                        // The same test for the warp's whole sweep (EH steps, rows first..first+EH*RPI-1): when it
                        // passes, the sweep is a sum of 2*EH tile values per lane, added to the bin in one go.
                        const int first = ix - lane_row;
                        bool swept = false;
                        {
                            const float t_lo = ldg(B.tab_sf[0] + first) + c_lo;
                            const float t_hi = ldg(B.tab_sf[0] + first + (EH * RPI - 1)) + c_hi;
                            const bool v0 = t_lo >= acc.slo[0] && t_hi < acc.shi[0];
                            const bool v1 = t_lo >= acc.slo[1] && t_hi < acc.shi[1];
                            if (v0 || v1) {
                                const T* lo = pt2 + ix * WZ;
                                const T* hi = pt2 + (NX - RPI - ix) * WZ;               // mirror of the SECOND row
                                const bool self = ix == 0;                             // row 0 pairs with itself
                                T ps = self ? lo[0] : lo[0] + hi[RPI * WZ];
#pragma unroll
                                for (int it = 1; it < EH; ++it) ps += lo[it * RPI * WZ] + hi[-(it - 1) * RPI * WZ];
                                const double pw = (double)(ps * wf);
                                const unsigned wn = (unsigned)EH * w2 - (self ? wt : 0u);
                                if (v0) { acc.sum[0] += pw; acc.cnt[0] += wn; }
                                else { acc.sum[1] += pw; acc.cnt[1] += wn; }
                                swept = true;
                            }
                        }
#pragma unroll 1
                        for (int it = 0; it < (swept ? 0 : EH); ++it) {
                            const int ixw = ix - lane_row;                       // first row of this warp's step
                            const float s_lo = ldg(B.tab_sf[0] + ixw) + c_lo, s_hi = ldg(B.tab_sf[0] + ixw + (RPI - 1)) + c_hi;
                            const bool u0 = s_lo >= acc.slo[0] && s_hi < acc.shi[0];
                            const bool u1 = s_lo >= acc.slo[1] && s_hi < acc.shi[1];
                            if (u0 || u1) {
                                const bool self = ix == 0;
                                const T plo = pt2[ix * WZ], phi = pt2[((NX - ix) & (NX - 1)) * WZ];
                                const double pw = (double)((self ? plo : plo + phi) * wf);
                                const unsigned wn = self ? wt : w2;
                                if (u0) { acc.sum[0] += pw; acc.cnt[0] += wn; }
                                else { acc.sum[1] += pw; acc.cnt[1] += wn; }
                            } else {
                                step(ix, pt2 + ix * WZ, pt2 + ((NX - ix) & (NX - 1)) * WZ, B.tab_sf[0] + ix, true, ix == 0);
                            }
                            ix += RPI;
                        }
                        // the Nyquist row goes to the last warp (the first one has the self-paired row 0)
                if (tid >= THREADS - 32) step(NX / 2, pt2 + (NX / 2) * WZ, pt2 + (NX / 2) * WZ, B.tab_sf[0] + NX / 2, q == S::TP - RPI, true);
                        return;
                    }
                }
#pragma unroll 1
                for (int it = 0; it < EH; ++it) {
                    step(ix, pt2 + ix * WZ, pt2 + ((NX - ix) & (NX - 1)) * WZ, B.tab_sf[0] + ix, true, ix == 0);   // row 0 pairs with itself
                    ix += RPI;
                }
                // the Nyquist row goes to the last warp (the first one has the self-paired row 0)
                if (tid >= THREADS - 32) step(NX / 2, pt2 + (NX / 2) * WZ, pt2 + (NX / 2) * WZ, B.tab_sf[0] + NX / 2, q == S::TP - RPI, true);
                return;
            }
        }
#endif
        const T* pt = ptile(smem) + ix0 * WZ + zl;
        const float* sfx = B.tab_sf[0] + ix0;
        const float* lfx = B.tab_losf + ix0;
#if T21_DEVICE_CODE
#pragma unroll 1
#endif
        for (int it = 0; it < E; ++it, pt += RPI * WZ, sfx += RPI, lfx += RPI) {
            const int ix = ix0 + it * RPI;
            const double pw = (double)(*pt * wf);
            bool valid = active;
            if (MU) valid = valid && !(line_fixed && (los == 0 || ix == zx));
#if T21_DEVICE_CODE
            if constexpr (THREADS % 32 == 0) {
                const float sh = ldg(sfx) + c_yz;
                float muh = 0.f;
                bool nyq = nyq_c;
                if (MU) {
                    float kl = klos_c;
                    if (los == 0) { kl = ldg(lfx); nyq = ix == nyx; }
                    muh = kl * fast_rsqrt(sh);
                }
                const float m_own = MU == 1 ? fabsf(muh) : muh;
                if (!warpbins_fast<(MU != 0)>(r.acc[0], sh, m_own, pw, wt, valid)) {
                    int w1 = 1;
                    const int bin = valid ? exact_bin(p, ix, iy, iz, has_twin, 0, sh, w1) : -1;
                    warpbins_slow(r.acc[0], hist, B, nb, bin, w1 == (int)wt ? pw : (double)*pt, (unsigned)w1, valid);
                }
                if constexpr (MU == 2) {                      // the twin goes to the mirrored mu bin
                    const bool tv = valid && has_twin;
                    const float m_tw = nyq ? muh : -muh;
                    if (!warpbins_fast<true>(r.acc[1], sh, m_tw, pw, 1u, tv)) {
                        int w1 = 1;
                        const int bin = tv ? exact_bin(p, ix, iy, iz, has_twin, 1, sh, w1) : -1;
                        warpbins_slow(r.acc[1], hist, B, nb, bin, pw, 1u, tv);
                    }
                }
                continue;
            }
#endif
            // per-thread path: CTAs smaller than a warp, and the CPU replay
            if (valid) {
                const ModeBins mb = classify(B, ix, iy, iz, has_twin);
                acc_add_thread(r.acc[0], hist, mb.own, (double)*pt, mb.own_w);
                if constexpr (MU == 2) acc_add_thread(r.acc[1], hist, mb.twin, (double)*pt, 1);
            }
        }
    }

    static T21_HD void epilogue_nd(const Params& p, int tid, int iy, int iy_loc, int iz, int zl, bool active, cx<T>* smem) {
        const T* pt = ptile(smem);
        const int q = tid / WZ;
        const int ix0 = (q / RPI) * (E * RPI) + (q % RPI);
        const bool has_twin = (iz != 0) && (iz != p.nz - iz);
        if (!active) return;
#if T21_DEVICE_CODE
#pragma unroll 1
#endif
        for (int it = 0; it < E; ++it) {
            const int ix = ix0 + it * RPI;
            double pw = (double)pt[ix * WZ + zl];
            if ((ix | iy | iz) == 0) pw = *dc_slot(smem);
            const double val = pw * p.scale;
            if (p.nd_half) {
                const long long oh = ((long long)ix * p.ny + iy_loc) * p.nd_half + iz;
                if (p.nd_f64) ((double*)p.nd_out)[oh] = val; else ((float*)p.nd_out)[oh] = (float)val;
                continue;
            }
            const int hy = p.ny / 2, hz = p.nz / 2;
            const int sx = ix + NX / 2 - (ix + NX / 2 >= NX ? NX : 0);
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
};

}  // namespace t21
