// Any-length fallbacks for the three passes: direct O(n^2) DFTs from a root table.
//
// The Stockham kernels cover power-of-two pencils from 8 to 4096.  Real data also comes in
// 250^3 / 300^3 / 600^3 grids and odd, non-cubic light-cone chunks, and
// tiny or degenerate axes (1-D / 2-D inputs are 3-D arrays with leading axes of length 1).
// These kernels keep the same workspace layout and argument structs, so every axis picks its
// own implementation.  They are still CUDA -- there is no CPU path -- but they make no attempt
// at speed-of-light: each output is a dot product with the n-th roots of unity (index k*j mod n,
// so the twiddles are exact table entries).
//
// Round 2: lengths whose prime factors are all <= 13 (250 = 2 5^3, 300 = 2^2 3 5^2, 600, 1000, odd light-cone chunks
// like 243 or 315 ...) no longer pay O(n^2): the tile a CTA already holds in shared memory is transformed there by a
// mixed-radix Stockham FFT (radix 4 / 2 butterflies hard-wired, odd prime radices as small DFTs whose twiddles are
// exact entries of the same root table), ping-ponging between two buffers.  O(n log n) per pencil, same layout, same
// epilogues; only lengths with a larger prime factor (127, 509 ...) keep the direct sums.
#include <string.h>
#include "t21_launch.cuh"

namespace t21 {

constexpr int kGenThreads = 256;

// ---- mixed-radix FFT of the pencils of a shared-memory tile -----------------------------------------------------------
struct Factors {
    int nf;             // 0: not smooth, use the direct sums
    int radix[24];
};

static Factors factorize(int n) {
    Factors f;
    memset(&f, 0, sizeof(f));
    if (n < 2) return f;
    int m = n, k = 0;
    while (m % 4 == 0) { f.radix[k++] = 4; m /= 4; }
    for (int p : {2, 3, 5, 7, 11, 13})
        while (m % p == 0) { f.radix[k++] = p; m /= p; }
    f.nf = m == 1 ? k : 0;
    return f;
}

// one butterfly of radix R on v[0..R-1] (already twiddled), natural order in and out.  Odd R: the R x R DFT with its
// twiddles w_R^(jk) read from the n-th root table at stride n / R (exact entries), R^2 complex multiply-adds.
template <int R, typename T>
__device__ __forceinline__ void small_dft(cx<T>* v, const cx<T>* __restrict__ roots, int n) {
    if constexpr (R == 2) {
        const cx<T> a = v[0], b = v[1];
        v[0] = cadd(a, b); v[1] = csub(a, b);
    } else if constexpr (R == 4) {
        const cx<T> a = cadd(v[0], v[2]), b = csub(v[0], v[2]), c = cadd(v[1], v[3]), d = csub(v[1], v[3]);
        const cx<T> dmi = cmake<T>(d.y, -d.x);                 // -i d
        v[0] = cadd(a, c); v[2] = csub(a, c); v[1] = cadd(b, dmi); v[3] = csub(b, dmi);
    } else if constexpr (R == 3) {
        // X1 = v0 - (v1 + v2)/2 - i s (v1 - v2), X2 = conj-symmetric partner, s = sin(pi/3)
        const T s3 = (T)0.86602540378443864676;
        const cx<T> t1 = cadd(v[1], v[2]), d = csub(v[1], v[2]);
        const cx<T> t2 = cmake<T>(v[0].x - (T)0.5 * t1.x, v[0].y - (T)0.5 * t1.y);
        const cx<T> q = cmake<T>(s3 * d.y, -s3 * d.x);                 // -i s d
        v[0] = cadd(v[0], t1); v[1] = cadd(t2, q); v[2] = csub(t2, q);
    } else if constexpr (R == 5) {
        const T c1 = (T)0.30901699437494742410, c2 = (T)-0.80901699437494742410;      // cos(2 pi/5), cos(4 pi/5)
        const T s1 = (T)0.95105651629515357212, s2 = (T)0.58778525229247312917;       // sin(2 pi/5), sin(4 pi/5)
        const cx<T> a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]), b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
        const cx<T> r1 = cmake<T>(v[0].x + c1 * a1.x + c2 * a2.x, v[0].y + c1 * a1.y + c2 * a2.y);
        const cx<T> r2 = cmake<T>(v[0].x + c2 * a1.x + c1 * a2.x, v[0].y + c2 * a1.y + c1 * a2.y);
        const cx<T> q1 = cmake<T>(s1 * b1.x + s2 * b2.x, s1 * b1.y + s2 * b2.y);
        const cx<T> q2 = cmake<T>(s2 * b1.x - s1 * b2.x, s2 * b1.y - s1 * b2.y);
        v[0] = cmake<T>(v[0].x + a1.x + a2.x, v[0].y + a1.y + a2.y);
        v[1] = cmake<T>(r1.x + q1.y, r1.y - q1.x);                     // r1 - i q1
        v[4] = cmake<T>(r1.x - q1.y, r1.y + q1.x);                     // r1 + i q1
        v[2] = cmake<T>(r2.x + q2.y, r2.y - q2.x);
        v[3] = cmake<T>(r2.x - q2.y, r2.y + q2.x);
    } else {
        cx<T> in[R];
#pragma unroll
        for (int j = 0; j < R; ++j) in[j] = v[j];
        const int step = n / R;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            cx<T> acc = in[0];
            int idx = 0;
#pragma unroll
            for (int j = 1; j < R; ++j) {
                idx += k * step;
                if (idx >= n) idx -= n;             // k step < n
                acc = cadd(acc, cmul(in[j], ldg(roots + idx)));
            }
            v[k] = acc;
        }
    }
}

template <int R, int NCOL, typename T>
__device__ __forceinline__ void mixed_stage(const cx<T>* __restrict__ a, cx<T>* __restrict__ b, int n, int ns,
                                            const cx<T>* __restrict__ roots, int tid, int nthreads) {
    const unsigned nb = (unsigned)n / R;                   // butterflies per pencil
    const unsigned tstep = (unsigned)n / ((unsigned)ns * R);   // root index of w_(ns R)^1
    for (unsigned w = tid; w < nb * NCOL; w += nthreads) {
        const unsigned jj = w / NCOL, col = w % NCOL;      // NCOL is a power of two
        const unsigned m = jj % (unsigned)ns;
        cx<T> v[R];
        unsigned idx = 0;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            cx<T> x = a[(jj + k * nb) * NCOL + col];
            if (k > 0 && m > 0) {
                idx += m * tstep;               // k m n / (ns R) < n: a single conditional subtraction keeps it in range
                if (idx >= (unsigned)n) idx -= (unsigned)n;
                x = cmul(x, ldg(roots + idx));
            }
            v[k] = x;
        }
        small_dft<R, T>(v, roots, n);
        const unsigned base = (jj - m) * R + m;
#pragma unroll
        for (int k = 0; k < R; ++k) b[(base + k * ns) * NCOL + col] = v[k];
    }
}

// transforms the NCOL pencils of tile a[n][NCOL] (forward, exp(-i ...)); b is scratch of the same size.  All threads of
// the CTA take part (CTA barriers inside).  Returns the buffer that holds the result (a or b), natural order.
template <int NCOL, typename T>
__device__ __forceinline__ cx<T>* mixed_fft(cx<T>* a, cx<T>* b, int n, const Factors& f, const cx<T>* __restrict__ roots) {
    int ns = 1;
    for (int s = 0; s < f.nf; ++s) {
        const int R = f.radix[s];
        const int tid = threadIdx.x, nt = blockDim.x;
        switch (R) {
            case 2: mixed_stage<2, NCOL, T>(a, b, n, ns, roots, tid, nt); break;
            case 3: mixed_stage<3, NCOL, T>(a, b, n, ns, roots, tid, nt); break;
            case 4: mixed_stage<4, NCOL, T>(a, b, n, ns, roots, tid, nt); break;
            case 5: mixed_stage<5, NCOL, T>(a, b, n, ns, roots, tid, nt); break;
            case 7: mixed_stage<7, NCOL, T>(a, b, n, ns, roots, tid, nt); break;
            case 11: mixed_stage<11, NCOL, T>(a, b, n, ns, roots, tid, nt); break;
            default: mixed_stage<13, NCOL, T>(a, b, n, ns, roots, tid, nt); break;
        }
        __syncthreads();
        cx<T>* t = a; a = b; b = t;
        ns *= R;
    }
    return a;
}

template <typename T> struct GenWZ { static constexpr int value = sizeof(T) == 4 ? 8 : 4; };

// ---- Z: real rows -> nz/2+1 complex ---------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kGenThreads)
gen_zpass_kernel(const T* __restrict__ in, cx<T>* __restrict__ out, long long nrows, int nz, int nzh, int pitch,
                 int ny, long long nxa, const T* __restrict__ offset, const cx<T>* __restrict__ roots) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* row = reinterpret_cast<T*>(smem_raw);
    const T off = offset ? ldg(offset) : (T)0;
    for (long long r = blockIdx.x; r < nrows; r += gridDim.x) {
        __syncthreads();
        for (int z = threadIdx.x; z < nz; z += kGenThreads) row[z] = in[r * nz + z] - off;
        __syncthreads();
        const long long x = r / ny, y = r - x * ny;
        constexpr int LW = WLay<T>::LW;
        const long long cstride = nxa * ny * LW;
        cx<T>* dst0 = out + (x * ny + y) * LW;
        for (int k = threadIdx.x; k < nzh; k += kGenThreads) {
            T sr = 0, si = 0;
            int idx = 0;                       // k*z mod nz
            for (int z = 0; z < nz; ++z) {
                const cx<T> w = ldg(roots + idx);
                sr += row[z] * w.x;
                si += row[z] * w.y;
                idx += k;
                if (idx >= nz) idx -= nz;
            }
            dst0[(long long)(k / LW) * cstride + (k % LW)] = cmake<T>(sr, si);
        }
    }
}

// smooth nz: 2 * RZ real rows at a time, packed in pairs as the real and imaginary parts of the RZ columns of a
// [nz][RZ] complex tile: one mixed-radix FFT transforms both rows of a pair, Z[k] = X1[k] + i X2[k], and they come apart
// as X1[k] = (Z[k] + conj Z[nz-k]) / 2, X2[k] = (Z[k] - conj Z[nz-k]) / 2i.  The load writes shared memory
// contiguously (element e of the tile, as floats, is sample z = e / (2 RZ) of row e % (2 RZ)).
constexpr int kGenRZ = 8;
template <typename T>
__global__ void __launch_bounds__(kGenThreads)
gen_zpass_fft_kernel(const T* __restrict__ in, cx<T>* __restrict__ out, long long nrows, int nz, int nzh, int ny, long long nxa,
                     const T* __restrict__ offset, const cx<T>* __restrict__ roots, const __grid_constant__ Factors f) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<T>* ta = reinterpret_cast<cx<T>*>(smem_raw);
    cx<T>* tb = ta + (size_t)nz * kGenRZ;
    const T off = offset ? ldg(offset) : (T)0;
    constexpr int LW = WLay<T>::LW;
    constexpr int NR = 2 * kGenRZ;                                        // real rows per tile
    const long long cstride = nxa * ny * LW;
    const long long ngroups = (nrows + NR - 1) / NR;
    for (long long g = blockIdx.x; g < ngroups; g += gridDim.x) {
        const long long r0 = g * NR;
        __syncthreads();
        T* taf = reinterpret_cast<T*>(ta);
        // a warp takes 4 consecutive z of 8 rows: 16-byte runs of global memory (the rest of each sector is an L1 hit on
        // the next trip), a 2-way bank conflict on the store instead of the 16-way one of a z-major deal
        for (int e = threadIdx.x; e < ((nz + 3) / 4) * 4 * NR; e += kGenThreads) {
            const int rr = (e & 7) | (((e >> 5) & 1) << 3), z = ((e >> 6) << 2) | ((e >> 3) & 3);
            const long long r = r0 + rr;
            if (z < nz) taf[z * NR + rr] = r < nrows ? in[r * nz + z] - off : (T)0;
        }
        __syncthreads();
        const cx<T>* res = mixed_fft<kGenRZ, T>(ta, tb, nz, f, roots);
        // lanes along the LW columns of a 64-byte piece, then along the row pairs: a store instruction writes whole pieces
        const int nkc = (nzh + LW - 1) / LW;
        for (int e = threadIdx.x; e < nkc * kGenRZ * LW; e += kGenThreads) {
            const int kl = e % LW, c = (e / LW) % kGenRZ, kc = e / (LW * kGenRZ);
            const int k = kc * LW + kl;
            if (k >= nzh) continue;
            const cx<T> zk = res[k * kGenRZ + c], zm = res[(k ? nz - k : 0) * kGenRZ + c];
            const cx<T> x1 = cmake<T>((T)0.5 * (zk.x + zm.x), (T)0.5 * (zk.y - zm.y));
            const cx<T> x2 = cmake<T>((T)0.5 * (zk.y + zm.y), (T)0.5 * (zm.x - zk.x));
            const long long ra = r0 + 2 * c;
            const long long kofs = (long long)kc * cstride + kl;
            // (row (x, y) sits at x * ny + y = its row number)
            if (ra < nrows) out[kofs + ra * LW] = x1;
            if (ra + 1 < nrows) out[kofs + (ra + 1) * LW] = x2;
        }
    }
}

// ---- Y: in-place strided pencils --------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kGenThreads)
gen_ypass_kernel(cx<T>* __restrict__ w, long long nouter, int ny, int nzh, int pitch,
                 const cx<T>* __restrict__ roots, const __grid_constant__ Factors f) {
    constexpr int WZ = GenWZ<T>::value;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<T>* tile = reinterpret_cast<cx<T>*>(smem_raw);      // [ny][WZ] (+ a second one when the length is smooth)
    const int nchunks = (nzh + WZ - 1) / WZ;
    const long long ntiles = nouter * nchunks;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int chunk = (int)(t % nchunks);
        const long long x = t / nchunks;
        // WZ == LW here: the tile (all y, one 64-byte piece, fixed x) is one contiguous block
        static_assert(GenWZ<T>::value == WLay<T>::LW, "generic tiles are one layout piece wide");
        cx<T>* base = w + ((long long)chunk * nouter + x) * ny * WZ;
        const long long stride = WZ;
        __syncthreads();
        for (int e = threadIdx.x; e < ny * WZ; e += kGenThreads) {
            const int y = e / WZ, zl = e % WZ;
            tile[e] = (chunk * WZ + zl < nzh) ? base[y * stride + zl] : cmake<T>(0, 0);
        }
        __syncthreads();
        if (f.nf > 0) {
            const cx<T>* res = mixed_fft<WZ, T>(tile, tile + (size_t)ny * WZ, ny, f, roots);
            for (int e = threadIdx.x; e < ny * WZ; e += kGenThreads) {
                const int zl = e % WZ;
                if (chunk * WZ + zl < nzh) base[(e / WZ) * stride + zl] = res[e];
            }
            continue;
        }
        for (int e = threadIdx.x; e < ny * WZ; e += kGenThreads) {
            const int k = e / WZ, zl = e % WZ;
            if (chunk * WZ + zl >= nzh) continue;
            T sr = 0, si = 0;
            int idx = 0;
            for (int y = 0; y < ny; ++y) {
                const cx<T> v = tile[y * WZ + zl], r = ldg(roots + idx);
                sr += v.x * r.x - v.y * r.y;
                si += v.x * r.y + v.y * r.x;
                idx += k;
                if (idx >= ny) idx -= ny;
            }
            base[k * stride + zl] = cmake<T>(sr, si);
        }
    }
}

// ---- X: pencils fused with power + binning or the n-D store -------------------------------------
// HK: 0 shared-memory histogram + per-CTA partials, 1 global atomics, 2 n-D store
template <typename T, int NF, int HK>
__global__ void __launch_bounds__(kGenThreads)
gen_xpass_kernel(const cx<T>* __restrict__ w0, const cx<T>* __restrict__ w1, int nx, int ny, int nz, int nzh,
                 int pitch, const cx<T>* __restrict__ roots, const T* off0, const T* off1, double ntot,
                 const __grid_constant__ t21_binspec B, int nbins, double* __restrict__ part_sum,
                 unsigned* __restrict__ part_cnt, double* gsum, unsigned long long* gcnt, void* nd_out, int nd_f64,
                 int nd_half, double scale, const __grid_constant__ Factors f) {
    constexpr int WZ = GenWZ<T>::value;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cx<T>* tile0 = reinterpret_cast<cx<T>*>(smem_raw);                 // [nx][WZ]
    cx<T>* tile1 = tile0 + (NF == 2 ? (size_t)nx * WZ : 0);
    cx<T>* scratch = tile0 + (size_t)NF * nx * WZ;                     // (smooth lengths only) ping-pong partner
    double* hs = reinterpret_cast<double*>(tile0 + (size_t)(NF + (f.nf > 0 ? 1 : 0)) * nx * WZ);
    unsigned* hc = reinterpret_cast<unsigned*>(hs + (HK == 0 ? nbins : 0));
    if (HK == 0) {
        for (int b = threadIdx.x; b < nbins; b += kGenThreads) { hs[b] = 0.0; hc[b] = 0u; }
    }
    SmemHist sh{hs, hc};
    GlobalHist gh{gsum, gcnt};
    RunAcc acc[2];
    run_init(acc[0]);
    run_init(acc[1]);
    const int nchunks = (nzh + WZ - 1) / WZ;
    const long long ntiles = (long long)ny * nchunks;
    const bool signed_mu = B.nmu > 0 && !B.absolute_mus;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int chunk = (int)(t % nchunks), iy = (int)(t / nchunks);
        __syncthreads();
        for (int e = threadIdx.x; e < nx * WZ; e += kGenThreads) {
            const int i = e / WZ, zl = e % WZ;
            const bool ok = chunk * WZ + zl < nzh;
            const long long o = (((long long)chunk * nx + i) * ny + iy) * WZ + zl;      // chunk-major, WZ == LW
            tile0[e] = ok ? w0[o] : cmake<T>(0, 0);
            if (NF == 2) tile1[e] = ok ? w1[o] : cmake<T>(0, 0);
        }
        __syncthreads();
        const cx<T>* res0 = tile0;
        const cx<T>* res1 = tile1;
        if (f.nf > 0) {
            // the transformed field ends in its own tile or in the scratch buffer; whichever is free serves the next field
            cx<T>* r0 = mixed_fft<WZ, T>(tile0, scratch, nx, f, roots);
            res0 = r0;
            if (NF == 2) res1 = mixed_fft<WZ, T>(tile1, r0 == tile0 ? scratch : tile0, nx, f, roots);
        }
        if (HK != 2 && f.nf > 0 && B.seg_tab != nullptr) {
            // Bins as index ranges (pencil segment table): powers go pencil-major into the buffer the transforms left
            // free, then every warp walks pencils range by range -- no per-mode classification, no per-mode atomics.
            cx<T>* freeb = (res0 != tile0 && (NF == 1 || res1 != tile0)) ? tile0
                         : ((NF == 2 && res0 != tile1 && res1 != tile1) ? tile1 : scratch);
            T* pw = reinterpret_cast<T*>(freeb);
            const int PPg = nx + 4;
            for (int e = threadIdx.x; e < nx * WZ; e += kGenThreads) {
                const int ix = e / WZ, zl = e % WZ, iz = chunk * WZ + zl;
                if (iz >= nzh) continue;
                const T ar = res0[e].x, ai = res0[e].y;
                const T br = NF == 2 ? res1[e].x : ar, bi = NF == 2 ? res1[e].y : ai;
                T pv = ar * br + ai * bi;
                if ((ix | iy | iz) == 0) {
                    const double c0 = off0 ? (double)ldg(off0) : 0.0;
                    const double c1 = (NF == 2 && off1) ? (double)ldg(off1) : c0;
                    const double dc = ((double)ar + c0 * ntot) * ((double)br + c1 * ntot) + (double)ai * (double)bi;
                    const int b0 = (int)(ldg(B.seg_tab) & 0xFFFFu) - 1;        // range 0 of pencil (0, 0)
                    if (b0 >= 0) { if (HK == 0) sh.add(b0, dc, 0u); else gh.add(b0, dc, 0u); }
                    pv = (T)0;                                                 // counted with its range, power added above
                }
                pw[zl * PPg + ix] = pv;
            }
            __syncthreads();
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            const size_t tstep = (size_t)ny * nzh * B.seg_stride;
            for (int pz = warp; pz < WZ; pz += kGenThreads / 32) {
                const int iz = chunk * WZ + pz;
                if (iz >= nzh) continue;
                const bool has_twin = (iz != 0) && (iz != nz - iz);
                const unsigned* e0 = B.seg_tab + ((size_t)iy * nzh + iz) * B.seg_stride + lane;
                const bool on = lane < B.seg_stride;
                const T* pp = pw + pz * PPg;
#define T21_GEN_WALK(LO, HI, ENT, WT)                                                                              \
                { if (HK == 0) walk_pencil_ranges<T, LO, HI>(pp, (unsigned)nx, ENT, lane, 0, 1, WT, sh);           \
                  else walk_pencil_ranges<T, LO, HI>(pp, (unsigned)nx, ENT, lane, 0, 1, WT, gh); }
                if (!signed_mu) {
                    T21_GEN_WALK(true, true, on ? ldg(e0) : kSegSentinel, has_twin ? 2u : 1u)
                } else {
                    T21_GEN_WALK(true, false, on ? ldg(e0) : kSegSentinel, 1u)
                    T21_GEN_WALK(false, true, on ? ldg(e0 + tstep) : kSegSentinel, 1u)
                    if (has_twin) {
                        T21_GEN_WALK(true, false, on ? ldg(e0 + 2 * tstep) : kSegSentinel, 1u)
                        T21_GEN_WALK(false, true, on ? ldg(e0 + 3 * tstep) : kSegSentinel, 1u)
                    }
                }
#undef T21_GEN_WALK
            }
            continue;
        }
        for (int e = threadIdx.x; e < nx * WZ; e += kGenThreads) {
            const int ix = e / WZ, zl = e % WZ, iz = chunk * WZ + zl;
            if (iz >= nzh) continue;
            T ar = 0, ai = 0, br = 0, bi = 0;
            int idx = 0;
            if (f.nf > 0) {
                ar = res0[e].x; ai = res0[e].y;
                if (NF == 2) { br = res1[e].x; bi = res1[e].y; }
            }
            for (int x = 0; x < (f.nf > 0 ? 0 : nx); ++x) {
                const cx<T> r = ldg(roots + idx);
                const cx<T> v = tile0[x * WZ + zl];
                ar += v.x * r.x - v.y * r.y;
                ai += v.x * r.y + v.y * r.x;
                if (NF == 2) {
                    const cx<T> u = tile1[x * WZ + zl];
                    br += u.x * r.x - u.y * r.y;
                    bi += u.x * r.y + u.y * r.x;
                }
                idx += ix;
                if (idx >= nx) idx -= nx;
            }
            if (NF == 1) { br = ar; bi = ai; }
            double pw;
            if ((ix | iy | iz) == 0) {
                const double c0 = off0 ? (double)ldg(off0) : 0.0;
                const double c1 = (NF == 2 && off1) ? (double)ldg(off1) : c0;
                pw = ((double)ar + c0 * ntot) * ((double)br + c1 * ntot) + (double)ai * (double)bi;
            } else {
                pw = (double)(ar * br + ai * bi);
            }
            const bool has_twin = (iz != 0) && (iz != nz - iz);
            if (HK == 2) {
                const double val = pw * scale;
                if (nd_half) {                  // the computed half spectrum, FFT order: [nx][ny][nd_half >= nz/2+1]
                    const long long oh = ((long long)ix * ny + iy) * nd_half + iz;
                    if (nd_f64) ((double*)nd_out)[oh] = val; else ((float*)nd_out)[oh] = (float)val;
                    continue;
                }
                const int hx = nx / 2, hy = ny / 2, hz = nz / 2;
                const int sx = (ix + hx) % nx, sy = (iy + hy) % ny, sz = (iz + hz) % nz;
                const long long o = ((long long)sx * ny + sy) * nz + sz;
                if (nd_f64) ((double*)nd_out)[o] = val; else ((float*)nd_out)[o] = (float)val;
                if (has_twin) {
                    const int ux = ((nx - ix) % nx + hx) % nx, uy = ((ny - iy) % ny + hy) % ny, uz = (nz - iz + hz) % nz;
                    const long long ot = ((long long)ux * ny + uy) * nz + uz;
                    if (nd_f64) ((double*)nd_out)[ot] = val; else ((float*)nd_out)[ot] = (float)val;
                }
            } else {
                const ModeBins mb = classify(B, ix, iy, iz, has_twin);
                if (HK == 0) {
                    run_add(acc[0], sh, mb.own, pw, mb.own_w);
                    if (signed_mu) run_add(acc[1], sh, mb.twin, pw, 1);
                } else {
                    run_add(acc[0], gh, mb.own, pw, mb.own_w);
                    if (signed_mu) run_add(acc[1], gh, mb.twin, pw, 1);
                }
            }
        }
    }
    if (HK == 0) {
        run_flush(acc[0], sh);
        run_flush(acc[1], sh);
        __syncthreads();
        for (int b = threadIdx.x; b < nbins; b += kGenThreads) {
            part_sum[(size_t)blockIdx.x * nbins + b] = hs[b];
            part_cnt[(size_t)blockIdx.x * nbins + b] = hc[b];
        }
    } else if (HK == 1) {
        run_flush(acc[0], gh);
        run_flush(acc[1], gh);
    }
}

// ---- launchers -------------------------------------------------------------------------------------
template <typename T>
static int gen_z(const ZArgs& a, cudaStream_t st) {
    const int nzh = a.nz / 2 + 1;
    const Factors f = factorize(a.nz);
    const size_t smem_fft = 2 * (size_t)a.nz * kGenRZ * sizeof(cx<T>);
    if (f.nf > 0 && smem_fft <= 200 * 1024 && !a.window) {
        auto kern = gen_zpass_fft_kernel<T>;
        if (int rc = ensure_smem(kern, smem_fft)) return rc;
        const long long ngroups = (a.nrows + 2 * kGenRZ - 1) / (2 * kGenRZ);
        long long g = ngroups < (long long)device_sm_count() * 8 ? ngroups : (long long)device_sm_count() * 8;
        if (g < 1) return 0;
        kern<<<(unsigned)g, kGenThreads, smem_fft, st>>>((const T*)a.in, (cx<T>*)a.out, a.nrows, a.nz, nzh, a.ny, a.nxa,
                                                         (const T*)a.offset, (const cx<T>*)a.twr, f);
        count_launch();
        T21_CUDA_OK(cudaGetLastError());
        return 0;
    }
    if (a.window) { set_error("windowed passes need power-of-two axes"); return T21_EINVAL; }
    const size_t smem = (size_t)a.nz * sizeof(T);
    auto kern = gen_zpass_kernel<T>;
    if (int rc = ensure_smem(kern, smem)) return rc;
    long long g = a.nrows < (long long)device_sm_count() * 16 ? a.nrows : (long long)device_sm_count() * 16;
    if (g < 1) return 0;
    kern<<<(unsigned)g, kGenThreads, smem, st>>>((const T*)a.in, (cx<T>*)a.out, a.nrows, a.nz, nzh, a.pitch, a.ny, a.nxa,
                                               (const T*)a.offset, (const cx<T>*)a.twr);
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    return 0;
}
int launch_zpass_generic(int dtype, const ZArgs& a, cudaStream_t st) {
    return dtype == T21_F32 ? gen_z<float>(a, st) : gen_z<double>(a, st);
}

template <typename T>
static int gen_y(const YArgs& a, cudaStream_t st) {
    constexpr int WZ = GenWZ<T>::value;
    size_t smem = (size_t)a.ny * WZ * sizeof(cx<T>);
    Factors f = factorize(a.ny);
    if (f.nf > 0 && 2 * smem <= 200 * 1024) smem *= 2; else f.nf = 0;      // smooth length: room for the ping-pong partner
    if (smem > 200 * 1024) { set_error("axis of length %d is too long for the generic pass", a.ny); return T21_EINVAL; }
    auto kern = gen_ypass_kernel<T>;
    if (int rc = ensure_smem(kern, smem)) return rc;
    const long long ntiles = a.nouter * ((a.nzh + WZ - 1) / WZ);
    long long g = ntiles < (long long)device_sm_count() * 8 ? ntiles : (long long)device_sm_count() * 8;
    if (g < 1) return 0;
    kern<<<(unsigned)g, kGenThreads, smem, st>>>((cx<T>*)a.w, a.nouter, a.ny, a.nzh, a.pitch, (const cx<T>*)a.tw, f);
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    return 0;
}
int launch_ypass_generic(int dtype, const YArgs& a, cudaStream_t st) {
    return dtype == T21_F32 ? gen_y<float>(a, st) : gen_y<double>(a, st);
}

template <typename T, int NF>
static int gen_x(XArgs& a, cudaStream_t st) {
    constexpr int WZ = GenWZ<T>::value;
    const bool is_bin = a.bins != nullptr;
    const int hk = !is_bin ? 2 : (a.hist_global ? 1 : 0);
    size_t smem = (size_t)NF * a.nx * WZ * sizeof(cx<T>);
    const size_t hist = hk == 0 ? (size_t)a.nbins * (sizeof(double) + sizeof(unsigned)) : 0;
    Factors f = factorize(a.nx);
    if (f.nf > 0 && smem + (size_t)a.nx * WZ * sizeof(cx<T>) + hist <= 200 * 1024) smem += (size_t)a.nx * WZ * sizeof(cx<T>);
    else f.nf = 0;
    smem += hist;
    if (smem > 200 * 1024) { set_error("axis of length %d is too long for the generic pass", a.nx); return T21_EINVAL; }
    const long long ntiles = (long long)a.ny * ((a.nzh + WZ - 1) / WZ);
    long long g = ntiles < (long long)device_sm_count() * 4 ? ntiles : (long long)device_sm_count() * 4;
    if (hk == 0 && g > a.max_grid) g = a.max_grid;
    if (g < 1) g = 1;
    t21_binspec dummy;
    memset(&dummy, 0, sizeof(dummy));
    t21_binspec B = is_bin ? *a.bins : dummy;
    {   // the segment table is used only when it matches the mode (four tables for signed mu, one otherwise)
        const bool is_signed = B.nmu > 0 && !B.absolute_mus;
        if (!x_use_segments() || B.seg_ntab != (is_signed ? 4 : 1) || B.seg_overflow != 0) B.seg_tab = nullptr;
    }
#define T21_GEN_X(HKV)                                                                                          \
    {                                                                                                           \
        auto kern = gen_xpass_kernel<T, NF, HKV>;                                                               \
        if (int rc = ensure_smem(kern, smem)) return rc;                                                        \
        kern<<<(unsigned)g, kGenThreads, smem, st>>>((const cx<T>*)a.w[0], (const cx<T>*)a.w[1], a.nx, a.ny, a.nz,  \
            a.nzh, a.pitch, (const cx<T>*)a.tw, (const T*)a.offset[0], (const T*)a.offset[1], a.ntot, B,        \
            a.nbins, a.part_sum, a.part_cnt, a.gsum, a.gcnt, a.nd_out, a.nd_f64, a.nd_half, a.scale, f);                      \
    }
    if (hk == 0) T21_GEN_X(0) else if (hk == 1) T21_GEN_X(1) else T21_GEN_X(2)
#undef T21_GEN_X
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    a.grid_used = (int)g;
    return 0;
}
int launch_xpass_generic(int dtype, XArgs& a, cudaStream_t st) {
    if (dtype == T21_F32) return a.nf == 1 ? gen_x<float, 1>(a, st) : gen_x<float, 2>(a, st);
    return a.nf == 1 ? gen_x<double, 1>(a, st) : gen_x<double, 2>(a, st);
}

}  // namespace t21
