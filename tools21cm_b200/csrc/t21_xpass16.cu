// X pass, warp-per-pencil form for nx = 512 (fp32; one field or two): no CTA barrier at all.
//
// 512 = 16 * 2 * 16.  A warp owns one pencil (one of the 8 columns of a 64-byte piece at one ky) from the first load to the
// histogram; the CTA's 8 warps only share the TMA-landed tile.
//   * the tile arrives as four column-PAIR sub-tiles ([512 rows][2 columns], TMA boxes of 256 rows x 16 B): a lane reads
//     rows lane + 32 k (k = 0..15) of its column at a 16-byte stride -- a 2-way bank conflict, against 16-way for the
//     plain [x][8] tile;
//   * lane j = 16 m0 + r transforms those 16 rows in registers (radix 16), lanes j and j ^ 16 combine (radix 2, one
//     shuffle per value: B[r][q0 + 16 q1] = C0[r][q0] + (-1)^q1 w32^q0 C1[r][q0]), a 32 x 16 transposition through a
//     warp-private strip of shared memory (pitch 17: conflict-free both ways, __syncwarp only) hands lane q the values
//     B[r][q], r = 0..15, which take their twiddle w512^(r q) and a second in-register radix-16 transform:
//     v[k'] = X[q + 32 k'] -- the layout the range walk of the nx = 1024 kernel works on, with 8 steps instead of 16;
//   * cross spectra keep the first field's 16 results in registers (32 of them) while the second field is transformed:
//     Re(F1 conj F2) per mode, then the same walk.  (At nx = 1024 the two fields' 64 values per lane do not fit.)
// Bins come from the pencil segment table(s) (t21_build_pencil_segments; four for signed mu); specs without one and tables
// with overflow marks take the CTA-per-tile kernel.  BASELINE configs 2 / 3 (512^3 cross spectra) run here.
#include <string.h>
#include "t21_launch.cuh"
#include "t21_tma.cuh"

namespace t21 {

namespace x16 {

constexpr int NX = 512;
constexpr int WZ = 8;
constexpr int THREADS = 256;
constexpr int NV = NX / 32;                           // values per lane
constexpr int SUB_BYTES = NX * 2 * 8;                 // one column-pair sub-tile
constexpr int TILE_BYTES = 4 * SUB_BYTES;             // per field
constexpr int XPITCH = 17;                            // transposition strip: 32 rows (q) of 16 (r) + 1
constexpr int XCH_BYTES = 32 * XPITCH * 8;            // 4352 per warp; reused by the walk (4 slots x 3 x 32 floats = 1536 B)

struct Params {
    int ny, nz, nzh, nchunks;
    long long ntiles;
    int iy0;
    int box_rows, xs_log2;
    const cx<float>* tw;      // w512^(r q), [16][32]
    const float* offset[2];
    double ntot;
    const unsigned* seg;
    int seg_stride;
    int seg_ntab;             // 1, or 4 for signed mu (own / twin bins for kx >= 0 and kx < 0)
    long long seg_pencils;
};

// tile buffers: one field keeps TWO tiles in shared memory (the next one is requested as soon as every warp has read the
// one before the current -- a whole tile's work ahead of its use); two fields have room for one (128 KB would leave a
// single CTA per SM), requested when every warp has read the current one.
template <int NF> struct Bufs { static constexpr int N = NF == 1 ? 2 : 1; };

#if T21_DEVICE_CODE

// close-range walk on 16 powers per lane: steps k = 0..7 hold |kx| = 32k + lane (pw[k]) and 32k + 32 - lane (pw[15 - k]).
// LO / HI: take the kx >= 0 / the kx < 0 half (signed mu walks them separately, with the own and the twin tables).
template <bool LO, bool HI, class Hist>
__device__ __forceinline__ void walk16(const float* pw, int lane, unsigned ent, float wf, unsigned wt, float* __restrict__ strip, Hist& hist) {
    const unsigned pos = ent >> 16, binp1 = ent & 0xFFFFu;
    const unsigned S = (unsigned)__popc(__ballot_sync(0xffffffffu, pos != 0xFFFFu));
    if (S == 1u && __shfl_sync(0xffffffffu, binp1, 0) == 0u) return;
    const unsigned mixed = __reduce_or_sync(0xffffffffu, ((unsigned)lane >= 1u && (unsigned)lane < S) ? 1u << ((pos - 1u) >> 5) : 0u);
    float cur = 0.f;
    unsigned j = 0u;
    unsigned nextpos = S > 1u ? __shfl_sync(0xffffffffu, pos, 1) : 0xFFFFu;
    auto close = [&](unsigned b) {
        float t = cur;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        const unsigned a = __shfl_sync(0xffffffffu, pos, j);
        const int bin = (int)__shfl_sync(0xffffffffu, binp1, j) - 1;
        if (lane == 0 && bin >= 0) hist.add(bin, (double)(t * wf), segment_modes(LO, HI, a, b, (unsigned)(NX / 2)) * wt);
    };
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        float run = 0.f;
        unsigned slot = 0u, ks = 0u;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int k = half * 4 + kk;
            if ((mixed >> k) & 1u) {
                float* d = strip + slot * 96u + lane;
                d[0] = run; d[32] = LO ? pw[k] : 0.f; d[64] = HI ? pw[15 - k] : 0.f;
                run = 0.f;
                ks |= (unsigned)k << (4u * slot);
                ++slot;
            } else {
                run += (LO ? pw[k] : 0.f) + (HI ? pw[15 - k] : 0.f);
            }
        }
#pragma unroll 1
        for (unsigned i = 0; i < slot; ++i) {
            const float* d = strip + i * 96u + lane;
            cur += d[0];
            float plo = d[32], phi = d[64];
            const unsigned k32 = ((ks >> (4u * i)) & 15u) << 5;
            const unsigned flo = k32 + (unsigned)lane, fhi = k32 + 32u - (unsigned)lane;
            while (nextpos <= k32 + 32u) {
                if (flo < nextpos) { cur += plo; plo = 0.f; }
                if (fhi < nextpos) { cur += phi; phi = 0.f; }
                close(nextpos);
                cur = 0.f;
                ++j;
                nextpos = j + 1u < S ? __shfl_sync(0xffffffffu, pos, (j + 1u) & 31u) : 0xFFFFu;
            }
            cur += plo + phi;
        }
        cur += run;
    }
    close((unsigned)(NX / 2 + 1));
}

// t[q0] = hi ? w32^q0 a[q0] : a[q0] (the twiddle constants are template parameters: folded at compile time)
template <int Q>
__device__ __forceinline__ void rot_all_impl(const cx<float>* a, cx<float>* t, bool hi) {
    if constexpr (Q < 16) {
        const cx<float> r = rot<32, Q, float>(a[Q]);
        t[Q] = hi ? r : a[Q];
        rot_all_impl<Q + 1>(a, t, hi);
    }
}
__device__ __forceinline__ void rot_all(const cx<float>* a, cx<float>* t, bool hi) { rot_all_impl<0>(a, t, hi); }

// one pencil: 16 rows per lane -> v[k'] = X[lane + 32 k']
__device__ __forceinline__ void pencil_fft(cx<float>* a, int lane, cx<float>* __restrict__ xb, const cx<float>* __restrict__ tw) {
    DFT<16, float>::run(a);                                           // C[q0] of lane j = 16 m0 + r
    const bool hi = lane >= 16;                                        // m0 = 1 (and, afterwards, q1 = 1)
    // radix 2 across the lane pair (j, j ^ 16): the m0 = 1 lane applies w32^q0 to its own value first
    {
        cx<float> t[16];
        rot_all(a, t, hi);
#pragma unroll
        for (int q0 = 0; q0 < 16; ++q0) {
            cx<float> o;
            o.x = __shfl_xor_sync(0xffffffffu, t[q0].x, 16);
            o.y = __shfl_xor_sync(0xffffffffu, t[q0].y, 16);
            a[q0] = hi ? csub(o, t[q0]) : cadd(t[q0], o);              // lane 16 q1 + r now holds B[r][q0 + 16 q1]
        }
    }
    // transposition: X[q][r], q = q0 + 16 q1 (row), r (column)
    const int r = lane & 15, q1 = lane >> 4;
    __syncwarp();
#pragma unroll
    for (int q0 = 0; q0 < 16; ++q0) xb[(q1 * 16 + q0) * XPITCH + r] = a[q0];
    __syncwarp();
#pragma unroll
    for (int rr = 0; rr < 16; ++rr) a[rr] = xb[lane * XPITCH + rr];   // lane q: B[rr][q]
#pragma unroll
    for (int rr = 1; rr < 16; ++rr) a[rr] = cmul(a[rr], ldg(tw + rr * 32 + lane));
    DFT<16, float>::run(a);                                           // a[k'] = X[lane + 32 k']
}

template <int NF, class Hist>
__device__ __forceinline__ void run_tiles(const CUtensorMap* map0, const CUtensorMap* map1, const Params& p, unsigned char* smem_raw,
                                          Hist& hist) {
    constexpr int NB = Bufs<NF>::N;
    constexpr int BAR_OFF = NB * NF * TILE_BYTES;
    constexpr int XCH_OFF = BAR_OFF + 128;
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw + BAR_OFF);     // bar[b]: tile buffer b is full
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    cx<float>* xb = reinterpret_cast<cx<float>*>(smem_raw + XCH_OFF + warp * XCH_BYTES);

    auto issue = [&](long long tile, int b) {
        const unsigned chunk = (unsigned)tile / (unsigned)p.ny;
        const int iy_loc = (int)((unsigned)tile - chunk * (unsigned)p.ny);
        mbar_expect_tx(bar + b, NF * TILE_BYTES);
#pragma unroll
        for (int f = 0; f < NF; ++f)
            for (int pr = 0; pr < 4; ++pr)
                for (int x = 0; x < NX; x += p.box_rows)
                    tma_load_5d(smem_raw + (b * NF + f) * TILE_BYTES + pr * SUB_BYTES + (size_t)x * 16, f == 0 ? map0 : map1, bar + b,
                                4 * pr, iy_loc, x & ((1 << p.xs_log2) - 1), (int)chunk, x >> p.xs_log2);
    };
    constexpr int kBatch = 8;
    auto tile_at = [&](long long i) { return ((i / kBatch) * gridDim.x + blockIdx.x) * kBatch + (i % kBatch); };
    long long it = 0;
    long long tile = tile_at(0);
    if (tid == 0) {
        if (tile < p.ntiles) issue(tile, 0);
        if (NB == 2 && tile_at(1) < p.ntiles) issue(tile_at(1), 1);
    }
    unsigned parity = 0;                                   // bit b: phase of bar[b] to wait for
    for (; tile < p.ntiles; tile = tile_at(++it)) {
        const int b = NB == 2 ? (int)(it & 1) : 0;
        const unsigned chunk = (unsigned)tile / (unsigned)p.ny;
        const int iy = (int)((unsigned)tile - chunk * (unsigned)p.ny) + p.iy0;
        const int iz = (int)chunk * WZ + warp;
        unsigned ent = kSegSentinel, ent1 = kSegSentinel, ent2 = kSegSentinel, ent3 = kSegSentinel;
        if (lane < p.seg_stride && iz < p.nzh) {
            const unsigned* e = p.seg + ((size_t)iy * p.nzh + iz) * p.seg_stride + lane;
            const size_t ts = (size_t)p.seg_pencils * p.seg_stride;
            ent = ldg(e);
            if (p.seg_ntab == 4) { ent1 = ldg(e + ts); ent2 = ldg(e + 2 * ts); ent3 = ldg(e + 3 * ts); }
        }
        cx<float> a[NF][NV];
        mbar_wait(bar + b, (parity >> b) & 1u);
        parity ^= 1u << b;
#pragma unroll
        for (int f = 0; f < NF; ++f) {
            const cx<float>* sub = reinterpret_cast<const cx<float>*>(smem_raw + (b * NF + f) * TILE_BYTES + (warp >> 1) * SUB_BYTES) + (warp & 1);
#pragma unroll
            for (int k = 0; k < NV; ++k) a[f][k] = sub[(lane + 32 * k) * 2];
        }
        // buffer b may be overwritten once every warp has its rows: the issuing warp waits for that, the others go on
        // (with two buffers a warp can be one tile ahead of the slowest one -- the tile after that is only requested once
        // everybody has read this one -- so consecutive tiles use different named barriers)
        fence_proxy_async();
        if (warp == 0) {
            named_sync(1 + b, THREADS);
            const long long next = tile_at(it + NB);
            if (lane == 0 && next < p.ntiles) issue(next, b);
        } else {
            named_arrive(1 + b, THREADS);
        }
        if (iz >= p.nzh) continue;
        float pw[NV];
        pencil_fft(a[0], lane, xb, p.tw);
        if constexpr (NF == 2) pencil_fft(a[1], lane, xb, p.tw);
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const cx<float> u = a[0][k], w = NF == 2 ? a[1][k] : a[0][k];
            pw[k] = u.x * w.x + u.y * w.y;
        }
        if (iy == 0 && iz == 0) {
            // DC of (x - c) plus c*N: power 0 in the flow (counted like any mode), true power added here with count 0
            const int b0 = (int)(__shfl_sync(0xffffffffu, ent, 0) & 0xFFFFu) - 1;
            if (lane == 0) {
                const double c0 = p.offset[0] ? (double)ldg(p.offset[0]) : 0.0;
                const double c1 = (NF == 2 && p.offset[1]) ? (double)ldg(p.offset[1]) : c0;
                const cx<float> u = a[0][0], w = NF == 2 ? a[1][0] : a[0][0];
                const double dc = ((double)u.x + c0 * p.ntot) * ((double)w.x + c1 * p.ntot) + (double)u.y * (double)w.y;
                pw[0] = 0.f;
                if (b0 >= 0) hist.add(b0, dc, 0u);
            }
        }
        const bool has_twin = (iz != 0) && (iz != p.nz - iz);
        const unsigned wt = has_twin ? 2u : 1u;
        __syncwarp();
        float* strip = reinterpret_cast<float*>(xb);
        if (p.seg_ntab == 1) {
            walk16<true, true>(pw, lane, ent, (float)wt, wt, strip, hist);
        } else {
            walk16<true, false>(pw, lane, ent, 1.f, 1u, strip, hist);
            walk16<false, true>(pw, lane, ent1, 1.f, 1u, strip, hist);
            walk16<true, false>(pw, lane, ent2, 1.f, 1u, strip, hist);
            walk16<false, true>(pw, lane, ent3, 1.f, 1u, strip, hist);
        }
    }
}

#endif  // T21_DEVICE_CODE

template <int NF>
__global__ void __launch_bounds__(THREADS, 2)
xpass16_kernel(const __grid_constant__ CUtensorMap map0, const __grid_constant__ CUtensorMap map1, const __grid_constant__ Params p, int nbins,
               double* __restrict__ part_sum, unsigned* __restrict__ part_cnt) {
#if T21_DEVICE_CODE
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    constexpr int NB = Bufs<NF>::N;
    constexpr int BAR_OFF = NB * NF * TILE_BYTES;
    constexpr int HIST_OFF = BAR_OFF + 128 + (THREADS / 32) * XCH_BYTES;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int b = 0; b < NB; ++b) mbar_init(smem_raw + BAR_OFF + 8 * b, 1);
        mbar_init_fence();
    }
    constexpr int NW = THREADS / 32;
    double* hs = reinterpret_cast<double*>(smem_raw + HIST_OFF);
    unsigned* hc = reinterpret_cast<unsigned*>(hs + (size_t)NW * nbins);
    for (int b = tid; b < NW * nbins; b += THREADS) { hs[b] = 0.0; hc[b] = 0u; }
    __syncthreads();
    WarpPrivHist hist{hs + (size_t)(tid >> 5) * nbins, hc + (size_t)(tid >> 5) * nbins};
    run_tiles<NF>(&map0, &map1, p, smem_raw, hist);
    __syncthreads();
    for (int b = tid; b < nbins; b += THREADS) {
        double s = 0.0;
        unsigned c = 0;
        for (int w = 0; w < NW; ++w) { s += hs[(size_t)w * nbins + b]; c += hc[(size_t)w * nbins + b]; }
        part_sum[(size_t)blockIdx.x * nbins + b] = s;
        part_cnt[(size_t)blockIdx.x * nbins + b] = c;
    }
#endif
}

template <int NF>
static int launch(XArgs& a, cudaStream_t st) {
    Params p;
    memset(&p, 0, sizeof(p));
    p.ny = a.ny; p.nz = a.nz; p.nzh = a.nzh; p.iy0 = a.iy0;
    p.nchunks = (a.nzh + WZ - 1) / WZ;
    p.ntiles = (long long)a.ny * p.nchunks;
    p.xs_log2 = a.xs_log2 >= 0 ? a.xs_log2 : 9;
    const int nxs = 1 << p.xs_log2;
    p.box_rows = nxs < 256 ? nxs : 256;
    p.tw = (const cx<float>*)a.tw32;
    p.offset[0] = (const float*)a.offset[0];
    p.offset[1] = (const float*)a.offset[1];
    p.ntot = a.ntot;
    p.seg = a.bins->seg_tab;
    p.seg_stride = a.bins->seg_stride;
    p.seg_ntab = a.bins->seg_ntab;
    p.seg_pencils = (long long)a.bins->n[1] * a.nzh;
    CUtensorMap map0, map1;
    if (int rc = make_wmap_f32(&map0, a.w[0], nxs, a.ny, a.pitch / WZ, NX / nxs, p.box_rows, 4)) return rc;
    if (int rc = make_wmap_f32(&map1, NF == 2 ? a.w[1] : a.w[0], nxs, a.ny, a.pitch / WZ, NX / nxs, p.box_rows, 4)) return rc;
    const size_t smem = (size_t)Bufs<NF>::N * NF * TILE_BYTES + 128 + (size_t)(THREADS / 32) * XCH_BYTES + (size_t)(THREADS / 32) * a.nbins * 12;
    auto kern = xpass16_kernel<NF>;
    if (int rc = ensure_smem(kern, smem)) return rc;
    int occ = 0;
    T21_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
    if (occ < 1) { set_error("x pass (nx = 512, warp per pencil) does not fit on an SM (smem %zu)", smem); return T21_EINVAL; }
    long long g = (long long)occ * device_sm_count();
    if (g > p.ntiles) g = p.ntiles;
    if (g > a.max_grid) g = a.max_grid;
    kern<<<(int)g, THREADS, smem, st>>>(map0, map1, p, a.nbins, a.part_sum, a.part_cnt);
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    a.grid_used = (int)g;
    return 0;
}

}  // namespace x16

bool xpass16_applicable(int dtype, const XArgs& a) {
    static const bool on = [] { const char* e = getenv("T21_PENCIL16"); return !(e && e[0] == '0'); }();
    if (!(on && dtype == T21_F32 && a.nx == x16::NX && (a.nf == 1 || a.nf == 2) && a.bins != nullptr && a.tw32 != nullptr &&
          a.pitch % x16::WZ == 0 && !a.hist_global))
        return false;
    const t21_binspec& B = *a.bins;
    const bool is_signed = B.nmu > 0 && !B.absolute_mus;
    return B.seg_tab != nullptr && B.seg_ntab == (is_signed ? 4 : 1) && B.seg_overflow == 0 && a.nbins <= 128;
}

int launch_xpass16_bin(XArgs& a, cudaStream_t st) {
    return a.nf == 2 ? x16::launch<2>(a, st) : x16::launch<1>(a, st);
}

}  // namespace t21
