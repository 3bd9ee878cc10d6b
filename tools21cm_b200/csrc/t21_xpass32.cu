// X pass, warp-per-pencil form (fp32, nx = 1024): TMA-landed tiles, radix-32 x radix-32 in registers with packed
// f32x2 butterflies, binning straight from the registers.
//
// Why (profiles/r01_ncu_final_xf_summary.txt): the 512-thread X pass of round 1 spent 33 % of its stall samples at
// CTA-wide barriers (five per tile, 16 values per thread, three stages) and had nothing in flight while a tile's
// global loads landed.  Here
//   * a tile (all 1024 x of the 8 columns of one 64-byte piece, one ky) is fetched by TMA (cp.async.bulk.tensor, four
//     boxes of 256 rows x 64 B) into shared memory under an mbarrier, issued as soon as the previous tile's data has
//     left the buffer -- the copy overlaps the second half of the previous tile's work and no register holds it;
//   * a thread holds 32 values: 1024 = 32 * 32 needs ONE exchange.  Stage 0 runs with lanes across the 8 columns
//     (conflict-free reads of the landed [x][8] tile) and stores transposed, pencil-major and padded
//     (i + i/32, pitch = 2 mod 16: conflict-free both ways); stage 1 then has one WARP per pencil, so nothing after
//     the exchange needs a CTA barrier;
//   * after stage 1 lane l holds kx = l + 32 k (k = 0..31) of its pencil: a step covers 32 consecutive kx at one
//     (ky, kz), so `c_yz = ky^2 + kz^2` is warp-uniform and the interval tests of the bin cache are made on
//     warp-uniform values for a whole pencil, or for the two steps that cover |kx| in [32k, 32k+32] -- no power
//     tile, no re-deal, no barrier;
//   * CTAs are 256 threads (8 pencils), two per SM; three 256-thread barriers per tile.
// Mode counts are decided exactly as in XPass (t21_passes.cuh): guard intervals first, classify() otherwise.
#include <string.h>
#include "t21_launch.cuh"
#include "t21_tma.cuh"

#ifndef T21_X32_BATCH
#define T21_X32_BATCH 8      // tiles per deal; measured at 1024^3: X 1.074 / 1.072 / 1.090 ms for 4 / 8 / 16
#endif

namespace t21 {

namespace x32 {

constexpr int NX = 1024;
constexpr int WZ = 8;                               // columns per tile = one 64-byte piece
constexpr int THREADS = 256;
constexpr int PITCH = NX + NX / 32 + 2;             // padded pencil, = 2 (mod 16): see the header comment
constexpr int BUF_BYTES = WZ * PITCH * 8;           // 67 712 >= the landed tile (65 536)
constexpr int BAR_OFF = (BUF_BYTES + 127) / 128 * 128;
constexpr int STAGE_OFF = BAR_OFF + 128;           // per-warp strips of 32 x 32 powers (shells epilogue)
constexpr int HIST_OFF = STAGE_OFF + (THREADS / 32) * 32 * 32 * 4;
constexpr int kMoments = 5;                         // multipole sums per k bin: P, P L2, P L4, L2^2, L4^2
using P = Plan<NX, 32>;
static_assert(P::NS == 2 && P::T == 32 && P::E == 32, "one warp per pencil");

struct Params {
    int ny, nz, nzh, nchunks;
    long long ntiles;
    int iy0;
    int box_rows;          // x planes per TMA box
    int xs_log2;           // log2(x planes per slab)
    const cx<float>* tw;   // Plan<1024, 32> stage twiddles
    const float* offset;   // mean offset (device scalar) or null
    double ntot;
    t21_binspec bins;
    const unsigned* seg;   // pencil segment table(s) (t21_build_pencil_segments) or null
    int seg_stride;
    long long seg_pencils; // pencils per table: ny_global * nzh
    int mp_los;            // multipoles: line-of-sight axis
    int seg_overflow;      // pencils of the table marked kSegOverflow (classified mode by mode)
    void* nd_out;          // NT = 6: out_half[NX][ny][nd_pitch], the scaled power of the computed half spectrum, FFT order
    int nd_f64, nd_pitch;
    double scale;
};

#if T21_DEVICE_CODE

// ---- shells (MU = 0): exact bin boundaries along the pencil -----------------------------------------------
// Along a pencil s = (kx^2 + ky^2) + kz^2 only grows with |kx| (rounding is monotonic), so bin b is the index range
// m_b <= |kx| < m_{b+1} with m_e = #{f : s(f) < thr[e]}.  Lane e finds m_e for edge e by bisection with the
// reference's own float64 sequence (10 probes of tab_s[0]); no fp32 estimate, no guard band, no look-up table.
// The pencil is then walked in steps of 64 modes (|kx| in [32k, 32k + 32]): steps no boundary falls into add
// their powers to an fp32 run; the others (a bit mask built with one warp reduction) give every lane its bin
// from the boundaries inside the step.  Sums leave the lanes through the two-entry warp cache (WarpBinCache).
template <class Hist>
__device__ __forceinline__ void wbc_add_uniform(WarpBinCache& a, Hist& h, int bin, double v, unsigned w) {
    if (bin == a.bin[0]) { a.sum[0] += v; a.cnt[0] += w; a.lru = 1; }
    else if (bin == a.bin[1]) { a.sum[1] += v; a.cnt[1] += w; a.lru = 0; }
    else if (a.lru == 0) {
        if (a.bin[0] >= 0) wbc_evict(a, h, 0);
        a.bin[0] = bin; a.sum[0] = v; a.cnt[0] = w; a.lru = 1;
    } else {
        if (a.bin[1] >= 0) wbc_evict(a, h, 1);
        a.bin[1] = bin; a.sum[1] = v; a.cnt[1] = w; a.lru = 0;
    }
}

// exact s of |kx| = f on this pencil: the reference's float64 sequence (kx^2 + ky^2) + kz^2 (power_spectrum.py:477)
__device__ __forceinline__ double s_exact(const double* __restrict__ t0, int f, double t1, double t2) {
    return dadd_rn(dadd_rn(ldg(t0 + f), t1), t2);
}

// m = min{ f in [0, NX/2] : s(f) >= thr }, NX/2 + 1 if there is none.  A closed-form fp32 guess (the k grid is
// uniform: kx^2 = f^2 * d) is verified -- and stepped, if it is off -- with the exact float64 values on both sides
// of the boundary; whatever has not settled after a few steps is bisected.  Every lane searches its own edge.
__device__ __forceinline__ int boundary_of(const double* __restrict__ t0, double t1, double t2, double thr, float inv_d) {
    const double x = thr - (t1 + t2);
    int g = 0;
    if (x > 0.0) {
        const float q = sqrtf((float)x * inv_d);
        g = q < (float)(NX / 2 + 1) ? (int)ceilf(q) : NX / 2 + 1;     // (NaN compares false: NX/2 + 1)
    }
    bool done = false;
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
        if (!done) {
            const bool hi_ok = g > NX / 2 || s_exact(t0, g, t1, t2) >= thr;
            const bool lo_ok = g == 0 || s_exact(t0, g - 1, t1, t2) < thr;
            if (!hi_ok) ++g; else if (!lo_ok) --g; else done = true;
        }
        if (__all_sync(0xffffffffu, done)) return g;
    }
    if (!done) {
        int lo = 0, hi = NX / 2 + 1;
#pragma unroll 1
        for (int it = 0; it < 10; ++it) {
            const int mid = (lo + hi) >> 1;
            const bool ge = s_exact(t0, mid, t1, t2) >= thr;
            if (lo < hi) { if (ge) hi = mid; else lo = mid + 1; }
        }
        g = lo;
    }
    return g;
}

template <class Hist>
__device__ __forceinline__ void epilogue_shells(const Params& p, const cx<float>* v, int lane, int iy, int iz, float* stage,
                                                WarpBinCache& c, Hist& hist) {
    const t21_binspec& B = p.bins;
    const int nk = B.nk;
    const bool has_twin = (iz != 0) && (iz != p.nz - iz);
    const unsigned wt = has_twin ? 2u : 1u;
    const float wf = (float)wt;
    constexpr unsigned kNone = 0xffffffffu;

    // boundaries: lane e <= nk owns edge e
    const float inv_d = 1.0f / (float)ldg(B.tab_s[0] + 1);               // 1 / (kx spacing)^2, for the boundary guess
    const int mm = boundary_of(B.tab_s[0], ldg(B.tab_s[1] + iy), ldg(B.tab_s[2] + iz), ldg(B.thr + (lane <= nk ? lane : nk)), inv_d);
    const int m = lane <= nk ? mm : (1 << 20);
    const int n0 = __popc(__ballot_sync(0xffffffffu, m <= 0));           // edges at or below |kx| = 0: bin = n0 - 1
    // distinct boundary positions inside the pencil, in order (warp-uniform)
    const unsigned inb = (m >= 1 && m <= NX / 2) ? (unsigned)m : kNone;
    const unsigned M1 = __reduce_min_sync(0xffffffffu, inb);
    const unsigned M2 = __reduce_min_sync(0xffffffffu, (M1 != kNone && inb > M1) ? inb : kNone);
    const unsigned M3 = __reduce_min_sync(0xffffffffu, (M2 != kNone && inb > M2) ? inb : kNone);

    float pw[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) pw[k] = v[k].x * v[k].x + v[k].y * v[k].y;
    if (iy == 0 && iz == 0 && lane == 0) {
        // DC of (x - c) plus c*N: the only sample the mean offset changes.  It stays in the flow with power 0 (so
        // that it is counted like any other mode); its true power goes to the histogram here, with count 0.
        const double c0 = p.offset ? (double)ldg(p.offset) : 0.0;
        const double ra = (double)v[0].x + c0 * p.ntot;
        const double dc = ra * ra + (double)v[0].y * (double)v[0].y;
        pw[0] = 0.f;
        if (n0 >= 1 && n0 <= nk) hist.add(n0 - 1, dc, 0u);
    }

    if (M3 == kNone) {
        // At most two boundaries cross the pencil (nearly always): three segments A | B | C of |kx|, summed in lane
        // registers.  A step of 64 modes (|kx| in [32k, 32k + 32]) no boundary falls into goes to one segment as a
        // whole (warp-uniform choice); in the one or two steps that hold a boundary every lane compares its own |kx|.
        const int K1 = M1 != kNone ? (int)((M1 - 1u) >> 5) : 99, K2 = M2 != kNone ? (int)((M2 - 1u) >> 5) : 99;
        float sA = 0.f, sB = 0.f, sC = 0.f;
        unsigned cA = 0u, cB = 0u, cC = 0u;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (k != K1 && k != K2) {
                const float pp = pw[k] + pw[31 - k];
                if (k < K1) { sA += pp; cA += 2u; }
                else if (k < K2) { sB += pp; cB += 2u; }
                else { sC += pp; cC += 2u; }
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const unsigned f = h ? 32u * k + 32u - lane : 32u * k + lane;      // |kx| of this lane's mode
                    const float pv = h ? pw[31 - k] : pw[k];
                    if (f < M1) { sA += pv; cA += 1u; }
                    else if (f < M2) { sB += pv; cB += 1u; }
                    else { sC += pv; cC += 1u; }
                }
            }
        }
        if (n0 >= 1 && n0 <= nk) wbc_add_uniform(c, hist, n0 - 1, (double)(sA * wf), cA * wt);
        if (M1 != kNone) {
            const int n1 = __popc(__ballot_sync(0xffffffffu, m <= (int)M1));
            if (n1 >= 1 && n1 <= nk) wbc_add_uniform(c, hist, n1 - 1, (double)(sB * wf), cB * wt);
        }
        if (M2 != kNone) {
            const int n2 = __popc(__ballot_sync(0xffffffffu, m <= (int)M2));
            if (n2 >= 1 && n2 <= nk) wbc_add_uniform(c, hist, n2 - 1, (double)(sC * wf), cC * wt);
        }
        return;
    }

    // Three or more boundaries (pencils close to the kx axis): the powers go through a warp-private strip of shared
    // memory so that the walk is a rolled loop (an unrolled walk with its mixed-step branches was 50 000
    // instructions and starved the instruction cache).
    const unsigned mixed = __reduce_or_sync(0xffffffffu, inb != kNone ? 1u << ((inb - 1u) >> 5) : 0u);
#pragma unroll
    for (int q = 0; q < 32; ++q) stage[q * 32 + lane] = pw[q];
    __syncwarp();
    int nle = n0;
    float run = 0.f;
    unsigned runc = 0u;
#pragma unroll 1
    for (int k = 0; k <= 16; ++k) {
        if (k < 16 && !((mixed >> k) & 1u)) {
            run += stage[k * 32 + lane] + stage[(31 - k) * 32 + lane];
            runc += 2u;
            continue;
        }
        if (runc && nle >= 1 && nle <= nk) wbc_add_uniform(c, hist, nle - 1, (double)(run * wf), runc * wt);
        run = 0.f; runc = 0u;
        if (k == 16) break;
        const int nhi = __popc(__ballot_sync(0xffffffffu, m <= 32 * k + 32));
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {
            const int kk = h ? 31 - k : k;
            const int f = h ? 32 * k + 32 - lane : 32 * k + lane;
            int ne = nle;
            for (int e = nle; e < nhi; ++e) ne += f >= __shfl_sync(0xffffffffu, m, e) ? 1 : 0;
            wbc_add(c, hist, (ne >= 1 && ne <= nk) ? ne - 1 : -1, (double)(stage[kk * 32 + lane] * wf), wt);
        }
        nle = nhi;
    }
    __syncwarp();
}

// ---- bins as index ranges from the pencil segment table (shells, |mu| and signed mu bins) ------------------------------
// Lane j holds entry j of the pencil's table: range j starts at |kx| = pos_j and falls in bin binp1_j - 1.  A step of
// 64 modes (|kx| in [32k, 32k + 32]) no range boundary falls into is added whole, as two packed FMAs per lane; a step
// that holds a boundary parks the running sum and its two powers in a lane-private strip of shared memory, and a
// rolled loop then closes the ranges in order: every lane adds its modes below the boundary, one shuffle tree sums the
// range over the warp, lane 0 adds it -- and the closed-form mode count of the range -- to the warp's histogram.
// LO / HI: take the kx >= 0 half (v[0..15], |kx| = 32k + lane) / the kx < 0 half (v[31-k], |kx| = 32k + 32 - lane).
// Signed mu walks the halves separately (mu changes sign with kx when the line of sight is x), once with the stored
// modes' own bins and once with their twins'.
template <bool LO, bool HI, class Hist>
__device__ __forceinline__ void walk_segments(const cx<float>* v, int lane, unsigned ent, float wf, unsigned wt,
                                              float* __restrict__ strip, Hist& hist) {
    const unsigned pos = ent >> 16, binp1 = ent & 0xFFFFu;
    const unsigned S = (unsigned)__popc(__ballot_sync(0xffffffffu, pos != 0xFFFFu));
    if (S == 1u && __shfl_sync(0xffffffffu, binp1, 0) == 0u) return;          // the whole pencil is in no bin
    const unsigned mixed = __reduce_or_sync(0xffffffffu, ((unsigned)lane >= 1u && (unsigned)lane < S) ? 1u << ((pos - 1u) >> 5) : 0u);
    float cur = 0.f;
    unsigned j = 0u;
    unsigned nextpos = S > 1u ? __shfl_sync(0xffffffffu, pos, 1) : 0xFFFFu;
    // close range j at boundary b (exclusive end): sum over the warp, lane 0 books it
    auto close = [&](unsigned b) {
        float t = cur;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        const unsigned a = __shfl_sync(0xffffffffu, pos, j);
        const int bin = (int)__shfl_sync(0xffffffffu, binp1, j) - 1;
        if (lane == 0 && bin >= 0) {
            const unsigned cnt = segment_modes(LO, HI, a, b, (unsigned)(NX / 2));   // kx = 0 and the Nyquist sample are single
            hist.add(bin, (double)(t * wf), cnt * wt);
        }
    };
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        cx<float> run2 = cmake<float>(0.f, 0.f);
        unsigned slot = 0u, ks = 0u;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            const int k = half * 8 + kk;
            if ((mixed >> k) & 1u) {
                float* d = strip + slot * 96u + lane;
                d[0] = run2.x + run2.y;
                d[32] = LO ? v[k].x * v[k].x + v[k].y * v[k].y : 0.f;
                d[64] = HI ? v[31 - k].x * v[31 - k].x + v[31 - k].y * v[31 - k].y : 0.f;
                run2 = cmake<float>(0.f, 0.f);
                ks |= (unsigned)k << (4u * slot);
                ++slot;
            } else {
                if (LO) run2 = fma2(v[k], v[k], run2);
                if (HI) run2 = fma2(v[31 - k], v[31 - k], run2);
            }
        }
        // (every lane reads back only what it wrote itself: no warp synchronisation needed)
#pragma unroll 1
        for (unsigned i = 0; i < slot; ++i) {
            const float* d = strip + i * 96u + lane;
            cur += d[0];
            float plo = d[32], phi = d[64];
            const unsigned k32 = ((ks >> (4u * i)) & 15u) << 5;
            const unsigned flo = k32 + (unsigned)lane, fhi = k32 + 32u - (unsigned)lane;
            while (nextpos <= k32 + 32u) {                  // warp-uniform: a boundary inside this step
                if (flo < nextpos) { cur += plo; plo = 0.f; }
                if (fhi < nextpos) { cur += phi; phi = 0.f; }
                close(nextpos);
                cur = 0.f;
                ++j;
                nextpos = j + 1u < S ? __shfl_sync(0xffffffffu, pos, (j + 1u) & 31u) : 0xFFFFu;
            }
            cur += plo + phi;
        }
        cur += run2.x + run2.y;
    }
    close((unsigned)(NX / 2 + 1));
}

// NT = 1: bins even in kx, one table; NT = 4: signed mu, tables {own kx>=0, own kx<0, twin kx>=0, twin kx<0}
// A pencil with more ranges than the table holds (a fine binning -- the reference's default is 100 bins -- on a pencil
// next to the kx axis): classify its modes one by one, as the CTA-per-tile kernel does, through the two-entry warp cache.
template <int NT, class Hist>
__device__ __noinline__ void epilogue_classify(const Params& p, float v0x, float v0y, int lane, int iy, int iz, float* __restrict__ strip,
                                               Hist& hist) {
    // (the caller has parked the 32 powers of every lane in the strip: the values stay in registers on its side)
    const bool has_twin = (iz != 0) && (iz != p.nz - iz);
    if (iy == 0 && iz == 0 && lane == 0) {
        const double c0 = p.offset ? (double)ldg(p.offset) : 0.0;
        const double ra = (double)v0x + c0 * p.ntot;
        const double dc = ra * ra + (double)v0y * (double)v0y;
        strip[0] = 0.f;                                    // counted in the flow, its power added here
        const int b0 = classify(p.bins, 0, 0, 0, false).own;
        if (b0 >= 0) hist.add(b0, dc, 0u);
    }
    WarpBinCache c;
    wbc_init(c);
#pragma unroll 1
    for (int q = 0; q < 32; ++q) {
        const ModeBins mb = classify(p.bins, lane + 32 * q, iy, iz, has_twin);
        const double pw = (double)strip[q * 32 + lane];
        wbc_add(c, hist, mb.own, pw * (double)mb.own_w, (unsigned)mb.own_w);
        if (NT == 4) wbc_add(c, hist, mb.twin, pw, 1u);
    }
    wbc_flush(c, hist);
}

template <int NT, bool OVF, class Hist>
__device__ __forceinline__ void epilogue_segments(const Params& p, cx<float>* v, int lane, int iy, int iz, const unsigned* ent,
                                                  float* __restrict__ strip, Hist& hist) {
    if constexpr (OVF) {                                  // (warp-uniform: entry 0 of table 0, or of any table for signed mu)
        bool ovf = false;
#pragma unroll
        for (int t = 0; t < NT; ++t) ovf = ovf || ent[t] == kSegOverflow;
        if (__any_sync(0xffffffffu, ovf)) {
#pragma unroll
            for (int q = 0; q < 32; ++q) strip[q * 32 + lane] = v[q].x * v[q].x + v[q].y * v[q].y;
            epilogue_classify<NT>(p, v[0].x, v[0].y, lane, iy, iz, strip, hist);
            return;
        }
    }
    if (iy == 0 && iz == 0) {
        // DC of (x - c) plus c*N: the only sample the mean offset changes.  It stays in the flow with power 0 (so
        // that it is counted like any other mode); its true power goes to the histogram here, with count 0.
        const int b0 = (int)(__shfl_sync(0xffffffffu, ent[0], 0) & 0xFFFFu) - 1;
        if (lane == 0) {
            const double c0 = p.offset ? (double)ldg(p.offset) : 0.0;
            const double ra = (double)v[0].x + c0 * p.ntot;
            const double dc = ra * ra + (double)v[0].y * (double)v[0].y;
            v[0] = cmake<float>(0.f, 0.f);
            if (b0 >= 0) hist.add(b0, dc, 0u);
        }
    }
    if constexpr (NT == 1) {
        const bool has_twin = (iz != 0) && (iz != p.nz - iz);
        const unsigned wt = has_twin ? 2u : 1u;
        walk_segments<true, true>(v, lane, ent[0], (float)wt, wt, strip, hist);
    } else {
        walk_segments<true, false>(v, lane, ent[0], 1.f, 1u, strip, hist);
        walk_segments<false, true>(v, lane, ent[1], 1.f, 1u, strip, hist);
        walk_segments<true, false>(v, lane, ent[2], 1.f, 1u, strip, hist);
        walk_segments<false, true>(v, lane, ent[3], 1.f, 1u, strip, hist);
    }
}

// ---- multipole moments per k bin (power_legendre.py:55-87) on the same ranges ----------------------------------------
// Five sums per range: P, P L2, P L4, L2^2, L4^2 with L2 = (3 mu^2 - 1)/2, L4 = 4.375 (mu^2 - 0.115587)(mu^2 - 0.741556)
// as the reference writes them (:57-60), mu^2 = k_los^2 / k^2 from fp32 (a weight, not a decision: the bin comes from
// the table).  k^2 = f^2 d + (ky^2 + kz^2) with f = |kx index| (exact up to 2^24) and d the squared kx spacing.  Modes
// with k < 0.001 have mu = NaN in the reference (:501) and here.  The steps of a quarter pencil that hold a range
// boundary park the running sums and their two powers; a rolled loop closes the ranges in order (five shuffle trees).
struct MomentHist {
    double* sum;      // [nk][5], this warp's copy
    unsigned* cnt;    // [nk]
};

__device__ __forceinline__ void epilogue_multipoles(const Params& p, cx<float>* v, int lane, int iy, int iz, unsigned ent,
                                                    float* __restrict__ strip, MomentHist& hist) {
    const t21_binspec& B = p.bins;
    const bool has_twin = (iz != 0) && (iz != p.nz - iz);
    const unsigned wt = has_twin ? 2u : 1u;
    const float wf = (float)wt;
    const unsigned pos = ent >> 16, binp1 = ent & 0xFFFFu;
    const unsigned S = (unsigned)__popc(__ballot_sync(0xffffffffu, pos != 0xFFFFu));
    if (S == 1u && __shfl_sync(0xffffffffu, binp1, 0) == 0u) return;          // the whole pencil is in no bin
    const unsigned mixed = __reduce_or_sync(0xffffffffu, ((unsigned)lane >= 1u && (unsigned)lane < S) ? 1u << ((pos - 1u) >> 5) : 0u);

    const float d2 = ldg(B.tab_sf[0] + 1);                                   // squared kx spacing
    const float sy = ldg(B.tab_sf[1] + iy), sz = ldg(B.tab_sf[2] + iz);
    const float cyz = sy + sz;
    const float klc = p.mp_los == 1 ? sy : sz;                               // k_los^2 when the line of sight is not x
    const bool los_x = p.mp_los == 0;
    const float snan = (float)B.s_nan;

    if (iy == 0 && iz == 0) {
        // DC of (x - c) plus c*N (see epilogue_segments): power 0 in the flow (its mu is NaN there, like the
        // reference's), true power straight to the P sum of its bin
        const int b0 = (int)__shfl_sync(0xffffffffu, binp1, 0) - 1;
        if (lane == 0) {
            const double c0 = p.offset ? (double)ldg(p.offset) : 0.0;
            const double ra = (double)v[0].x + c0 * p.ntot;
            const double dc = ra * ra + (double)v[0].y * (double)v[0].y;
            v[0] = cmake<float>(0.f, 0.f);
            if (b0 >= 0) hist.sum[b0 * kMoments] += dc;
        }
    }

    // contribution of the mode |kx index| = f with power pw to five running sums
    auto add_mode = [&](unsigned f, float pw, float* a) {
        const float sx = (float)(f * f) * d2;
        const float s = sx + cyz;
        float mu2 = __fdividef(los_x ? sx : klc, s);
        mu2 = s >= snan ? mu2 : __int_as_float(0x7fc00000);
        const float l2 = fmaf(1.5f, mu2, -0.5f);
        const float l4 = 4.375f * (mu2 - 0.115587f) * (mu2 - 0.741556f);
        a[0] += pw; a[1] = fmaf(pw, l2, a[1]); a[2] = fmaf(pw, l4, a[2]); a[3] = fmaf(l2, l2, a[3]); a[4] = fmaf(l4, l4, a[4]);
    };

    float cur[kMoments] = {0.f, 0.f, 0.f, 0.f, 0.f};
    unsigned j = 0u;
    unsigned nextpos = S > 1u ? __shfl_sync(0xffffffffu, pos, 1) : 0xFFFFu;
    auto close = [&](unsigned b) {
        float t[kMoments];
#pragma unroll
        for (int q = 0; q < kMoments; ++q) {
            t[q] = cur[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t[q] += __shfl_xor_sync(0xffffffffu, t[q], o);
            cur[q] = 0.f;
        }
        const unsigned a = __shfl_sync(0xffffffffu, pos, j);
        const int bin = (int)__shfl_sync(0xffffffffu, binp1, j) - 1;
        if (lane == 0 && bin >= 0) {
#pragma unroll
            for (int q = 0; q < kMoments; ++q) hist.sum[bin * kMoments + q] += (double)(t[q] * wf);
            hist.cnt[bin] += segment_modes(true, true, a, b, (unsigned)(NX / 2)) * wt;
        }
    };
#pragma unroll
    for (int quarter = 0; quarter < 4; ++quarter) {
        float run[kMoments] = {0.f, 0.f, 0.f, 0.f, 0.f};
        unsigned slot = 0u, ks = 0u;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            const int k = quarter * 4 + kk;
            const float plo = v[k].x * v[k].x + v[k].y * v[k].y;
            const float phi = v[31 - k].x * v[31 - k].x + v[31 - k].y * v[31 - k].y;
            if ((mixed >> k) & 1u) {
                float* d = strip + slot * 224u + lane;                       // 7 values per slot and lane
#pragma unroll
                for (int q = 0; q < kMoments; ++q) { d[32 * q] = run[q]; run[q] = 0.f; }
                d[160] = plo;
                d[192] = phi;
                ks |= (unsigned)k << (4u * slot);
                ++slot;
            } else {
                add_mode(32u * k + (unsigned)lane, plo, run);
                add_mode(32u * k + 32u - (unsigned)lane, phi, run);
            }
        }
#pragma unroll 1
        for (unsigned i = 0; i < slot; ++i) {
            const float* d = strip + i * 224u + lane;
#pragma unroll
            for (int q = 0; q < kMoments; ++q) cur[q] += d[32 * q];
            const float plo = d[160], phi = d[192];
            const unsigned k32 = ((ks >> (4u * i)) & 15u) << 5;
            const unsigned flo = k32 + (unsigned)lane, fhi = k32 + 32u - (unsigned)lane;
            bool lo_open = true, hi_open = true;
            while (nextpos <= k32 + 32u) {                  // warp-uniform: a boundary inside this step
                if (lo_open && flo < nextpos) { add_mode(flo, plo, cur); lo_open = false; }
                if (hi_open && fhi < nextpos) { add_mode(fhi, phi, cur); hi_open = false; }
                close(nextpos);
                ++j;
                nextpos = j + 1u < S ? __shfl_sync(0xffffffffu, pos, (j + 1u) & 31u) : 0xFFFFu;
            }
            if (lo_open) add_mode(flo, plo, cur);
            if (hi_open) add_mode(fhi, phi, cur);
        }
#pragma unroll
        for (int q = 0; q < kMoments; ++q) cur[q] += run[q];
    }
    close((unsigned)(NX / 2 + 1));
}

template <int NT, bool OVF, class Hist>
__device__ __forceinline__ void run_tiles(const CUtensorMap* map, const Params& p, unsigned char* smem_raw, Hist& hist, int nbins) {
    cx<float>* buf = reinterpret_cast<cx<float>*>(smem_raw);
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw + BAR_OFF);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int zl0 = tid & (WZ - 1), j0 = tid >> 3;          // stage 0: lanes across the columns

    WarpBinCache wbc;
    if constexpr (NT == 0) wbc_init(wbc);

    auto issue = [&](long long tile) {
        const unsigned chunk = (unsigned)tile / (unsigned)p.ny;
        const int iy_loc = (int)((unsigned)tile - chunk * (unsigned)p.ny);
        mbar_expect_tx(bar, NX * WZ * 8);
        for (int x = 0; x < NX; x += p.box_rows)
            tma_load_5d(smem_raw + (size_t)x * (WZ * 8), map, bar, 0, iy_loc, x & ((1 << p.xs_log2) - 1), (int)chunk, x >> p.xs_log2);
    };

    // Tiles are dealt to the CTAs in batches of kBatch consecutive tiles (ky runs fastest): consecutive pencils of a
    // warp then cross nearly the same bins, so the warp's cached bins survive from one pencil to the next, while the
    // expensive tiles (pencils near the kx axis cross many bin edges) are spread over all CTAs.  (One contiguous
    // range per CTA left the SMs idle 41 % of the time: profiles/r02_x32_imbalance.txt.)
    constexpr int kBatch = T21_X32_BATCH;
    auto tile_at = [&](long long i) { return ((i / kBatch) * gridDim.x + blockIdx.x) * kBatch + (i % kBatch); };
    long long it = 0;
    long long tile = tile_at(0);
    if (tid == 0 && tile < p.ntiles) issue(tile);
    unsigned parity = 0;
    for (; tile < p.ntiles; tile = tile_at(++it)) {
        const unsigned chunk = (unsigned)tile / (unsigned)p.ny;
        const int iy = (int)((unsigned)tile - chunk * (unsigned)p.ny) + p.iy0;
        cx<float> v[32];
        const int iz = (int)chunk * WZ + warp;
        // this pencil's ranges: asked for now, needed after the two butterfly stages
        constexpr int NTAB = NT == 4 ? 4 : 1;             // NT: 0 in-kernel shells, 1 one table, 4 signed mu, 5 multipoles, 6 n-D store
        unsigned ent[NTAB];
#pragma unroll
        for (int t = 0; t < NTAB; ++t) {
            ent[t] = kSegSentinel;
            if (NT > 0 && NT < 6 && lane < p.seg_stride && iz < p.nzh)
                ent[t] = ldg(p.seg + ((size_t)t * p.seg_pencils + (size_t)iy * p.nzh + iz) * p.seg_stride + lane);
        }
        mbar_wait(bar, parity);
        parity ^= 1u;
        // stage 0: rows j0 + 32 k of column zl0 from the landed [x][8] tile
        stage_load<P, 0>(v, j0, [&](int i) { return buf[i * WZ + zl0]; });
        DFT<32, float>::run(v);
        __syncthreads();                                   // every thread has its inputs: the tile may be overwritten
        {
            cx<float>* dst = buf + zl0 * PITCH + 33 * j0;  // outputs 32 j0 + k, padded
#pragma unroll
            for (int k = 0; k < 32; ++k) dst[k] = v[k];
        }
        __syncthreads();
        // stage 1: warp = pencil, lane = butterfly; inputs lane + 32 k
        {
            const cx<float>* src = buf + warp * PITCH + lane;
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = src[33 * k];
        }
        // The buffer is free once every warp has its inputs: only the issuing warp waits for that (bar.sync), the
        // others just say so (bar.arrive) and go on; the next tile lands under the rest of this one.
        fence_proxy_async();                               // (generic-proxy accesses above, async-proxy writes below)
        if (warp == 0) {
            named_sync(1, THREADS);
            const long long next = tile_at(it + 1);
            if (lane == 0 && next < p.ntiles) issue(next);
        } else {
            named_arrive(1, THREADS);
        }
#pragma unroll
        for (int k = 1; k < 32; ++k) v[k] = cmul(v[k], ldg(p.tw + (k - 1) * 32 + lane));
        DFT<32, float>::run(v);                            // v[k] = X[lane + 32 k]
        if constexpr (NT == 7) {
            // n-D outputs, pencil-major: out_t[ky][kz][kx] -- a warp writes its pencil as 32 runs of 128 bytes; the pass
            // that assembles the fftshifted cube transposes tiles of it through shared memory (t21_nd_assemble_t).
            // (Storing [kx][ky][kz] from here writes 32-byte pieces 2 MB apart: 2.97 ms per 1024^3 cube against 1.1.)
            if (iz < p.nzh) {
                const size_t base = ((size_t)(iy - p.iy0) * p.nzh + iz) * NX + lane;
                double dc = 0.0;
                const bool is_dc = iy == 0 && iz == 0 && lane == 0;
                if (is_dc) {
                    const double c0 = p.offset ? (double)ldg(p.offset) : 0.0;
                    const double ra = (double)v[0].x + c0 * p.ntot;
                    dc = ra * ra + (double)v[0].y * (double)v[0].y;
                }
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    double val = (double)(v[k].x * v[k].x + v[k].y * v[k].y) * p.scale;
                    if (k == 0 && is_dc) val = dc * p.scale;
                    if (p.nd_f64) ((double*)p.nd_out)[base + 32 * k] = val;
                    else ((float*)p.nd_out)[base + 32 * k] = (float)val;
                }
            }
        } else if constexpr (NT == 6) {
            // n-D outputs: powers go to a [kx][8 + 1] tile in shared memory (pitch 9: a warp's 32 kx fall on 32 banks),
            // then every thread writes whole 32-byte runs -- the 8 columns of this piece at one kx -- of the output
            float* st = reinterpret_cast<float*>(smem_raw + STAGE_OFF);
            double* dc_slot = reinterpret_cast<double*>(smem_raw + BAR_OFF + 64);
            if (iz < p.nzh) {
                if (iy == 0 && iz == 0 && lane == 0) {
                    const double c0 = p.offset ? (double)ldg(p.offset) : 0.0;
                    const double ra = (double)v[0].x + c0 * p.ntot;
                    *dc_slot = ra * ra + (double)v[0].y * (double)v[0].y;
                }
#pragma unroll
                for (int k = 0; k < 32; ++k) st[(lane + 32 * k) * 9 + warp] = v[k].x * v[k].x + v[k].y * v[k].y;
            }
            __syncthreads();
            const int ncols = p.nzh - (int)chunk * WZ < WZ ? p.nzh - (int)chunk * WZ : WZ;
            const int iy_loc = iy - p.iy0;
#pragma unroll 1
            for (int kx = tid; kx < NX; kx += THREADS) {
                const float* row = st + kx * 9;
                const size_t o = ((size_t)kx * p.ny + iy_loc) * p.nd_pitch + (size_t)chunk * WZ;
                double val[WZ];
#pragma unroll
                for (int c = 0; c < WZ; ++c) val[c] = (double)row[c] * p.scale;
                if (kx == 0 && iy == 0 && chunk == 0) val[0] = *dc_slot * p.scale;
                if (p.nd_f64) {
                    double* d = (double*)p.nd_out + o;
                    if (ncols == WZ) {
#pragma unroll
                        for (int c = 0; c < WZ; c += 2) *reinterpret_cast<double2*>(d + c) = make_double2(val[c], val[c + 1]);
                    } else {
                        for (int c = 0; c < ncols; ++c) d[c] = val[c];
                    }
                } else {
                    float* d = (float*)p.nd_out + o;
                    if (ncols == WZ) {
                        *reinterpret_cast<float4*>(d) = make_float4((float)val[0], (float)val[1], (float)val[2], (float)val[3]);
                        *reinterpret_cast<float4*>(d + 4) = make_float4((float)val[4], (float)val[5], (float)val[6], (float)val[7]);
                    } else {
                        for (int c = 0; c < ncols; ++c) d[c] = (float)val[c];
                    }
                }
            }
        } else if (iz < p.nzh) {
            float* stage = reinterpret_cast<float*>(smem_raw + STAGE_OFF) + warp * 1024;
            if constexpr (NT == 5) epilogue_multipoles(p, v, lane, iy, iz, ent[0], stage, hist);
            else if constexpr (NT > 0) epilogue_segments<NT, OVF>(p, v, lane, iy, iz, ent, stage, hist);
            else epilogue_shells(p, v, lane, iy, iz, stage, wbc, hist);
        }
    }
    if constexpr (NT == 0) wbc_flush(wbc, hist);
}

#endif  // T21_DEVICE_CODE

// The histogram (at most 31 bins) lives in shared memory, one copy per warp: only lane 0 of the owning warp writes
// it, in program order; warps are summed in order here and CTAs in order by reduce_partials_kernel, so the result
// is bit-reproducible run to run.
// OVF: the table carries overflow marks (pencils classified mode by mode); the common kernels are compiled without that path
template <int NT, bool OVF>
__global__ void __launch_bounds__(THREADS, 2)
xpass32_kernel(const __grid_constant__ CUtensorMap map, const __grid_constant__ Params p, int nbins,
               double* __restrict__ part_sum, unsigned* __restrict__ part_cnt) {
#if T21_DEVICE_CODE
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(smem_raw + BAR_OFF, 1);
        mbar_init_fence();
    }
    constexpr int NW = THREADS / 32;
    double* hs = reinterpret_cast<double*>(smem_raw + HIST_OFF);
    unsigned* hc = reinterpret_cast<unsigned*>(hs + (size_t)NW * nbins);
    for (int b = tid; b < NW * nbins; b += THREADS) { hs[b] = 0.0; hc[b] = 0u; }
    __syncthreads();
    WarpPrivHist hist{hs + (size_t)(tid >> 5) * nbins, hc + (size_t)(tid >> 5) * nbins};
    run_tiles<NT, OVF>(&map, p, smem_raw, hist, nbins);
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

// multipole moments: per-warp private [nk][5] sums + [nk] counts, same fixed summation order
__global__ void __launch_bounds__(THREADS, 2)
xpass32_moments_kernel(const __grid_constant__ CUtensorMap map, const __grid_constant__ Params p, int nk,
                       double* __restrict__ part_sum, unsigned* __restrict__ part_cnt) {
#if T21_DEVICE_CODE
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(smem_raw + BAR_OFF, 1);
        mbar_init_fence();
    }
    constexpr int NW = THREADS / 32;
    const int ns = nk * kMoments;
    double* hs = reinterpret_cast<double*>(smem_raw + HIST_OFF);
    unsigned* hc = reinterpret_cast<unsigned*>(hs + (size_t)NW * ns);
    for (int b = tid; b < NW * ns; b += THREADS) hs[b] = 0.0;
    for (int b = tid; b < NW * nk; b += THREADS) hc[b] = 0u;
    __syncthreads();
    MomentHist hist{hs + (size_t)(tid >> 5) * ns, hc + (size_t)(tid >> 5) * nk};
    run_tiles<5, false>(&map, p, smem_raw, hist, ns);
    __syncthreads();
    for (int b = tid; b < ns; b += THREADS) {
        double s = 0.0;
        for (int w = 0; w < NW; ++w) s += hs[(size_t)w * ns + b];
        part_sum[(size_t)blockIdx.x * ns + b] = s;
    }
    for (int b = tid; b < nk; b += THREADS) {
        unsigned c = 0;
        for (int w = 0; w < NW; ++w) c += hc[(size_t)w * nk + b];
        part_cnt[(size_t)blockIdx.x * nk + b] = c;
    }
#endif
}

// n-D outputs: no histogram, the power of the computed half spectrum is stored
struct NoHist32 {
    __device__ void add(int, double, unsigned) {}
};
constexpr int kStoreSmem = STAGE_OFF + NX * 9 * 4;
template <int NT>
__global__ void __launch_bounds__(THREADS, 2)
xpass32_store_kernel(const __grid_constant__ CUtensorMap map, const __grid_constant__ Params p) {
#if T21_DEVICE_CODE
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    if (threadIdx.x == 0) {
        mbar_init(smem_raw + BAR_OFF, 1);
        mbar_init_fence();
    }
    __syncthreads();
    NoHist32 hist;
    run_tiles<NT, false>(&map, p, smem_raw, hist, 0);
#endif
}

// T21_X32_SEGMENTS=0 keeps the in-kernel boundary search even when the spec carries a table (A/B runs); read once
static bool use_segments() {
    static const bool on = [] { const char* e = getenv("T21_X32_SEGMENTS"); return !(e && e[0] == '0'); }();
    return on;
}

static int launch(XArgs& a, cudaStream_t st, int moments_los = -1) {
    Params p;
    memset(&p, 0, sizeof(p));
    p.ny = a.ny; p.nz = a.nz; p.nzh = a.nzh; p.iy0 = a.iy0;
    p.nchunks = (a.nzh + WZ - 1) / WZ;
    p.ntiles = (long long)a.ny * p.nchunks;
    p.xs_log2 = a.xs_log2 >= 0 ? a.xs_log2 : 10;
    const int nxs = 1 << p.xs_log2;
    p.box_rows = nxs < 256 ? nxs : 256;
    p.tw = (const cx<float>*)a.tw32;
    p.offset = (const float*)a.offset[0];
    p.ntot = a.ntot;
    p.bins = *a.bins;
    p.seg = a.bins->seg_tab;
    p.seg_stride = a.bins->seg_stride;
    p.seg_pencils = (long long)a.bins->n[1] * a.nzh;
    p.seg_overflow = a.bins->seg_overflow;
    const int nt = (p.seg != nullptr && use_segments()) ? a.bins->seg_ntab : 0;
    if (nt != 0 && nt != 1 && nt != 4) { set_error("bad segment table count %d", nt); return T21_EINVAL; }
    p.mp_los = moments_los;
    if (moments_los >= 0) {
        if (nt != 1 || a.bins->nmu != 0) { set_error("the fused multipole pass needs a shell spec with one segment table"); return T21_EINVAL; }
        const int nk = a.bins->nk;
        const size_t smem_m = HIST_OFF + (size_t)(THREADS / 32) * nk * (kMoments * 8 + 4);
        auto kern_m = xpass32_moments_kernel;
        if (int rc = ensure_smem(kern_m, smem_m)) return rc;
        int occ_m = 0;
        T21_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_m, kern_m, THREADS, smem_m));
        if (occ_m < 1) { set_error("x pass (multipoles) does not fit on an SM (smem %zu)", smem_m); return T21_EINVAL; }
        CUtensorMap map_m;
        if (int rc = make_wmap_f32(&map_m, a.w[0], nxs, a.ny, a.pitch / WZ, NX / nxs, p.box_rows)) return rc;
        long long gm = (long long)occ_m * device_sm_count();
        if (gm > p.ntiles) gm = p.ntiles;
        if (gm > a.max_grid) gm = a.max_grid;
        kern_m<<<(int)gm, THREADS, smem_m, st>>>(map_m, p, nk, a.part_sum, a.part_cnt);
        count_launch();
        T21_CUDA_OK(cudaGetLastError());
        a.grid_used = (int)gm;
        return 0;
    }
    CUtensorMap map;
    if (int rc = make_wmap_f32(&map, a.w[0], nxs, a.ny, a.pitch / WZ, NX / nxs, p.box_rows)) return rc;

    const size_t smem = HIST_OFF + (size_t)(THREADS / 32) * a.nbins * 12;
    const bool ovf = nt != 0 && p.seg_overflow != 0;
    auto kern = nt == 0 ? xpass32_kernel<0, false>
              : (nt == 1 ? (ovf ? xpass32_kernel<1, true> : xpass32_kernel<1, false>)
                         : (ovf ? xpass32_kernel<4, true> : xpass32_kernel<4, false>));
    if (int rc = ensure_smem(kern, smem)) return rc;
    int occ = 0;
    T21_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, smem));
    if (occ < 1) { set_error("x pass (warp per pencil) does not fit on an SM (smem %zu)", smem); return T21_EINVAL; }
    long long g = (long long)occ * device_sm_count();
    if (g > p.ntiles) g = p.ntiles;
    if (g > a.max_grid) g = a.max_grid;
    kern<<<(int)g, THREADS, smem, st>>>(map, p, a.nbins, a.part_sum, a.part_cnt);
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    a.grid_used = (int)g;
    return 0;
}

}  // namespace x32

bool xpass32_applicable(int dtype, const XArgs& a) {
    // One field, spherical shells, one lane per bin edge; everything else keeps the CTA-per-tile kernel.  ((k, mu)
    // bins were tried here with the guard-interval cache of XPass: a warp that walks a whole pencil crosses ~7
    // (k, mu) bins per pencil and a two-entry cache thrashes -- 9.1 ms against 6.0 ms at 1024^3; it needs the
    // boundary form with host-built mu thresholds, which is not written.)
    if (!(dtype == T21_F32 && a.nx == x32::NX && a.nf == 1 && a.bins != nullptr && a.tw32 != nullptr && a.pitch % x32::WZ == 0 &&
          !a.hist_global))
        return false;
    if (a.bins->seg_tab != nullptr && x32::use_segments()) {     // bins as table-driven index ranges
        const bool is_signed = a.bins->nmu > 0 && !a.bins->absolute_mus;
        return a.nbins <= 128 && a.bins->seg_ntab == (is_signed ? 4 : 1) && (a.bins->seg_overflow == 0 || a.bins->thr != nullptr);
    }
    return a.bins->nmu == 0 && a.bins->nk <= 31;
}

int launch_xpass32_bin(XArgs& a, cudaStream_t st) {
    return x32::launch(a, st);
}

// the scaled power of the computed half spectrum (t21_ps_half, t21_slab_stage2_nd): fp32 transform, nx = 1024, one field
bool xpass32_store_applicable(int dtype, const XArgs& a) {
    static const bool on = [] { const char* e = getenv("T21_PENCIL32_ND"); return !(e && e[0] == '0'); }();
    // (the [kx][ky][kz] form writes 32-byte pieces 2 MB apart whoever does it: 2.97 ms here, 2.93 ms on the CTA-per-tile
    // kernel, which therefore keeps it unless T21_PENCIL32_HALF=1; the pencil-major form is this kernel's own)
    static const bool half_too = [] { const char* e = getenv("T21_PENCIL32_HALF"); return e && e[0] == '1'; }();
    return on && dtype == T21_F32 && a.nx == x32::NX && a.nf == 1 && a.nd_out != nullptr && a.tw32 != nullptr && a.pitch % x32::WZ == 0 &&
           (a.nd_half == -1 || (half_too && a.nd_half > 0 && a.nd_half % x32::WZ == 0));
}
// nd_half > 0: out_half[kx][ky][nd_half]; nd_half == -1: pencil-major out_t[ky][nz/2+1][kx]
int launch_xpass32_store(XArgs& a, cudaStream_t st) {
    using namespace x32;
    Params p;
    memset(&p, 0, sizeof(p));
    p.ny = a.ny; p.nz = a.nz; p.nzh = a.nzh; p.iy0 = a.iy0;
    p.nchunks = (a.nzh + WZ - 1) / WZ;
    p.ntiles = (long long)a.ny * p.nchunks;
    p.xs_log2 = a.xs_log2 >= 0 ? a.xs_log2 : 10;
    const int nxs = 1 << p.xs_log2;
    p.box_rows = nxs < 256 ? nxs : 256;
    p.tw = (const cx<float>*)a.tw32;
    p.offset = (const float*)a.offset[0];
    p.ntot = a.ntot;
    p.nd_out = a.nd_out; p.nd_f64 = a.nd_f64; p.nd_pitch = a.nd_half; p.scale = a.scale;
    CUtensorMap map;
    if (int rc = make_wmap_f32(&map, a.w[0], nxs, a.ny, a.pitch / WZ, NX / nxs, p.box_rows)) return rc;
    auto kern = a.nd_half == -1 ? xpass32_store_kernel<7> : xpass32_store_kernel<6>;
    if (int rc = ensure_smem(kern, (size_t)kStoreSmem)) return rc;
    int occ = 0;
    T21_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, THREADS, (size_t)kStoreSmem));
    if (occ < 1) { set_error("x pass (n-D store) does not fit on an SM"); return T21_EINVAL; }
    long long g = (long long)occ * device_sm_count();
    if (g > p.ntiles) g = p.ntiles;
    kern<<<(int)g, THREADS, kStoreSmem, st>>>(map, p);
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    a.grid_used = (int)g;
    return 0;
}

// multipole moments fused into the X pass: fp32, nx = 1024, one field, a shell spec that carries its segment table
bool xpass32_moments_applicable(int dtype, const XArgs& a) {
    return dtype == T21_F32 && a.nx == x32::NX && a.nf == 1 && a.bins != nullptr && a.tw32 != nullptr && a.pitch % x32::WZ == 0 &&
           a.bins->seg_tab != nullptr && a.bins->seg_ntab == 1 && a.bins->seg_overflow == 0 && a.bins->nmu == 0 && a.bins->nk <= 64 &&
           x32::use_segments();
}
int launch_xpass32_moments(XArgs& a, int los_axis, cudaStream_t st) {
    return x32::launch(a, st, los_axis);
}

}  // namespace t21
