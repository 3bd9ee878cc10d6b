// In-register radix butterflies and the staged Stockham pencil FFT.
//
// A length-N pencil (N a power of two) is transformed by T = N/E threads, each
// holding E complex values in registers.  Stage s applies E/R_s radix-R_s
// butterflies per thread; between stages values are exchanged through shared
// memory in Stockham (self-sorting) order, so the last stage produces natural
// frequency order with no bit reversal:
//
//   butterfly jj in [0, N/R):  in[jj + k*N/R] * w^(k*(jj mod Ns))  -> DFT_R ->
//                              out[(jj - jj mod Ns)*R + jj mod Ns + k*Ns]
//
// Twiddles between stages come from per-stage tables laid out [k-1][m]
// (m = jj mod Ns fastest) so consecutive butterflies read consecutive entries.
// Tables are built on the host in float64 and rounded once (t21_plan_create).
#pragma once
#include "t21_common.cuh"

namespace t21 {

// ---- compile-time constants: exp(-2*pi*i*idx/32) --------------------------------
constexpr double kCos32[9] = {1.0,
                              0.98078528040323044912618223613424,
                              0.92387953251128675612818318939679,
                              0.83146961230254523707878837761791,
                              0.70710678118654752440084436210485,
                              0.55557023301960222474283081394853,
                              0.38268343236508977172845998403040,
                              0.19509032201612826784828486847702,
                              0.0};
constexpr double cos32(int i) {
    i &= 31;
    if (i > 16) i = 32 - i;
    return i > 8 ? -kCos32[16 - i] : kCos32[i];
}
constexpr double sin32(int i) { return cos32((i + 24) & 31); }

// v *= exp(-2*pi*i*K/R).  K is a template parameter so the constants are evaluated by the
// front end (device code may not index a host constexpr array with a run-time value).
template <int R, int K, typename T>
T21_HD cx<T> rot(cx<T> v) {
    constexpr int idx = K * (32 / R);  // in 32nds of a turn
    if constexpr (idx == 0) {
        return v;
    } else if constexpr (idx == 8) {
        return cmake<T>(v.y, -v.x);                        // * -i
    } else if constexpr (idx == 4) {
        constexpr T h = (T)0.70710678118654752440084436210485;
        return cscale(cadd(v, cmake<T>(v.y, -v.x)), h);        // * (1-i)/sqrt2 = ((x + y) h, (y - x) h)
    } else if constexpr (idx == 12) {
        constexpr T h = (T)0.70710678118654752440084436210485;
        return cscale(csub(cmake<T>(v.y, -v.x), v), h);        // * (-1-i)/sqrt2 = ((y - x) h, -(x + y) h)
    } else {
        constexpr T c = (T)cos32(idx), s = (T)sin32(idx);
        // (x + iy)(c - is) = (xc + ys) + i(yc - xs)
        return cmul(v, cmake<T>(c, -s));
    }
}

template <int R, typename T, int K = 0>
T21_HD void dft_combine(cx<T>* v, const cx<T>* e, const cx<T>* o) {
    if constexpr (K < R / 2) {
        const cx<T> t = rot<R, K, T>(o[K]);
        v[K] = cadd(e[K], t);
        v[K + R / 2] = csub(e[K], t);
        dft_combine<R, T, K + 1>(v, e, o);
    }
}

template <int R, typename T>
struct DFT {
    // in place, natural order in and out; forward sign exp(-i...)
    static T21_HD void run(cx<T>* v) {
        static_assert(R >= 4 && R <= 32 && is_pow2(R), "radix");
        cx<T> e[R / 2], o[R / 2];
#pragma unroll
        for (int i = 0; i < R / 2; ++i) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
        DFT<R / 2, T>::run(e);
        DFT<R / 2, T>::run(o);
        dft_combine<R, T>(v, e, o);
    }
};
template <typename T>
struct DFT<2, T> {
    static T21_HD void run(cx<T>* v) {
        cx<T> a = v[0], b = v[1];
        v[0] = cadd(a, b);
        v[1] = csub(a, b);
    }
};
template <typename T>
struct DFT<1, T> {
    static T21_HD void run(cx<T>*) {}
};

// ---- compile-time plan ------------------------------------------------------------
// EMAX = values per thread (= largest radix), 16 or 32.  32 halves the number of shared-memory
// exchanges and barriers and lets a CTA be half as many threads (more CTAs per SM for the same
// shared memory) at the price of ~86 registers per thread; each pass picks its own (t21_internal.h).
template <int N, int EMAX = 16>
struct Plan {
    static_assert(is_pow2(N) && N >= 2, "pencil length must be a power of two");
    static_assert(EMAX == 16 || EMAX == 32, "EMAX");
    static constexpr int E = N < EMAX ? N : EMAX;  // values per thread
    static constexpr int LEN = N;
    static constexpr int T = N / E;            // threads per pencil
    static constexpr int radix(int s) {        // radix of stage s (0 past the end)
        int rem = N;
        for (int i = 0; i < s; ++i) rem /= (rem < E ? rem : E);
        return rem <= 1 ? 0 : (rem < E ? rem : E);
    }
    static constexpr int ns(int s) {           // product of the radices before stage s
        int p = 1;
        for (int i = 0; i < s; ++i) p *= radix(i);
        return p;
    }
    static constexpr int stages() {
        int s = 0;
        while (radix(s) != 0) ++s;
        return s;
    }
    static constexpr int NS = stages();
    static constexpr int tw_off(int s) {       // offset of stage s in the twiddle table
        int o = 0;
        for (int i = 1; i < s; ++i) o += (radix(i) - 1) * ns(i);
        return o;
    }
    static constexpr int TW_TOTAL = tw_off(NS);
    // three-stage plans 16*16*r: the middle stage can run in place (one butterfly per thread, whose 16 inputs
    // and 16 outputs are the same 16 locations), which saves the barrier between its loads and its stores
    static constexpr bool kInPlace1 = NS == 3 && radix(0) == 16 && radix(1) == 16 && E == 16;
};

// number of twiddle entries for a run-time N (host side of t21_plan_create)
inline int plan_tw_total(int n, int emax) {
    int e = n < emax ? n : emax, rem = n, nsv = 1, tot = 0, s = 0;
    while (rem > 1) {
        int r = rem < e ? rem : e;
        if (s > 0) tot += (r - 1) * nsv;
        nsv *= r; rem /= r; ++s;
    }
    return tot;
}

// Walks the rows a thread touches in one stage, in the order the unrolled loops visit them
// (butterflies b outer, points k inner), with two 64-bit adds per access instead of an index
// multiply: element (b, k) sits at p0 + b*step_b' + k*step_k.
template <typename T>
struct RowCursor {
    T* p;
    long long step_k;     // between consecutive points of one butterfly
    long long step_wrap;  // from the last point of butterfly b to the first point of butterfly b+1
    int k, radix;
    T21_HD RowCursor(T* p0, long long sk, long long sb, int r)
        : p(p0), step_k(sk), step_wrap(sb - (long long)(r - 1) * sk), k(0), radix(r) {}
    T21_HD T* next() {
        T* cur = p;
        if (++k == radix) { k = 0; p += step_wrap; } else { p += step_k; }
        return cur;
    }
};

// ---- one stage, split into its load and its compute+store halves -------------------
// ld(i) returns element i of this thread's pencil; st(i, v, slot) stores element i, which
// lives in register slot `slot` (a compile-time constant after unrolling).
// PERM (stage 2 after an in-place stage 1, see Plan::kInPlace1): element e = jj + k*STRIDE of the Stockham
// order was left at position 16*(e >> 8) + (e & 15) + T*((e >> 4) & 15) by the thread that produced it.
template <class P, int S, bool PERM = false, typename T, class LoadFn>
T21_HD void stage_load(cx<T>* v, int j, LoadFn&& ld) {
    constexpr int R = P::radix(S), NB = P::E / R, STRIDE = P::LEN / R;
    static_assert(!PERM || (S == 2 && P::kInPlace1), "permuted loads follow the in-place stage");
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int jj = j + b * P::T;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            if constexpr (PERM) v[b * R + k] = ld(16 * k + (jj & 15) + P::T * (jj >> 4));   // STRIDE == 256
            else v[b * R + k] = ld(jj + k * STRIDE);
        }
    }
}

// INPLACE (stage 1 only): output k goes back to where input k came from (jj + k*STRIDE) instead of to its
// Stockham position, so the thread touches no location another thread reads or writes in this stage and its
// loads and stores need no barrier between them; stage 2 then loads with PERM.
template <class P, int S, bool INPLACE = false, typename T, class StoreFn>
T21_HD void stage_compute(cx<T>* v, int j, const cx<T>* __restrict__ tw, StoreFn&& st) {
    constexpr int R = P::radix(S), NB = P::E / R, NSV = P::ns(S), OFF = P::tw_off(S);
    static_assert(!INPLACE || (S == 1 && P::kInPlace1), "only the middle stage of a 16*16*r plan runs in place");
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        const int jj = j + b * P::T;
        const int m = jj & (NSV - 1);
        if (NSV > 1) {
#pragma unroll
            for (int k = 1; k < R; ++k)
                v[b * R + k] = cmul(v[b * R + k], ldg(tw + OFF + (k - 1) * NSV + m));
        }
        DFT<R, T>::run(v + b * R);
        const int base = (jj - m) * R + m;
#pragma unroll
        for (int k = 0; k < R; ++k) {
            if constexpr (INPLACE) st(jj + k * (P::LEN / R), v[b * R + k], b * R + k);
            else st(base + k * NSV, v[b * R + k], b * R + k);
        }
    }
}

}  // namespace t21
