// Y pass (in-place strided pencils) instantiations.
#include <string.h>
#include "t21_launch.cuh"

namespace t21 {

template <typename T>
static int launch_ypass_t(const YArgs& a, cudaStream_t st) {
#define YCASE(N)                                                                         \
    {                                                                                    \
        constexpr int WZ = WZFor<T, N>::value;                                           \
        const int nchunks = (a.nzh + WZ - 1) / WZ;                                       \
        if (a.peers) {                                                                   \
            using KP = YPass<N, T, WZ, 2>;                                               \
            typename KP::Params pp;                                                      \
            memset(&pp, 0, sizeof(pp));                                                  \
            pp.w = (cx<T>*)a.w; pp.nzh = a.nzh; pp.pitch = a.pitch; pp.nchunks = nchunks; \
            pp.cstride = a.nouter * a.ny * WLay<T>::LW; pp.tw = (const cx<T>*)a.tw;      \
            pp.nyl_log2 = a.nyl_log2;                                                    \
            pp.peer_cstride = a.nx_global * (a.ny / a.npeers) * WLay<T>::LW;             \
            pp.nouter = a.nouter; pp.xg0 = a.xg0;                                        \
            for (int r = 0; r < a.npeers; ++r) pp.peer[r] = (cx<T>*)a.peers[r];          \
            return launch_tiles<KP, T>(pp, a.nouter * nchunks, st);                      \
        }                                                                                \
        if (a.send) {                                                                    \
            using KS = YPass<N, T, WZ, 1>;                                               \
            typename KS::Params ps;                                                      \
            memset(&ps, 0, sizeof(ps));                                                  \
            ps.w = (cx<T>*)a.w; ps.nzh = a.nzh; ps.pitch = a.pitch; ps.nchunks = nchunks; \
            ps.cstride = a.nouter * a.ny * WLay<T>::LW; ps.tw = (const cx<T>*)a.tw;      \
            ps.send = (cx<T>*)a.send;                                                    \
            ps.nyl_log2 = a.nyl_log2; ps.nouter = a.nouter;                              \
            return launch_tiles<KS, T>(ps, a.nouter * nchunks, st);                      \
        }                                                                                \
        using K = YPass<N, T, WZ>;                                                       \
        typename K::Params p;                                                            \
        memset(&p, 0, sizeof(p));                                                        \
        p.w = (cx<T>*)a.w; p.nzh = a.nzh; p.pitch = a.pitch; p.nchunks = nchunks;        \
        p.cstride = (a.nxa ? a.nxa : a.nouter) * a.ny * WLay<T>::LW; p.tw = (const cx<T>*)a.tw; \
        p.nouter = a.nouter;                                                             \
        return launch_tiles<K, T>(p, a.nouter * nchunks, st);                            \
    }
    T21_FOR_FAST_N(a.ny, YCASE)
#undef YCASE
    return 0;
}

int launch_ypass(int dtype, const YArgs& a, cudaStream_t st) {
    return dtype == T21_F32 ? launch_ypass_t<float>(a, st) : launch_ypass_t<double>(a, st);
}

}  // namespace t21
