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
        if (a.send) {                                                                    \
            using KS = YPass<N, T, WZ, true>;                                            \
            typename KS::Params ps{(cx<T>*)a.w, a.nzh, a.pitch, nchunks, a.xb, (const cx<T>*)a.tw, \
                                   (cx<T>*)a.send, a.nyl_log2, a.nouter};                \
            return launch_tiles<KS, T>(ps, a.nouter * nchunks, st);                      \
        }                                                                                \
        using K = YPass<N, T, WZ>;                                                       \
        typename K::Params p{(cx<T>*)a.w, a.nzh, a.pitch, nchunks, a.xb, (const cx<T>*)a.tw,      \
                             (cx<T>*)a.send, a.nyl_log2, a.nouter};                        \
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
