// Z pass (real rows -> half spectrum) instantiations.
#include <string.h>
#include "t21_launch.cuh"

namespace t21 {

template <typename T>
static int launch_zpass_t(const ZArgs& a, cudaStream_t st) {
#define ZCASE(N)                                                                                   \
    {                                                                                              \
        using K = ZPass<N, T>;                                                                     \
        typename K::Params p{(const T*)a.in, (cx<T>*)a.out, a.nrows, a.pitch, a.ny,                \
                             a.nxa * a.ny * WLay<T>::LW, (const T*)a.offset,                        \
                             (const cx<T>*)a.tw, (const cx<T>*)a.twr, 0, {0, 0, 0}, {0, 0, 0}};     \
        if (a.window) {                                                                            \
            if (a.nrows > 2147483647LL) { set_error("cube too large for a windowed pass"); return T21_EINVAL; } \
            p.win_on = 1;                                                                          \
            for (int d = 0; d < 3; ++d) { p.win_lo[d] = a.window[2 * d]; p.win_hi[d] = a.window[2 * d + 1]; } \
        }                                                                                          \
        return launch_tiles<K, T>(p, (a.nrows + K::ROWS - 1) / K::ROWS, st);                       \
    }
    T21_FOR_FAST_N(a.nz, ZCASE)
#undef ZCASE
    return 0;
}

int launch_zpass(int dtype, const ZArgs& a, cudaStream_t st) {
    return dtype == T21_F32 ? launch_zpass_t<float>(a, st) : launch_zpass_t<double>(a, st);
}

}  // namespace t21
