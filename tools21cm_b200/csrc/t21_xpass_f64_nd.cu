// X pass instantiations: double, MODE_ND (auto and cross).
#include <string.h>
#include "t21_launch.cuh"

namespace t21 {

int launch_xpass_f64_nd(XArgs& a, cudaStream_t st) {
    using T = double;
#define XCASE(N)                                                                  \
    {                                                                             \
        constexpr int WZ = WZForX<T, N>::value;                                   \
        if (a.nf == 1) return launch_x<XPass<N, T, WZ, 1, MODE_ND>, T>(a, st);      \
        return launch_x<XPass<N, T, WZForX2<T, N>::value, 2, MODE_ND>, T>(a, st);                     \
    }
    T21_FOR_FAST_N(a.nx, XCASE)
#undef XCASE
    return 0;
}

}  // namespace t21
