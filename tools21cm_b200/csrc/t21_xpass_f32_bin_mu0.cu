// X pass instantiations: float, MODE_BIN, mu mode 0 (0 shells, 1 |mu|, 2 signed mu); auto and cross.
#include <string.h>
#include "t21_launch.cuh"

namespace t21 {

int launch_xpass_f32_bin_mu0(XArgs& a, cudaStream_t st) {
    using T = float;
#define XCASE(N)                                                                          \
    {                                                                                     \
        constexpr int WZ = WZForX<T, N>::value;                                           \
        if (a.nf == 1) return launch_x<XPass<N, T, WZ, 1, MODE_BIN, 0>, T>(a, st);    \
        return launch_x<XPass<N, T, WZForX2<T, N>::value, 2, MODE_BIN, 0>, T>(a, st);                   \
    }
    T21_FOR_FAST_N(a.nx, XCASE)
#undef XCASE
    return 0;
}

}  // namespace t21
