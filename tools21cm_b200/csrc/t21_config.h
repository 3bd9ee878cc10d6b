// Compile-time tuning choices shared by the kernels, the launchers and the CPU replay
// (no CUDA headers here: tests/emu includes it with a plain C++ compiler).
#pragma once

namespace t21 {

// widest pow2 pencil the Stockham kernels are instantiated for; other lengths take the
// generic DFT kernels (t21_generic.cu)
constexpr int kMinFast = 8;
constexpr int kMaxFast = 4096;
inline bool fast_len(int n) { return n >= kMinFast && n <= kMaxFast && (n & (n - 1)) == 0; }

// columns of the half spectrum a strided CTA owns (WZ * sizeof(complex) contiguous bytes per row)
// values per thread (largest radix) of each pass's pencil plan, by pencil length.  Measured on B200 at
// 1024^3: radix-32 plans save an exchange but need ~86 registers; with 2 CTAs/SM of 256
// threads they are no faster (Y 1.92 vs 1.99 ms, Z 1.91 vs 1.74 ms), with 3 CTAs/SM they spill
// (Y 2.94 ms).  So every pass uses 16; the knobs stay for other sizes / later rounds.
#ifndef T21_EMAX_Y
#define T21_EMAX_Y 16
#endif
#ifndef T21_EMAX_X
#define T21_EMAX_X 16
#endif
constexpr int emax_z(int) { return 16; }                                            // applied to nz/2
constexpr int emax_y(int n) { return (n == 512 || n == 1024) ? T21_EMAX_Y : 16; }
constexpr int emax_x(int n) { return (n == 512 || n == 1024) ? T21_EMAX_X : 16; }

// 16*16*r pencil plans run their middle stage in place (t21_fft.cuh, Plan::kInPlace1)
#ifndef T21_INPLACE_STAGE1
#define T21_INPLACE_STAGE1 1
#endif

#ifndef T21_WZ_LONG
#define T21_WZ_LONG 8
#endif
template <typename T, int N>
struct WZFor {
    static constexpr int value = sizeof(T) == 4 ? (N <= 512 ? 16 : (N <= 2048 ? (N == 1024 ? T21_WZ_LONG : 8) : 4))
                                                : (N <= 512 ? 8 : (N <= 2048 ? 4 : 2));
};

// The X pass may own fewer columns per CTA than the Y pass (half a 64-byte piece): smaller CTAs, more of
// them per SM, barriers that stall fewer warps.  The pieces' parts are consecutive tiles.  Measured at 1024^3:
// 4 columns 2.71 ms against 2.31 ms for 8, so the auto spectrum keeps full pieces.
#ifndef T21_WZ_X
#define T21_WZ_X T21_WZ_LONG
#endif
template <typename T, int N>
struct WZForX {
    static constexpr int value = (sizeof(T) == 4 && N == 1024) ? T21_WZ_X : WZFor<T, N>::value;
};

// Cross spectra (two fields) at 2048: a 1024-thread CTA leaves 64 registers per thread for two fields' worth of
// values; half-width tiles (512 threads, 128 registers) avoid the spills: 2048^3 cross on one GPU 99.6 -> 96.6 ms.
#ifndef T21_WZ_X2_2048
#define T21_WZ_X2_2048 4
#endif
template <typename T, int N>
struct WZForX2 {
    static constexpr int value = (sizeof(T) == 4 && N == 2048) ? T21_WZ_X2_2048 : WZForX<T, N>::value;
};

}  // namespace t21
