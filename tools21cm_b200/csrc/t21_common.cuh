// Shared host/device helpers for the t21ps kernels (sm_100a).
//
// Kernel bodies are written as barrier-separated *phases* over host/device
// templates: on the GPU a kernel runs phase 0, __syncthreads(), phase 1, ...;
// tests/emu replays the same phases thread by thread on the CPU to verify the
// index arithmetic (no GPU in the build container).  Nothing in the emulator
// is linked into the product library.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define T21_HD __host__ __device__ __forceinline__
#define T21_D __device__ __forceinline__
#else
#define T21_HD inline
#define T21_D inline
#endif

#if defined(__CUDA_ARCH__)
#define T21_DEVICE_CODE 1
#else
#define T21_DEVICE_CODE 0
#endif

namespace t21 {

template <typename T>
struct alignas(2 * sizeof(T)) cx {
    T x, y;
};

template <typename T> T21_HD cx<T> cmake(T x, T y) { cx<T> r; r.x = x; r.y = y; return r; }
template <typename T> T21_HD cx<T> cadd(cx<T> a, cx<T> b) { return cmake<T>(a.x + b.x, a.y + b.y); }
template <typename T> T21_HD cx<T> csub(cx<T> a, cx<T> b) { return cmake<T>(a.x - b.x, a.y - b.y); }
template <typename T> T21_HD cx<T> cmul(cx<T> a, cx<T> b) {
    return cmake<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}

// read-only global load (LDG through L1 on the device, plain load on the host)
template <typename T> T21_HD T ldg(const T* p) {
#if T21_DEVICE_CODE
    return __ldg(p);
#else
    return *p;
#endif
}
template <typename T> T21_HD cx<T> ldg(const cx<T>* p);
template <> T21_HD cx<float> ldg<float>(const cx<float>* p) {
#if T21_DEVICE_CODE
    float2 v = __ldg(reinterpret_cast<const float2*>(p));
    return cmake<float>(v.x, v.y);
#else
    return *p;
#endif
}
template <> T21_HD cx<double> ldg<double>(const cx<double>* p) {
#if T21_DEVICE_CODE
    double2 v = __ldg(reinterpret_cast<const double2*>(p));
    return cmake<double>(v.x, v.y);
#else
    return *p;
#endif
}

// ask for the line holding p to be brought into L2 (no register, no stall); no-op on the host
T21_HD void prefetch_l2(const void* p) {
#if T21_DEVICE_CODE
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// IEEE double add that the compiler may never contract into an FMA
T21_HD double dadd_rn(double a, double b) {
#if T21_DEVICE_CODE
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}

constexpr bool is_pow2(int n) { return n > 0 && (n & (n - 1)) == 0; }
constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }

}  // namespace t21
