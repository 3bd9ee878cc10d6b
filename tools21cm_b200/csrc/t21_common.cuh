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

// Packed fp32 pairs (sm_100a: add/mul/fma.rn.f32x2 -> FADD2 / FMUL2 / FFMA2, one issue slot for both halves of a
// complex value; ptxas folds the swaps, per-half negations and scalar broadcasts below into operand modifiers).
#ifndef T21_PACKED_F32
#define T21_PACKED_F32 1
#endif
#if T21_DEVICE_CODE && T21_PACKED_F32
#define T21_HAVE_F32X2 1
T21_D unsigned long long pk(cx<float> a) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a.x), "f"(a.y)); return r; }
T21_D cx<float> upk(unsigned long long v) { cx<float> r; asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v)); return r; }
T21_D cx<float> add2(cx<float> a, cx<float> b) { unsigned long long d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b))); return upk(d); }
T21_D cx<float> sub2(cx<float> a, cx<float> b) { unsigned long long d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b))); return upk(d); }
T21_D cx<float> mul2(cx<float> a, cx<float> b) { unsigned long long d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pk(a)), "l"(pk(b))); return upk(d); }
T21_D cx<float> fma2(cx<float> a, cx<float> b, cx<float> c) { unsigned long long d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(pk(a)), "l"(pk(b)), "l"(pk(c))); return upk(d); }
#else
#define T21_HAVE_F32X2 0
#endif

template <typename T> T21_HD cx<T> cadd(cx<T> a, cx<T> b) { return cmake<T>(a.x + b.x, a.y + b.y); }
template <typename T> T21_HD cx<T> csub(cx<T> a, cx<T> b) { return cmake<T>(a.x - b.x, a.y - b.y); }
template <typename T> T21_HD cx<T> cmul(cx<T> a, cx<T> b) {
    return cmake<T>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
template <> T21_HD cx<float> cadd<float>(cx<float> a, cx<float> b) {
#if T21_HAVE_F32X2
    return add2(a, b);
#else
    return cmake<float>(a.x + b.x, a.y + b.y);
#endif
}
template <> T21_HD cx<float> csub<float>(cx<float> a, cx<float> b) {
#if T21_HAVE_F32X2
    return sub2(a, b);
#else
    return cmake<float>(a.x - b.x, a.y - b.y);
#endif
}
template <typename T> T21_HD cx<T> cscale(cx<T> a, T h) { return cmake<T>(a.x * h, a.y * h); }
template <> T21_HD cx<float> cscale<float>(cx<float> a, float h) {
#if T21_HAVE_F32X2
    return mul2(a, cmake<float>(h, h));
#else
    return cmake<float>(a.x * h, a.y * h);
#endif
}
// (ax + i ay)(bx + i by) = (ax, ay) * (bx, bx) + (ay, ax) * (-by, by)
template <> T21_HD cx<float> cmul<float>(cx<float> a, cx<float> b) {
#if T21_HAVE_F32X2
    return fma2(a, cmake<float>(b.x, b.x), mul2(cmake<float>(a.y, a.x), cmake<float>(-b.y, b.y)));
#else
    return cmake<float>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
#endif
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
