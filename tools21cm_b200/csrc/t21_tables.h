// Host-side construction of the twiddle tables (float64 math, rounded once to T).
#pragma once
#include <cmath>
#include <vector>
#include "t21_fft.cuh"

namespace t21 {

// exp(-2*pi*i*num/den) evaluated with the argument reduced to the first octant so that
// exact symmetries (e.g. num/den = 1/8) are exact in the table.
inline void unit_root(long long num, long long den, double& c, double& s) {
    num %= den;
    if (num < 0) num += den;
    // angle = 2*pi*num/den; reduce to [0, pi/4]
    long long oct = (8 * num) / den;           // octant 0..7
    long long rnum = 8 * num - oct * den;      // remainder in units of (2*pi/8)/den
    const double PI = 3.14159265358979323846264338327950288;
    double a = (PI / 4.0) * ((double)rnum / (double)den);   // in [0, pi/4)
    double ca, sa;
    if (oct & 1) { a = PI / 4.0 - a; ca = std::sin(a); sa = std::cos(a); }
    else { ca = std::cos(a); sa = std::sin(a); }
    // (ca, sa) = (cos, sin) of the angle within the quadrant pair; rotate by octant/2 quarter turns
    double cq, sq;
    switch ((oct >> 1) & 3) {
        case 0: cq = ca; sq = sa; break;
        case 1: cq = -sa; sq = ca; break;
        case 2: cq = -ca; sq = -sa; break;
        default: cq = sa; sq = -ca; break;
    }
    c = cq;
    s = -sq;  // exp(-i theta) = cos - i sin
}

// per-stage tables of Plan<n, emax>: stage s>=1, entry [(k-1)*Ns + m] = exp(-2 pi i k m / (Ns*R))
template <typename T>
std::vector<cx<T>> build_stage_twiddles(int n, int emax) {
    std::vector<cx<T>> tab;
    const int e = n < emax ? n : emax;
    int rem = n, nsv = 1, s = 0;
    while (rem > 1) {
        const int r = rem < e ? rem : e;
        if (s > 0) {
            for (int k = 1; k < r; ++k)
                for (int m = 0; m < nsv; ++m) {
                    double c, sn;
                    unit_root((long long)k * m, (long long)nsv * r, c, sn);
                    tab.push_back(cmake<T>((T)c, (T)sn));
                }
        }
        nsv *= r; rem /= r; ++s;
    }
    if (tab.empty()) tab.push_back(cmake<T>((T)1, (T)0));
    return tab;
}

// exp(-2 pi i k / n), k = 0..count-1
template <typename T>
std::vector<cx<T>> build_roots(int n, int count) {
    std::vector<cx<T>> tab(count);
    for (int k = 0; k < count; ++k) {
        double c, s;
        unit_root(k, n, c, s);
        tab[k] = cmake<T>((T)c, (T)s);
    }
    return tab;
}

}  // namespace t21
