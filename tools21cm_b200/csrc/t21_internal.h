// Internal, type-erased interfaces between the translation units of libt21ps.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/t21ps.h"
#include "t21_config.h"

namespace t21 {

struct ZArgs {
    const void* in; void* out; long long nrows; int nz; int pitch; int ny; long long nxa;   // nxa: x planes in `out`
    const void* offset; const void* tw; const void* twr;
    const int* window;    // null, or {x0, x1, y0, y1, z0, z1}: samples outside are read as zero (mask fused into the load)
};
struct YArgs {
    void* w; long long nouter; int ny; int nzh; int pitch; const void* tw;
    void* send; int nyl_log2;          // slab mode: packed all-to-all send buffer (or null)
    void* const* peers; int npeers; long long xg0; long long nx_global;   // fused exchange: peer y-slab buffers (or null)
    long long nxa;                     // x planes in the allocation `w` points into (0: nouter) -- plane-group launches
};
struct XArgs {
    const void* w[2]; int nf; int nx, ny, nz, nzh, pitch; int xs_log2; int iy0;   // xs_log2: log2(x planes per slab), <0 = one slab
    const void* tw; const void* offset[2]; double ntot;
    const void* tw32;                        // Plan<nx, 32> stage twiddles (warp-per-pencil kernels) or null
    const t21_binspec* bins;                 // MODE_BIN
    void* nd_out; int nd_f64; int nd_half; double scale;  // MODE_ND
    // histogram plumbing (MODE_BIN)
    int nbins; int hist_global;              // 1: accumulate straight into gsum/gcnt
    double* part_sum; unsigned* part_cnt;    // [grid][nbins] per-CTA partials
    double* gsum; unsigned long long* gcnt;
    int max_grid;                            // capacity of the partial buffers, in CTAs
    int grid_used;                           // out
};

int launch_zpass(int dtype, const ZArgs& a, cudaStream_t st);
int launch_ypass(int dtype, const YArgs& a, cudaStream_t st);
int launch_xpass_f32_bin_mu0(XArgs& a, cudaStream_t st);   // one translation unit per mu mode
int launch_xpass_f32_bin_mu1(XArgs& a, cudaStream_t st);
int launch_xpass_f32_bin_mu2(XArgs& a, cudaStream_t st);
int launch_xpass_f64_bin_mu0(XArgs& a, cudaStream_t st);
int launch_xpass_f64_bin_mu1(XArgs& a, cudaStream_t st);
int launch_xpass_f64_bin_mu2(XArgs& a, cudaStream_t st);
int launch_xpass_f32_nd(XArgs& a, cudaStream_t st);
int launch_xpass_f64_nd(XArgs& a, cudaStream_t st);
// t21_xpass32.cu: warp-per-pencil X pass (fp32, nx = 1024, one field), TMA-staged
bool xpass32_applicable(int dtype, const XArgs& a);
int launch_xpass32_bin(XArgs& a, cudaStream_t st);
bool xpass32_store_applicable(int dtype, const XArgs& a);
int launch_xpass32_store(XArgs& a, cudaStream_t st);
bool xpass32_moments_applicable(int dtype, const XArgs& a);
int launch_xpass32_moments(XArgs& a, int los_axis, cudaStream_t st);   // part_sum [grid][nk][5], part_cnt [grid][nk]

// t21_xpass16.cu: warp-per-pencil X pass for nx = 512 (fp32, one or two fields, a spec with one segment table)
bool xpass16_applicable(int dtype, const XArgs& a);
int launch_xpass16_bin(XArgs& a, cudaStream_t st);

inline int launch_xpass_bin(int dtype, XArgs& a, cudaStream_t st) {
    if (xpass32_applicable(dtype, a)) return launch_xpass32_bin(a, st);
    if (xpass16_applicable(dtype, a)) return launch_xpass16_bin(a, st);
    const int mu = (a.bins && a.bins->nmu > 0) ? (a.bins->absolute_mus ? 1 : 2) : 0;
    if (dtype == T21_F32)
        return mu == 0 ? launch_xpass_f32_bin_mu0(a, st) : (mu == 1 ? launch_xpass_f32_bin_mu1(a, st) : launch_xpass_f32_bin_mu2(a, st));
    return mu == 0 ? launch_xpass_f64_bin_mu0(a, st) : (mu == 1 ? launch_xpass_f64_bin_mu1(a, st) : launch_xpass_f64_bin_mu2(a, st));
}

// t21_generic.cu: any-length direct-DFT versions (tw / twr hold the n-th roots of unity there)
int launch_zpass_generic(int dtype, const ZArgs& a, cudaStream_t st);
int launch_ypass_generic(int dtype, const YArgs& a, cudaStream_t st);
int launch_xpass_generic(int dtype, XArgs& a, cudaStream_t st);

// t21_aux.cu
int launch_mean_estimate(int dtype, const void* x, long long n, void* offset_out, cudaStream_t st);
int launch_reduce_partials(const double* part_sum, const unsigned* part_cnt, int nparts, int nbins,
                           double* sums_out, long long* counts_out, cudaStream_t st);
int launch_zero_hist(double* sums, long long* counts, int nbins, cudaStream_t st);
int launch_bin_only(const void* p, int dtype, const t21_binspec& spec, double* part_sum, unsigned* part_cnt,
                    int max_grid, double* gsum, long long* gcnt, int* grid_used, cudaStream_t st);

int launch_bin_cyl(const void* p, int dtype, int n0, int n1, int n2, int nu_axis, const int* perp_tab,
                   const int* par_tab, int nperp, int npar, double* part_sum, unsigned* part_cnt, int max_grid,
                   double* gsum, long long* gcnt, int* grid_used, cudaStream_t st);

int launch_cyl_rows(const void* p, int dtype, int n0, int n1, int nz, int pitch, int nu_axis, const int* perp_tab, const int* par_tab,
                    int nperp, int npar, double* part_sum, int max_grid, double* sums_out, cudaStream_t st);
int launch_bin_multipoles(const void* p, int dtype, const t21_binspec& spec, int pitch, int los, int exclude_zero,
                          double* part_sum, unsigned* part_cnt, int max_grid, int* grid_used, cudaStream_t st);
int launch_reduce_moments(const double* part_sum, const unsigned* part_cnt, int nparts, int nk, double* sums_out,
                          long long* counts_out, cudaStream_t st);

int launch_pencil_segments(const t21_binspec& spec, int which, unsigned* out, int stride, int* max_seg_dev, cudaStream_t st);

int launch_pencil_segments_cyl(int n0, int n1, int n2, int nu_axis, const int* perp_tab, const int* par_tab, int npar,
                               unsigned* out, int stride, int* max_seg_dev, cudaStream_t st);

// error plumbing (t21_api.cu)
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch();

constexpr int kSmemHistMaxBins = 2048;   // CTA histograms up to this many bins live in shared memory
int device_sm_count();

}  // namespace t21

#define T21_CUDA_OK(expr)                                              \
    do {                                                               \
        cudaError_t e__ = (expr);                                      \
        if (e__ != cudaSuccess) return ::t21::cuda_fail(e__, #expr);   \
    } while (0)
