// C ABI of libt21ps.so (declared in include/t21ps.h): plans, workspace layout, and the
// three entry points that chain the passes on the caller's stream.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <vector>

#include "t21_internal.h"
#include "t21_tables.h"

namespace t21 {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
int cuda_fail(cudaError_t e, const char* what) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return T21_ECUDA;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// SM count of the current device, cached per device ordinal (atomics: calls may come from several host threads)
int device_sm_count() {
    constexpr int kMaxDev = 64;
    static std::atomic<int> cache[kMaxDev];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev >= 0 && dev < kMaxDev) {
        const int hit = cache[dev].load(std::memory_order_relaxed);
        if (hit > 0) return hit;
    }
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    if (dev >= 0 && dev < kMaxDev) cache[dev].store(n, std::memory_order_relaxed);
    return n;
}

// T21_PENCIL32=0 keeps the round-1 CTA-per-tile kernels everywhere (A/B runs); read once
static bool use_pencil32() {
    static const bool on = [] { const char* e = getenv("T21_PENCIL32"); return !(e && e[0] == '0'); }();
    return on;
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- optional per-pass timing (bench.py's roofline block) ------------------------------
// When enabled (per host thread), every public call of that thread records a CUDA event on the caller's stream
// after each pass; t21_profile_last() turns them into per-slot durations.  Off by default.
enum { SLOT_MEAN = 0, SLOT_Z, SLOT_Y, SLOT_X, SLOT_REDUCE, SLOT_COUNT };
struct Prof {
    bool on = false;
    static constexpr int kMax = 16;
    cudaEvent_t ev[kMax + 1];
    int slot[kMax];
    int n = 0;          // segments recorded in the last call
    bool made = false;
};
static thread_local Prof g_prof;   // per host thread: concurrent callers do not share event slots

static void prof_begin(cudaStream_t st) {
    if (!g_prof.on) return;
    if (!g_prof.made) {
        for (int i = 0; i <= Prof::kMax; ++i) cudaEventCreate(&g_prof.ev[i]);
        g_prof.made = true;
    }
    g_prof.n = 0;
    cudaEventRecord(g_prof.ev[0], st);
}
static void prof_mark(cudaStream_t st, int slot) {
    if (!g_prof.on || g_prof.n >= Prof::kMax) return;
    g_prof.slot[g_prof.n] = slot;
    cudaEventRecord(g_prof.ev[++g_prof.n], st);
}

}  // namespace t21

using namespace t21;

struct t21_plan {
    int n[3];
    int dtype;
    int device;
    int nzh, pitch;
    int fast[3];          // axis handled by the Stockham kernels
    void* tw[3];          // stage twiddles per axis (z: for nz/2), device
    void* twr;            // exp(-2 pi i k / nz), k = 0..nz/2, device
    void* tw32[3];        // Plan<n, 32> stage twiddles of the warp-per-pencil kernels (fp32, n = 1024), or null
    int max_grid;         // capacity (in CTAs) of the partial-histogram buffers
};

struct WsLayout {
    size_t w[2], offs, part_sum, part_cnt, total;
    int nbins_cap;
};

static WsLayout ws_layout(const t21_plan* p, int nf, int nbins) {
    WsLayout L;
    memset(&L, 0, sizeof(L));
    const size_t esz = (p->dtype == T21_F32 ? 8 : 16);
    const size_t wbytes = align_up((size_t)p->n[0] * p->n[1] * p->pitch * esz, 256);
    size_t o = 0;
    for (int f = 0; f < nf; ++f) { L.w[f] = o; o += wbytes; }
    L.offs = o; o += 256;
    L.nbins_cap = nbins > kSmemHistMaxBins ? 0 : nbins;
    L.part_sum = o; o += align_up((size_t)p->max_grid * L.nbins_cap * sizeof(double), 256);
    L.part_cnt = o; o += align_up((size_t)p->max_grid * L.nbins_cap * sizeof(unsigned), 256);
    L.total = o;
    return L;
}

template <typename T>
static int upload(const std::vector<cx<T>>& v, void** out) {
    T21_CUDA_OK(cudaMalloc(out, v.size() * sizeof(cx<T>)));
    T21_CUDA_OK(cudaMemcpy(*out, v.data(), v.size() * sizeof(cx<T>), cudaMemcpyHostToDevice));
    return 0;
}

template <typename T>
static int build_tables(t21_plan* p) {
    for (int d = 0; d < 3; ++d) {
        if (p->fast[d]) {
            const int len = d == 2 ? p->n[2] / 2 : p->n[d];
            const int emax = d == 2 ? emax_z(len) : (d == 1 ? emax_y(len) : emax_x(len));
            if (int rc = upload<T>(build_stage_twiddles<T>(len, emax), &p->tw[d])) return rc;
        } else if (d < 2) {
            if (int rc = upload<T>(build_roots<T>(p->n[d], p->n[d]), &p->tw[d])) return rc;   // generic pass
        }
    }
    if (sizeof(T) == 4 && use_pencil32())
        for (int d = 0; d < 2; ++d)
            if (p->n[d] == 1024)
                if (int rc = upload<T>(build_stage_twiddles<T>(1024, 32), &p->tw32[d])) return rc;
    if (sizeof(T) == 4 && use_pencil32() && p->n[0] == 512) {
        // the nx = 512 warp-per-pencil kernel: w512^(r q), r = 0..15 (rows), q = 0..31
        std::vector<cx<T>> tw(16 * 32);
        for (int r = 0; r < 16; ++r)
            for (int q = 0; q < 32; ++q) {
                const double ang = -2.0 * 3.14159265358979323846264338327950288 * (double)(r * q) / 512.0;
                tw[r * 32 + q] = cx<T>{(T)cos(ang), (T)sin(ang)};
            }
        if (int rc = upload<T>(tw, &p->tw32[0])) return rc;
    }
    if (p->fast[2]) {
        if (int rc = upload<T>(build_roots<T>(p->n[2], p->n[2] / 2 + 1), &p->twr)) return rc;
    } else {
        if (int rc = upload<T>(build_roots<T>(p->n[2], p->n[2]), &p->twr)) return rc;          // generic pass
    }
    return 0;
}

extern "C" {

int t21_version(void) { return 100; }
const char* t21_last_error(void) { return g_err; }
unsigned long long t21_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int t21_plan_create(t21_plan** out, int nx, int ny, int nz, int dtype, int device) {
    if (!out) { set_error("plan pointer is null"); return T21_EINVAL; }
    *out = nullptr;
    if (nx < 1 || ny < 1 || nz < 1) { set_error("bad shape %dx%dx%d", nx, ny, nz); return T21_EINVAL; }
    if (dtype != T21_F32 && dtype != T21_F64) { set_error("bad dtype %d", dtype); return T21_EINVAL; }
    int cur = 0;
    T21_CUDA_OK(cudaGetDevice(&cur));
    if (cur != device) T21_CUDA_OK(cudaSetDevice(device));
    t21_plan* p = new t21_plan();
    memset(p, 0, sizeof(*p));
    p->n[0] = nx; p->n[1] = ny; p->n[2] = nz;
    p->dtype = dtype;
    p->device = device;
    p->nzh = nz / 2 + 1;
    p->pitch = (int)align_up((size_t)p->nzh, 16);
    p->fast[0] = fast_len(nx);
    p->fast[1] = fast_len(ny);
    p->fast[2] = fast_len(nz);
    p->max_grid = device_sm_count() * 8;
    int rc = dtype == T21_F32 ? build_tables<float>(p) : build_tables<double>(p);
    if (cur != device) cudaSetDevice(cur);
    if (rc) { t21_plan_destroy(p); return rc; }
    *out = p;
    return T21_OK;
}

int t21_plan_destroy(t21_plan* p) {
    if (!p) return T21_OK;
    for (int d = 0; d < 3; ++d) if (p->tw[d]) cudaFree(p->tw[d]);
    for (int d = 0; d < 3; ++d) if (p->tw32[d]) cudaFree(p->tw32[d]);
    if (p->twr) cudaFree(p->twr);
    delete p;
    return T21_OK;
}

size_t t21_workspace_bytes(const t21_plan* p, int n_fields, int nbins) {
    if (!p || n_fields < 0 || n_fields > 2 || nbins < 0) return 0;
    return ws_layout(p, n_fields, nbins).total;
}

// z and y passes of every field; leaves the half spectra in the workspace
static int forward_zy(const t21_plan* p, const void* const x[2], int nf, int subtract_mean, unsigned char* ws,
                      const WsLayout& L, const void* offs[2], cudaStream_t st, const int* window = nullptr) {
    const size_t rsz = p->dtype == T21_F32 ? 4 : 8;
    const long long ntot = (long long)p->n[0] * p->n[1] * p->n[2];
    for (int f = 0; f < nf; ++f) {
        offs[f] = nullptr;
        if (subtract_mean) {
            void* o = ws + L.offs + f * rsz;
            if (int rc = launch_mean_estimate(p->dtype, x[f], ntot, o, st)) return rc;
            offs[f] = o;
            prof_mark(st, SLOT_MEAN);
        }
        // (Running Z and Y slab by slab to keep the half spectrum in L2 between them was measured and
        // is slower: neither pass is DRAM-bandwidth bound and every extra launch costs ~5 us.)
        ZArgs z{x[f], ws + L.w[f], (long long)p->n[0] * p->n[1], p->n[2], p->pitch, p->n[1], (long long)p->n[0], offs[f], p->tw[2], p->twr, window};
        if (int rc = p->fast[2] ? launch_zpass(p->dtype, z, st) : launch_zpass_generic(p->dtype, z, st)) return rc;
        prof_mark(st, SLOT_Z);
        YArgs y{ws + L.w[f], (long long)p->n[0], p->n[1], p->nzh, p->pitch, p->tw[1], nullptr, 0, nullptr, 0, 0, 0};
        if (int rc = p->fast[1] ? launch_ypass(p->dtype, y, st) : launch_ypass_generic(p->dtype, y, st)) return rc;
        prof_mark(st, SLOT_Y);
    }
    return 0;
}

// the current device must be the plan's (the slab entry points take raw buffers, so this is their only guard)
static int check_device(const t21_plan* p) {
    int cur = 0;
    T21_CUDA_OK(cudaGetDevice(&cur));
    if (cur != p->device) { set_error("plan was made for device %d, current device is %d", p->device, cur); return T21_EINVAL; }
    return 0;
}

static int check_common(const t21_plan* p, const void* x1, void* ws, size_t need, size_t have) {
    if (!p || !x1 || !ws) { set_error("null plan, input or workspace"); return T21_EINVAL; }
    if (have < need) { set_error("workspace too small: %zu < %zu bytes", have, need); return T21_ENOMEM; }
    if (((uintptr_t)x1 & 15) || ((uintptr_t)ws & 255)) { set_error("input must be 16-byte and workspace 256-byte aligned"); return T21_EINVAL; }
    int cur = 0;
    T21_CUDA_OK(cudaGetDevice(&cur));
    if (cur != p->device) { set_error("plan was made for device %d, current device is %d", p->device, cur); return T21_EINVAL; }
    return 0;
}

int t21_ps_binned(const t21_plan* p, const void* x1, const void* x2, int subtract_mean, const t21_binspec* spec,
                  double* sums_out, long long* counts_out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!spec || !sums_out || !counts_out) { set_error("null spec or outputs"); return T21_EINVAL; }
    const int nf = x2 ? 2 : 1;
    const int nbins = spec->nk * (spec->nmu > 0 ? spec->nmu : 1);
    if (nbins < 1) { set_error("no bins"); return T21_EINVAL; }
    if (p && (spec->n[0] != p->n[0] || spec->n[1] != p->n[1] || spec->n[2] != p->n[2])) {
        set_error("bin spec shape does not match the plan"); return T21_EINVAL;
    }
    if (x2 && ((uintptr_t)x2 & 15)) { set_error("second input must be 16-byte aligned"); return T21_EINVAL; }
    const WsLayout L = p ? ws_layout(p, nf, nbins) : WsLayout();
    if (int rc = check_common(p, x1, workspace, L.total, workspace_bytes)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* ws = (unsigned char*)workspace;
    const void* x[2] = {x1, x2};
    const void* offs[2] = {nullptr, nullptr};
    prof_begin(st);
    if (int rc = forward_zy(p, x, nf, subtract_mean, ws, L, offs, st)) return rc;

    XArgs a;
    memset(&a, 0, sizeof(a));
    a.w[0] = ws + L.w[0];
    a.w[1] = nf == 2 ? ws + L.w[1] : nullptr;
    a.nf = nf;
    a.nx = p->n[0]; a.ny = p->n[1]; a.nz = p->n[2]; a.nzh = p->nzh; a.pitch = p->pitch; a.xs_log2 = -1;
    a.tw = p->tw[0]; a.tw32 = p->tw32[0];
    a.offset[0] = offs[0]; a.offset[1] = offs[1];
    a.ntot = (double)p->n[0] * p->n[1] * p->n[2];
    a.bins = spec;
    a.nbins = nbins;
    a.hist_global = L.nbins_cap == 0;
    a.part_sum = (double*)(ws + L.part_sum);
    a.part_cnt = (unsigned*)(ws + L.part_cnt);
    a.gsum = sums_out;
    a.gcnt = (unsigned long long*)counts_out;
    a.max_grid = p->max_grid;
    if (!spec->thr && !(p->fast[0] && xpass32_applicable(p->dtype, a))) {
        set_error("a table-only bin spec (no thresholds) can only run on the warp-per-pencil X pass"); return T21_EINVAL;
    }
    if (a.hist_global)
        if (int rc = launch_zero_hist(sums_out, counts_out, nbins, st)) return rc;
    int rc = !p->fast[0] ? launch_xpass_generic(p->dtype, a, st)
                         : launch_xpass_bin(p->dtype, a, st);
    if (rc) return rc;
    prof_mark(st, SLOT_X);
    if (!a.hist_global) {
        rc = launch_reduce_partials(a.part_sum, a.part_cnt, a.grid_used, nbins, sums_out, counts_out, st);
        prof_mark(st, SLOT_REDUCE);
    }
    return rc;
}

// power_spectrum_multipoles (power_legendre.py:4-90) in ONE pipeline: the Z / Y passes, then the warp-per-pencil X pass
// whose epilogue adds, per k bin, sum P, sum P L2, sum P L4, sum L2^2, sum L4^2 and the mode count -- no power array is
// written.  Needs what that kernel needs (fp32, nx = 1024) and a shell spec carrying its pencil segment table, built
// with spec->los_axis / spec->exclude_zero set (t21_build_pencil_segments drops the k_perp = 0 line then); returns
// T21_EINVAL otherwise and the caller takes t21_ps_half + t21_bin_multipoles_half.  sums_out[nk][5], counts_out[nk];
// workspace: t21_workspace_bytes(plan, 1, 5 * nk).
int t21_ps_multipoles(const t21_plan* p, const void* x1, int subtract_mean, const t21_binspec* spec, int los_axis,
                      double* sums_out, long long* counts_out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!spec || !sums_out || !counts_out) { set_error("null spec or outputs"); return T21_EINVAL; }
    if (los_axis < 0 || los_axis > 2 || spec->nmu != 0 || spec->nk < 1) { set_error("bad multipole spec"); return T21_EINVAL; }
    if (p && (spec->n[0] != p->n[0] || spec->n[1] != p->n[1] || spec->n[2] != p->n[2])) {
        set_error("bin spec shape does not match the plan"); return T21_EINVAL;
    }
    const int nk = spec->nk;
    const WsLayout L = p ? ws_layout(p, 1, nk * 5) : WsLayout();
    if (int rc = check_common(p, x1, workspace, L.total, workspace_bytes)) return rc;
    if (!(p->fast[0] && p->fast[1] && p->fast[2]) || L.nbins_cap == 0) { set_error("fused multipoles need power-of-two axes"); return T21_EINVAL; }
    XArgs a;
    memset(&a, 0, sizeof(a));
    unsigned char* ws = (unsigned char*)workspace;
    a.w[0] = ws + L.w[0];
    a.nf = 1;
    a.nx = p->n[0]; a.ny = p->n[1]; a.nz = p->n[2]; a.nzh = p->nzh; a.pitch = p->pitch; a.xs_log2 = -1;
    a.tw = p->tw[0]; a.tw32 = p->tw32[0];
    a.ntot = (double)p->n[0] * p->n[1] * p->n[2];
    a.bins = spec;
    a.nbins = nk * 5;
    a.part_sum = (double*)(ws + L.part_sum);
    a.part_cnt = (unsigned*)(ws + L.part_cnt);
    a.max_grid = p->max_grid;
    if (!xpass32_moments_applicable(p->dtype, a)) { set_error("fused multipoles: fp32, nx = 1024 and a spec with a segment table only"); return T21_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    const void* x[2] = {x1, nullptr};
    const void* offs[2] = {nullptr, nullptr};
    prof_begin(st);
    if (int rc = forward_zy(p, x, 1, subtract_mean, ws, L, offs, st)) return rc;
    a.offset[0] = offs[0];
    if (int rc = launch_xpass32_moments(a, los_axis, st)) return rc;
    prof_mark(st, SLOT_X);
    const int rc = launch_reduce_moments(a.part_sum, a.part_cnt, a.grid_used, nk, sums_out, counts_out, st);
    prof_mark(st, SLOT_REDUCE);
    return rc;
}

// ---- the estimator fed from HOST memory, copy overlapped with the first two passes ----------------------------------------
// A numpy cube (the reference's calling convention) costs one PCIe copy of 4 N^3 bytes, ~78 ms at 1024^3 against 4 ms of
// kernels.  The Z and Y passes work plane by plane along x, so they can run on the x planes that have already landed
// while the rest of the cube is still on the link: the cube is copied in groups of planes on an internal copy stream
// (through the pinned ring for pageable memory, straight for pinned memory), the caller's stream waits for each group's
// event and runs Z and Y on it; only the last group's passes and the X pass are left after the copy.
namespace t21 {
struct HostFeed {
    cudaStream_t copy = nullptr;
    std::vector<cudaEvent_t> ev;
    cudaEvent_t start = nullptr;
    std::mutex mu;            // one feed at a time per device (the copy stream and its events are shared)
};
static HostFeed g_feed[64];
}  // namespace t21

int t21_ps_binned_host(const t21_plan* p, const void* x_host, int host_is_pinned, int subtract_mean, const t21_binspec* spec,
                       double* sums_out, long long* counts_out, void* x_dev, void* workspace, size_t workspace_bytes,
                       int threads, void* stream) {
    if (!spec || !sums_out || !counts_out || !x_host || !x_dev) { set_error("null argument"); return T21_EINVAL; }
    const int nbins = spec->nk * (spec->nmu > 0 ? spec->nmu : 1);
    if (nbins < 1) { set_error("no bins"); return T21_EINVAL; }
    if (p && (spec->n[0] != p->n[0] || spec->n[1] != p->n[1] || spec->n[2] != p->n[2])) {
        set_error("bin spec shape does not match the plan"); return T21_EINVAL;
    }
    const WsLayout L = p ? ws_layout(p, 1, nbins) : WsLayout();
    if (int rc = check_common(p, x_dev, workspace, L.total, workspace_bytes)) return rc;
    if (!(p->fast[0] && p->fast[1] && p->fast[2]) || p->device < 0 || p->device >= 64) {
        set_error("the host-fed estimator needs power-of-two axes"); return T21_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* ws = (unsigned char*)workspace;
    const size_t rsz = p->dtype == T21_F32 ? 4 : 8, esz = 2 * rsz, lw = 64 / esz;
    const size_t plane = (size_t)p->n[1] * p->n[2] * rsz;
    static const size_t group_bytes = [] {                                    // ~128 MB per group; T21_HOST_FEED_GROUP_MB for tests
        const char* e = getenv("T21_HOST_FEED_GROUP_MB");
        const long mb = e ? atol(e) : 0;
        return (size_t)(mb >= 1 && mb <= 4096 ? mb : 128) << 20;
    }();
    int gp = (int)std::max<size_t>(1, group_bytes / plane);                    // x planes per group
    if (gp > p->n[0]) gp = p->n[0];
    const int ngroups = (p->n[0] + gp - 1) / gp;

    HostFeed& F = g_feed[p->device];
    std::lock_guard<std::mutex> lock(F.mu);
    if (!F.copy) {
        T21_CUDA_OK(cudaStreamCreateWithFlags(&F.copy, cudaStreamNonBlocking));
        T21_CUDA_OK(cudaEventCreateWithFlags(&F.start, cudaEventDisableTiming));
    }
    while ((int)F.ev.size() < ngroups) {
        cudaEvent_t e;
        T21_CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        F.ev.push_back(e);
    }
    // the copy may only start once the caller's earlier work on `stream` (which may still use x_dev) is done
    T21_CUDA_OK(cudaEventRecord(F.start, st));
    T21_CUDA_OK(cudaStreamWaitEvent(F.copy, F.start, 0));
    prof_begin(st);
    const void* off = nullptr;
    for (int g = 0; g < ngroups; ++g) {
        const int x0 = g * gp, np = std::min(gp, p->n[0] - x0);
        unsigned char* dst = (unsigned char*)x_dev + (size_t)x0 * plane;
        const unsigned char* src = (const unsigned char*)x_host + (size_t)x0 * plane;
        if (host_is_pinned) T21_CUDA_OK(cudaMemcpyAsync(dst, src, (size_t)np * plane, cudaMemcpyHostToDevice, F.copy));
        else if (int rc = t21_h2d_staged(dst, src, (size_t)np * plane, threads, F.copy)) return rc;
        T21_CUDA_OK(cudaEventRecord(F.ev[g], F.copy));
        T21_CUDA_OK(cudaStreamWaitEvent(st, F.ev[g], 0));
        if (g == 0 && subtract_mean) {
            // c only has to be close to the mean (the DC sample is repaired exactly): the first group's sample will do
            void* o = ws + L.offs;
            if (int rc = launch_mean_estimate(p->dtype, dst, (long long)np * p->n[1] * p->n[2], o, st)) return rc;
            off = o;
        }
        unsigned char* wg = ws + L.w[0] + (size_t)x0 * p->n[1] * lw * esz;
        ZArgs z{dst, wg, (long long)np * p->n[1], p->n[2], p->pitch, p->n[1], (long long)p->n[0], off, p->tw[2], p->twr, nullptr};
        if (int rc = launch_zpass(p->dtype, z, st)) return rc;
        YArgs y{wg, (long long)np, p->n[1], p->nzh, p->pitch, p->tw[1], nullptr, 0, nullptr, 0, 0, 0, (long long)p->n[0]};
        if (int rc = launch_ypass(p->dtype, y, st)) return rc;
    }
    prof_mark(st, SLOT_Z);
    XArgs a;
    memset(&a, 0, sizeof(a));
    a.w[0] = ws + L.w[0];
    a.nf = 1;
    a.nx = p->n[0]; a.ny = p->n[1]; a.nz = p->n[2]; a.nzh = p->nzh; a.pitch = p->pitch; a.xs_log2 = -1;
    a.tw = p->tw[0]; a.tw32 = p->tw32[0];
    a.offset[0] = off;
    a.ntot = (double)p->n[0] * p->n[1] * p->n[2];
    a.bins = spec;
    a.nbins = nbins;
    a.hist_global = L.nbins_cap == 0;
    a.part_sum = (double*)(ws + L.part_sum);
    a.part_cnt = (unsigned*)(ws + L.part_cnt);
    a.gsum = sums_out;
    a.gcnt = (unsigned long long*)counts_out;
    a.max_grid = p->max_grid;
    if (!spec->thr && !xpass32_applicable(p->dtype, a)) {
        set_error("a table-only bin spec (no thresholds) can only run on the warp-per-pencil X pass"); return T21_EINVAL;
    }
    if (a.hist_global)
        if (int rc = launch_zero_hist(sums_out, counts_out, nbins, st)) return rc;
    int rc = launch_xpass_bin(p->dtype, a, st);
    if (rc) return rc;
    prof_mark(st, SLOT_X);
    if (!a.hist_global) {
        rc = launch_reduce_partials(a.part_sum, a.part_cnt, a.grid_used, nbins, sums_out, counts_out, st);
        prof_mark(st, SLOT_REDUCE);
    }
    return rc;
}

static int binned_windowed(const t21_plan* p, const void* x1, const int* window, const t21_binspec* spec, double* sums_out,
                           long long* counts_out, unsigned char* ws, const WsLayout& L, cudaStream_t st);

// power_spectrum_1d of `nwindows` cubically masked copies of ONE cube (position_dependent_power_spectrum.py:65-77:
// `c1 = cube1 * _W_L(...)` then power_spectrum_1d(c1), once per sub-box): the mask is fused into the Z pass's load, so
// no masked cube is ever written and only the window's samples are read; the launches of all windows are queued by
// this one call.  windows[i] = {x0, x1, y0, y1, z0, z1} (host array); sums_out[nwindows][nbins]; the mode counts do
// not depend on the window and are written once.  The FFT sees the raw samples (no mean offset: a masked cube's mean
// is dominated by the mask).
int t21_ps_binned_windows(const t21_plan* p, const void* x1, const int* windows, int nwindows, const t21_binspec* spec,
                          double* sums_out, long long* counts_out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!spec || !sums_out || !counts_out || !windows || nwindows < 1) { set_error("null or empty argument"); return T21_EINVAL; }
    if (!p || !(p->fast[0] && p->fast[1] && p->fast[2])) { set_error("windowed batches need power-of-two axes"); return T21_EINVAL; }
    const int nbins = spec->nk * (spec->nmu > 0 ? spec->nmu : 1);
    if (nbins < 1 || nbins > kSmemHistMaxBins) { set_error("windowed batches take 1..%d bins", kSmemHistMaxBins); return T21_EINVAL; }
    if (spec->n[0] != p->n[0] || spec->n[1] != p->n[1] || spec->n[2] != p->n[2]) { set_error("bin spec shape does not match the plan"); return T21_EINVAL; }
    const WsLayout L = ws_layout(p, 1, nbins);
    if (int rc = check_common(p, x1, workspace, L.total, workspace_bytes)) return rc;
    for (int i = 0; i < nwindows; ++i)
        for (int d = 0; d < 3; ++d)
            if (windows[6 * i + 2 * d] < 0 || windows[6 * i + 2 * d + 1] > p->n[d] || windows[6 * i + 2 * d] > windows[6 * i + 2 * d + 1]) {
                set_error("window %d is outside the cube", i); return T21_EINVAL;
            }
    cudaStream_t st = (cudaStream_t)stream;
    for (int i = 0; i < nwindows; ++i)
        if (int rc = binned_windowed(p, x1, windows + 6 * i, spec, sums_out + (size_t)i * nbins, counts_out, (unsigned char*)workspace, L, st))
            return rc;
    return T21_OK;
}

static int binned_windowed(const t21_plan* p, const void* x1, const int* window, const t21_binspec* spec, double* sums_out,
                           long long* counts_out, unsigned char* ws, const WsLayout& L, cudaStream_t st) {
    const void* x[2] = {x1, nullptr};
    const void* offs[2] = {nullptr, nullptr};
    if (int rc = forward_zy(p, x, 1, 0, ws, L, offs, st, window)) return rc;
    const int nbins = spec->nk * (spec->nmu > 0 ? spec->nmu : 1);
    XArgs a;
    memset(&a, 0, sizeof(a));
    a.w[0] = ws + L.w[0];
    a.nf = 1;
    a.nx = p->n[0]; a.ny = p->n[1]; a.nz = p->n[2]; a.nzh = p->nzh; a.pitch = p->pitch; a.xs_log2 = -1;
    a.tw = p->tw[0]; a.tw32 = p->tw32[0];
    a.ntot = (double)p->n[0] * p->n[1] * p->n[2];
    a.bins = spec;
    a.nbins = nbins;
    a.part_sum = (double*)(ws + L.part_sum);
    a.part_cnt = (unsigned*)(ws + L.part_cnt);
    a.gsum = sums_out;
    a.gcnt = (unsigned long long*)counts_out;
    a.max_grid = p->max_grid;
    if (int rc = launch_xpass_bin(p->dtype, a, st)) return rc;
    return launch_reduce_partials(a.part_sum, a.part_cnt, a.grid_used, nbins, sums_out, counts_out, st);
}

static int ps_nd_impl(const t21_plan* p, const void* x1, const void* x2, int subtract_mean, double scale, void* out,
                      int out_dtype, int half, void* workspace, size_t workspace_bytes, void* stream);

int t21_ps_nd(const t21_plan* p, const void* x1, const void* x2, int subtract_mean, double scale, void* out,
              int out_dtype, void* workspace, size_t workspace_bytes, void* stream) {
    return ps_nd_impl(p, x1, x2, subtract_mean, scale, out, out_dtype, 0, workspace, workspace_bytes, stream);
}

// The scaled power of the computed half spectrum only, FFT order: out[nx][ny][nz/2+1].  It is what the estimators
// that bin on something other than (k, mu) boxes read back (t21_bin_cyl_half, t21_bin_multipoles_half): a quarter
// of the bytes of the full fftshifted float64 array, and no Hermitian-twin stores.  Rows are t21_half_pitch(plan) long.
int t21_ps_half(const t21_plan* p, const void* x1, const void* x2, int subtract_mean, double scale, void* out_half,
                int out_dtype, void* workspace, size_t workspace_bytes, void* stream) {
    return ps_nd_impl(p, x1, x2, subtract_mean, scale, out_half, out_dtype, 1, workspace, workspace_bytes, stream);
}

// The scaled power of the computed half spectrum PENCIL-MAJOR, out_t[ny][nz/2+1][nx]: what the warp-per-pencil X pass can
// write with full 128-byte runs (fp32 transform, nx = 1024, one field; T21_EINVAL otherwise).  t21_nd_assemble_t turns
// it into the fftshifted cube of power_spectrum_nd (power_spectrum.py:45-54).
int t21_ps_half_t(const t21_plan* p, const void* x1, int subtract_mean, double scale, void* out_t, int out_dtype,
                  void* workspace, size_t workspace_bytes, void* stream) {
    return ps_nd_impl(p, x1, nullptr, subtract_mean, scale, out_t, out_dtype, -1, workspace, workspace_bytes, stream);
}

static int ps_nd_impl(const t21_plan* p, const void* x1, const void* x2, int subtract_mean, double scale, void* out,
                      int out_dtype, int half, void* workspace, size_t workspace_bytes, void* stream) {
    if (!out) { set_error("null output"); return T21_EINVAL; }
    if (out_dtype != T21_F32 && out_dtype != T21_F64) { set_error("bad output dtype"); return T21_EINVAL; }
    const int nf = x2 ? 2 : 1;
    if (x2 && ((uintptr_t)x2 & 15)) { set_error("second input must be 16-byte aligned"); return T21_EINVAL; }
    const WsLayout L = p ? ws_layout(p, nf, 0) : WsLayout();
    if (int rc = check_common(p, x1, workspace, L.total, workspace_bytes)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* ws = (unsigned char*)workspace;
    const void* x[2] = {x1, x2};
    const void* offs[2] = {nullptr, nullptr};
    prof_begin(st);
    if (int rc = forward_zy(p, x, nf, subtract_mean, ws, L, offs, st)) return rc;
    XArgs a;
    memset(&a, 0, sizeof(a));
    a.w[0] = ws + L.w[0];
    a.w[1] = nf == 2 ? ws + L.w[1] : nullptr;
    a.nf = nf;
    a.nx = p->n[0]; a.ny = p->n[1]; a.nz = p->n[2]; a.nzh = p->nzh; a.pitch = p->pitch; a.xs_log2 = -1;
    a.tw = p->tw[0]; a.tw32 = p->tw32[0];
    a.offset[0] = offs[0]; a.offset[1] = offs[1];
    a.ntot = (double)p->n[0] * p->n[1] * p->n[2];
    a.nd_out = out; a.nd_f64 = out_dtype == T21_F64; a.scale = scale; a.nd_half = half > 0 ? t21_half_pitch(p) : (half < 0 ? -1 : 0);
    a.max_grid = p->max_grid;
    if (half < 0 && !(p->fast[0] && xpass32_store_applicable(p->dtype, a))) {
        set_error("the pencil-major half spectrum needs the warp-per-pencil X pass (fp32, nx = 1024, one field)"); return T21_EINVAL;
    }
    const int rc = !p->fast[0] ? launch_xpass_generic(p->dtype, a, st)
                               : (xpass32_store_applicable(p->dtype, a) ? launch_xpass32_store(a, st)
                               : (p->dtype == T21_F32 ? launch_xpass_f32_nd(a, st) : launch_xpass_f64_nd(a, st)));
    prof_mark(st, SLOT_X);
    return rc;
}

int t21_build_pencil_segments(const t21_binspec* spec, uint32_t* out, int stride, int* max_segments_dev, void* stream) {
    if (!spec || !out || !max_segments_dev) { set_error("null argument to t21_build_pencil_segments"); return T21_EINVAL; }
    if (stride != 8 && stride != 16 && stride != 32) { set_error("segment stride must be 8, 16 or 32"); return T21_EINVAL; }
    if (spec->zero_idx[0] != 0 || spec->n[0] < 2 || spec->n[0] >= 65535) { set_error("segment tables need a spec in FFT index order"); return T21_EINVAL; }
    if (!spec->thr || !spec->k_lut || !spec->tab_s[0] || !spec->tab_s[1] || !spec->tab_s[2] || !spec->tab_sf[0] || !spec->tab_sf[1] || !spec->tab_sf[2]) {
        set_error("segment tables are built from a complete bin spec (thresholds, look-up and k tables)"); return T21_EINVAL;
    }
    if (spec->nk * (spec->nmu > 0 ? spec->nmu : 1) >= 65535) { set_error("too many bins for a segment table"); return T21_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    T21_CUDA_OK(cudaMemsetAsync(max_segments_dev, 0, 2 * sizeof(int), st));
    const bool is_signed = spec->nmu > 0 && !spec->absolute_mus;
    if (!is_signed) return launch_pencil_segments(*spec, (spec->nmu == 0 && spec->exclude_zero) ? 5 : 0, out, stride, max_segments_dev, st);
    const size_t per_table = (size_t)spec->n[1] * (spec->n[2] / 2 + 1) * stride;
    for (int w = 1; w <= 4; ++w)
        if (int rc = launch_pencil_segments(*spec, w, out + (size_t)(w - 1) * per_table, stride, max_segments_dev, st)) return rc;
    return T21_OK;
}

int t21_build_pencil_segments_cyl(int n0, int n1, int n2, int nu_axis, const int* perp_tab, const int* par_tab, int nperp, int npar,
                                  uint32_t* out, int stride, int* max_segments_dev, void* stream) {
    if (!perp_tab || !par_tab || !out || !max_segments_dev) { set_error("null argument to t21_build_pencil_segments_cyl"); return T21_EINVAL; }
    if (stride != 8 && stride != 16 && stride != 32) { set_error("segment stride must be 8, 16 or 32"); return T21_EINVAL; }
    if (n0 < 2 || n0 >= 65535 || n1 < 1 || n2 < 1 || nu_axis < 0 || nu_axis > 2 || nperp < 1 || npar < 1 || (long long)nperp * npar >= 65535) {
        set_error("bad argument to t21_build_pencil_segments_cyl"); return T21_EINVAL;
    }
    cudaStream_t st = (cudaStream_t)stream;
    T21_CUDA_OK(cudaMemsetAsync(max_segments_dev, 0, 2 * sizeof(int), st));
    return launch_pencil_segments_cyl(n0, n1, n2, nu_axis, perp_tab, par_tab, npar, out, stride, max_segments_dev, st);
}

int t21_profile(int enable) {
    g_prof.on = enable != 0;
    return T21_OK;
}

// ms_out[0..4] = milliseconds spent in {mean estimate, z pass, y pass, x pass, reduce} during the
// last profiled call (summed over fields).  Blocks until that call has finished.
int t21_profile_last(float* ms_out, int cap) {
    if (!ms_out || cap < SLOT_COUNT) { set_error("need room for %d slots", (int)SLOT_COUNT); return T21_EINVAL; }
    for (int i = 0; i < cap; ++i) ms_out[i] = 0.f;
    if (!g_prof.made || g_prof.n == 0) return 0;
    T21_CUDA_OK(cudaEventSynchronize(g_prof.ev[g_prof.n]));
    for (int i = 0; i < g_prof.n; ++i) {
        float ms = 0.f;
        T21_CUDA_OK(cudaEventElapsedTime(&ms, g_prof.ev[i], g_prof.ev[i + 1]));
        ms_out[g_prof.slot[i]] += ms;
    }
    return SLOT_COUNT;
}

// ---- one cube sharded over P GPUs as x slabs --------------------------------------
// stage 1 (local):  Z and Y passes of this rank's x slab [nx/P][ny][nz] -> w_out, chunk-major (nx/P planes of ny rows)
// exchange (caller): all-to-all over NCCL re-partitions the half spectrum into y slabs (nx planes of ny/P rows)
// stage 2 (local):  X pass + binning of this rank's y slab; the caller all-reduces sums and counts.
int t21_mean_estimate(const void* x, long long n, int dtype, void* out_dev, void* stream) {
    if (!x || !out_dev || n < 1 || (dtype != T21_F32 && dtype != T21_F64)) { set_error("bad argument to t21_mean_estimate"); return T21_EINVAL; }
    return launch_mean_estimate(dtype, x, n, out_dev, (cudaStream_t)stream);
}

int t21_slab_pitch(const t21_plan* p) { return p ? p->pitch : 0; }
// row length (elements) of the half-spectrum power arrays: nz/2+1 rounded up to 8, so that the 8-column runs the
// X pass stores are whole 32-byte sectors (an odd row length made every run straddle two: +100 % DRAM traffic)
int t21_half_pitch(const t21_plan* p) { return p ? (p->nzh + 7) / 8 * 8 : 0; }

int t21_slab_stage1(const t21_plan* p, const void* x_slab, const void* offset_or_null, void* w_out, void* send_out,
                    int nranks, void* stream) {
    if (!p || !x_slab || !w_out) { set_error("null argument"); return T21_EINVAL; }
    if (int rc = check_device(p)) return rc;
    if (!(p->fast[1] && p->fast[2])) { set_error("slab mode needs power-of-two y and z axes"); return T21_EINVAL; }
    if (((uintptr_t)x_slab & 15) || ((uintptr_t)w_out & 15) || ((uintptr_t)send_out & 15)) {
        set_error("buffers must be 16-byte aligned"); return T21_EINVAL;
    }
    int nyl_log2 = 0;
    if (send_out) {
        if (nranks < 1 || (nranks & (nranks - 1)) || p->n[1] % nranks) {
            set_error("rank count %d must be a power of two dividing ny = %d", nranks, p->n[1]); return T21_EINVAL;
        }
        while ((p->n[1] / nranks) >> (nyl_log2 + 1)) ++nyl_log2;
    }
    cudaStream_t st = (cudaStream_t)stream;
    prof_begin(st);
    ZArgs z{x_slab, w_out, (long long)p->n[0] * p->n[1], p->n[2], p->pitch, p->n[1], (long long)p->n[0], offset_or_null, p->tw[2], p->twr, nullptr};
    if (int rc = launch_zpass(p->dtype, z, st)) return rc;
    prof_mark(st, SLOT_Z);
    YArgs y{w_out, (long long)p->n[0], p->n[1], p->nzh, p->pitch, p->tw[1], send_out, nyl_log2, nullptr, 0, 0, 0};
    if (int rc = launch_ypass(p->dtype, y, st)) return rc;
    prof_mark(st, SLOT_Y);
    return T21_OK;
}

// stage 1 with the exchange fused in: the Y pass stores every transformed row straight into the y-slab
// buffer of the rank that owns it (peer_cols[s] = rank s's chunk-major y-slab buffer, nx planes of ny/P rows, peer-mapped over NVLink),
// so there is no send buffer and no collective; x0_global = index of this slab's first x plane in the cube.
// The caller orders "all ranks finished stage 1" before stage 2 (a tiny all-reduce on the same stream).
int t21_slab_stage1_p2p(const t21_plan* p, const void* x_slab, const void* offset_or_null, void* w_scratch,
                        void* const* peer_cols, int nranks, long long x0_global, void* stream) {
    if (!p || !x_slab || !w_scratch || !peer_cols) { set_error("null argument"); return T21_EINVAL; }
    if (int rc = check_device(p)) return rc;
    if (((uintptr_t)x_slab & 15) || ((uintptr_t)w_scratch & 15)) { set_error("buffers must be 16-byte aligned"); return T21_EINVAL; }
    if (!(p->fast[1] && p->fast[2])) { set_error("slab mode needs power-of-two y and z axes"); return T21_EINVAL; }
    if (nranks < 1 || nranks > 16 || (nranks & (nranks - 1)) || p->n[1] % nranks) {
        set_error("rank count %d must be a power of two <= 16 dividing ny = %d", nranks, p->n[1]); return T21_EINVAL;
    }
    for (int r = 0; r < nranks; ++r)
        if (!peer_cols[r] || ((uintptr_t)peer_cols[r] & 15)) { set_error("bad peer buffer %d", r); return T21_EINVAL; }
    int nyl_log2 = 0;
    while ((p->n[1] / nranks) >> (nyl_log2 + 1)) ++nyl_log2;
    cudaStream_t st = (cudaStream_t)stream;
    prof_begin(st);
    ZArgs z{x_slab, w_scratch, (long long)p->n[0] * p->n[1], p->n[2], p->pitch, p->n[1], (long long)p->n[0], offset_or_null, p->tw[2], p->twr, nullptr};
    if (int rc = launch_zpass(p->dtype, z, st)) return rc;
    prof_mark(st, SLOT_Z);
    YArgs y{w_scratch, (long long)p->n[0], p->n[1], p->nzh, p->pitch, p->tw[1], nullptr, nyl_log2, peer_cols, nranks, x0_global,
            (long long)p->n[0] * nranks};
    if (int rc = launch_ypass(p->dtype, y, st)) return rc;
    prof_mark(st, SLOT_Y);
    return T21_OK;
}

// Buffers other processes store into (the fused exchange).  They are plain cudaMalloc allocations so that
// their CUDA IPC handle refers to the buffer itself (offset 0); a peer opens the handle with ITS device
// current, which maps the memory for that device's kernels and enables peer access over NVLink.
int t21_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
    if (!ptr || !handle64 || bytes == 0) { set_error("null argument"); return T21_EINVAL; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    T21_CUDA_OK(cudaMalloc(ptr, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, *ptr);
    if (e != cudaSuccess) { cudaFree(*ptr); *ptr = nullptr; return cuda_fail(e, "cudaIpcGetMemHandle"); }
    memcpy(handle64, &h, 64);
    return T21_OK;
}
int t21_peer_open(const unsigned char* handle64, void** ptr) {
    if (!ptr || !handle64) { set_error("null argument"); return T21_EINVAL; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    T21_CUDA_OK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return T21_OK;
}
int t21_peer_close(void* ptr) {
    if (ptr) T21_CUDA_OK(cudaIpcCloseMemHandle(ptr));
    return T21_OK;
}
int t21_peer_free(void* ptr) {
    if (ptr) T21_CUDA_OK(cudaFree(ptr));
    return T21_OK;
}

// p: plan of shape (nx, ny/P, nz); w1/w2: this rank's y slab(s), chunk-major, as n_xslabs consecutive x slabs
// (1 for the fused exchange, P for buffers received from an NCCL all-to-all); iy0: global index of
// local row 0; spec: bin spec of the GLOBAL cube; offsets: the common mean offsets used in stage 1.
int t21_slab_stage2(const t21_plan* p, const void* w1, const void* w2, int n_xslabs, int iy0, int ny_global,
                    const void* offset1, const void* offset2, const t21_binspec* spec, double* sums_out, long long* counts_out,
                    void* workspace, size_t workspace_bytes, void* stream) {
    if (!p || !w1 || !spec || !sums_out || !counts_out || !workspace) { set_error("null argument"); return T21_EINVAL; }
    if (int rc = check_device(p)) return rc;
    if (((uintptr_t)w1 & 63) || ((uintptr_t)w2 & 63) || ((uintptr_t)workspace & 255)) {
        set_error("y slabs must be 64-byte and the workspace 256-byte aligned"); return T21_EINVAL;
    }
    if (!p->fast[0]) { set_error("slab mode needs a power-of-two x axis"); return T21_EINVAL; }
    if (spec->n[0] != p->n[0] || spec->n[2] != p->n[2] || spec->n[1] != ny_global || iy0 < 0 || iy0 + p->n[1] > ny_global) {
        set_error("bin spec / slab geometry mismatch"); return T21_EINVAL;
    }
    if (n_xslabs < 1 || (n_xslabs & (n_xslabs - 1)) || p->n[0] % n_xslabs) { set_error("bad x slab count %d", n_xslabs); return T21_EINVAL; }
    int xs_log2 = 0;
    while ((p->n[0] / n_xslabs) >> (xs_log2 + 1)) ++xs_log2;
    const int nf = w2 ? 2 : 1;
    const int nbins = spec->nk * (spec->nmu > 0 ? spec->nmu : 1);
    WsLayout L = ws_layout(p, 0, nbins);
    if (workspace_bytes < L.total) { set_error("workspace too small: %zu < %zu bytes", workspace_bytes, L.total); return T21_ENOMEM; }
    cudaStream_t st = (cudaStream_t)stream;
    unsigned char* ws = (unsigned char*)workspace;
    XArgs a;
    memset(&a, 0, sizeof(a));
    a.w[0] = w1; a.w[1] = w2; a.nf = nf;
    a.nx = p->n[0]; a.ny = p->n[1]; a.nz = p->n[2]; a.nzh = p->nzh; a.pitch = p->pitch; a.xs_log2 = xs_log2; a.iy0 = iy0;
    a.tw = p->tw[0]; a.tw32 = p->tw32[0];
    a.offset[0] = offset1; a.offset[1] = offset2;
    a.ntot = (double)p->n[0] * ny_global * p->n[2];
    a.bins = spec; a.nbins = nbins;
    a.hist_global = L.nbins_cap == 0;
    a.part_sum = (double*)(ws + L.part_sum);
    a.part_cnt = (unsigned*)(ws + L.part_cnt);
    a.gsum = sums_out; a.gcnt = (unsigned long long*)counts_out;
    a.max_grid = p->max_grid;
    if (!g_prof.on || g_prof.n == 0) prof_begin(st);
    if (a.hist_global)
        if (int rc = launch_zero_hist(sums_out, counts_out, nbins, st)) return rc;
    int rc = launch_xpass_bin(p->dtype, a, st);
    if (rc) return rc;
    prof_mark(st, SLOT_X);
    if (!a.hist_global) {
        rc = launch_reduce_partials(a.part_sum, a.part_cnt, a.grid_used, nbins, sums_out, counts_out, st);
        prof_mark(st, SLOT_REDUCE);
    }
    return rc;
}

// n-D output of a slab-sharded cube: the scaled power of this rank's rows, HALF spectrum only,
// FFT order: out[nx][ny/P][nz/2+1].  The caller assembles the fftshifted rows (the kz < 0 half of a
// row is the mirrored half of row -ky, which another rank holds; tools21cm_b200/distributed.py).
int t21_slab_stage2_nd(const t21_plan* p, const void* w1, const void* w2, int n_xslabs, int iy0, int ny_global,
                       const void* offset1, const void* offset2, double scale, void* out_half, int out_dtype,
                       void* stream) {
    if (!p || !w1 || !out_half) { set_error("null argument"); return T21_EINVAL; }
    if (int rc = check_device(p)) return rc;
    if (n_xslabs < 1 || (n_xslabs & (n_xslabs - 1)) || p->n[0] % n_xslabs) { set_error("bad x slab count %d", n_xslabs); return T21_EINVAL; }
    int xs_log2 = 0;
    while ((p->n[0] / n_xslabs) >> (xs_log2 + 1)) ++xs_log2;
    if (!p->fast[0]) { set_error("slab mode needs a power-of-two x axis"); return T21_EINVAL; }
    if (out_dtype != T21_F32 && out_dtype != T21_F64) { set_error("bad output dtype"); return T21_EINVAL; }
    if (iy0 < 0 || iy0 + p->n[1] > ny_global) { set_error("slab geometry mismatch"); return T21_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    XArgs a;
    memset(&a, 0, sizeof(a));
    a.w[0] = w1; a.w[1] = w2; a.nf = w2 ? 2 : 1;
    a.nx = p->n[0]; a.ny = p->n[1]; a.nz = p->n[2]; a.nzh = p->nzh; a.pitch = p->pitch; a.xs_log2 = xs_log2; a.iy0 = iy0;
    a.tw = p->tw[0]; a.tw32 = p->tw32[0];
    a.offset[0] = offset1; a.offset[1] = offset2;
    a.ntot = (double)p->n[0] * ny_global * p->n[2];
    a.nd_out = out_half; a.nd_f64 = out_dtype == T21_F64; a.nd_half = t21_half_pitch(p); a.scale = scale;
    a.max_grid = p->max_grid;
    const int rc = xpass32_store_applicable(p->dtype, a) ? launch_xpass32_store(a, st)
                 : (p->dtype == T21_F32 ? launch_xpass_f32_nd(a, st) : launch_xpass_f64_nd(a, st));
    prof_mark(st, SLOT_X);
    return rc;
}

size_t t21_bin_only_workspace_bytes(const t21_binspec* spec) {
    if (!spec) return 0;
    const int nbins = spec->nk * (spec->nmu > 0 ? spec->nmu : 1);
    if (nbins > kSmemHistMaxBins) return 256;
    const size_t g = (size_t)device_sm_count() * 8;
    return align_up(g * nbins * sizeof(double), 256) + align_up(g * nbins * sizeof(unsigned), 256);
}

int t21_bin_only(const void* p_shifted, int dtype, const t21_binspec* spec, double* sums_out, long long* counts_out,
                 void* workspace, size_t workspace_bytes, void* stream) {
    if (!p_shifted || !spec || !sums_out || !counts_out || !workspace) { set_error("null argument"); return T21_EINVAL; }
    if (dtype != T21_F32 && dtype != T21_F64) { set_error("bad dtype %d", dtype); return T21_EINVAL; }
    const int nbins = spec->nk * (spec->nmu > 0 ? spec->nmu : 1);
    if (nbins < 1) { set_error("no bins"); return T21_EINVAL; }
    const size_t need = t21_bin_only_workspace_bytes(spec);
    if (workspace_bytes < need) { set_error("workspace too small: %zu < %zu bytes", workspace_bytes, need); return T21_ENOMEM; }
    cudaStream_t st = (cudaStream_t)stream;
    const int max_grid = device_sm_count() * 8;
    double* part_sum = (double*)workspace;
    unsigned* part_cnt = (unsigned*)((unsigned char*)workspace + align_up((size_t)max_grid * nbins * sizeof(double), 256));
    const bool global = nbins > kSmemHistMaxBins;
    if (global)
        if (int rc = launch_zero_hist(sums_out, counts_out, nbins, st)) return rc;
    int grid = 0;
    if (int rc = launch_bin_only(p_shifted, dtype, *spec, part_sum, part_cnt, max_grid, sums_out, counts_out, &grid, st))
        return rc;
    if (!global) return launch_reduce_partials(part_sum, part_cnt, grid, nbins, sums_out, counts_out, st);
    return T21_OK;
}

size_t t21_bin_cyl_workspace_bytes(int nperp, int npar) {
    const long long nbins = (long long)nperp * npar;
    if (nbins < 1) return 0;
    if (nbins > kSmemHistMaxBins) return 256;
    const size_t g = (size_t)device_sm_count() * 8;
    return align_up(g * nbins * sizeof(double), 256) + align_up(g * nbins * sizeof(unsigned), 256);
}

int t21_bin_cyl(const void* p_shifted, int dtype, int n0, int n1, int n2, int nu_axis, const int* perp_tab,
                const int* par_tab, int nperp, int npar, double* sums_out, long long* counts_out, void* workspace,
                size_t workspace_bytes, void* stream) {
    if (!p_shifted || !perp_tab || !par_tab || !sums_out || !counts_out || !workspace) { set_error("null argument"); return T21_EINVAL; }
    if (dtype != T21_F32 && dtype != T21_F64) { set_error("bad dtype %d", dtype); return T21_EINVAL; }
    if (nu_axis < 0 || nu_axis > 2 || n0 < 1 || n1 < 1 || n2 < 1 || nperp < 1 || npar < 1) { set_error("bad geometry"); return T21_EINVAL; }
    const int nbins = nperp * npar;
    const size_t need = t21_bin_cyl_workspace_bytes(nperp, npar);
    if (workspace_bytes < need) { set_error("workspace too small: %zu < %zu bytes", workspace_bytes, need); return T21_ENOMEM; }
    cudaStream_t st = (cudaStream_t)stream;
    const int max_grid = device_sm_count() * 8;
    double* part_sum = (double*)workspace;
    unsigned* part_cnt = (unsigned*)((unsigned char*)workspace + align_up((size_t)max_grid * nbins * sizeof(double), 256));
    const bool global = nbins > kSmemHistMaxBins;
    if (global)
        if (int rc = launch_zero_hist(sums_out, counts_out, nbins, st)) return rc;
    int grid = 0;
    if (int rc = launch_bin_cyl(p_shifted, dtype, n0, n1, n2, nu_axis, perp_tab, par_tab, nperp, npar, part_sum, part_cnt,
                                max_grid, sums_out, counts_out, &grid, st))
        return rc;
    if (!global) return launch_reduce_partials(part_sum, part_cnt, grid, nbins, sums_out, counts_out, st);
    return T21_OK;
}

// the same SUMS from the computed half spectrum of t21_ps_half ([nx][ny][row_pitch], FFT order; tables in FFT order
// over the full axes).  Mode counts of separable bins are formed on the host from the two tables.
int t21_bin_cyl_half(const void* p_half, int dtype, int nx, int ny, int nz, int row_pitch, int nu_axis, const int* perp_tab,
                     const int* par_tab, int nperp, int npar, double* sums_out, void* workspace, size_t workspace_bytes,
                     void* stream) {
    if (!p_half || !perp_tab || !par_tab || !sums_out || !workspace) { set_error("null argument"); return T21_EINVAL; }
    if (dtype != T21_F32 && dtype != T21_F64) { set_error("bad dtype %d", dtype); return T21_EINVAL; }
    if (nu_axis < 0 || nu_axis > 2 || nx < 1 || ny < 1 || nz < 1 || nperp < 1 || npar < 1 || row_pitch < nz / 2 + 1) {
        set_error("bad geometry"); return T21_EINVAL;
    }
    const size_t need = t21_bin_cyl_workspace_bytes(nperp, npar);
    if (workspace_bytes < need) { set_error("workspace too small: %zu < %zu bytes", workspace_bytes, need); return T21_ENOMEM; }
    return launch_cyl_rows(p_half, dtype, nx, ny, nz, row_pitch, nu_axis, perp_tab, par_tab, nperp, npar, (double*)workspace,
                           device_sm_count() * 8, sums_out, (cudaStream_t)stream);
}

size_t t21_bin_multipoles_workspace_bytes(int nk) {
    if (nk < 1) return 0;
    const size_t g = (size_t)device_sm_count() * 4;
    return align_up(g * nk * 5 * sizeof(double), 256) + align_up(g * nk * sizeof(unsigned), 256);
}

int t21_bin_multipoles_half(const void* p_half, int dtype, int row_pitch, const t21_binspec* spec, int los_axis, int exclude_zero,
                            double* sums_out, long long* counts_out, void* workspace, size_t workspace_bytes, void* stream) {
    if (!p_half || !spec || !sums_out || !counts_out || !workspace) { set_error("null argument"); return T21_EINVAL; }
    if (dtype != T21_F32 && dtype != T21_F64) { set_error("bad dtype %d", dtype); return T21_EINVAL; }
    if (spec->nmu != 0 || spec->nk < 1 || los_axis < 0 || los_axis > 2 || !spec->tab_losf) {
        set_error("multipoles need a shell bin spec that carries the line-of-sight table"); return T21_EINVAL;
    }
    const size_t need = t21_bin_multipoles_workspace_bytes(spec->nk);
    if (workspace_bytes < need) { set_error("workspace too small: %zu < %zu bytes", workspace_bytes, need); return T21_ENOMEM; }
    cudaStream_t st = (cudaStream_t)stream;
    const int max_grid = device_sm_count() * 4;
    double* part_sum = (double*)workspace;
    unsigned* part_cnt = (unsigned*)((unsigned char*)workspace + align_up((size_t)max_grid * spec->nk * 5 * sizeof(double), 256));
    int grid = 0;
    if (row_pitch < spec->n[2] / 2 + 1) { set_error("row pitch %d is shorter than nz/2+1", row_pitch); return T21_EINVAL; }
    if (int rc = launch_bin_multipoles(p_half, dtype, *spec, row_pitch, los_axis, exclude_zero, part_sum, part_cnt, max_grid, &grid, st))
        return rc;
    return launch_reduce_moments(part_sum, part_cnt, grid, spec->nk, sums_out, counts_out, st);
}

}  // extern "C"
