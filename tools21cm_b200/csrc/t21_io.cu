// The input side of the path: getting cubes from host memory / C2Ray-format files onto the
// device in the layout the first pass reads, without the host round trips the reference makes.
//
//  * t21_h2d_staged: pageable host memory -> device through a pinned ring, filled by several host threads while the
//    previous chunk's DMA runs.  A numpy array handed to the estimators is pageable: one blocking cudaMemcpy of it
//    stages through the driver's own small bounce buffers on one thread (~10 GB/s); this keeps PCIe busy instead
//    (52-54 GB/s against 55.6 GB/s from pinned memory).
//  * t21_f2c: C2Ray / CubeP3M blobs are Fortran ordered (x fastest; xfrac_file.py:62, density_file.py:51
//    `reshape(..., order='F')`); the passes read C order (z fastest).  Tiled transpose through shared memory.
//  * t21_c2ray_dt: calc_dt (temperature.py:51-80, _dt :215-222) FUSED with that transpose: reads the ionised fraction and
//    the raw density blobs as they sit in the files, writes dT = factor * (1 - x_i) * rho in C order, ready for the
//    Z pass -- the float64 copies, the cgs density and the two transposed arrays of the reference are never formed.
#include <immintrin.h>
#include <omp.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <mutex>
#include "t21_internal.h"

namespace t21 {

// ---- staged host -> device copy ---------------------------------------------------------------------------------
constexpr int kRing = 4;
// Chunk and store kind were measured on the B200 host (16 cores, 60 MB L3, tools/r2_h2d.py; pinned copy 55.6 GB/s):
//   32 MB chunks, streaming stores (round 2 as first built): 30-37 GB/s -- source read + ring write + the DMA's read
//   of the ring are three passes over host DRAM, and the VM's memory bandwidth is what runs out;
//   8 MB chunks, plain stores: 52-54 GB/s -- the 32 MB ring stays in the last-level cache, the DMA engine's reads are
//   served from it, and only the source is read from DRAM.  1-2 MB chunks pay the per-chunk launch (33 GB/s).
static size_t kChunk = 8u << 20;       // T21_H2D_CHUNK_MB overrides it (read once, at the first copy)

struct StageRing {
    unsigned char* buf[kRing] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev[kRing];
    bool used[kRing] = {false, false, false, false};
    bool ok = false;
    std::mutex mu;
};
static StageRing g_ring;

// copy into the pinned ring.  With streaming stores the lines are not pulled into the cache first; with plain stores
// (default) the ring is small enough to live in the last-level cache, where the DMA engine finds it
static bool g_nt_stores = false;       // T21_H2D_NT=1: streaming stores (for rings larger than the last-level cache)
static void stream_copy(unsigned char* dst, const unsigned char* src, size_t n) {
#if defined(__AVX2__)
    size_t i = 0;
    if (g_nt_stores && ((uintptr_t)dst & 31) == 0) {
        for (; i + 128 <= n; i += 128) {
            const __m256i a = _mm256_loadu_si256((const __m256i*)(src + i)), b = _mm256_loadu_si256((const __m256i*)(src + i + 32));
            const __m256i c = _mm256_loadu_si256((const __m256i*)(src + i + 64)), d = _mm256_loadu_si256((const __m256i*)(src + i + 96));
            _mm256_stream_si256((__m256i*)(dst + i), a);
            _mm256_stream_si256((__m256i*)(dst + i + 32), b);
            _mm256_stream_si256((__m256i*)(dst + i + 64), c);
            _mm256_stream_si256((__m256i*)(dst + i + 96), d);
        }
        _mm_sfence();
    }
    if (i < n) memcpy(dst + i, src + i, n - i);
#else
    memcpy(dst, src, n);
#endif
}

static int ring_init() {
    if (g_ring.ok) return 0;
    // several ranks on one host (torchrun sets LOCAL_WORLD_SIZE): their rings together no longer fit the last-level
    // cache, so each goes back to 32 MB chunks written with streaming stores (8 ranks: 43 cubes/s against 16)
    if (const char* e = getenv("LOCAL_WORLD_SIZE")) { if (atoi(e) > 1) { g_nt_stores = true; kChunk = 32u << 20; } }
    if (const char* e = getenv("T21_H2D_NT")) g_nt_stores = e[0] != '0';
    if (const char* e = getenv("T21_H2D_CHUNK_MB")) { const long mb = atol(e); if (mb >= 1 && mb <= 1024) kChunk = (size_t)mb << 20; }
    for (int i = 0; i < kRing; ++i) {
        T21_CUDA_OK(cudaHostAlloc((void**)&g_ring.buf[i], kChunk, cudaHostAllocDefault));
        T21_CUDA_OK(cudaEventCreateWithFlags(&g_ring.ev[i], cudaEventDisableTiming));
    }
    g_ring.ok = true;
    return 0;
}

// ---- Fortran order -> C order, optionally fused with the brightness temperature ---------------------------------
// in: [nz][ny][nx] as stored (x fastest) = element (x, y, z) of the logical cube; out[x][y][z].  One 32 x 32 (x, z)
// tile per CTA and y: rows of 32 x are read coalesced, rows of 32 z written coalesced.
template <typename TX, typename TD, typename TO, bool DT>
__global__ void __launch_bounds__(256)
f2c_kernel(const TX* __restrict__ xin, const TD* __restrict__ din, TO* __restrict__ out, int nx, int ny, int nz, double factor,
           int neutral) {
    __shared__ TO tile[32][33];
    const int x0 = blockIdx.x * 32, z0 = blockIdx.y * 32, y = blockIdx.z;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 32 x 8
#pragma unroll
    for (int r = 0; r < 32; r += 8) {
        const int z = z0 + ty + r, x = x0 + tx;
        if (z < nz && x < nx) {
            const long long o = ((long long)z * ny + y) * nx + x;
            if (DT) {
                // the file may hold the neutral fraction: the reference forms 1 - x in the FILE's precision first
                // (xfrac_file.py:64-65, a float32 array for old-format files), then casts to float64 (temperature.py:66)
                const double xi = neutral ? (double)((TX)1 - xin[o]) : (double)xin[o];
                tile[ty + r][tx] = (TO)(factor * (1.0 - xi) * (double)din[o]);      // temperature.py:221
            } else {
                tile[ty + r][tx] = (TO)xin[o];
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 32; r += 8) {
        const int x = x0 + ty + r, z = z0 + tx;
        if (x < nx && z < nz) out[((long long)x * ny + y) * nz + z] = tile[tx][ty + r];
    }
}

}  // namespace t21

using namespace t21;

extern "C" {

int t21_h2d_staged(void* dst_dev, const void* src_host, size_t bytes, int threads, void* stream) {
    if (!dst_dev || !src_host) { set_error("null argument"); return T21_EINVAL; }
    if (bytes == 0) return T21_OK;
    std::lock_guard<std::mutex> lock(g_ring.mu);
    if (int rc = ring_init()) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    int nt = threads > 0 ? threads : std::min(32, std::max(1, omp_get_num_procs()));
    const unsigned char* src = (const unsigned char*)src_host;
    unsigned char* dst = (unsigned char*)dst_dev;
    int slot = 0;
    for (size_t off = 0; off < bytes; off += kChunk, slot = (slot + 1) % kRing) {
        const size_t n = std::min(kChunk, bytes - off);
        if (g_ring.used[slot]) T21_CUDA_OK(cudaEventSynchronize(g_ring.ev[slot]));     // its last DMA has read the buffer
        unsigned char* b = g_ring.buf[slot];
        const size_t piece = (n / nt + 4095) / 4096 * 4096 + 4096;
#pragma omp parallel for num_threads(nt) schedule(static)
        for (int t = 0; t < nt; ++t) {
            const size_t a = (size_t)t * piece;
            if (a < n) stream_copy(b + a, src + off + a, std::min(piece, n - a));
        }
        T21_CUDA_OK(cudaMemcpyAsync(dst + off, b, n, cudaMemcpyHostToDevice, st));
        T21_CUDA_OK(cudaEventRecord(g_ring.ev[slot], st));
        g_ring.used[slot] = true;
    }
    return T21_OK;
}

#define T21_F2C_LAUNCH(TX, TD, TO, DT)                                                                                 \
    f2c_kernel<TX, TD, TO, DT><<<grid, 256, 0, (cudaStream_t)stream>>>((const TX*)xin, (const TD*)din, (TO*)out, nx, ny, nz, factor, neutral)

int t21_f2c(const void* in_f, int dtype, int nx, int ny, int nz, void* out_c, int out_dtype, void* stream) {
    if (!in_f || !out_c || nx < 1 || ny < 1 || nz < 1 || ny > 65535 || (nz + 31) / 32 > 65535) { set_error("bad argument to t21_f2c"); return T21_EINVAL; }
    const dim3 grid((nx + 31) / 32, (nz + 31) / 32, ny);
    const void* xin = in_f; const void* din = nullptr; void* out = out_c; const double factor = 1.0; const int neutral = 0;
    if (dtype == T21_F32 && out_dtype == T21_F32) T21_F2C_LAUNCH(float, float, float, false);
    else if (dtype == T21_F32) T21_F2C_LAUNCH(float, float, double, false);
    else if (out_dtype == T21_F32) T21_F2C_LAUNCH(double, float, float, false);
    else T21_F2C_LAUNCH(double, float, double, false);
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    return T21_OK;
}

int t21_c2ray_dt(const void* xfrac_f, int xfrac_dtype, int neutral, const void* dens_f, int dens_dtype, double factor, int nx,
                 int ny, int nz, void* out_c, int out_dtype, void* stream) {
    if (!xfrac_f || !dens_f || !out_c || nx < 1 || ny < 1 || nz < 1 || ny > 65535 || (nz + 31) / 32 > 65535) { set_error("bad argument to t21_c2ray_dt"); return T21_EINVAL; }
    const dim3 grid((nx + 31) / 32, (nz + 31) / 32, ny);
    const void* xin = xfrac_f; const void* din = dens_f; void* out = out_c;
    const int key = (xfrac_dtype == T21_F64 ? 4 : 0) | (dens_dtype == T21_F64 ? 2 : 0) | (out_dtype == T21_F64 ? 1 : 0);
    switch (key) {
        case 0: T21_F2C_LAUNCH(float, float, float, true); break;
        case 1: T21_F2C_LAUNCH(float, float, double, true); break;
        case 2: T21_F2C_LAUNCH(float, double, float, true); break;
        case 3: T21_F2C_LAUNCH(float, double, double, true); break;
        case 4: T21_F2C_LAUNCH(double, float, float, true); break;
        case 5: T21_F2C_LAUNCH(double, float, double, true); break;
        case 6: T21_F2C_LAUNCH(double, double, float, true); break;
        default: T21_F2C_LAUNCH(double, double, double, true); break;
    }
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    return T21_OK;
}

}  // extern "C"
