// Microbenchmark (not part of the library): how fast can TMA pull X-pass tiles -- 1024 rows of 64 B, 64 KB apart --
// and Y-pass tiles (one contiguous 64 KB block) into shared memory?  Decides the staging scheme of the pencil kernels.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_tile tma_tile.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, void* bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// MODE 0: X tiles through a 4-D tensor map (16 floats, y, x, chunk), 4 boxes of 256 rows per tile
// MODE 1: Y tiles: one 64 KB bulk copy
// MODE 2: X tiles with plain LDG.64 by all threads (what the round-1 kernel does), summed so the loads stay
// NBUF buffers of 64 KB per CTA, all kept in flight; `consume` = every thread reads 128 B of the landed tile.
template <int MODE, int NBUF>
__global__ void __launch_bounds__(256) fetch_kernel(const __grid_constant__ CUtensorMap map, const float* w, int nx, int ny, int nch,
                                                   long long ntiles, float* sink, int consume) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[NBUF];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int b = 0; b < NBUF; ++b) mbar_init(&bars[b], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    float acc = 0.f;
    auto issue = [&](long long tile, int b) {
        unsigned char* dst = smem + (size_t)b * 65536;
        if (MODE == 0) {
            const int c = (int)(tile / ny), y = (int)(tile % ny);
            mbar_expect_tx(&bars[b], 65536);
            for (int q = 0; q < nx / 256; ++q) tma_load_4d(dst + q * 16384, &map, &bars[b], 0, y, q * 256, c);
        } else {
            mbar_expect_tx(&bars[b], 65536);
            bulk_load(dst, w + tile * 16384, 65536, &bars[b]);
        }
    };
    if (MODE == 2) {
        for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int c = (int)(tile / ny), y = (int)(tile % ny);
            const float2* base = reinterpret_cast<const float2*>(w) + ((long long)c * nx * ny + y) * 8 + (tid & 7);
            for (int i = tid >> 3; i < nx; i += 32 * 4) {
                float2 v0 = __ldg(base + (long long)i * ny * 8), v1 = __ldg(base + (long long)(i + 32) * ny * 8);
                float2 v2 = __ldg(base + (long long)(i + 64) * ny * 8), v3 = __ldg(base + (long long)(i + 96) * ny * 8);
                acc += v0.x + v1.y + v2.x + v3.y;
            }
        }
        if (acc == 12345.f) sink[0] = acc;
        return;
    }
    // prologue: fill the ring
    long long next = blockIdx.x;
    if (tid == 0)
        for (int b = 0; b < NBUF && next + (long long)b * gridDim.x < ntiles; ++b) issue(next + (long long)b * gridDim.x, b);
    int b = 0;
    unsigned phase = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        mbar_wait(&bars[b], phase);
        if (consume) {
            const float4* t = reinterpret_cast<const float4*>(smem + (size_t)b * 65536);
            for (int i = tid; i < 4096; i += 256 * 4) { float4 v = t[i]; acc += v.x + v.w; }
        }
        __syncthreads();                       // everyone is done with buffer b
        const long long refill = tile + (long long)NBUF * gridDim.x;
        if (tid == 0 && refill < ntiles) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(refill, b);
        }
        if (++b == NBUF) { b = 0; phase ^= 1; }
    }
    if (acc == 12345.f) sink[0] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE, int NBUF>
static void run(const char* name, const CUtensorMap& map, const float* w, int nx, int ny, int nch, int ctas_per_sm, int consume, float* sink) {
    auto kern = fetch_kernel<MODE, NBUF>;
    const size_t smem = MODE == 2 ? 0 : (size_t)NBUF * 65536;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 256, smem));
    if (occ < ctas_per_sm) { printf("%-28s skipped (occupancy %d < %d)\n", name, occ, ctas_per_sm); return; }
    const long long ntiles = (long long)ny * nch;
    const int grid = 148 * ctas_per_sm;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) kern<<<grid, 256, smem>>>(map, w, nx, ny, nch, ntiles, sink, consume);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    const int reps = 10;
    for (int i = 0; i < reps; ++i) kern<<<grid, 256, smem>>>(map, w, nx, ny, nch, ntiles, sink, consume);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= reps;
    const double bytes = (double)ntiles * 65536.0;
    printf("%-28s ctas/SM %d  bufs %d  consume %d : %.3f ms  %.0f GB/s\n", name, ctas_per_sm, NBUF, consume, ms, bytes / ms / 1e6);
}

int main() {
    const int nx = 1024, ny = 1024, nch = 66;           // 1024^3 half spectrum, chunk-major: [nch][nx][ny][8] complex64
    const size_t bytes = (size_t)nch * nx * ny * 64;
    float* w = nullptr; float* sink = nullptr;
    CK(cudaMalloc(&w, bytes));
    CK(cudaMalloc(&sink, 64));
    CK(cudaMemset(w, 0, bytes));
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 1; }
    CUtensorMap map;
    cuuint64_t dims[4] = {16, (cuuint64_t)ny, (cuuint64_t)nx, (cuuint64_t)nch};
    cuuint64_t strides[3] = {64, (cuuint64_t)ny * 64, (cuuint64_t)nx * ny * 64};
    cuuint32_t box[4] = {16, 1, 256, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, w, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
    for (int consume = 0; consume < 2; ++consume) {
        run<0, 1>("x tiles, tma 4d", map, w, nx, ny, nch, 1, consume, sink);
        run<0, 1>("x tiles, tma 4d", map, w, nx, ny, nch, 2, consume, sink);
        run<0, 1>("x tiles, tma 4d", map, w, nx, ny, nch, 3, consume, sink);
        run<0, 2>("x tiles, tma 4d", map, w, nx, ny, nch, 1, consume, sink);
        run<0, 3>("x tiles, tma 4d", map, w, nx, ny, nch, 1, consume, sink);
        run<1, 1>("y tiles, bulk 64K", map, w, nx, ny, nch, 1, consume, sink);
        run<1, 1>("y tiles, bulk 64K", map, w, nx, ny, nch, 2, consume, sink);
        run<1, 1>("y tiles, bulk 64K", map, w, nx, ny, nch, 3, consume, sink);
        run<1, 3>("y tiles, bulk 64K", map, w, nx, ny, nch, 1, consume, sink);
    }
    run<2, 1>("x tiles, ldg.64", map, w, nx, ny, nch, 4, 0, sink);
    run<2, 1>("x tiles, ldg.64", map, w, nx, ny, nch, 8, 0, sink);
    printf("done\n");
    return 0;
}
