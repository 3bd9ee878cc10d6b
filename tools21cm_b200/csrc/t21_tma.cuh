// sm_100a asynchronous-copy plumbing shared by the warp-per-pencil kernels: mbarriers, TMA tensor loads
// (cp.async.bulk.tensor -> UTMALDG), bulk copies (cp.async.bulk -> UBLKCP) and the host-side tensor-map encoder.
//
// Why TMA here (profiles/r02_ubench_tma_tile.txt, B200): an X-pass tile is 1024 rows of 64 B that sit 64 KB apart.
// Two CTAs per SM that each keep ONE such tile in flight through a 4-D tensor map pull 6.5 TB/s (the measured
// copy peak is 6.56 TB/s); the same tiles fetched with LDG.64 by 8 CTAs x 256 threads per SM reach 5.8 TB/s and
// need every warp's registers to do it.  A Y-pass tile (one contiguous 64 KB block) moves at 7.3 TB/s as a
// single bulk copy.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace t21 {

#if defined(__CUDACC__)
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
// makes the initialised barrier visible to the async proxy (TMA completes transactions on it)
__device__ __forceinline__ void mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// one arrival that also announces `bytes` of asynchronous copies to wait for
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// blocks until the phase with the given parity has completed (try_wait suspends the warp in hardware)
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// orders this thread's earlier generic-proxy accesses to shared memory before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, void* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void bulk_load(void* dst, const void* src, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_store(void* dst, const void* src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// waits until the sources of all but the N most recent bulk-store groups have been read (the buffers may be reused)
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }

// named barrier among `nthreads` threads of the CTA (ids 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
#endif

// Tensor map of a chunk-major half spectrum, complex64: element (slab s, piece c, x_lo, y, l) of
// [nslab][nch][nxs][nya][8 complex] as a 5-D fp32 tensor (16 floats, nya, nxs, nch, nslab); the box is
// `box_rows` x planes of one (y, piece): box_rows * 64 bytes.  Returns 0 or a T21_E* code (t21_api.cu).
// box_floats: 16 (the whole 64-byte piece) or 4 (one pair of columns: the nx = 512 kernel lands a tile as four column pairs).
int make_wmap_f32(CUtensorMap* out, const void* w, int nxs, int nya, int nch, int nslab, int box_rows, int box_floats = 16);

}  // namespace t21
