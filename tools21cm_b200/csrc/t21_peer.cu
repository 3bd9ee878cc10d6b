// Small cross-GPU exchanges over peer-mapped memory (NVLink / NVSwitch), without a collective library: an all-gather
// of a few KB that doubles as a barrier.  It carries the three tiny exchanges of a slab-sharded power spectrum -- the
// common mean offset, "every rank has finished its peer stores" (the fused re-partition of the Y pass), and the
// per-bin sums and counts -- which used to be four NCCL all-reduces of ~20-40 us each on a step of ~1 ms.
//
// Every rank owns one buffer (t21_peer_alloc, mapped by the peers with t21_peer_open):
//     flags [2][nranks] u64 at offset 0            flags[e & 1][r] = e once rank r's payload of epoch e has landed
//     mail  [2][nranks][slot_bytes] at kMailOff    mail[e & 1][r]  = rank r's payload of epoch e
// signal: copy the payload into slot (e & 1, rank) of EVERY rank's mail (peer stores), fence at system scope, then
// publish e in every rank's flags.  wait: spin (with a time limit) until all nranks flags of this epoch's parity
// have reached e.  Two slots by parity are enough: a rank cannot start epoch e + 2 before it has seen every
// peer's flag of epoch e + 1, which those peers only raise after they are done reading epoch e.
// Both kernels are stream-ordered after the caller's earlier work (e.g. the Y pass's peer stores): the system-scope
// fence in `signal` orders those stores before the flag.  A rank only waits on flags its peers raise unconditionally,
// so the kernels of different GPUs cannot deadlock each other.
#include "t21_internal.h"

namespace t21 {

constexpr size_t kMailOff = 4096;

__global__ void __launch_bounds__(256)
peer_signal_kernel(void* const* __restrict__ peers_dev, int nranks, int rank, unsigned long long epoch,
                   const unsigned* __restrict__ payload, int nwords, int slot_bytes) {
    unsigned char* dst = static_cast<unsigned char*>(peers_dev[blockIdx.x]);
    const int par = (int)(epoch & 1ull);
    unsigned* mail = reinterpret_cast<unsigned*>(dst + kMailOff + ((size_t)par * nranks + rank) * slot_bytes);
    for (int i = threadIdx.x; i < nwords; i += blockDim.x) mail[i] = payload[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned long long* flag = reinterpret_cast<volatile unsigned long long*>(dst) + (size_t)par * nranks + rank;
        *flag = epoch;
        __threadfence_system();
    }
}

__global__ void __launch_bounds__(32)
peer_wait_kernel(const void* mine, int nranks, unsigned long long epoch, int* __restrict__ timed_out, long long limit_cycles) {
    const int par = (int)(epoch & 1ull);
    const volatile unsigned long long* flags = reinterpret_cast<const volatile unsigned long long*>(mine) + (size_t)par * nranks;
    const long long t0 = clock64();
    for (int r = threadIdx.x; r < nranks; r += 32) {
        while (flags[r] < epoch) {
            __nanosleep(200);
            if (clock64() - t0 > limit_cycles) { if (timed_out) *timed_out = 1; break; }
        }
    }
    __threadfence_system();
}

// ---- fftshifted n-D power of a slab-sharded cube, assembled with peer stores --------------------------------------
// Every rank holds the scaled power of its rows ky as the computed half spectrum half[nx][nyl][pitch] (FFT order,
// kz = 0 .. nz/2).  The kz < 0 half of output row ky is the kz > 0 half of row -ky read at -kx (Hermitian symmetry of a
// real cube), and row -ky lives on another rank: so each rank writes the kz >= 0 part of its own output rows and
// STORES the mirrored part straight into the output of the rank that owns row -ky, over NVLink -- one pass over the
// half spectrum, no send / receive buffers, no roll / flip / cat passes (44 ms -> see DESIGN.md section 5 at 2048^3).
// out[nx][nyl][nz] is fftshifted along x and z; local row j is global row iy0 + j (the caller reports the shifted y
// index of each row).  One warp per (kx, row): 128-byte runs both ways (the mirrored run is written by lanes in
// descending order and starts one element later so that it is segment-aligned).
template <typename T>
__global__ void __launch_bounds__(256)
nd_assemble_kernel(const T* __restrict__ half, int nx, int nyl, int nz, int pitch, int iy0, int ny, void* const* __restrict__ outs,
                   int rank, int shift_y) {
    const int lane = threadIdx.x & 31;
    const long long nrows = (long long)nx * nyl;
    const long long w0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nw = (long long)gridDim.x * (blockDim.x >> 5);
    const int cz = nz / 2, nneg = (nz - 1) / 2;
    T* out_mine = static_cast<T*>(outs[rank]);
    for (long long r = w0; r < nrows; r += nw) {
        const int ix = (int)(r / nyl), j = (int)(r - (long long)ix * nyl);
        const int sx = (ix + nx / 2) % nx, sxm = ((nx - ix) % nx + nx / 2) % nx;
        const int iym = (ny - (iy0 + j)) % ny, rm = iym / nyl, jm = iym - rm * nyl;
        const T* src = half + r * pitch;
        // one rank holding every row (shift_y): the rows are fftshifted along y as well
        const int jo = shift_y ? (j + ny / 2) % ny : j, jmo = shift_y ? (jm + ny / 2) % ny : jm;
        T* lo = out_mine + ((long long)sx * nyl + jo) * nz;
        T* mi = static_cast<T*>(outs[rm]) + ((long long)sxm * nyl + jmo) * nz;
        for (int kz = lane; kz <= nneg; kz += 32) lo[cz + kz] = src[kz];
        if (nz % 2 == 0 && lane == 0) lo[0] = src[nz / 2];                     // the Nyquist plane sits at index 0
        for (int kz = 1 + lane; kz <= nneg; kz += 32) mi[cz - kz] = src[kz];
    }
}

// The same assembly from a PENCIL-MAJOR half spectrum, half_t[ny][nz/2+1][nx] (what the warp-per-pencil X pass writes
// with full 128-byte runs): 32 x 32 (kz, kx) tiles go through shared memory, so both the reads (along kx) and the writes
// (along kz, for the mode and, mirrored, for its Hermitian twin) are 128-byte runs.  One rank, all three axes shifted.
template <typename T, int TZ, int TX>
__global__ void __launch_bounds__(256)
nd_assemble_t_kernel(const T* __restrict__ half_t, int nx, int ny, int nz, T* __restrict__ out) {
    // tile[kz][kx]: TZ x TX modes of one ky.  Reads: TX consecutive kx per kz (runs of TX elements); writes: TZ
    // consecutive kz per kx (runs of TZ elements, twice -- the mode and its twin).  TZ > TX: the writes are 2/3 of the traffic.
    __shared__ T tile[TZ][TX + 1];
    const int nzh = nz / 2 + 1, hx = nx / 2, hz = nz / 2;
    const int ix0 = blockIdx.x * TX, iz0 = blockIdx.y * TZ, iy = blockIdx.z;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // (row bases once per warp and row, wrap-arounds by conditional subtraction: a run-time modulo per element made the
    // first version of this kernel instruction bound)
    for (int izl = warp; izl < TZ; izl += 8) {
        const int iz = iz0 + izl;
        if (iz >= nzh) break;
        const T* src = half_t + ((size_t)iy * nzh + iz) * nx + ix0;
        for (int ixl = lane; ixl < TX; ixl += 32)
            if (ix0 + ixl < nx) tile[izl][ixl] = src[ixl];
    }
    __syncthreads();
    int sy = iy + ny / 2; if (sy >= ny) sy -= ny;
    int uy = (iy ? ny - iy : 0) + ny / 2; if (uy >= ny) uy -= ny;
    for (int ixl = warp; ixl < TX; ixl += 8) {
        const int ix = ix0 + ixl;
        if (ix >= nx) break;
        int sx = ix + hx; if (sx >= nx) sx -= nx;
        int ux = (ix ? nx - ix : 0) + hx; if (ux >= nx) ux -= nx;
        T* lo = out + ((size_t)sx * ny + sy) * nz;
        T* mi = out + ((size_t)ux * ny + uy) * nz;
        for (int izl = lane; izl < TZ; izl += 32) {
            const int iz = iz0 + izl;
            if (iz >= nzh) break;
            const T val = tile[izl][ixl];
            int sz = iz + hz; if (sz >= nz) sz -= nz;
            lo[sz] = val;
            if (iz != 0 && 2 * iz != nz) mi[hz - iz] = val;       // the twin (-kx, -ky, -kz): kz index nz - iz, shifted
        }
    }
}

}  // namespace t21

using namespace t21;

extern "C" {

int t21_nd_assemble_t(const void* half_t, int dtype, int nx, int ny, int nz, void* out, void* stream) {
    if (!half_t || !out || nx < 1 || ny < 1 || nz < 1 || ny > 65535 || (dtype != T21_F32 && dtype != T21_F64)) {
        set_error("bad argument to t21_nd_assemble_t"); return T21_EINVAL;
    }
    const int nzh = nz / 2 + 1;
    if (dtype == T21_F32) {
        const dim3 grid((nx + 63) / 64, (nzh + 127) / 128, ny);
        nd_assemble_t_kernel<float, 128, 64><<<grid, 256, 0, (cudaStream_t)stream>>>((const float*)half_t, nx, ny, nz, (float*)out);
    } else {
        const dim3 grid((nx + 63) / 64, (nzh + 63) / 64, ny);
        nd_assemble_t_kernel<double, 64, 64><<<grid, 256, 0, (cudaStream_t)stream>>>((const double*)half_t, nx, ny, nz, (double*)out);
    }
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    return T21_OK;
}

size_t t21_peer_comm_bytes(int nranks, int slot_bytes) {
    if (nranks < 1 || slot_bytes < 4) return 0;
    return kMailOff + 2 * (size_t)nranks * slot_bytes;
}
size_t t21_peer_comm_mail_offset(void) { return kMailOff; }

int t21_peer_signal(void* const* peers_dev, int nranks, int rank, unsigned long long epoch, const void* payload_dev,
                    int nbytes, int slot_bytes, void* stream) {
    if (!peers_dev || nranks < 1 || nranks > 64 || rank < 0 || rank >= nranks || epoch == 0 || nbytes < 0 || nbytes > slot_bytes ||
        (nbytes & 3) || (nbytes > 0 && !payload_dev) || 2 * (size_t)nranks * 8 > kMailOff) {
        set_error("bad argument to t21_peer_signal"); return T21_EINVAL;
    }
    peer_signal_kernel<<<nranks, 256, 0, (cudaStream_t)stream>>>(peers_dev, nranks, rank, epoch, (const unsigned*)payload_dev,
                                                                nbytes / 4, slot_bytes);
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    return T21_OK;
}

int t21_peer_wait(const void* my_buf, int nranks, unsigned long long epoch, int* timed_out_dev, double timeout_s, void* stream) {
    if (!my_buf || nranks < 1 || nranks > 64 || epoch == 0) { set_error("bad argument to t21_peer_wait"); return T21_EINVAL; }
    const long long limit = (long long)((timeout_s > 0 ? timeout_s : 10.0) * 1.9e9);
    peer_wait_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(my_buf, nranks, epoch, timed_out_dev, limit);
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    return T21_OK;
}

int t21_nd_assemble(const void* half, int dtype, int nx, int nyl, int nz, int row_pitch, int iy0, int ny_global,
                    void* const* outs_dev, int nranks, int rank, int shift_y, void* stream) {
    if (!half || !outs_dev || nx < 1 || nyl < 1 || nz < 1 || row_pitch < nz / 2 + 1 || nranks < 1 || rank < 0 || rank >= nranks ||
        ny_global != nyl * nranks || iy0 != rank * nyl || (dtype != T21_F32 && dtype != T21_F64) || (shift_y && nranks != 1)) {
        set_error("bad argument to t21_nd_assemble"); return T21_EINVAL;
    }
    long long g = ((long long)nx * nyl + 7) / 8;
    const long long cap = (long long)device_sm_count() * 8;
    if (g > cap) g = cap;
    if (dtype == T21_F32)
        nd_assemble_kernel<float><<<(int)g, 256, 0, (cudaStream_t)stream>>>((const float*)half, nx, nyl, nz, row_pitch, iy0, ny_global, outs_dev, rank, shift_y);
    else
        nd_assemble_kernel<double><<<(int)g, 256, 0, (cudaStream_t)stream>>>((const double*)half, nx, nyl, nz, row_pitch, iy0, ny_global, outs_dev, rank, shift_y);
    count_launch();
    T21_CUDA_OK(cudaGetLastError());
    return T21_OK;
}

}  // extern "C"
