/* t21ps -- C ABI of the B200-native power-spectrum estimator.
 *
 * The reference (tools21cm v2.4.0) is pure Python: its "FFI" for this path is the
 * set of module functions in src/tools21cm/power_spectrum.py.  Each entry point
 * below is what a binding for one of those functions would call; the Python host
 * side in tools21cm_b200/power_spectrum.py binds them with ctypes and keeps the
 * reference signatures.  Plain pointers and sizes only; no torch types.
 *
 * Conventions
 *   - every `const void*` / `void*` data pointer is a DEVICE pointer owned by the
 *     caller (the Python side allocates them as torch tensors); the library
 *     allocates nothing on the device except the immutable twiddle tables of a plan;
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous with
 *     respect to the host and ordered on that stream;
 *   - functions return 0 on success or a negative T21_E* code and never throw;
 *     t21_last_error() returns a thread-local message for the last failure.
 */
#ifndef T21PS_H
#define T21PS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define T21_API __attribute__((visibility("default")))
#else
#define T21_API
#endif

#define T21_F32 0
#define T21_F64 1

#define T21_OK 0
#define T21_EINVAL -1   /* bad argument */
#define T21_ECUDA -2    /* CUDA runtime error (message in t21_last_error) */
#define T21_ENOMEM -3   /* workspace too small */

#define T21_LUT_OUTSIDE 0xFFFFu /* lookup cell lies outside every bin */
#define T21_LUT_EXACT 0xFFFEu   /* lookup cell touches an edge: use the float64 path */

typedef struct t21_plan t21_plan;

/* Everything the binning code needs, built on the host by tools21cm_b200/binspec.py
 * with the reference's float64 expressions:
 *   tab_s[d][i]  addend of s along axis d (k_d^2; reference power_spectrum.py:473-477)
 *   thr[b]       bin b  <=>  thr[b] <= s < thr[b+1], s = (tab_s0 + tab_s1) + tab_s2
 *                (edges of _get_kbins :505-519 mapped through sqrt; last one strict or
 *                 not depending on np.histogram :128 vs mu_binning :424 semantics)
 *   tab_los      signed k along the line of sight (_get_mu :481-502)
 *   mu_edges     half-open mu bins (:406-409, :428)
 *   *_lut        fp32 fast-path tables; T21_LUT_EXACT cells fall back to float64.
 *   *_lo_f/_hi_f fp32 intervals strictly inside each bin (guard band >> fp32 error).
 * Indices are FFT order for the fused kernels, fftshift order for t21_bin_only. */
typedef struct t21_binspec {
    int32_t n[3];
    int32_t nk;
    int32_t nmu;          /* 0: spherical shells only */
    int32_t los_axis;     /* 0..2, or -1 when nmu == 0 */
    int32_t absolute_mus;
    int32_t exclude_zero; /* drop the k_perp = 0 line (_get_nonzero_idx :563-575) */
    int32_t zero_idx[3];  /* index of k=0 along each axis */
    int32_t nyq_idx[3];   /* index of the self-conjugate Nyquist sample, -1 if none */
    const double* tab_s[3];
    const float* tab_sf[3];
    const double* tab_los;
    const float* tab_losf;
    const double* thr;
    const double* mu_edges;
    double s_nan;         /* modes with s < s_nan have mu = NaN (:501) */
    const uint16_t* k_lut;
    int32_t k_cells;
    float k_lut_c0;       /* cell = floor(log2(s) * k_lut_inv + k_lut_c0) */
    float k_lut_inv;
    const uint16_t* mu_lut;
    int32_t mu_cells;     /* cell = floor((mu + 1) * mu_cells / 2) */
    /* no-lookup fast path: fp32 s^ in [k_lo_f[b], k_hi_f[b]) guarantees the exact s is in k bin b
     * (b = nk / nk+1: below / above every bin); same for mu^ and [mu_lo_f[m], mu_hi_f[m]) */
    const float* k_lo_f;
    const float* k_hi_f;
    const float* mu_lo_f;
    const float* mu_hi_f;
    /* Optional pencil segment table of the warp-per-pencil X pass (t21_build_pencil_segments), or NULL.
     * seg_tab[(iy * (n[2]/2+1) + iz) * seg_stride + j] = (p_j << 16) | (bin_j + 1): along the x pencil (iy, iz) the
     * modes with p_j <= |kx index| < p_{j+1} fall in flat bin bin_j (-1: none); p_0 = 0; unused entries 0xFFFF0000. */
    const uint32_t* seg_tab;
    int32_t seg_stride;   /* entries per pencil: 8, 16 or 32 */
    int32_t seg_ntab;     /* 1 (shells, |mu|) or 4 (signed mu: own bin for kx >= 0, own for kx < 0, twin's bin for each) */
    int32_t seg_overflow; /* pencils marked 0xFFFE0000 in entry 0 (more ranges than seg_stride): only the warp-per-pencil
                           * X pass takes such a table, and classifies those pencils mode by mode */
} t21_binspec;

/* Bins are index ranges along an x pencil: for fixed (ky, kz), k only grows with |kx| and so does |mu| when the line of
 * sight is x (it only falls otherwise), and a mode and its mirror -kx share a bin unless mu is signed.  This builds, once
 * per (shape, bin spec), the table of those ranges for every pencil of the computed half spectrum -- with the very
 * classification every other kernel uses per mode (float64 sequence of power_spectrum.py:477,498 where the fp32
 * look-up cannot decide) -- so that the X pass adds whole ranges instead of classifying 10^9 modes per cube.
 * `spec` must be in FFT index order.  out: T * ny * (nz/2+1) * stride uint32 (device), T = 1 table for shells and |mu|
 * bins (even in kx), T = 4 consecutive tables for signed mu (the stored mode's own bin for kx = +f and for kx = -f, then
 * the bin of its Hermitian twin for each; seg_ntab).  max_segments_dev (two device ints): [0] the largest number of ranges
 * any pencil needs, stride + 1 if some pencil needs more than `stride`; [1] how many pencils do -- entry 0 of those is the
 * overflow mark 0xFFFE0000 (put the count in seg_overflow). */
T21_API int t21_build_pencil_segments(const t21_binspec* spec, uint32_t* out, int stride, int* max_segments_dev, void* stream);

/* The same table for the (k_perp, k_par) bins of power_spectrum_2d (power_spectrum.py:177-254), from the separable tables of
 * t21_bin_cyl_half (FFT index order; flat bin = perp * npar + par).  A t21_binspec that carries only n, nk = nperp * npar,
 * nmu = 0 and this table (every other pointer NULL) makes t21_ps_binned the fused form of that estimator; such a
 * table-only spec is refused (T21_EINVAL) wherever the warp-per-pencil X pass does not apply. */
T21_API int t21_build_pencil_segments_cyl(int n0, int n1, int n2, int nu_axis, const int* perp_tab, const int* par_tab, int nperp,
                                          int npar, uint32_t* out, int stride, int* max_segments_dev, void* stream);

/* Plan for real cubes of shape (nx, ny, nz), C order.  dtype T21_F32 / T21_F64 is the
 * arithmetic type of the transform.  Replaces the implicit scipy.fftpack plan behind
 * power_spectrum.py:45,85-86. */
T21_API int t21_plan_create(t21_plan** plan, int nx, int ny, int nz, int dtype, int device);
T21_API int t21_plan_destroy(t21_plan* plan);

/* Bytes of caller-provided workspace t21_ps_binned (nbins = nk * max(nmu, 1)) or
 * t21_ps_nd (nbins = 0) need for n_fields (1 = auto, 2 = cross) input cubes. */
T21_API size_t t21_workspace_bytes(const t21_plan* plan, int n_fields, int nbins);

/* power_spectrum_1d / cross_power_spectrum_1d / power_spectrum_mu / cross_power_spectrum_mu
 * (power_spectrum.py:137-174, 257-289, 292-332, 335-377) without the final division:
 * R2C 3-D FFT of x1 (and x2), |F|^2 or Re(F1 conj F2), binned on the fly.
 *   sums_out[nbins]   float64 unscaled sum of power per bin (bin = imu*nk + ik)
 *   counts_out[nbins] int64 number of modes per bin
 * subtract_mean != 0: transform x - c with c ~ mean(x) estimated on the device and repair the
 * DC sample analytically (only changes rounding; DC keeps its true value). */
T21_API int t21_ps_binned(const t21_plan* plan, const void* x1, const void* x2_or_null, int subtract_mean,
                  const t21_binspec* spec, double* sums_out, long long* counts_out,
                  void* workspace, size_t workspace_bytes, void* stream);

/* t21_ps_binned fed from HOST memory (one field; power-of-two axes): the cube is copied in groups of x planes on an
 * internal copy stream -- through the pinned ring of t21_h2d_staged for pageable memory (host_is_pinned == 0), straight
 * for pinned memory -- while `stream` runs the Z and Y passes on the groups that have landed; only the X pass is left
 * after the copy.  x_dev: device buffer of the cube's size (scratch for the caller afterwards).  This is what
 * power_spectrum_1d(numpy_array) (power_spectrum.py:137-174 with its float64 copy at :45) costs end to end. */
T21_API int t21_ps_binned_host(const t21_plan* plan, const void* x_host, int host_is_pinned, int subtract_mean,
                               const t21_binspec* spec, double* sums_out, long long* counts_out, void* x_dev,
                               void* workspace, size_t workspace_bytes, int threads, void* stream);

/* The batch behind integrated_bispectrum / power_spectrum_response (position_dependent_power_spectrum.py:45-88): the
 * binned auto power spectrum of `nwindows` cubically masked copies of one cube, `cube * W_L` (:65-77,90-103), with the
 * mask fused into the first pass's load -- no masked cube is written, only the window's samples are read, and one
 * call queues every launch.  windows: HOST array [nwindows][6] = {x0, x1, y0, y1, z0, z1}; sums_out[nwindows][nbins];
 * counts_out[nbins] (the same for every window).  Power-of-two axes, at most 2048 bins, no mean offset. */
T21_API int t21_ps_binned_windows(const t21_plan* plan, const void* x1, const int* windows, int nwindows,
                                  const t21_binspec* spec, double* sums_out, long long* counts_out, void* workspace,
                                  size_t workspace_bytes, void* stream);

/* power_spectrum_nd / cross_power_spectrum_nd (power_spectrum.py:24-54, 57-95): the full
 * fftshifted power array, multiplied by `scale`, written as out_dtype (T21_F32/T21_F64). */
T21_API int t21_ps_nd(const t21_plan* plan, const void* x1, const void* x2_or_null, int subtract_mean,
              double scale, void* out, int out_dtype, void* workspace, size_t workspace_bytes,
              void* stream);

/* One cube sharded over P GPUs as x slabs (configs 4 and 5 of BASELINE.json; the reference has no
 * distributed path, this replaces the single-process fftn of power_spectrum.py:45,85-86 when the
 * cube is spread over ranks).  Half-spectrum buffers are "chunk major": [pitch/LW][x planes][y rows][LW]
 * complex with LW = 8 (fp32) / 4 (fp64), i.e. pitch * planes * rows elements (pitch from t21_slab_pitch).
 * stage1: Z and Y passes of this rank's slab [nx/P][ny][nz] into w_out (offset = device scalar
 * subtracted from every sample, the SAME value on every rank, or NULL).  With send_out != NULL the Y pass stores straight
 * into the packed all-to-all send buffer (nranks blocks; block s = the chunk-major part for rank s; w_out is
 * then scratch), so the re-partition needs no separate pack pass.  After an NCCL all-to-all the receive buffer
 * is n_xslabs = P consecutive x slabs, each chunk-major; stage2 (plan of shape (nx, ny/P, nz)) runs
 * the X pass fused with the binning for rows iy0 .. iy0+ny/P of the global cube; the caller
 * all-reduces sums_out / counts_out.  t21_workspace_bytes(plan, 0, nbins) sizes stage2's workspace. */
T21_API int t21_slab_pitch(const t21_plan* plan);
/* row length (elements) of the half-spectrum power arrays of t21_ps_half / t21_slab_stage2_nd: nz/2+1 rounded up to 8 */
T21_API int t21_half_pitch(const t21_plan* plan);
T21_API int t21_slab_stage1(const t21_plan* plan_slab, const void* x_slab, const void* offset_or_null,
                            void* w_out, void* send_out_or_null, int nranks, void* stream);
/* stage 1 with the re-partition FUSED into the Y pass: every transformed row is stored straight into the
 * y-slab buffer of the rank that owns it.  peer_cols[s] = rank s's chunk-major y-slab buffer (nx planes of ny/P rows) mapped into this
 * process (CUDA IPC / peer access over NVLink; peer_cols[own rank] is the local one); x0_global = global index
 * of this slab's first x plane.  No send buffer, no collective: the caller only has to order "every rank has
 * finished stage 1" before stage 2 (e.g. a one-element all-reduce on the same stream). */
/* buffers of the fused exchange: t21_peer_alloc makes one (cudaMalloc) and returns its 64-byte CUDA IPC handle;
 * a peer process calls t21_peer_open on that handle with its own device current. */
T21_API int t21_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64);
T21_API int t21_peer_open(const unsigned char* handle64, void** ptr);
T21_API int t21_peer_close(void* ptr);
T21_API int t21_peer_free(void* ptr);
/* A few-KB all-gather over peer-mapped memory that doubles as a barrier (no collective library): every rank owns a
 * buffer of t21_peer_comm_bytes() made with t21_peer_alloc (zero it once) and maps the others'.  t21_peer_signal
 * copies this rank's payload into slot (epoch & 1, rank) of every rank's mail area and then raises its flag there;
 * t21_peer_wait returns (on the stream) once all nranks payloads of `epoch` have landed in the caller's own buffer at
 * mail_offset + ((epoch & 1) * nranks + r) * slot_bytes.  Epochs start at 1 and grow by 1 per exchange on every rank.
 * peers_dev is a DEVICE array of the nranks mapped buffer addresses (own one included).  Work queued on the stream
 * before the signal (the Y pass's peer stores) is visible to every rank that has passed the matching wait. */
T21_API size_t t21_peer_comm_bytes(int nranks, int slot_bytes);
T21_API size_t t21_peer_comm_mail_offset(void);
T21_API int t21_peer_signal(void* const* peers_dev, int nranks, int rank, unsigned long long epoch, const void* payload_dev,
                            int nbytes, int slot_bytes, void* stream);
T21_API int t21_peer_wait(const void* my_buf, int nranks, unsigned long long epoch, int* timed_out_dev, double timeout_s,
                          void* stream);
/* Assembles the fftshifted n-D power of a slab-sharded cube from every rank's half spectrum (t21_slab_stage2_nd):
 * half[nx][ny/P][row_pitch] -> outs_dev[r] = rank r's output [nx][ny/P][nz] (peer-mapped; a DEVICE array of nranks
 * addresses).  A rank writes the kz >= 0 part of its own rows and stores the mirrored kz < 0 part straight into the
 * output of the rank that owns row -ky.  Follow it with a barrier (t21_peer_signal / t21_peer_wait) before reading.
 * Rows keep their FFT order along y (the caller knows which ky each local row is); with ONE rank, shift_y != 0 shifts
 * that axis too, which makes `outs_dev[0]` the complete fftshifted array of power_spectrum_nd (power_spectrum.py:45). */
T21_API int t21_nd_assemble(const void* half, int dtype, int nx, int nyl, int nz, int row_pitch, int iy0, int ny_global,
                            void* const* outs_dev, int nranks, int rank, int shift_y, void* stream);
/* c ~ mean(x) of n elements from a strided sample, written to a device scalar of x's type (the offset of the passes) */
T21_API int t21_mean_estimate(const void* x, long long n, int dtype, void* out_dev, void* stream);
T21_API int t21_slab_stage1_p2p(const t21_plan* plan_slab, const void* x_slab, const void* offset_or_null,
                                void* w_scratch, void* const* peer_cols, int nranks, long long x0_global,
                                void* stream);
T21_API int t21_slab_stage2(const t21_plan* plan_cols, const void* w1, const void* w2_or_null, int n_xslabs,
                            int iy0, int ny_global, const void* offset1, const void* offset2,
                            const t21_binspec* spec, double* sums_out, long long* counts_out,
                            void* workspace, size_t workspace_bytes, void* stream);

/* power_spectrum_nd / cross_power_spectrum_nd of a slab-sharded cube: the scaled power of this rank's
 * rows as the computed HALF spectrum in FFT order, out_half[nx][ny/P][t21_half_pitch(plan)].  The caller assembles
 * the fftshifted rows: the kz < 0 half of row ky is the mirrored half of row -ky, held by another rank. */
T21_API int t21_slab_stage2_nd(const t21_plan* plan_cols, const void* w1, const void* w2_or_null, int n_xslabs,
                               int iy0, int ny_global, const void* offset1, const void* offset2, double scale,
                               void* out_half, int out_dtype, void* stream);

/* radial_average / mu_binning (power_spectrum.py:98-134, 380-435) on a caller-supplied
 * fftshifted array of spec->n elements (1-/2-D arrays are padded with leading 1s). */
T21_API int t21_bin_only(const void* p_shifted, int dtype, const t21_binspec* spec, double* sums_out,
                 long long* counts_out, void* workspace, size_t workspace_bytes, void* stream);
T21_API size_t t21_bin_only_workspace_bytes(const t21_binspec* spec);

/* Per-pass timing for benchmarks: t21_profile(1) makes every call record CUDA events on the
 * caller's stream between passes; t21_profile_last fills ms_out[0..4] with the milliseconds of
 * {mean estimate, z pass, y pass, x pass, reduce} of the last call and returns 5. */
T21_API int t21_profile(int enable);
T21_API int t21_profile_last(float* ms_out, int cap);

/* power_spectrum_2d (power_spectrum.py:177-254): (k_perp, k_par) means of a fftshifted 3-D array.
 * perp_tab[n_a * n_b] / par_tab[n_nu] give the perpendicular / parallel bin (or -1) of every index
 * pair across / index along nu_axis; they are built on the host with the reference's own numpy / scipy
 * digitisation (binspec.build_cyl_tables).  sums_out / counts_out have nperp * npar entries. */
T21_API int t21_bin_cyl(const void* p_shifted, int dtype, int n0, int n1, int n2, int nu_axis,
                        const int* perp_tab, const int* par_tab, int nperp, int npar, double* sums_out,
                        long long* counts_out, void* workspace, size_t workspace_bytes, void* stream);
T21_API size_t t21_bin_cyl_workspace_bytes(int nperp, int npar);

/* The scaled power of the computed HALF spectrum only, FFT order: out_half[nx][ny][t21_half_pitch(plan)] (columns >= nz/2+1 are padding; a quarter of the bytes
 * of the full float64 array of t21_ps_nd, no Hermitian-twin stores).  It feeds the estimators whose bins are not
 * (k, mu) boxes, so that power_spectrum_2d (power_spectrum.py:177-254) and power_spectrum_multipoles
 * (power_legendre.py:4-90) never materialise the fftshifted float64 cube the reference builds (:45-46). */
T21_API int t21_ps_half(const t21_plan* plan, const void* x1, const void* x2_or_null, int subtract_mean, double scale,
                        void* out_half, int out_dtype, void* workspace, size_t workspace_bytes, void* stream);
/* The same PENCIL-MAJOR, out_t[ny][nz/2+1][nx] -- the layout the warp-per-pencil X pass writes with full 128-byte runs (fp32
 * transform, nx = 1024, one field; T21_EINVAL otherwise) -- and the pass that turns it into the complete fftshifted cube
 * of power_spectrum_nd (power_spectrum.py:45-54), transposing 32 x 32 tiles through shared memory. */
T21_API int t21_ps_half_t(const t21_plan* plan, const void* x1, int subtract_mean, double scale, void* out_t, int out_dtype,
                          void* workspace, size_t workspace_bytes, void* stream);
T21_API int t21_nd_assemble_t(const void* half_t, int dtype, int nx, int ny, int nz, void* out, void* stream);
/* t21_bin_cyl on that half spectrum (rows of `row_pitch` elements): perp_tab / par_tab in FFT index order and over
 * the full axes; a mode whose Hermitian twin is not stored stands for both (weight 2).  sums_out only: the mode
 * counts of separable bins are a product of two table histograms and are formed on the host. */
T21_API int t21_bin_cyl_half(const void* p_half, int dtype, int nx, int ny, int nz, int row_pitch, int nu_axis,
                             const int* perp_tab, const int* par_tab, int nperp, int npar, double* sums_out,
                             void* workspace, size_t workspace_bytes, void* stream);
/* power_spectrum_multipoles (power_legendre.py:55-87): per k bin of `spec` (shells, FFT order, half-open last bin,
 * tab_los / tab_losf filled for los_axis) sums_out[nk][5] = sum P, sum P L2, sum P L4, sum L2^2, sum L4^2 over the
 * modes of the bin (twins included) and counts_out[nk] = number of modes. */
T21_API int t21_bin_multipoles_half(const void* p_half, int dtype, int row_pitch, const t21_binspec* spec, int los_axis,
                                    int exclude_zero, double* sums_out, long long* counts_out, void* workspace,
                                    size_t workspace_bytes, void* stream);
T21_API size_t t21_bin_multipoles_workspace_bytes(int nk);
/* power_spectrum_multipoles (power_legendre.py:4-90) as ONE pipeline: Z / Y passes, then the warp-per-pencil X pass whose
 * epilogue adds the five sums above per k bin -- no power array is written.  fp32, nx = 1024, and `spec` (shells, FFT
 * order, half-open last bin, los_axis / exclude_zero set) must carry the segment table t21_build_pencil_segments made
 * for it; T21_EINVAL otherwise (take t21_ps_half + t21_bin_multipoles_half).  sums_out[nk][5], counts_out[nk];
 * workspace t21_workspace_bytes(plan, 1, 5 * nk). */
T21_API int t21_ps_multipoles(const t21_plan* plan, const void* x1, int subtract_mean, const t21_binspec* spec, int los_axis,
                              double* sums_out, long long* counts_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- input side ----
 * t21_h2d_staged: pageable HOST memory -> device through a pinned ring filled by `threads` host threads (0: choose) while
 * the previous chunk's DMA runs; returns when the last chunk is queued on `stream`.
 * t21_f2c: a Fortran-ordered blob (x fastest: xfrac_file.py:62, density_file.py:51) -> C order [nx][ny][nz].
 * t21_c2ray_dt: calc_dt (temperature.py:51-80, :215-222) fused with that transpose: out[x][y][z] =
 *   factor * (1 - x_i) * rho from the two blobs as stored (neutral != 0: the first blob holds 1 - x_i). */
T21_API int t21_h2d_staged(void* dst_dev, const void* src_host, size_t bytes, int threads, void* stream);
T21_API int t21_f2c(const void* in_f, int dtype, int nx, int ny, int nz, void* out_c, int out_dtype, void* stream);
T21_API int t21_c2ray_dt(const void* xfrac_f, int xfrac_dtype, int neutral, const void* dens_f, int dens_dtype, double factor,
                         int nx, int ny, int nz, void* out_c, int out_dtype, void* stream);

T21_API const char* t21_last_error(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
T21_API unsigned long long t21_launch_count(void);
T21_API int t21_version(void);

#ifdef __cplusplus
}
#endif
#endif /* T21PS_H */
