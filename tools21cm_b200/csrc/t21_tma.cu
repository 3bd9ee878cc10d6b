// Host side of the TMA path: cuTensorMapEncodeTiled is reached through the runtime's driver entry point, so the
// library needs no link-time dependency on libcuda.
#include "t21_internal.h"
#include "t21_tma.cuh"

namespace t21 {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

int make_wmap_f32(CUtensorMap* out, const void* w, int nxs, int nya, int nch, int nslab, int box_rows, int box_floats) {
    EncodeTiledFn enc = encode_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this driver"); return T21_ECUDA; }
    if (((uintptr_t)w & 63) || box_rows < 1 || box_rows > 256 || nxs % box_rows || (box_floats != 4 && box_floats != 16)) {
        set_error("bad tensor-map geometry"); return T21_EINVAL;
    }
    const cuuint64_t dims[5] = {16, (cuuint64_t)nya, (cuuint64_t)nxs, (cuuint64_t)nch, (cuuint64_t)nslab};
    const cuuint64_t strides[4] = {64, (cuuint64_t)nya * 64, (cuuint64_t)nxs * nya * 64, (cuuint64_t)nch * nxs * nya * 64};
    const cuuint32_t box[5] = {(cuuint32_t)box_floats, 1, (cuuint32_t)box_rows, 1, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<void*>(w), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return T21_ECUDA; }
    return 0;
}

}  // namespace t21
