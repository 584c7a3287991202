// wdm_tmap.h -- host-side TMA descriptor (CUtensorMap) construction without linking libcuda:
// the driver entry point is resolved through the runtime (cudaGetDriverEntryPoint).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace wdm {

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

inline PFN_tmapEncodeTiled tmap_encode_fn() {
    static PFN_tmapEncodeTiled fn = []() -> PFN_tmapEncodeTiled {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<PFN_tmapEncodeTiled>(p);
    }();
    return fn;
}

// dims/strides innermost first; strides[i] is the byte stride of dimension i+1 (rank-1 entries).
// Returns 0 on success, a CUresult (>0) on failure, -1 if the entry point is unavailable.
inline int make_tmap(CUtensorMap* out, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swz,
                     CUtensorMapL2promotion l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     const uint32_t* elem_strides = nullptr) {
    PFN_tmapEncodeTiled fn = tmap_encode_fn();
    if (!fn) return -1;
    cuuint64_t d[5], s[4];
    cuuint32_t b[5], e[5];
    for (int i = 0; i < rank; ++i) {
        d[i] = dims[i];
        b[i] = box[i];
        e[i] = elem_strides ? elem_strides[i] : 1;
        if (i + 1 < rank) s[i] = strides_bytes[i];
    }
    CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                    l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return (int)r;
}

}  // namespace wdm
