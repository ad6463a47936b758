// Host-side construction of TMA tensor maps (bf16, 128-byte swizzle), shared by the tcgen05 kernels.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace distb200 {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

inline int make_map(CUtensorMap* tm, const void* base, int rank, const long long* dims, const long long* strides_elems,
             const int* box, const char* what) {
    EncodeTiledFn fn = encode_fn();
    DISTB200_REQUIRE(fn != nullptr, "gemm(tcgen05): cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = (cuuint64_t)dims[i];
        bdim[i] = (cuuint32_t)box[i];
        estr[i] = 1;
        DISTB200_REQUIRE(dims[i] >= 1 && box[i] >= 1 && box[i] <= 256, "gemm(tcgen05): bad %s dim %d: size %lld box %d", what, i,
                        dims[i], box[i]);
    }
    for (int i = 1; i < rank; ++i) {
        long long s = strides_elems[i] * 2;
        if (dims[i] == 1 && s < 16) s = (i > 1 ? (long long)gstr[i - 2] : 16);   // unused dimension: any legal stride
        if (dims[i] == 1 && s % 16 != 0) s = 16;
        DISTB200_REQUIRE(s % 16 == 0 && s > 0, "gemm(tcgen05): %s stride %d (%lld bytes) must be a positive multiple of 16", what, i, s);
        gstr[i - 1] = (cuuint64_t)s;
    }
    DISTB200_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "gemm(tcgen05): %s base is not 16-byte aligned", what);
    CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bdim, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DISTB200_REQUIRE(r == CUDA_SUCCESS, "gemm(tcgen05): cuTensorMapEncodeTiled(%s) failed with %d", what, (int)r);
    return 0;
}

}  // namespace distb200
