// Shared helpers for the distb200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/distb200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "distb200 is written for sm_100a (B200) only"
#endif

namespace distb200 {

// ---- error reporting across the C ABI (no exceptions) ------------------------------------------
void set_error(const char* fmt, ...);
int  check_launch(const char* what);   // cudaPeekAtLastError -> code + message

#define DISTB200_REQUIRE(cond, ...)                  \
    do {                                             \
        if (!(cond)) {                               \
            ::distb200::set_error(__VA_ARGS__);      \
            return 1;                                \
        }                                            \
    } while (0)

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float quick_gelu(float x) {
    // x * sigmoid(1.702 x)  (models/base/clip.py:199-201)
    return x / (1.0f + __expf(-1.702f * x));
}

__device__ __forceinline__ float to_float(float v) { return v; }
__device__ __forceinline__ float to_float(bf16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_float(float v);
template <> __device__ __forceinline__ float from_float<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_float<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 p = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&p);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

inline int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// implemented in gemm_tcgen05.cu / gemm_simt.cu
int gemm_tcgen05_launch(const distb200_gemm_desc& d, cudaStream_t stream);
int gemm_simt_launch(const distb200_gemm_desc& d, cudaStream_t stream);
int wgrad_simt_launch(const distb200_wgrad_desc& d, cudaStream_t stream);
int wgrad_tcgen05_launch(const distb200_wgrad_desc& d, cudaStream_t stream);

}  // namespace distb200
