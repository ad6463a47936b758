// Shared helpers for the distb200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

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

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------------
// Every kernel of the library can be launched with cudaLaunchAttributeProgrammaticStreamSerialization and starts with
// grid_dep_sync(): its CTAs may become resident (and run their prologue: barrier init, TMEM allocation, descriptor
// prefetch) while the previous kernel of the stream drains, and they touch global memory only after that kernel has
// completed and flushed.  Measured on B200 inside the CUDA graph (B/16 8+16f, 276 launches): 15.60 ms with PDL vs
// 15.33 ms without - the persistent one-CTA-per-SM kernels leave no room for the next grid until they exit, so only the
// launch bookkeeping is added.  The attribute is therefore OFF by default (DISTB200_PDL=1 enables it); without it the
// device-side instructions are no-ops.
__device__ __forceinline__ void grid_dep_sync() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

inline bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("DISTB200_PDL");
        on = (e && atoi(e) != 0) ? 1 : 0;
    }
    return on != 0;
}

template <typename... KArgs, typename... Args>
inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);        // errors surface through check_launch()
}
#define DISTB200_LAUNCH(kernel, grid, block, smem, stream, ...) \
    ::distb200::launch_kernel(kernel, dim3(grid), dim3(block), (size_t)(smem), (cudaStream_t)(stream), __VA_ARGS__)

// Per-DEVICE caches (a process may drive several GPUs: attributes set with cudaFuncSetAttribute and the SM count belong to the
// current device, not to the process).
constexpr int DISTB200_MAX_DEVICES = 64;
inline int current_device() {
    int dev = 0;
    cudaGetDevice(&dev);
    return (dev >= 0 && dev < DISTB200_MAX_DEVICES) ? dev : 0;
}

inline int sm_count() {
    static int n[DISTB200_MAX_DEVICES] = {};
    const int dev = current_device();
    if (n[dev] == 0) {
        cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
        if (n[dev] <= 0) n[dev] = 148;
    }
    return n[dev];
}

// implemented in gemm_tcgen05.cu / gemm_simt.cu
int gemm_tcgen05_launch(const distb200_gemm_desc& d, cudaStream_t stream);
int gemm_simt_launch(const distb200_gemm_desc& d, cudaStream_t stream);
int wgrad_simt_launch(const distb200_wgrad_desc& d, cudaStream_t stream);
int wgrad_tcgen05_launch(const distb200_wgrad_desc& d, cudaStream_t stream);

}  // namespace distb200
