// Microbenchmark: raw tcgen05.mma issue rate (no loads, no epilogue). nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_peak mma_peak.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../dist_b200/csrc/ptx.cuh"
using namespace distb200;

template <int CTAS>
__global__ void __launch_bounds__(128, 1) mma_rate(int iters, int n, int stages, int mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ __align__(8) uint64_t dummy[8];
    __shared__ uint32_t slot;
    const uint32_t base = (ptx::smem_u32(smem) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = CTAS == 2 ? ptx::cluster_ctarank() : 0;
    for (int i = threadIdx.x; i < 48 * 1024 / 4 * stages; i += blockDim.x) reinterpret_cast<uint32_t*>(smem + (base - ptx::smem_u32(smem)))[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { ptx::mbar_init(ptx::smem_u32(&bar), 1); for (int i = 0; i < 8; ++i) ptx::mbar_init(ptx::smem_u32(&dummy[i]), 1); ptx::fence_barrier_init(); }
    ptx::fence_proxy_async();
    if (warp == 0) { if (CTAS == 2) { ptx::tmem_alloc_2sm(ptx::smem_u32(&slot), 512); ptx::tmem_relinquish_2sm(); } else { ptx::tmem_alloc(ptx::smem_u32(&slot), 512); ptx::tmem_relinquish(); } }
    ptx::tc_fence_before();
    if (CTAS == 2) ptx::cluster_sync(); else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = slot;
    if (warp == 1 && lane == 0 && rank == 0) {
        const uint32_t idesc = ptx::umma_idesc_bf16(128 * CTAS, n);
        for (int it = 0; it < iters; ++it) {
            const uint32_t sa = base + (uint32_t)(it % stages) * 48 * 1024;
            const uint64_t da = ptx::umma_desc_k_sw128(sa), db = ptx::umma_desc_k_sw128(sa + 16384);
            for (int ks = 0; ks < 4; ++ks) {
                if (CTAS == 2) ptx::mma_f16_ss_2sm(tmem + (it & 1) * 256, da + 2 * ks, db + 2 * ks, idesc, 1);
                else ptx::mma_f16_ss(tmem + (it & 1) * 256, da + 2 * ks, db + 2 * ks, idesc, 1);
            }
            if (mode >= 1) { if (CTAS == 2) ptx::mma_commit_2sm(ptx::smem_u32(&dummy[it & 7]), 1); else ptx::mma_commit(ptx::smem_u32(&dummy[it & 7])); }
            if (mode >= 2 && it >= 4) ptx::mbar_wait(ptx::smem_u32(&dummy[(it - 4) & 7]), ((it - 4) >> 3) & 1);   // wait for the commit of 4 k-blocks ago
        }
        if (CTAS == 2) ptx::mma_commit_2sm(ptx::smem_u32(&bar), 1); else ptx::mma_commit(ptx::smem_u32(&bar));
        ptx::mbar_wait(ptx::smem_u32(&bar), 0);
    }
    ptx::tc_fence_before();
    if (CTAS == 2) ptx::cluster_sync(); else __syncthreads();
    if (warp == 0) { ptx::tc_fence_after(); if (CTAS == 2) ptx::tmem_dealloc_2sm(tmem, 512); else ptx::tmem_dealloc(tmem, 512); }
}

template <int CTAS>
void run(int n, int stages, int mode = 0) {
    const int iters = 4000, smem = 48 * 1024 * stages + 2048;
    cudaFuncSetAttribute(mma_rate<CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = CTAS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        cudaLaunchKernelEx(&cfg, mma_rate<CTAS>, iters, n, stages, mode);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
    }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double flops = 2.0 * 128 * n * 64 * (double)iters * 148;
    printf("mode=%d ctas=%d N=%d stages=%d: %.3f ms  %.1f TFLOP/s  (%s)\n", mode, CTAS, n, stages, ms, flops / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    run<1>(256, 4, 0); run<1>(256, 4, 1); run<1>(256, 4, 2); run<2>(256, 4, 0); run<2>(256, 4, 1); run<2>(256, 4, 2);
    return 0;
}
