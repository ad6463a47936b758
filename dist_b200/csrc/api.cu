// C ABI entry points that dispatch between implementations, plus error reporting.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace distb200 {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        cudaGetLastError();
        return 2;
    }
    return 0;
}

int attention_simt_launch(const void* qkv, void* out, int frames, int tokens, int heads, int dtype, bool causal, cudaStream_t stream);
int attention_tc_launch(const void* qkv, void* out, int frames, int tokens, int heads, cudaStream_t stream);

}  // namespace distb200

using namespace distb200;

extern "C" int distb200_version(void) { return DISTB200_VERSION; }
extern "C" int distb200_arch(void) { return 100; }
extern "C" const char* distb200_last_error(void) { return g_error; }

extern "C" int distb200_gemm(const distb200_gemm_desc* desc, void* stream) {
    DISTB200_REQUIRE(desc != nullptr, "gemm: null descriptor");
    const distb200_gemm_desc& d = *desc;
    DISTB200_REQUIRE(d.a && d.b, "gemm: null operand");
    DISTB200_REQUIRE(d.out || d.out2, "gemm: no output");
    DISTB200_REQUIRE(d.dtype == DISTB200_F32 || d.dtype == DISTB200_BF16, "gemm: unknown dtype %d", d.dtype);
    DISTB200_REQUIRE(d.num_taps >= 1 && d.num_taps <= DISTB200_MAX_TAPS, "gemm: num_taps=%d out of range", d.num_taps);
    DISTB200_REQUIRE(d.groups >= 0 && d.rows_per_group >= 0 && d.n >= 0 && d.k >= 1, "gemm: bad sizes");
    DISTB200_REQUIRE(d.out_rep >= 1, "gemm: out_rep must be >= 1");
    DISTB200_REQUIRE(d.group_dim == 0 || d.group_dim == 2 || d.group_dim == 3, "gemm: group_dim must be 2 or 3");
    cudaStream_t st = (cudaStream_t)stream;
    if (d.impl == DISTB200_IMPL_SIMT || d.dtype == DISTB200_F32) {
        DISTB200_REQUIRE(d.impl == DISTB200_IMPL_AUTO || d.impl == DISTB200_IMPL_SIMT, "gemm: the tcgen05 kernels need bf16 operands");
        DISTB200_REQUIRE(!d.stat_partials, "gemm: stat_partials is emitted by the tcgen05 epilogue only");
        return gemm_simt_launch(d, st);
    }
    return gemm_tcgen05_launch(d, st);
}

extern "C" int distb200_gemm_wgrad(const distb200_wgrad_desc* desc, void* stream) {
    DISTB200_REQUIRE(desc != nullptr, "gemm_wgrad: null descriptor");
    const distb200_wgrad_desc& d = *desc;
    DISTB200_REQUIRE(d.x && d.dy && d.dw, "gemm_wgrad: null pointer");
    DISTB200_REQUIRE(d.dtype == DISTB200_F32 || d.dtype == DISTB200_BF16, "gemm_wgrad: unknown dtype %d", d.dtype);
    DISTB200_REQUIRE(d.num_taps >= 1 && d.num_taps <= DISTB200_MAX_TAPS, "gemm_wgrad: num_taps=%d out of range", d.num_taps);
    DISTB200_REQUIRE(d.groups >= 0 && d.rows_per_group >= 0 && d.n >= 1 && d.k >= 1, "gemm_wgrad: bad sizes");
    DISTB200_REQUIRE(d.group_dim == 0 || d.group_dim == 2 || d.group_dim == 3, "gemm_wgrad: group_dim must be 2 or 3");
    DISTB200_REQUIRE(d.ld_dw >= d.k, "gemm_wgrad: ld_dw too small");
    cudaStream_t st = (cudaStream_t)stream;
    if (d.impl == DISTB200_IMPL_SIMT || d.dtype == DISTB200_F32) {
        DISTB200_REQUIRE(d.impl == DISTB200_IMPL_AUTO || d.impl == DISTB200_IMPL_SIMT, "gemm_wgrad: the tcgen05 kernel needs bf16 operands");
        return wgrad_simt_launch(d, st);
    }
    return wgrad_tcgen05_launch(d, st);
}

extern "C" int distb200_attention(const void* qkv, void* out, int32_t frames, int32_t tokens, int32_t heads, int32_t dtype,
                                  int32_t impl, void* stream) {
    if (frames == 0) return 0;
    DISTB200_REQUIRE(qkv && out, "attention: null pointer");
    DISTB200_REQUIRE(tokens >= 1 && heads >= 1, "attention: bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DISTB200_BF16 && impl != DISTB200_IMPL_SIMT) return attention_tc_launch(qkv, out, frames, tokens, heads, st);
    DISTB200_REQUIRE(impl != DISTB200_IMPL_TCGEN05, "attention: the tensor-core kernel needs bf16");
    return attention_simt_launch(qkv, out, frames, tokens, heads, dtype, false, st);
}

extern "C" int distb200_attention_causal(const void* qkv, void* out, int32_t seqs, int32_t tokens, int32_t heads, int32_t dtype, void* stream) {
    if (seqs == 0) return 0;
    DISTB200_REQUIRE(qkv && out, "attention_causal: null pointer");
    DISTB200_REQUIRE(tokens >= 1 && heads >= 1, "attention_causal: bad sizes");
    DISTB200_REQUIRE(dtype == DISTB200_F32 || dtype == DISTB200_BF16, "attention_causal: unknown dtype %d", dtype);
    return attention_simt_launch(qkv, out, seqs, tokens, heads, dtype, true, (cudaStream_t)stream);
}
