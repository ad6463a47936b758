// SIMT multi-head self attention over the tokens of one frame (fp32 parity path; also the cross-check
// for the tensor-core kernel).  One block per (frame, head): K and V of the head are staged in shared
// memory as fp32, each warp owns a strided set of queries; softmax statistics in fp32.
#include "common.cuh"

namespace distb200 {

namespace {

constexpr int HD = 64;
constexpr int ATT_WARPS = 8;

// CAUSAL: query i sees the keys j <= i only - the additive upper-triangular -inf mask of the CLIP text transformer
// (clip.py:404-410); masked scores contribute exp(-inf) = 0 exactly, so they are simply not visited.
template <typename T, bool CAUSAL>
__global__ void __launch_bounds__(ATT_WARPS * 32) attention_simt_kernel(const T* __restrict__ qkv, T* __restrict__ out, int tokens, int heads) {
    grid_dep_sync();
    extern __shared__ float smem[];
    const int D = heads * HD;
    const int f = blockIdx.x / heads, h = blockIdx.x % heads;
    float* Ks = smem;                                  // [tokens][HD+1]
    float* Vs = Ks + (size_t)tokens * (HD + 1);        // [tokens][HD]
    float* Qs = Vs + (size_t)tokens * HD;              // [warps][HD]
    float* Ps = Qs + ATT_WARPS * HD;                   // [warps][tokens]
    const T* base = qkv + (long long)f * tokens * 3 * D + h * HD;
    for (int i = threadIdx.x; i < tokens * HD; i += blockDim.x) {
        const int j = i / HD, dd = i % HD;
        Ks[j * (HD + 1) + dd] = to_float(base[(long long)j * 3 * D + D + dd]);
        Vs[j * HD + dd] = to_float(base[(long long)j * 3 * D + 2 * D + dd]);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* qs = Qs + warp * HD;
    float* ps = Ps + (size_t)warp * tokens;
    for (int i = warp; i < tokens; i += ATT_WARPS) {
        qs[lane] = to_float(base[(long long)i * 3 * D + lane]) * 0.125f;
        qs[lane + 32] = to_float(base[(long long)i * 3 * D + lane + 32]) * 0.125f;
        __syncwarp();
        const int keys = CAUSAL ? i + 1 : tokens;
        float mx = -INFINITY;
        for (int j = lane; j < keys; j += 32) {
            const float* kr = Ks + j * (HD + 1);
            float s = 0.f;
#pragma unroll 16
            for (int dd = 0; dd < HD; ++dd) s = fmaf(qs[dd], kr[dd], s);
            ps[j] = s;
            mx = fmaxf(mx, s);
        }
        mx = warp_max(mx);
        float den = 0.f;
        for (int j = lane; j < keys; j += 32) {
            const float e = __expf(ps[j] - mx);
            ps[j] = e;
            den += e;
        }
        den = warp_sum(den);
        __syncwarp();
        float o0 = 0.f, o1 = 0.f;
        for (int j = 0; j < keys; ++j) {
            const float pj = ps[j];
            o0 = fmaf(pj, Vs[j * HD + lane], o0);
            o1 = fmaf(pj, Vs[j * HD + lane + 32], o1);
        }
        const float inv = 1.f / den;
        T* o = out + ((long long)f * tokens + i) * D + h * HD;
        o[lane] = from_float<T>(o0 * inv);
        o[lane + 32] = from_float<T>(o1 * inv);
        __syncwarp();
    }
}

}  // namespace

template <typename T, bool CAUSAL>
static void launch_variant(const void* qkv, void* out, unsigned grid, size_t smem, int tokens, int heads, cudaStream_t stream) {
    static bool done_dev[DISTB200_MAX_DEVICES] = {};
    bool& done = done_dev[current_device()];
    if (!done) { cudaFuncSetAttribute(attention_simt_kernel<T, CAUSAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); done = true; }
    DISTB200_LAUNCH((attention_simt_kernel<T, CAUSAL>), grid, ATT_WARPS * 32, smem, stream, (const T*)qkv, (T*)out, tokens, heads);
}

int attention_simt_launch(const void* qkv, void* out, int frames, int tokens, int heads, int dtype, bool causal, cudaStream_t stream) {
    const size_t smem = ((size_t)tokens * (HD + 1) + (size_t)tokens * HD + ATT_WARPS * HD + (size_t)ATT_WARPS * tokens) * sizeof(float);
    DISTB200_REQUIRE(smem <= 227 * 1024, "attention(simt): %d tokens need %zu bytes of shared memory", tokens, smem);
    const unsigned grid = (unsigned)frames * heads;
    if (dtype == DISTB200_F32) {
        if (causal) launch_variant<float, true>(qkv, out, grid, smem, tokens, heads, stream);
        else launch_variant<float, false>(qkv, out, grid, smem, tokens, heads, stream);
    } else {
        if (causal) launch_variant<bf16, true>(qkv, out, grid, smem, tokens, heads, stream);
        else launch_variant<bf16, false>(qkv, out, grid, smem, tokens, heads, stream);
    }
    return check_launch("attention_simt");
}

}  // namespace distb200
