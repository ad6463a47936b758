// Fused softmax attention for the ViT (197 / 257 tokens, head dim 64) on tcgen05 tensor cores.
//
// One CTA per (frame, head, 128-query tile); 2 CTAs are co-resident per SM when the keys fit 256 TMEM columns,
// so one CTA's softmax overlaps the other's loads and MMAs.
//   warp 4 (one lane)  TMA: Q tile, all K rows and all V rows of the head, straight out of the packed qkv
//                      activation [frames, tokens, 3*D] (rows past the last token read as zero);
//                      S = Q K^T   tcgen05.mma SS, fp32 accumulator in TMEM columns [0, keys_pad)
//                      O = P V     tcgen05.mma TS: A = P read from TMEM, B = V tile used MN-major (no transpose)
//   warps 0..3         one query row per thread: row max, exp2, row sum in fp32 from tcgen05.ld; the bf16
//                      probabilities are written back over S with tcgen05.st (S never leaves the SM);
//                      epilogue scales O by 1/rowsum and stores bf16 [frames, tokens, D].
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"

namespace distb200 {

int attention_simt_launch(const void* qkv, void* out, int frames, int tokens, int heads, int dtype, cudaStream_t stream);

namespace {

constexpr int HD = 64;
constexpr int QT = 128;                 // query rows per CTA
constexpr int BOX_ROWS = 64;            // rows per TMA box
constexpr int ATT_THREADS = 160;

struct alignas(64) AttArgs {
    CUtensorMap tm;                     // qkv as (3*D, tokens, frames), box (64, 64, 1), 128B swizzle
    bf16* out;
    int tokens, heads, q_tiles;
    int keys_pad;                       // tokens rounded up to 16
    int kv_rows;                        // tokens rounded up to BOX_ROWS (smem rows per K / V tile)
    int tmem_cols, o_col;
};

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]),
          "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}

// A operand from TMEM (probabilities), B from shared memory
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

__global__ void __launch_bounds__(ATT_THREADS) attention_tc_kernel(const __grid_constant__ AttArgs args) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[4];       // load, s_ready, p_ready, o_ready
    __shared__ uint32_t tmem_slot;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qt = blockIdx.x % args.q_tiles;
    const int fh = blockIdx.x / args.q_tiles;
    const int h = fh % args.heads, f = fh / args.heads;
    const int D = args.heads * HD;

    const uint32_t s_q = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t s_k = s_q + QT * HD * 2;
    const uint32_t s_v = s_k + (uint32_t)args.kv_rows * HD * 2;
    const uint32_t bar_load = ptx::smem_u32(&bars[0]), bar_s = ptx::smem_u32(&bars[1]);
    const uint32_t bar_p = ptx::smem_u32(&bars[2]), bar_o = ptx::smem_u32(&bars[3]);

    if (warp == 4) {
        if (lane == 0) {
            ptx::prefetch_tensormap(&args.tm);
            ptx::mbar_init(bar_load, 1);
            ptx::mbar_init(bar_s, 1);
            ptx::mbar_init(bar_p, 128);
            ptx::mbar_init(bar_o, 1);
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), (uint32_t)args.tmem_cols);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const int q0 = qt * QT;

    if (warp == 4) {
        if (lane == 0) {
            const int kv_boxes = args.kv_rows / BOX_ROWS;
            const int q_boxes = QT / BOX_ROWS;
            ptx::mbar_arrive_expect_tx(bar_load, (uint32_t)((q_boxes + 2 * kv_boxes) * BOX_ROWS * HD * 2));
            for (int i = 0; i < q_boxes; ++i)
                ptx::tma_load_3d(s_q + i * BOX_ROWS * HD * 2, &args.tm, bar_load, h * HD, q0 + i * BOX_ROWS, f);
            for (int i = 0; i < kv_boxes; ++i) {
                ptx::tma_load_3d(s_k + i * BOX_ROWS * HD * 2, &args.tm, bar_load, D + h * HD, i * BOX_ROWS, f);
                ptx::tma_load_3d(s_v + i * BOX_ROWS * HD * 2, &args.tm, bar_load, 2 * D + h * HD, i * BOX_ROWS, f);
            }
            ptx::mbar_wait(bar_load, 0);
            ptx::tc_fence_after();
            // ---- S = Q K^T : M=128, N=keys_pad (split at 256), K=64 ----
            const uint64_t dq = ptx::umma_desc_k_sw128(s_q);
            for (int n0 = 0; n0 < args.keys_pad; n0 += 256) {
                const int nn = min(256, args.keys_pad - n0);
                const uint32_t idesc = ptx::umma_idesc_bf16(QT, nn);
                const uint64_t dk = ptx::umma_desc_k_sw128(s_k + (uint32_t)n0 * HD * 2);
                for (int ks = 0; ks < HD / 16; ++ks)
                    ptx::mma_f16_ss(tmem + (uint32_t)n0, dq + (uint64_t)(2 * ks), dk + (uint64_t)(2 * ks), idesc, ks != 0);
            }
            ptx::mma_commit(bar_s);
            // ---- O = P V : A = P in TMEM (bf16 pairs, 8 columns per 16 keys), B = V rows [16ks, 16ks+16) MN-major ----
            ptx::mbar_wait(bar_p, 0);
            ptx::tc_fence_after();
            const uint32_t idesc_pv = ptx::umma_idesc_bf16(QT, HD) | (1u << 16);      // B is MN-major
            const int ksteps = args.keys_pad / 16;
            for (int ks = 0; ks < ksteps; ++ks) {
                const uint64_t dv = ptx::umma_desc_k_sw128(s_v + (uint32_t)ks * 16 * HD * 2);
                mma_f16_ts(tmem + (uint32_t)args.o_col, tmem + (uint32_t)(8 * ks), dv, idesc_pv, ks != 0);
            }
            ptx::mma_commit(bar_o);
        }
    } else {
        // ---- softmax: thread = one query row = one TMEM lane ----
        const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
        const int row = q0 + warp * 32 + lane;
        ptx::mbar_wait(bar_s, 0);
        ptx::tc_fence_after();
        const float c = 0.125f * 1.4426950408889634f;     // 1/sqrt(64) * log2(e)
        const int N = args.tokens;
        float mx = -INFINITY;
        for (int c0 = 0; c0 < args.keys_pad; c0 += 16) {
            uint32_t v[16];
            ptx::tmem_ld16(lane_base + (uint32_t)c0, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (c0 + i < N) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
        const float mxs = mx * c;
        float sum = 0.f;
        for (int c0 = 0; c0 < args.keys_pad; c0 += 32) {
            uint32_t v[32];
            const bool two = c0 + 16 < args.keys_pad;
            ptx::tmem_ld16(lane_base + (uint32_t)c0, v);
            if (two) ptx::tmem_ld16(lane_base + (uint32_t)c0 + 16u, v + 16);
            ptx::tmem_ld_wait();
            uint32_t p[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                const int k0 = c0 + 2 * i;
                float e0 = (k0 < N && (i < 8 || two)) ? exp2f(fmaf(__uint_as_float(v[2 * i]), c, -mxs)) : 0.f;
                float e1 = (k0 + 1 < N && (i < 8 || two)) ? exp2f(fmaf(__uint_as_float(v[2 * i + 1]), c, -mxs)) : 0.f;
                // the row sum must match what the tensor core sees: sum the bf16-rounded probabilities
                const __nv_bfloat162 pr = __floats2bfloat162_rn(e0, e1);
                sum += __bfloat162float(pr.x) + __bfloat162float(pr.y);
                p[i] = *reinterpret_cast<const uint32_t*>(&pr);
            }
            // P chunk covers keys [c0, c0+32) = 16 packed columns starting at c0/2 (always behind the S read front)
            tmem_st16(lane_base + (uint32_t)(c0 >> 1), p);
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        ptx::mbar_arrive(bar_p);
        // ---- epilogue ----
        ptx::mbar_wait(bar_o, 0);
        ptx::tc_fence_after();
        const float inv = 1.f / sum;
        uint32_t o[64];
        ptx::tmem_ld32(lane_base + (uint32_t)args.o_col, o);
        ptx::tmem_ld32(lane_base + (uint32_t)args.o_col + 32u, o + 32);
        ptx::tmem_ld_wait();
        if (row < N) {
            uint4* dst = reinterpret_cast<uint4*>(args.out + ((long long)f * N + row) * D + h * HD);
#pragma unroll
            for (int i = 0; i < 8; ++i)
                dst[i] = make_uint4(pack_bf16x2(__uint_as_float(o[8 * i]) * inv, __uint_as_float(o[8 * i + 1]) * inv),
                                    pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv),
                                    pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv),
                                    pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv));
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, (uint32_t)args.tmem_cols);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

int attention_tc_launch(const void* qkv, void* out, int frames, int tokens, int heads, cudaStream_t stream) {
    const int keys_pad = (tokens + 15) / 16 * 16;
    if (keys_pad > 512 - 64 || ((uintptr_t)qkv & 15) || ((uintptr_t)out & 15))
        return attention_simt_launch(qkv, out, frames, tokens, heads, DISTB200_BF16, stream);

    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            encode = reinterpret_cast<EncodeTiledFn>(p);
        DISTB200_REQUIRE(encode != nullptr, "attention(tcgen05): cuTensorMapEncodeTiled is not available");
    }
    AttArgs args;
    const int D = heads * HD;
    cuuint64_t gdim[3] = {(cuuint64_t)(3 * D), (cuuint64_t)tokens, (cuuint64_t)frames};
    cuuint64_t gstr[2] = {(cuuint64_t)(3 * D) * 2, (cuuint64_t)(3 * D) * 2 * (cuuint64_t)tokens};
    cuuint32_t box[3] = {HD, BOX_ROWS, 1}, estr[3] = {1, 1, 1};
    CUresult r = encode(&args.tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv), gdim, gstr, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DISTB200_REQUIRE(r == CUDA_SUCCESS, "attention(tcgen05): cuTensorMapEncodeTiled failed with %d", (int)r);
    args.out = reinterpret_cast<bf16*>(out);
    args.tokens = tokens;
    args.heads = heads;
    args.q_tiles = (tokens + QT - 1) / QT;
    args.keys_pad = keys_pad;
    args.kv_rows = (tokens + BOX_ROWS - 1) / BOX_ROWS * BOX_ROWS;
    args.o_col = (keys_pad / 2 + 31) / 32 * 32;
    int need = args.o_col + HD > keys_pad ? args.o_col + HD : keys_pad;
    int cols = 32;
    while (cols < need) cols *= 2;
    args.tmem_cols = cols;
    const int smem = QT * HD * 2 + 2 * args.kv_rows * HD * 2 + 1024;
    static int smem_set = 0;
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        DISTB200_REQUIRE(e == cudaSuccess, "attention(tcgen05): cannot reserve %d bytes of shared memory: %s", smem, cudaGetErrorString(e));
        smem_set = smem;
    }
    const long long grid = (long long)frames * heads * args.q_tiles;
    DISTB200_REQUIRE(grid < 2147483647LL, "attention(tcgen05): grid too large");
    attention_tc_kernel<<<(unsigned)grid, ATT_THREADS, smem, stream>>>(args);
    return check_launch("attention_tc");
}

}  // namespace distb200
