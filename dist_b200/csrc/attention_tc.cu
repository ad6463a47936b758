// Fused softmax attention for the ViT (197 / 257 tokens, head dim 64) on tcgen05 tensor cores.
//
// Persistent kernel, one CTA per SM, work item = (frame, head): K and V of the head are loaded once and shared by
// the item's 128-row query tiles.  Roles:
//   warp 8  (one lane)  TMA producer: K/V of the next item (2 smem stages) and Q tiles (ring of 4) straight out of
//                       the packed qkv activation [frames, tokens, 3*D]; rows past the last token read as zero
//   warp 9  (one lane)  S = Q K^T   tcgen05.mma SS -> fp32 in a TMEM slot (keys_pad columns)
//   warp 10 (one lane)  O = P V     tcgen05.mma TS: A = P read from TMEM, B = V used MN-major (no transpose)
//                       (two issuing threads, so that neither product waits behind the other one's barrier)
//   warps 0-3 / 4-7     two softmax groups, one per TMEM slot (ping-pong): one query row per thread; row max, exp2,
//                       row sum in fp32 from tcgen05.ld; bf16 probabilities are written back over S with tcgen05.st
//                       (S and P never leave the SM); epilogue scales O by 1/rowsum and stores bf16.
// With 257 tokens a slot needs 272 columns and only one fits the 512-column TMEM: group 1 idles and the S MMA,
// softmax and PV MMA of a tile run back to back (the loads still run ahead).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace distb200 {

int attention_simt_launch(const void* qkv, void* out, int frames, int tokens, int heads, int dtype, bool causal, cudaStream_t stream);

namespace {

constexpr int HD = 64;
constexpr int QT = 128;                 // query rows per tile
constexpr int Q_RING = 4;
#ifndef DISTB200_ATT_KV_STAGES
#define DISTB200_ATT_KV_STAGES 2
#endif
constexpr int KV_STAGES = DISTB200_ATT_KV_STAGES;
constexpr int ATT_THREADS = 352;
#ifndef DISTB200_ATT_HELD
#define DISTB200_ATT_HELD 1                // measured (B200, 197 / 257 tokens): 0 -> 0.117 / 0.897 ms, 1 -> 0.104 / 0.828 ms, 2 -> 0.131 / 0.986 ms (spills
                                          // at the 168-register cap of 11 warps), 3 -> 0.145 / 1.441 ms
#endif
#ifndef DISTB200_ATT_SETMAXNREG
#define DISTB200_ATT_SETMAXNREG 0         // tried: ptxas then spills in every role (even HELD = 1: 424 bytes), slower
#endif
constexpr int HELD = DISTB200_ATT_HELD;       // 32-column score chunks kept in registers between the two softmax passes
constexpr uint32_t Q_TILE_BYTES = QT * HD * 2;

struct alignas(64) AttArgs {
    CUtensorMap tm64;                   // qkv as (3*D, tokens, frames), box (64, 64, 1), 128B swizzle
    CUtensorMap tm16;                   // same tensor, box (64, 16, 1) for the ragged tail of K / V
    bf16* out;
    int tokens, heads, q_tiles;
    int keys_pad;                       // keys on the tensor-core path (multiple of 16)
    int keys_ld;                        // K / V rows held in shared memory (tokens rounded up to 16)
    int odd;                            // 1: the last key (index keys_pad) is handled on the CUDA cores, see below
    int single;                         // 1: the item's last tile holds ONE query row (tokens % 128 == 1): that row is computed on the CUDA cores
    int cosign;                         // odd | single: the softmax threads read Q / K / V from shared memory and co-sign their release
    int slot_cols, o_col, n_slots;
    int split_col;                      // S columns [0, split_col) are consumed before O (which aliases the S tail) may be written
    long long items;                    // frames * heads
    int dbg;                            // -DDISTB200_ATT_PROBES + DISTB200_ATT_DBG: 1 no max pass, 2 no exponentials, 4 no O read-out, 8 no P write-back, 16 no loads
};

// Bottleneck probes and a clock64 timeline (tools/trace_attention.py), compiled in only with -DDISTB200_ATT_PROBES.
#ifdef DISTB200_ATT_PROBES
#define ATT_PROBE(bit) (args.dbg & (bit))
__device__ long long* g_att_trace = nullptr;      // [cta][tile < 64][16]
#define ATT_STAMP(g, slot) do { if (g_att_trace && (g) < 64) g_att_trace[((long long)blockIdx.x * 64 + (g)) * 16 + (slot)] = clock64(); } while (0)
#else
#define ATT_PROBE(bit) false
#define ATT_STAMP(g, slot) do { } while (0)
#endif

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

__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// 2^x for x <= 0 on the FMA / ALU pipes only (no MUFU, no F2I): round-to-nearest split with the 1.5 * 2^23 constant, degree-3 polynomial
// for 2^f on [-0.5, 0.5] (max relative error 7.7e-5, the probabilities are rounded to bf16 afterwards), exponent added as an integer.
// Both softmax groups of an SM are in the exponential pass together for most of an item and MUFU.EX2 runs at 16 / clock / SM: every
// DISTB200_ATT_POLY-th exponential goes here instead (the FlashAttention-4 trick).
#ifndef DISTB200_ATT_POLY
#define DISTB200_ATT_POLY 0
#endif
__device__ __forceinline__ float poly_exp2(float x) {
    x = fmaxf(x, -125.0f);
    const float t = x + 12582912.0f;
    const float f = x - (t - 12582912.0f);
    float p = fmaf(f, 0.05508868f, 0.24260405f);
    p = fmaf(p, f, 0.69327623f);
    p = fmaf(p, f, 0.99992895f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// idx = position of the element in its (fully unrolled) chunk: the choice folds at compile time
__device__ __forceinline__ float exp2_sel(float x, int idx) {
    if (DISTB200_ATT_POLY > 0 && (idx % (DISTB200_ATT_POLY > 0 ? DISTB200_ATT_POLY : 1)) == DISTB200_ATT_POLY - 1) return poly_exp2(x);
    return fast_exp2(x);
}

// One query row against all keys of the item on the CUDA cores (see the call site): the 128 threads of a softmax group, Q row / K / V
// tiles in their 128-byte-swizzled shared-memory layout.  The query row is re-read (broadcast) per key instead of living in 64
// registers: the main loop runs at the register cap.
__device__ __forceinline__ void single_row_tile(const uint8_t* qrow, const uint8_t* kbase, const uint8_t* vbase, float* P, float* red, float* osc,
                                                int NR, int grp, bf16* dst) {
    const int lane = threadIdx.x & 31, w4 = (threadIdx.x >> 5) & 3, t = w4 * 32 + lane;
    const float c = 0.125f * 1.4426950408889634f;
    float mx = -INFINITY;
#pragma unroll 1
    for (int k = t; k < NR; k += 128) {
        const uint4* kr = reinterpret_cast<const uint4*>(kbase + k * 128);
        float a0 = 0.f, a1 = 0.f;
#pragma unroll 2
        for (int j = 0; j < 8; ++j) {
            const uint4 ka = kr[j ^ (k & 7)], qa = reinterpret_cast<const uint4*>(qrow)[j];       // row 0 of the Q tile: not swizzled
            const uint32_t kw[4] = {ka.x, ka.y, ka.z, ka.w}, qw[4] = {qa.x, qa.y, qa.z, qa.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                a0 = fmaf(__uint_as_float(qw[e] << 16), __uint_as_float(kw[e] << 16), a0);
                a1 = fmaf(__uint_as_float(qw[e] & 0xffff0000u), __uint_as_float(kw[e] & 0xffff0000u), a1);
            }
        }
        const float sc = a0 + a1;
        P[k] = sc;
        mx = fmaxf(mx, sc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) red[w4] = mx;
    asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    const float mxs1 = mx * c;
    float part = 0.f;
#pragma unroll 1
    for (int k = t; k < NR; k += 128) {
        const float e = fast_exp2(fmaf(P[k], c, -mxs1));
        part += e;
        P[k] = __bfloat162float(__float2bfloat16_rn(e));      // the tensor-core path multiplies bf16 probabilities too
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if (lane == 0) red[4 + w4] = part;
    asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
    const float inv1 = 1.f / ((red[4] + red[5]) + (red[6] + red[7]));
    const int dg = t & 7, kg = t >> 3;                       // 8 output columns x one of 16 key groups
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll 2
    for (int k = kg; k < NR; k += 16) {
        const float pk = P[k];
        const uint4 va = reinterpret_cast<const uint4*>(vbase + k * 128)[dg ^ (k & 7)];
        const uint32_t vw[4] = {va.x, va.y, va.z, va.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            acc[2 * e] = fmaf(pk, __uint_as_float(vw[e] << 16), acc[2 * e]);
            acc[2 * e + 1] = fmaf(pk, __uint_as_float(vw[e] & 0xffff0000u), acc[2 * e + 1]);
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) osc[kg * HD + 8 * dg + e] = acc[e];
    asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");
    if (t < HD) {
        float o = 0.f;
#pragma unroll
        for (int j = 0; j < 16; ++j) o += osc[j * HD + t];
        dst[t] = __float2bfloat16_rn(o * inv1);
    }
    asm volatile("bar.sync %0, 128;" ::"r"(grp + 1) : "memory");       // the scratch is free for this group's next single-row tile
}

struct Bars {   // shared-memory addresses of the mbarriers
    uint32_t kv_full, kv_empty, q_full, q_empty, s_full, p_full, o_full, slot_free, p_early;
};

template <bool SINGLE>          // SINGLE: the item's last tile holds one query row and is computed on the CUDA cores (args.single)
__global__ void __launch_bounds__(ATT_THREADS, 1) attention_tc_kernel(const __grid_constant__ AttArgs args) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_mem[2 * KV_STAGES + 2 * Q_RING + 10];
    __shared__ uint32_t tmem_slot;
    __shared__ float sr_p[SINGLE ? 2 : 1][SINGLE ? 288 : 1], sr_red[SINGLE ? 2 : 1][8], sr_o[SINGLE ? 2 : 1][SINGLE ? 16 : 1][HD];      // single-row tiles: scores / probabilities, reductions, partial outputs

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = args.heads * HD;
    const uint32_t kv_bytes = (uint32_t)args.keys_ld * HD * 2;           // one of K / V
    const uint32_t s_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t s_q = s_base;                                          // Q_RING tiles
    const uint32_t s_kv = s_q + Q_RING * Q_TILE_BYTES;                    // KV_STAGES x (K | V)
    Bars b;
    b.kv_full = ptx::smem_u32(&bar_mem[0]);
    b.kv_empty = b.kv_full + 8 * KV_STAGES;
    b.q_full = b.kv_empty + 8 * KV_STAGES;
    b.q_empty = b.q_full + 8 * Q_RING;
    b.s_full = b.q_empty + 8 * Q_RING;
    b.p_full = b.s_full + 16;
    b.o_full = b.p_full + 16;
    b.slot_free = b.o_full + 16;
    b.p_early = b.slot_free + 16;

    if (warp == 9) {
        if (lane == 0) {
            ptx::prefetch_tensormap(&args.tm64);
            ptx::prefetch_tensormap(&args.tm16);
            // odd-key mode: the epilogue of an item's last tile still reads the odd V row, so its 128 threads co-sign the release of K / V
            for (int i = 0; i < KV_STAGES; ++i) { ptx::mbar_init(b.kv_full + 8 * i, 1); ptx::mbar_init(b.kv_empty + 8 * i, args.cosign ? 129 : 1); }
            // odd-key mode: the softmax threads read their Q row from shared memory, so they co-sign the release of a Q tile
            for (int i = 0; i < Q_RING; ++i) { ptx::mbar_init(b.q_full + 8 * i, 1); ptx::mbar_init(b.q_empty + 8 * i, args.cosign ? 129 : 1); }
            for (int i = 0; i < 2; ++i) {
                ptx::mbar_init(b.s_full + 8 * i, 1);
                ptx::mbar_init(b.p_full + 8 * i, 128);
                ptx::mbar_init(b.o_full + 8 * i, 1);
                ptx::mbar_init(b.slot_free + 8 * i, 128);
                ptx::mbar_init(b.p_early + 8 * i, 128);
            }
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem = tmem_slot;
    grid_dep_sync();          // PDL: the prologue above overlaps the previous kernel's tail
    const int QTn = args.q_tiles;
    const int ns = args.n_slots;
    long long my_items = (args.items - blockIdx.x + gridDim.x - 1) / gridDim.x;
    if ((long long)blockIdx.x >= args.items) my_items = 0;
    const long long total = my_items * QTn;             // tiles of this CTA, g = item_index * q_tiles + qt

    // Register split (setmaxnreg): the three issuing warps need a few dozen registers, the softmax warps keep HELD x 32 scores per
    // thread between their two passes.  11 warps x 168 >= 8 x 200 + 3 x 80.
#if DISTB200_ATT_SETMAXNREG
    if (warp >= 8) asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
#endif
    if (warp == 8) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            const int boxes64 = args.keys_ld / 64, tail16 = (args.keys_ld % 64) / 16;
            long long g = 0;
            int it = 0;
            for (long long item = blockIdx.x; item < args.items; item += gridDim.x, ++it) {
                const int h = (int)(item % args.heads), f = (int)(item / args.heads);
                const int st = it % KV_STAGES;
                const uint32_t ph = (uint32_t)(it / KV_STAGES) & 1u;
                ptx::mbar_wait(b.kv_empty + 8 * st, ph ^ 1u);
                if (ATT_PROBE(16)) ptx::mbar_arrive(b.kv_full + 8 * st); else ptx::mbar_arrive_expect_tx(b.kv_full + 8 * st, 2 * kv_bytes);
                const uint32_t sk = s_kv + (uint32_t)st * 2 * kv_bytes, sv = sk + kv_bytes;
                for (int i = 0; i < (ATT_PROBE(16) ? 0 : boxes64); ++i) {
                    ptx::tma_load_3d(sk + i * 64 * HD * 2, &args.tm64, b.kv_full + 8 * st, D + h * HD, i * 64, f);
                    ptx::tma_load_3d(sv + i * 64 * HD * 2, &args.tm64, b.kv_full + 8 * st, 2 * D + h * HD, i * 64, f);
                }
                for (int i = 0; i < (ATT_PROBE(16) ? 0 : tail16); ++i) {
                    const int r = boxes64 * 64 + i * 16;
                    ptx::tma_load_3d(sk + r * HD * 2, &args.tm16, b.kv_full + 8 * st, D + h * HD, r, f);
                    ptx::tma_load_3d(sv + r * HD * 2, &args.tm16, b.kv_full + 8 * st, 2 * D + h * HD, r, f);
                }
                for (int qt = 0; qt < QTn; ++qt, ++g) {
                    const int qs = (int)(g % Q_RING);
                    const uint32_t qph = (uint32_t)(g / Q_RING) & 1u;
                    ptx::mbar_wait(b.q_empty + 8 * qs, qph ^ 1u);
                    if (ATT_PROBE(16)) { ptx::mbar_arrive(b.q_full + 8 * qs); continue; }
                    ptx::mbar_arrive_expect_tx(b.q_full + 8 * qs, Q_TILE_BYTES);
                    ptx::tma_load_3d(s_q + qs * Q_TILE_BYTES, &args.tm64, b.q_full + 8 * qs, h * HD, qt * QT, f);
                    ptx::tma_load_3d(s_q + qs * Q_TILE_BYTES + 64 * HD * 2, &args.tm64, b.q_full + 8 * qs, h * HD, qt * QT + 64, f);
                }
            }
        }
    } else if (warp == 9) {
        // ===================== S = Q K^T issuer =====================
        if (lane == 0) {
            for (uint32_t g = 0; g < (uint32_t)total; ++g) {
                const int it = (int)(g / (uint32_t)QTn), qt = (int)(g % (uint32_t)QTn);
                const int st = it % KV_STAGES, qs = (int)(g % Q_RING), slot = (int)(g % (uint32_t)ns);
                if (qt == 0) ptx::mbar_wait(b.kv_full + 8 * st, (uint32_t)(it / KV_STAGES) & 1u);
                ptx::mbar_wait(b.q_full + 8 * qs, (uint32_t)(g / Q_RING) & 1u);
                ATT_STAMP(g, 0);
                ptx::mbar_wait(b.slot_free + 8 * slot, ((uint32_t)(g / ns) & 1u) ^ 1u);
                ptx::tc_fence_after();
                ATT_STAMP(g, 1);
                const uint32_t sk = s_kv + (uint32_t)st * 2 * kv_bytes;
                const uint64_t dq = ptx::umma_desc_k_sw128(s_q + qs * Q_TILE_BYTES);
                const uint32_t ts = tmem + (uint32_t)(slot * args.slot_cols);
                const bool single = SINGLE && qt == QTn - 1;               // computed by the softmax group on the CUDA cores: no MMA, same hand-shakes
                for (int n0 = 0; n0 < (single ? 0 : args.keys_pad); n0 += 256) {
                    const int nn = min(256, args.keys_pad - n0);
                    const uint32_t idesc = ptx::umma_idesc_bf16(QT, nn);
                    const uint64_t dk = ptx::umma_desc_k_sw128(sk + (uint32_t)n0 * HD * 2);
                    for (int ks = 0; ks < HD / 16; ++ks)
                        ptx::mma_f16_ss(ts + (uint32_t)n0, dq + (uint64_t)(2 * ks), dk + (uint64_t)(2 * ks), idesc, ks != 0);
                }
                ptx::mma_commit(b.s_full + 8 * slot);
                ptx::mma_commit(b.q_empty + 8 * qs);
            }
        }
    } else if (warp == 10) {
        // ===================== O = P V issuer =====================
        if (lane == 0) {
            const uint32_t idesc_pv = ptx::umma_idesc_bf16(QT, HD) | (1u << 16);      // B (= V) is MN-major
            const int ksteps_pv = args.keys_pad / 16;
            for (uint32_t g = 0; g < (uint32_t)total; ++g) {
                const int it = (int)(g / (uint32_t)QTn), qt = (int)(g % (uint32_t)QTn);
                const int st = it % KV_STAGES, slot = (int)(g % (uint32_t)ns);
                // The probabilities arrive in two batches: p_early as soon as the softmax has read every S column that the O
                // accumulator aliases (so that most of P V overlaps the remaining exponentials), p_full when the row is done.
                const uint32_t ph = (uint32_t)(g / ns) & 1u;
                const uint32_t sv = s_kv + (uint32_t)st * 2 * kv_bytes + kv_bytes;
                const uint32_t ts = tmem + (uint32_t)(slot * args.slot_cols);
                const int ks_split = args.split_col < args.keys_pad ? args.split_col / 16 : 0;
                const bool single = SINGLE && qt == QTn - 1;
                if (ks_split > 0) {
                    ptx::mbar_wait(b.p_early + 8 * slot, ph);
                    ptx::tc_fence_after();
                    for (int ks = 0; ks < (single ? 0 : ks_split); ++ks) {
                        const uint64_t dv = ptx::umma_desc_k_sw128(sv + (uint32_t)ks * 16 * HD * 2);
                        mma_f16_ts(ts + (uint32_t)args.o_col, ts + (uint32_t)(8 * ks), dv, idesc_pv, ks != 0);
                    }
                }
                ptx::mbar_wait(b.p_full + 8 * slot, ph);
                ptx::tc_fence_after();
                ATT_STAMP(g, 2);
                for (int ks = ks_split; ks < (single ? 0 : ksteps_pv); ++ks) {
                    const uint64_t dv = ptx::umma_desc_k_sw128(sv + (uint32_t)ks * 16 * HD * 2);
                    mma_f16_ts(ts + (uint32_t)args.o_col, ts + (uint32_t)(8 * ks), dv, idesc_pv, ks != 0);
                }
                ptx::mma_commit(b.o_full + 8 * slot);
                if (qt == QTn - 1) ptx::mma_commit(b.kv_empty + 8 * st);        // K and V of the item are no longer needed
            }
        }
    } else if (warp < 8) {
        // ===================== softmax groups =====================
        const int grp = warp >> 2;                         // 0 or 1 = TMEM slot
        const int w4 = warp & 3;
        if (grp < ns) {
            const uint32_t ts = tmem + ((uint32_t)(w4 * 32) << 16) + (uint32_t)(grp * args.slot_cols);
            const float c = 0.125f * 1.4426950408889634f;  // 1/sqrt(64) * log2(e)
            const int N = args.tokens - args.odd;          // keys on the tensor-core path
            const int NR = args.tokens;                    // query rows
            const int full_end = (N / 32) * 32;            // keys [0, full_end) need no validity predicate
            uint32_t use = 0;
            // 32-bit index arithmetic (the host guarantees items * q_tiles < 2^31): 64-bit divisions cost ~10 % of a tile here
            const uint32_t total32 = (uint32_t)total, qtn = (uint32_t)QTn, heads32 = (uint32_t)args.heads;
            for (uint32_t g = (uint32_t)grp; g < total32; g += (uint32_t)ns, ++use) {
                const uint32_t it = g / qtn;
                const int qt = (int)(g - it * qtn);
                const uint32_t item = blockIdx.x + it * gridDim.x;
                const uint32_t f32i = item / heads32;
                const int h = (int)(item - f32i * heads32), f = (int)f32i;
                const uint32_t ph = use & 1u;
                const int row = qt * QT + w4 * 32 + lane;
                bool warp_has_rows = qt * QT + w4 * 32 < NR;
                if (w4 == 0 && lane == 0) ATT_STAMP(g, 4);
                ptx::mbar_wait(b.s_full + 8 * grp, ph);
                ptx::tc_fence_after();
                if (w4 == 0 && lane == 0) ATT_STAMP(g, 5);
                float sum = 0.f;
                float s_x = -INFINITY, p_x = 0.f;
                if (SINGLE && qt == (int)qtn - 1) {
                    // A tile with ONE query row (257 = 2 x 128 + 1) would cost a full S-MMA -> softmax -> PV-MMA chain with one busy
                    // thread.  The group's 128 threads compute the row on the CUDA cores instead, from the swizzled Q / K / V tiles in
                    // shared memory: thread t scores keys t, t+128, ...; (8 columns x 16 key groups) for P V.  The mbarrier hand-shakes
                    // of a normal tile are kept (the issuers skip their MMAs), so slot and ring phases stay in step.
                    const int st = (int)(it % KV_STAGES), qs = (int)(g % Q_RING);
                    const uint8_t* kbase = smem_raw + (s_kv - ptx::smem_u32(smem_raw)) + (uint32_t)st * 2 * kv_bytes;
                    single_row_tile(smem_raw + (s_q - ptx::smem_u32(smem_raw)) + qs * Q_TILE_BYTES, kbase, kbase + kv_bytes, sr_p[grp], sr_red[grp],
                                    &sr_o[grp][0][0], NR, grp, args.out + ((long long)f * NR + qt * QT) * D + h * HD);
                    warp_has_rows = false;                    // from here on the tile behaves like one without rows
                }
                if (args.cosign) {
                    // The last key does not fit the TMEM budget of two slots (257 keys -> 272 fp32 columns): its score is one
                    // 64-term dot product per row on the CUDA cores, straight from the swizzled Q / K tiles in shared memory.
                    if (args.odd && warp_has_rows) {
                        const int st = (int)(it % KV_STAGES), qs = (int)(g % Q_RING);
                        const int r = w4 * 32 + lane;
                        const uint4* qrow = reinterpret_cast<const uint4*>(smem_raw + (s_q - ptx::smem_u32(smem_raw)) + qs * Q_TILE_BYTES + r * 128);
                        const uint4* krow = reinterpret_cast<const uint4*>(smem_raw + (s_kv - ptx::smem_u32(smem_raw)) + (uint32_t)st * 2 * kv_bytes +
                                                                           (uint32_t)args.keys_pad * 128);
                        float acc0 = 0.f, acc1 = 0.f;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint4 qa = qrow[j ^ (r & 7)], ka = krow[j];          // row keys_pad is a multiple of 8: not swizzled
                            const uint32_t qw[4] = {qa.x, qa.y, qa.z, qa.w}, kw[4] = {ka.x, ka.y, ka.z, ka.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                acc0 = fmaf(__uint_as_float(qw[e] << 16), __uint_as_float(kw[e] << 16), acc0);
                                acc1 = fmaf(__uint_as_float(qw[e] & 0xffff0000u), __uint_as_float(kw[e] & 0xffff0000u), acc1);
                            }
                        }
                        s_x = acc0 + acc1;
                    }
                    ptx::mbar_arrive(b.q_empty + 8 * (uint32_t)(g % Q_RING));
                }
                if (warp_has_rows) {
                    // pass 1: row maximum over the valid keys
                    float mx = s_x;
                    if (ATT_PROBE(1)) mx = 4.0f;
                    // The kernel is bound by TMEM read bandwidth (tcgen05.ld: ~64 B / clock / SM; every score is read in both passes,
                    // tools/trace_attention.py): the first HELD chunks of the row stay in registers between the passes.
                    uint32_t hold[HELD][32];
#pragma unroll
                    for (int hc = 0; hc < HELD; ++hc) {
                        if (32 * hc < full_end) {
                            ptx::tmem_ld32(ts + (uint32_t)(32 * hc), hold[hc]);
                            ptx::tmem_ld_wait();
                            float m0 = __uint_as_float(hold[hc][0]), m1 = __uint_as_float(hold[hc][1]);
#pragma unroll
                            for (int i = 2; i < 32; i += 2) {
                                m0 = fmaxf(m0, __uint_as_float(hold[hc][i]));
                                m1 = fmaxf(m1, __uint_as_float(hold[hc][i + 1]));
                            }
                            mx = fmaxf(mx, fmaxf(m0, m1));
                        }
                    }
                    for (int c0 = 32 * HELD; c0 < (ATT_PROBE(1) ? 0 : full_end); c0 += 32) {
                        uint32_t v[32];
                        ptx::tmem_ld32(ts + (uint32_t)c0, v);
                        ptx::tmem_ld_wait();
                        float m0 = __uint_as_float(v[0]), m1 = __uint_as_float(v[1]);
#pragma unroll
                        for (int i = 2; i < 32; i += 2) {
                            m0 = fmaxf(m0, __uint_as_float(v[i]));
                            m1 = fmaxf(m1, __uint_as_float(v[i + 1]));
                        }
                        mx = fmaxf(mx, fmaxf(m0, m1));
                    }
                    if (full_end < args.keys_pad) {
                        uint32_t v[32];
                        const bool two = full_end + 16 < args.keys_pad;
                        ptx::tmem_ld16(ts + (uint32_t)full_end, v);
                        if (two) ptx::tmem_ld16(ts + (uint32_t)full_end + 16u, v + 16);
                        ptx::tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (full_end + i < N && (i < 16 || two)) mx = fmaxf(mx, __uint_as_float(v[i]));
                    }
                    if (w4 == 0 && lane == 0) ATT_STAMP(g, 6);
                    // pass 2: p = 2^(s*c - max*c), bf16 pairs written back over S; fp32 row sum
                    const float mxs = mx * c;
                    float s0 = 0.f, s1 = 0.f;
#pragma unroll
                    for (int hc = 0; hc < HELD; ++hc) {
                        if (32 * hc < full_end) {
                            uint32_t p[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                const float e0 = exp2_sel(fmaf(__uint_as_float(hold[hc][2 * i]), c, -mxs), 2 * i);
                                const float e1 = exp2_sel(fmaf(__uint_as_float(hold[hc][2 * i + 1]), c, -mxs), 2 * i + 1);
                                s0 += e0;
                                s1 += e1;
                                p[i] = pack_bf16x2(e0, e1);
                            }
                            tmem_st16(ts + (uint32_t)(16 * hc), p);
                        }
                    }
                    for (int c0 = 32 * HELD; c0 < full_end; c0 += 32) {
                        uint32_t v[32];
                        ptx::tmem_ld32(ts + (uint32_t)c0, v);
                        ptx::tmem_ld_wait();
                        uint32_t p[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float a0 = fmaf(__uint_as_float(v[2 * i]), c, -mxs), a1 = fmaf(__uint_as_float(v[2 * i + 1]), c, -mxs);
                            const float e0 = ATT_PROBE(2) ? a0 : exp2_sel(a0, 2 * i);
                            const float e1 = ATT_PROBE(2) ? a1 : exp2_sel(a1, 2 * i + 1);
                            s0 += e0;
                            s1 += e1;
                            p[i] = pack_bf16x2(e0, e1);
                        }
                        // keys [c0, c0+32) -> 16 packed columns at c0/2: always behind the S read front of this lane
                        if (!ATT_PROBE(8)) tmem_st16(ts + (uint32_t)(c0 >> 1), p);
                        if (c0 + 32 == args.split_col && args.split_col < args.keys_pad) {      // first batch of P is complete
                            ptx::tmem_st_wait();
                            ptx::tc_fence_before();
                            ptx::mbar_arrive(b.p_early + 8 * grp);
                        }
                    }
                    if (full_end < args.keys_pad) {
                        uint32_t v[32];
                        const bool two = full_end + 16 < args.keys_pad;
                        ptx::tmem_ld16(ts + (uint32_t)full_end, v);
                        if (two) ptx::tmem_ld16(ts + (uint32_t)full_end + 16u, v + 16);
                        ptx::tmem_ld_wait();
                        uint32_t p[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int k0 = full_end + 2 * i;
                            const bool in = i < 8 || two;
                            const float e0 = (in && k0 < N) ? fast_exp2(fmaf(__uint_as_float(v[2 * i]), c, -mxs)) : 0.f;
                            const float e1 = (in && k0 + 1 < N) ? fast_exp2(fmaf(__uint_as_float(v[2 * i + 1]), c, -mxs)) : 0.f;
                            s0 += e0;
                            s1 += e1;
                            p[i] = pack_bf16x2(e0, e1);
                        }
                        tmem_st16(ts + (uint32_t)(full_end >> 1), p);
                    }
                    if (args.odd) p_x = fast_exp2(fmaf(s_x, c, -mxs));
                    sum = s0 + s1 + p_x;
                    ptx::tmem_st_wait();
                } else if (args.split_col < args.keys_pad) {
                    ptx::mbar_arrive(b.p_early + 8 * grp);             // a warp without rows still signs the early batch
                }
                ptx::tc_fence_before();
                if (w4 == 0 && lane == 0) ATT_STAMP(g, 7);
                ptx::mbar_arrive(b.p_full + 8 * grp);
                ptx::mbar_wait(b.o_full + 8 * grp, ph);
                ptx::tc_fence_after();
                if (w4 == 0 && lane == 0) ATT_STAMP(g, 8);
                if (warp_has_rows && !ATT_PROBE(4)) {
                    const float inv = 1.f / sum;
                    bf16* dst_row = args.out + ((long long)f * NR + row) * D + h * HD;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint32_t o[32];
                        ptx::tmem_ld32(ts + (uint32_t)args.o_col + 32u * half, o);
                        ptx::tmem_ld_wait();
                        if (args.odd) {
                            const int st = (int)(it % KV_STAGES);
                            const uint4* vrow = reinterpret_cast<const uint4*>(smem_raw + (s_kv - ptx::smem_u32(smem_raw)) + (uint32_t)st * 2 * kv_bytes +
                                                                               kv_bytes + (uint32_t)args.keys_pad * 128) + 4 * half;
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const uint4 va = vrow[j];
                                const uint32_t vw[4] = {va.x, va.y, va.z, va.w};
#pragma unroll
                                for (int e = 0; e < 4; ++e) {
                                    o[8 * j + 2 * e] = __float_as_uint(fmaf(p_x, __uint_as_float(vw[e] << 16), __uint_as_float(o[8 * j + 2 * e])));
                                    o[8 * j + 2 * e + 1] = __float_as_uint(fmaf(p_x, __uint_as_float(vw[e] & 0xffff0000u), __uint_as_float(o[8 * j + 2 * e + 1])));
                                }
                            }
                        }
                        if (row < NR) {
                            uint4* dst = reinterpret_cast<uint4*>(dst_row + 32 * half);
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                dst[i] = make_uint4(pack_bf16x2(__uint_as_float(o[8 * i]) * inv, __uint_as_float(o[8 * i + 1]) * inv),
                                                    pack_bf16x2(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv),
                                                    pack_bf16x2(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv),
                                                    pack_bf16x2(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv));
                        }
                    }
                }
                if (args.cosign && qt == (int)qtn - 1) ptx::mbar_arrive(b.kv_empty + 8 * (uint32_t)(it % KV_STAGES));
                ptx::tc_fence_before();
                if (w4 == 0 && lane == 0) ATT_STAMP(g, 9);
                ptx::mbar_arrive(b.slot_free + 8 * grp);
            }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 9) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem, 512);
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

}  // namespace

#ifdef DISTB200_ATT_PROBES
extern "C" int distb200_debug_att_trace(long long* buf) {
    cudaMemcpyToSymbol(g_att_trace, &buf, sizeof(buf));
    return 0;
}
#endif
int attention_tc_launch(const void* qkv, void* out, int frames, int tokens, int heads, cudaStream_t stream) {
    int keys_pad = (tokens + 15) / 16 * 16;
    const int keys_ld = keys_pad;
    if (keys_pad > 288 || ((uintptr_t)qkv & 15) || ((uintptr_t)out & 15))
        return attention_simt_launch(qkv, out, frames, tokens, heads, DISTB200_BF16, false, stream);

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
    cuuint32_t box64[3] = {HD, 64, 1}, box16[3] = {HD, 16, 1}, estr[3] = {1, 1, 1};
    CUresult r = encode(&args.tm64, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv), gdim, gstr, box64, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DISTB200_REQUIRE(r == CUDA_SUCCESS, "attention(tcgen05): cuTensorMapEncodeTiled failed with %d", (int)r);
    r = encode(&args.tm16, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(qkv), gdim, gstr, box16, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    DISTB200_REQUIRE(r == CUDA_SUCCESS, "attention(tcgen05): cuTensorMapEncodeTiled failed with %d", (int)r);
    args.out = reinterpret_cast<bf16*>(out);
    args.tokens = tokens;
    args.heads = heads;
    args.q_tiles = (tokens + QT - 1) / QT;
    // 257 tokens (ViT-L/14): 272 fp32 score columns per slot would leave room for ONE slot in the 512-column TMEM and
    // serialise S-MMA -> softmax -> PV-MMA -> epilogue.  With the single odd key taken off the tensor core, two slots of 256 fit.
    args.odd = 0;
    {
        const int oc = (keys_pad / 2 + 31) / 32 * 32, sc = oc + HD > keys_pad ? oc + HD : keys_pad;
        const int kp2 = keys_pad - 16, oc2 = (kp2 / 2 + 31) / 32 * 32, sc2 = oc2 + HD > kp2 ? oc2 + HD : kp2;
        static const int odd_env = getenv("DISTB200_ATT_ODD") ? atoi(getenv("DISTB200_ATT_ODD")) : 1;
        if (odd_env && tokens % 16 == 1 && kp2 >= 32 && 2 * sc > 512 && 2 * sc2 <= 512) {
            args.odd = 1;
            keys_pad = kp2;
        }
    }
    args.single = (tokens % QT == 1 && tokens > QT) ? 1 : 0;
    {
        static const int single_env = getenv("DISTB200_ATT_SINGLE") ? atoi(getenv("DISTB200_ATT_SINGLE")) : 1;
        if (!single_env) args.single = 0;
    }
    args.cosign = (args.odd || args.single) ? 1 : 0;
    args.keys_ld = keys_ld;
    args.keys_pad = keys_pad;
    args.o_col = (keys_pad / 2 + 31) / 32 * 32;
    args.slot_cols = args.o_col + HD > keys_pad ? args.o_col + HD : keys_pad;
    args.n_slots = 2 * args.slot_cols <= 512 ? 2 : 1;
    args.split_col = (args.o_col + HD + 31) / 32 * 32;       // first 32-column chunk boundary past the O region
    if (args.split_col > (tokens - args.odd) / 32 * 32 || args.o_col >= keys_pad) args.split_col = keys_pad;      // no chunk boundary there: one batch
    args.items = (long long)frames * heads;
    args.dbg = 0;
#ifdef DISTB200_ATT_PROBES
    args.dbg = getenv("DISTB200_ATT_DBG") ? atoi(getenv("DISTB200_ATT_DBG")) : 0;
#endif
    DISTB200_REQUIRE(args.items * args.q_tiles < (1ll << 31), "attention(tcgen05): too many tiles");
    const int smem = Q_RING * (int)Q_TILE_BYTES + KV_STAGES * 2 * keys_ld * HD * 2 + 1024;
    static int smem_set_dev[DISTB200_MAX_DEVICES] = {};
    int& smem_set = smem_set_dev[current_device()];
    if (smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(attention_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        DISTB200_REQUIRE(e == cudaSuccess, "attention(tcgen05): cannot reserve %d bytes of shared memory: %s", smem, cudaGetErrorString(e));
        smem_set = smem;
    }
    const long long grid = args.items < sm_count() ? args.items : sm_count();
    if (args.single) DISTB200_LAUNCH(attention_tc_kernel<true>, (unsigned)grid, ATT_THREADS, smem, stream, args);
    else DISTB200_LAUNCH(attention_tc_kernel<false>, (unsigned)grid, ATT_THREADS, smem, stream, args);
    return check_launch("attention_tc");
}

}  // namespace distb200
