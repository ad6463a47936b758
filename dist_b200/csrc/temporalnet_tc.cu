// distb200_temporalnet: one fused TemporalNet block (dist.py:48-65) per launch on sm_100a.
//
//   xe = x + upsample_alpha(u)            (integration->temporal add of the previous DiST layer, dist.py:105,231)
//   y  = LayerNorm_C(xe)                  z = q(conv(3,1,1)(y) + b1)        out = q(xe + conv(1,3,3)(z) + b2)
//
// Unit of work = (clip, band of image rows, frame).  A CTA pair (cta_group::2) runs two units in lock step - one per CTA - so
// that every tcgen05.mma is 256 x C x 16 and each CTA keeps only HALF of the output channels of the twelve weight taps in
// shared memory: both weight sets stay resident for the whole launch (C = 96: 110 KB per CTA) and nothing but activations
// moves through L2.  A CTA walks consecutive frames of one (clip, band), so every frame band is LayerNorm'd once.
//
// Shared memory (per CTA), all MMA operands in the un-swizzled K-major layout (8-row x 16-byte core matrices, SBO = 128, i.e.
// row r of k-chunk kc at  base + kc * LBO + r * 16 ):
//   weights   12 taps x [C/8 chunks][C/2 rows][16 B]
//   ring      3 slots x [C/8][128 rows][16 B]   LayerNorm'd bf16 rows of frames tau-1, tau, tau+1 (band + one halo row either side,
//                                               positions in image order: row m = (r - r0 + 1) * g + c)
//   z         [C/8][z_rows][16 B]               conv(3,1,1) output in a zero-padded pitch-(g+2) layout: position (lr, c) of the
//                                               band (lr = 0 is the halo row above) at row lr * (g+2) + c + 1.  Pad columns
//                                               and rows outside the frame hold zeros, so spatial tap (i, j) of the (1,3,3)
//                                               convolution is the SAME tile with its start address moved by (i*(g+2) + j - 1) rows.
//   staging   8 x 2 KB                          transposes of the output epilogue
// TMEM: conv(3,1,1) accumulators at columns [0, 2C), conv(1,3,3) accumulators at [2C, 4C) (two stages each).
//
// Warps: 0-7 epilogue (TMEM lane quadrant = warp & 3, channel half = warp >> 2), 8-14 LayerNorm producers, 15 MMA issuer
// (leader CTA only).  Per iteration `it` the issuer runs conv311(it) then conv133(it-1); the epilogue warps turn the conv311
// accumulator into z (bias, QuickGELU, bf16) and then finish unit it-1 (bias, residual, QuickGELU, fp32 + bf16 stores) while the
// tensor core is busy with the next unit; the LayerNorm warps fill the ring one frame ahead.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace distb200 {

namespace {

constexpr int TN_EPI_WARPS = 8;
constexpr int TN_LN_WARPS = 7;                // 8 + 7 + 1 = 16 warps: four per scheduler, 128 registers per thread (with 227 KB of shared memory
                                              // carved out there is no L1 to speak of: a spilled register costs an L2 round trip)
constexpr int TN_MMA_WARP = TN_EPI_WARPS + TN_LN_WARPS;
constexpr int TN_THREADS = (TN_MMA_WARP + 1) * 32;
constexpr int TN_M = 128;                     // accumulator rows per CTA
constexpr uint32_t TN_LBO_A = TN_M * 16 + 64; // k-chunk pitch of a ring slot: +64 bytes so that the four k-chunks a warp's LayerNorm store
                                              // touches fall on different banks (ncu: 8-way conflicts on every STS.64 without it)
constexpr int TN_STAGE_BYTES = 2048;          // per epilogue warp: 32 rows x 16 fp32 columns
constexpr int TN_TMEM_COLS = 512;
// mbarriers (8 bytes each, same offsets in both CTAs)
constexpr int B_LN_FULL = 0;                  // leader: LayerNorm warps of both CTAs filled the ring for iteration `it`
constexpr int B_EPI1 = 1;                     // leader: z of iteration j written (and its conv311 accumulator read) in both CTAs
constexpr int B_ACC2_FREE = 2;                // leader [2]: conv133 accumulator stage read by both CTAs
constexpr int B_C311 = 4;                     // each CTA [2]: conv311(it) completed (tcgen05.commit, multicast)
constexpr int B_C133 = 6;                     // each CTA [2]: conv133(j) completed
constexpr int B_W_READY = 8;                  // leader: weights / biases / zeroed z of both CTAs are in shared memory
constexpr int TN_NBAR = 9;

struct alignas(16) TnArgs {
    distb200_temporalnet_desc d;
    int nb;            // bands per frame
    int br;            // rows of the largest band
    int W;             // pitch of the padded z layout = g + 2
    int z_rows;        // (br + 2) * W
    int P;             // g * g
    int ts;            // frames / alpha
    int n_per;         // iterations (units per CTA)
    int units;         // clips * nb * frames
    int band_base, band_rem;   // band b has band_base + (b < band_rem) image rows
    uint32_t off_ln, off_z, off_stg, off_par, off_bar;
};

// Timeline trace / bottleneck probes, compiled in only with -DDISTB200_TN_TRACE (tools/trace_temporalnet.py): clock64 stamps of one
// lane per role and iteration into a global buffer [cta][it][32], and DISTB200_TN_DBG bits that remove one component at a time
// (1 no MMA issue, 2 no LayerNorm loads, 4 no residual loads / output stores, 8 no z epilogue math, 16 CTA-scope barrier arrives).
#ifdef DISTB200_TN_TRACE
__device__ long long* g_tn_trace = nullptr;
__device__ int g_tn_dbg = 0;
#define TN_STAMP(it, slot) do { if (g_tn_trace && (it) < 64) g_tn_trace[((long long)blockIdx.x * 64 + (it)) * 32 + (slot)] = clock64(); } while (0)
#define TN_PROBE(bit) (g_tn_dbg & (bit))
#else
#define TN_STAMP(it, slot) do { } while (0)
#define TN_PROBE(bit) false
#endif

struct Unit {
    bool valid;
    int clip, band, tau, r0, nrows, frame;
    int index;
};

// unit index = (clip * nb + band) * T + tau: consecutive units of a CTA are consecutive frames of one band
__device__ __forceinline__ Unit decode_unit(const TnArgs& a, int u) {
    Unit un;
    un.valid = u < a.units;
    const uint32_t T = (uint32_t)a.d.frames;
    const uint32_t uu = un.valid ? (uint32_t)u : 0u;
    const uint32_t cb = uu / T;
    un.tau = (int)(uu - cb * T);
    un.clip = (int)(cb / (uint32_t)a.nb);
    const int band = (int)cb - un.clip * a.nb;
    un.band = band;
    un.index = u;
    un.nrows = a.band_base + (band < a.band_rem ? 1 : 0);
    un.r0 = band * a.band_base + (band < a.band_rem ? band : a.band_rem);
    un.frame = un.clip * (int)T + un.tau;
    return un;
}

// the unit after `un` (units of a CTA are consecutive): no divisions on the per-iteration path
__device__ __forceinline__ Unit next_unit(const TnArgs& a, Unit un) {
    un.index += 1;
    if (un.index >= a.units) { un.valid = false; return un; }      // keeps the fields of the last valid unit (valid addresses)
    if (!un.valid) return un;
    un.tau += 1;
    un.frame += 1;
    if (un.tau == a.d.frames) {
        un.tau = 0;
        un.band += 1;
        un.frame -= a.d.frames;
        if (un.band == a.nb) { un.band = 0; un.clip += 1; un.frame += a.d.frames; }
        un.nrows = a.band_base + (un.band < a.band_rem ? 1 : 0);
        un.r0 = un.band * a.band_base + (un.band < a.band_rem ? un.band : a.band_rem);
    }
    return un;
}

__device__ __forceinline__ float gelu_fast(float x) {        // bf16 destinations: one MUFU.TANH (see gemm_tcgen05.cu)
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
    return x * fmaf(0.5f, t, 0.5f);
}

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ uint2 ldg_bf4(const bf16* p) { return __ldg(reinterpret_cast<const uint2*>(p)); }      // four bf16 values
__device__ __forceinline__ void add_bf4(float4& v, const uint2 w) {
    v.x += __uint_as_float(w.x << 16); v.y += __uint_as_float(w.x & 0xffff0000u);
    v.z += __uint_as_float(w.y << 16); v.w += __uint_as_float(w.y & 0xffff0000u);
}

// ---- LayerNorm producers -------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_l2(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// Arrive on a barrier of the leader CTA.  Default: the plain arrive (release at CTA scope) that the CTA-pair GEMM uses for its
// accumulator hand-back - the data a remote arrive publishes here was written into the ARRIVING CTA's own shared memory / read from
// its own TMEM and is consumed by that CTA's tensor core.  -DDISTB200_TN_STRICT_FENCES selects release.cluster, which ptxas
// turns into MEMBAR.ALL.GPU + ERRBAR in front of every arrive (1 - 2 k cycles each, three of them on the per-unit critical path).
__device__ __forceinline__ void arrive_leader(uint32_t cluster_addr) {
#ifdef DISTB200_TN_STRICT_FENCES
    ptx::mbar_arrive_release_cluster(cluster_addr);
#else
    ptx::mbar_arrive_cluster(cluster_addr);
#endif
}
__device__ __forceinline__ void wait_leader(uint32_t bar, uint32_t parity) {
#ifdef DISTB200_TN_STRICT_FENCES
    ptx::mbar_wait_acquire_cluster(bar, parity);
#else
    ptx::mbar_wait(bar, parity);
#endif
}

// The frames LayerNorm'd for unit `un` in iteration `it` are window positions j = j0 .. 2 (frame tau - 1 + j into ring slot
// (it + 2 + j) % 3): a new (clip, band) fills the whole window (j0 = 0), otherwise only frame tau + 1 is new (j0 = 2).  A frame
// outside the clip is a slot of zeros: the convolution pads LayerNorm's OUTPUT.  (No arrays: local memory has no L1 here.)
__device__ __forceinline__ int ln_first_job(const Unit& un, int it) { return !un.valid ? 3 : ((it == 0 || un.tau == 0) ? 0 : 2); }

// Eight lanes per row (lane & 7 = 16-byte piece inside every 128-byte third of the row), four rows per warp and pass, the
// seven warps cover the 128 slot rows in five passes of 28 rows.  A frame is done in three rounds (passes 0-1, 2-3, 4) through two
// register buffers: the loads of round r+1 are in flight while round r is normalised, so only the first round's latency is
// exposed - and that one overlaps the wait for the ring slot.  One lane asks L2 for the rows of the NEXT iteration's frames
// (cp.async.bulk.prefetch.L2: the band rows of a frame are contiguous), which makes the loads themselves L2 hits.
template <int C, bool HAS_U>
__device__ __forceinline__ void ln_role(const TnArgs& a, uint8_t* smem, int wl, int lane, int first) {
    constexpr int NJ = C / 32;
    constexpr int KC = C / 8;
    constexpr uint32_t LBO_A = TN_LBO_A;
    constexpr uint32_t SLOT = KC * LBO_A;
    constexpr int RP = 4 * TN_LN_WARPS;              // rows per pass
    const distb200_temporalnet_desc& d = a.d;
    const int g = d.grid, T = d.frames, P = a.P;
    const int sub = lane >> 3, cl = lane & 7;
    const int mrow = wl * 4 + sub;                   // row of this lane in pass 0
    const float* par = reinterpret_cast<const float*>(smem + a.off_par) + 4 * cl;
    const uint32_t bar0 = ptx::smem_u32(smem + a.off_bar);
    const uint32_t ln_full = ptx::mapa(bar0 + 8u * B_LN_FULL, 0);
    const float inv_c = 1.0f / (float)C;
    const uint32_t st_off = (uint32_t)(cl >> 1) * LBO_A + (uint32_t)(cl & 1) * 8u;      // channels 4 cl .. +3 of every 32: k-chunk cl/2, half cl&1

    auto prefetch_unit = [&](int it, const Unit& un) {               // one lane: the rows that LN(it) will read
        if (it >= a.n_per) return;
        int lo = (un.r0 - 1) * g, hi = lo + (un.nrows + 2) * g;
        lo = lo < 0 ? 0 : lo;
        hi = hi > P ? P : hi;
        const int j0 = ln_first_job(un, it);
        for (int j = j0; j < 3; ++j) {
            const int sigma = un.tau - 1 + j;
            if (sigma < 0 || sigma >= T) continue;
            prefetch_l2(d.x + ((long long)(un.clip * T + sigma) * P + lo) * C, (uint32_t)((hi - lo) * C * 4));
            if (HAS_U && (j == j0 || sigma % d.alpha == 0))
                prefetch_l2(reinterpret_cast<const bf16*>(d.u) + ((long long)(un.clip * a.ts + sigma / d.alpha) * P + lo) * C, (uint32_t)((hi - lo) * C * 2));
        }
    };
    Unit un = decode_unit(a, first);
    if (wl == 0 && lane == 0 && !TN_PROBE(32)) prefetch_unit(0, un);

    for (int it = 0; it < a.n_per; ++it) {
        if (it > 0) un = next_unit(a, un);
        if (wl == 0 && lane == 0 && !TN_PROBE(32)) prefetch_unit(it + 1, next_unit(a, un));
        // every slot written in iteration `it` was last read by conv311(it - 1)
        bool waited = it == 0;
        const uint32_t wbar = bar0 + 8u * (uint32_t)(B_C311 + ((it - 1) & 1)), wpar = (uint32_t)((it - 1) >> 1) & 1u;
        const int rows_valid = (un.nrows + 2) * g, pos0 = (un.r0 - 1) * g;
        // rows of this lane: m = mrow + RP * pass; a row exists when m < rows_valid and its position lies inside the frame
        const int p_lo = pos0 < 0 ? -pos0 : 0;                               // first slot row inside the frame
        const int p_hi = rows_valid < P - pos0 ? rows_valid : P - pos0;      // one past the last one
        for (int j = ln_first_job(un, it); j < 3; ++j) {
            uint8_t* slot = smem + a.off_ln + (uint32_t)((it + 2 + j) % 3) * SLOT + st_off;
            const int sigma = un.tau - 1 + j;
            if (sigma < 0 || sigma >= T) {
                if (!waited) { ptx::mbar_wait(wbar, wpar); waited = true; }
#pragma unroll
                for (int pass = 0; pass < 5; ++pass) {
                    const int m = mrow + RP * pass;
                    if (m < rows_valid) {
#pragma unroll
                        for (int jj = 0; jj < NJ; ++jj) *reinterpret_cast<uint2*>(slot + (uint32_t)(4 * jj) * LBO_A + (uint32_t)m * 16u) = make_uint2(0u, 0u);
                    }
                }
                continue;
            }
            const float* xf = d.x + ((long long)(un.clip * T + sigma) * P + pos0) * C + 4 * cl;
            const bf16* uf = HAS_U ? reinterpret_cast<const bf16*>(d.u) + ((long long)(un.clip * a.ts + sigma / d.alpha) * P + pos0) * C + 4 * cl : nullptr;
            float4 va[2][NJ], vb[2][NJ];
            uint2 wa[HAS_U ? 2 : 1][NJ], wb[HAS_U ? 2 : 1][NJ];
            // rows outside the frame are loaded from a clamped address (branch-free) and never stored
            auto load = [&](int pass, float4* v, uint2* w) {
                int m = mrow + RP * pass;
                m = m < p_lo ? p_lo : (m >= p_hi ? p_hi - 1 : m);
                const long long off = (long long)m * C;
#pragma unroll
                for (int jj = 0; jj < NJ; ++jj) {
                    v[jj] = TN_PROBE(2) ? make_float4(1.f, 2.f, 3.f, 4.f) : ldg4(xf + off + 32 * jj);
                    if constexpr (HAS_U) w[jj] = TN_PROBE(2) ? make_uint2(0u, 0u) : ldg_bf4(uf + off + 32 * jj);
                }
            };
            auto norm_store = [&](int pass, float4* v, uint2* w) {
                const int m = mrow + RP * pass;
                float s = 0.f;
#pragma unroll
                for (int jj = 0; jj < NJ; ++jj) {
                    if constexpr (HAS_U) add_bf4(v[jj], w[jj]);
                    s += (v[jj].x + v[jj].y) + (v[jj].z + v[jj].w);
                }
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                s += __shfl_xor_sync(0xffffffffu, s, 4);
                const float mean = s * inv_c;
                float qq = 0.f;
#pragma unroll
                for (int jj = 0; jj < NJ; ++jj) {
                    v[jj].x -= mean; v[jj].y -= mean; v[jj].z -= mean; v[jj].w -= mean;
                    qq += fmaf(v[jj].x, v[jj].x, v[jj].y * v[jj].y) + fmaf(v[jj].z, v[jj].z, v[jj].w * v[jj].w);
                }
                qq += __shfl_xor_sync(0xffffffffu, qq, 1);
                qq += __shfl_xor_sync(0xffffffffu, qq, 2);
                qq += __shfl_xor_sync(0xffffffffu, qq, 4);
                const float rstd = rsqrtf(qq * inv_c + d.eps);
                const bool st = m >= p_lo && m < p_hi;
#pragma unroll
                for (int jj = 0; jj < NJ; ++jj) {
                    const float4 gm = *reinterpret_cast<const float4*>(par + 32 * jj);
                    const float4 bt = *reinterpret_cast<const float4*>(par + C + 32 * jj);
                    const uint2 pk = make_uint2(pack_bf16x2(fmaf(v[jj].x * rstd, gm.x, bt.x), fmaf(v[jj].y * rstd, gm.y, bt.y)),
                                                pack_bf16x2(fmaf(v[jj].z * rstd, gm.z, bt.z), fmaf(v[jj].w * rstd, gm.w, bt.w)));
                    if (st) *reinterpret_cast<uint2*>(slot + (uint32_t)(4 * jj) * LBO_A + (uint32_t)m * 16u) = pk;
                }
            };
            load(0, va[0], wa[0]);
            load(1, va[1], wa[HAS_U ? 1 : 0]);
            load(2, vb[0], wb[0]);
            load(3, vb[1], wb[HAS_U ? 1 : 0]);
            if (!waited) {
                if (wl == 0 && lane == 0) TN_STAMP(it, 0);
                ptx::mbar_wait(wbar, wpar);
                waited = true;
                if (wl == 0 && lane == 0) TN_STAMP(it, 1);
            }
            norm_store(0, va[0], wa[0]);
            norm_store(1, va[1], wa[HAS_U ? 1 : 0]);
            load(4, va[0], wa[0]);
            norm_store(2, vb[0], wb[0]);
            norm_store(3, vb[1], wb[HAS_U ? 1 : 0]);
            norm_store(4, va[0], wa[0]);
        }
        if (!waited) ptx::mbar_wait(wbar, wpar);       // keep every warp within one phase of the others (the barrier counts arrivals per phase)
        if (wl == 0 && lane == 0) TN_STAMP(it, 2);
        ptx::fence_proxy_async();                      // generic-proxy stores -> visible to the tensor core's reads
        __syncwarp();
        if (lane == 0) arrive_leader(ln_full);
        if (wl == 0 && lane == 0) TN_STAMP(it, 3);
    }
}

// ---- MMA issuer (leader CTA) -----------------------------------------------------------------------------------------
template <int C>
__device__ __forceinline__ void mma_role(const TnArgs& a, uint8_t* smem, uint32_t tmem_base) {
    constexpr int KC = C / 8, KS = C / 16;
    constexpr uint32_t LBO_A = TN_LBO_A, SLOT = KC * LBO_A;
    constexpr uint32_t LBO_B = (C / 2) * 16, TAPB = KC * LBO_B;
    const uint32_t LBO_Z = (uint32_t)a.z_rows * 16u;
    const uint32_t s0 = ptx::smem_u32(smem);
    const uint32_t bar0 = s0 + a.off_bar;
    const uint32_t idesc = ptx::umma_idesc_bf16(2 * TN_M, C);
    wait_leader(bar0 + 8u * B_W_READY, 0);            // the weights of both CTAs are in shared memory
    for (int it = 0; it <= a.n_per; ++it) {
        if (it < a.n_per) {
            // ---- conv(3,1,1) of iteration it: taps = ring slots (it+2, it, it+1) % 3 = frames tau-1, tau, tau+1
            if (threadIdx.x % 32 == 0) TN_STAMP(it, 4);
            wait_leader(bar0 + 8u * B_LN_FULL, (uint32_t)it & 1u);
            ptx::tc_fence_after();
            if (threadIdx.x % 32 == 0) TN_STAMP(it, 5);
            if (ptx::elect_one()) {
                const uint32_t acc = tmem_base + (uint32_t)((it & 1) * C);
#pragma unroll
                for (int k = 0; k < (TN_PROBE(1) ? 0 : 3); ++k) {
                    const uint32_t sa = s0 + a.off_ln + (uint32_t)((it + 2 + k) % 3) * SLOT;
                    const uint32_t sb = s0 + (uint32_t)k * TAPB;
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks)
                        ptx::mma_f16_ss_2sm(acc, ptx::umma_desc_k_nosw(sa + 2u * ks * LBO_A, LBO_A, 128), ptx::umma_desc_k_nosw(sb + 2u * ks * LBO_B, LBO_B, 128),
                                            idesc, (k | ks) != 0);
                }
                ptx::mma_commit_2sm(bar0 + 8u * (uint32_t)(B_C311 + (it & 1)), 3);
            }
            __syncwarp();
        }
        if (it >= 1) {
            // ---- conv(1,3,3) of iteration j: nine row-shifted views of z
            const int j = it - 1;
            if (threadIdx.x % 32 == 0) TN_STAMP(j, 6);
            wait_leader(bar0 + 8u * B_EPI1, (uint32_t)j & 1u);
            wait_leader(bar0 + 8u * (uint32_t)(B_ACC2_FREE + (j & 1)), ((uint32_t)(j >> 1) & 1u) ^ 1u);
            ptx::tc_fence_after();
            if (threadIdx.x % 32 == 0) TN_STAMP(j, 7);
            if (ptx::elect_one()) {
                const uint32_t acc = tmem_base + (uint32_t)(2 * C + (j & 1) * C);
#pragma unroll
                for (int t = 0; t < (TN_PROBE(1) ? 0 : 9); ++t) {
                    const uint32_t sa = s0 + a.off_z + (uint32_t)((t / 3) * a.W + (t % 3) - 1) * 16u;      // tap (0,0) starts one row before z (a pad-column output)
                    const uint32_t sb = s0 + (uint32_t)(3 + t) * TAPB;
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks)
                        ptx::mma_f16_ss_2sm(acc, ptx::umma_desc_k_nosw(sa + 2u * ks * LBO_Z, LBO_Z, 128), ptx::umma_desc_k_nosw(sb + 2u * ks * LBO_B, LBO_B, 128),
                                            idesc, (t | ks) != 0);
                }
                ptx::mma_commit_2sm(bar0 + 8u * (uint32_t)(B_C133 + (j & 1)), 3);
            }
            __syncwarp();
        }
    }
}

// ---- epilogue warps --------------------------------------------------------------------------------------------------
// Straight-line code: every load runs from a clamped (always valid) address and only the stores are predicated, so that the
// sixteen values a lane owns per chunk are independent instruction chains (two epilogue warps share a scheduler with two
// LayerNorm warps: instruction-level parallelism is what hides the ALU / MUFU / shared-memory latencies here).
template <int C, bool HAS_U>
__device__ __forceinline__ void epi_role(const TnArgs& a, uint8_t* smem, uint32_t tmem_base, int warp, int lane, int first) {
    constexpr int NCH = C / 32;             // 16-column chunks per warp (its half of the channels)
    const distb200_temporalnet_desc& d = a.d;
    const int g = d.grid, P = a.P, W = a.W;
    const int quad = warp & 3, hf = warp >> 2;
    const uint32_t LBO_Z = (uint32_t)a.z_rows * 16u;
    const float* par = reinterpret_cast<const float*>(smem + a.off_par);
    const float* b1s = par + 2 * C + hf * (C / 2);
    const float* b2s = par + 3 * C + hf * (C / 2);
    float4* stg = reinterpret_cast<float4*>(smem + a.off_stg + warp * TN_STAGE_BYTES);
    const uint32_t bar0 = ptx::smem_u32(smem + a.off_bar);
    const uint32_t epi1_done = ptx::mapa(bar0 + 8u * B_EPI1, 0);
    const uint32_t acc2_free = ptx::mapa(bar0 + 8u * B_ACC2_FREE, 0);
    const uint32_t tq = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(hf * (C / 2));
    // thread = row domain (conv311 accumulator -> z): accumulator row m1 = position (lr1, c1) of the haloed band
    const int m1 = quad * 32 + lane;
    const int lr1 = m1 / g, c1 = m1 - lr1 * g;
    const bool in_z = m1 < (a.br + 2) * g;
    uint8_t* zrow = smem + a.off_z + (uint32_t)(hf * (C / 16)) * LBO_Z + (uint32_t)(lr1 * W + c1 + 1) * 16u;
    // column domain (conv133 accumulator -> global): four lanes per row, eight rows per pass; accumulator row m2 = padded
    // position (lo, cc) of the band, cc = 0 and cc = g + 1 being pad columns.  Row offsets are the same for every unit.
    const int rs = lane >> 2, l4 = lane & 3;
    int lo_[4];                              // band row of the lane's row in pass p, or a huge value for pad columns (never stored)
    uint32_t of1[4], of2[4];                 // element offsets of that row from the band's first position, pitch C / ld_out2
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int m2 = quad * 32 + 8 * p + rs;
        const int lo = m2 / W, cc = m2 - lo * W;
        const bool real = cc >= 1 && cc <= g && lo < a.br;
        const int rel = real ? lo * g + cc - 1 : 0;             // pad rows read row 0 of the band (valid memory) and store nothing
        lo_[p] = real ? lo : (1 << 20);
        of1[p] = (uint32_t)(rel * C + hf * (C / 2) + 4 * l4);
        of2[p] = (uint32_t)rel * (uint32_t)d.ld_out2 + (uint32_t)(hf * (C / 2) + 4 * l4);
    }
    const uint32_t of1_0 = (uint32_t)(hf * (C / 2) + 4 * l4);
    const uint32_t st_w = (uint32_t)(lane * 4), st_x = (uint32_t)((lane >> 1) & 3);       // staging: 16-byte piece j of row r at slot j ^ ((r >> 1) & 3)

    // per-unit bases (64-bit, warp-uniform) of the finished unit
    const float* px = d.x;
    const bf16* pu = reinterpret_cast<const bf16*>(d.u);
    float* po = d.out;
    bf16* p2 = reinterpret_cast<bf16*>(d.out2);
    int prev_nrows = 0;
    bool prev_valid = false;
    auto set_prev = [&](const Unit& un) {
        const long long band0 = (long long)un.frame * P + un.r0 * g;
        px = d.x + band0 * C;
        if constexpr (HAS_U) pu = reinterpret_cast<const bf16*>(d.u) + ((long long)(un.clip * a.ts + un.tau / d.alpha) * P + un.r0 * g) * C;
        po = d.out ? d.out + band0 * C : nullptr;
        if (d.out2) {
            long long row2;
            int col2 = 0;
            if (d.out2_gdiv > 0) {
                const int qd = un.frame / d.out2_gdiv;
                row2 = (long long)qd * d.out2_gstride + d.out2_roff + un.r0 * g;
                col2 = (un.frame - qd * d.out2_gdiv) * d.out2_cstep;
            } else {
                row2 = band0;
            }
            p2 = reinterpret_cast<bf16*>(d.out2) + row2 * d.ld_out2 + col2;
        }
        prev_nrows = un.nrows;
        prev_valid = un.valid && !TN_PROBE(4);
    };
    float4 rx[4];
    uint2 ru[HAS_U ? 4 : 1];
    auto load_res = [&](int ch) {
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            // rows past the end of a short band would lie in the next band / frame (past the tensor for the last frame): row 0 instead
            const uint32_t o = (lo_[p] < prev_nrows ? of1[p] : of1_0) + 16 * ch;
            rx[p] = TN_PROBE(4) ? make_float4(0.f, 0.f, 0.f, 0.f) : ldg4(px + o);
            if constexpr (HAS_U) ru[p] = TN_PROBE(4) ? make_uint2(0u, 0u) : ldg_bf4(pu + o);
        }
    };

    Unit un = decode_unit(a, first);
    for (int it = 0; it <= a.n_per; ++it) {
        if (it > 0) un = next_unit(a, un);
        if (it >= 1) load_res(0);                       // residual of the unit about to be finished: in flight during the z epilogue

        if (it < a.n_per) {
            if (warp == 0 && lane == 0) TN_STAMP(it, 8);
            ptx::mbar_wait(bar0 + 8u * (uint32_t)(B_C311 + (it & 1)), (uint32_t)(it >> 1) & 1u);
            if (warp == 0 && lane == 0) TN_STAMP(it, 9);
            if (it >= 1) ptx::mbar_wait(bar0 + 8u * (uint32_t)(B_C133 + ((it - 1) & 1)), (uint32_t)((it - 1) >> 1) & 1u);   // z is free again
            ptx::tc_fence_after();
            if (warp == 0 && lane == 0) TN_STAMP(it, 10);
            // ---- conv311 accumulator -> z (bf16, padded layout); rows outside the frame are zeros, not q(b1)
            const int pos = (un.r0 - 1) * g + m1;
            const bool ok = un.valid && m1 < (un.nrows + 2) * g && pos >= 0 && pos < P;
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                uint32_t acc[16];
                ptx::tmem_ld16(tq + (uint32_t)((it & 1) * C + 16 * ch), acc);
                ptx::tmem_ld_wait();
                if (!TN_PROBE(8)) {
#pragma unroll
                    for (int h8 = 0; h8 < 2; ++h8) {
                        const float4 ba = *reinterpret_cast<const float4*>(b1s + 16 * ch + 8 * h8), bb = *reinterpret_cast<const float4*>(b1s + 16 * ch + 8 * h8 + 4);
                        const uint32_t* v = acc + 8 * h8;
                        uint4 pk;
                        pk.x = pack_bf16x2(gelu_fast(__uint_as_float(v[0]) + ba.x), gelu_fast(__uint_as_float(v[1]) + ba.y));
                        pk.y = pack_bf16x2(gelu_fast(__uint_as_float(v[2]) + ba.z), gelu_fast(__uint_as_float(v[3]) + ba.w));
                        pk.z = pack_bf16x2(gelu_fast(__uint_as_float(v[4]) + bb.x), gelu_fast(__uint_as_float(v[5]) + bb.y));
                        pk.w = pack_bf16x2(gelu_fast(__uint_as_float(v[6]) + bb.z), gelu_fast(__uint_as_float(v[7]) + bb.w));
                        if (!ok) pk = make_uint4(0u, 0u, 0u, 0u);
                        if (in_z) *reinterpret_cast<uint4*>(zrow + (uint32_t)(2 * ch + h8) * LBO_Z) = pk;
                    }
                }
            }
            if (warp == 0 && lane == 0) TN_STAMP(it, 14);
            ptx::fence_proxy_async();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) arrive_leader(epi1_done);
            if (warp == 0 && lane == 0) TN_STAMP(it, 11);
        } else {
            ptx::mbar_wait(bar0 + 8u * (uint32_t)(B_C133 + ((it - 1) & 1)), (uint32_t)((it - 1) >> 1) & 1u);
            ptx::tc_fence_after();
        }

        if (it >= 1) {
            // ---- conv133 accumulator of unit it-1 -> out / out2
            const int j = it - 1;
#pragma unroll 1
            for (int ch = 0; ch < NCH; ++ch) {
                uint32_t acc[16];
                ptx::tmem_ld16(tq + (uint32_t)(2 * C + (j & 1) * C + 16 * ch), acc);
                ptx::tmem_ld_wait();
                if (warp == 0 && lane == 0) TN_STAMP(j, 16 + 4 * ch);
                if (ch == NCH - 1) {                     // the accumulator stage is in registers: release it
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) arrive_leader(acc2_free + 8u * (uint32_t)(j & 1));
                    if (warp == 0 && lane == 0) TN_STAMP(j, 12);
                }
                // transpose 32 rows x 16 columns through shared memory
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                    stg[st_w + ((uint32_t)jj ^ st_x)] = make_float4(__uint_as_float(acc[4 * jj]), __uint_as_float(acc[4 * jj + 1]),
                                                                  __uint_as_float(acc[4 * jj + 2]), __uint_as_float(acc[4 * jj + 3]));
                __syncwarp();
                const float4 bias = *reinterpret_cast<const float4*>(b2s + 16 * ch + 4 * l4);
                float4 v[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int rr = 8 * p + rs;
                    v[p] = stg[rr * 4 + (l4 ^ ((rr >> 1) & 3))];
                    float4 r = rx[p];
                    if constexpr (HAS_U) add_bf4(r, ru[p]);
                    v[p].x += bias.x + r.x; v[p].y += bias.y + r.y; v[p].z += bias.z + r.z; v[p].w += bias.w + r.w;
                }
                __syncwarp();                            // the staging tile is rewritten by the next chunk
                if (warp == 0 && lane == 0) TN_STAMP(j, 17 + 4 * ch);
                if (ch + 1 < NCH) load_res(ch + 1);      // next chunk's residual: in flight during the activation and the stores
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    float4 y;
                    y.x = gelu_fast(v[p].x);
                    y.y = gelu_fast(v[p].y);
                    y.z = gelu_fast(v[p].z);
                    y.w = gelu_fast(v[p].w);
                    const bool ok = prev_valid && lo_[p] < prev_nrows;
                    if (ok && po) *reinterpret_cast<float4*>(po + of1[p] + 16 * ch) = y;
                    if (ok && d.out2) *reinterpret_cast<uint2*>(p2 + of2[p] + 16 * ch) = make_uint2(pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
                }
                if (warp == 0 && lane == 0) TN_STAMP(j, 19 + 4 * ch);
            }
            if (warp == 0 && lane == 0) TN_STAMP(j, 13);
        }
        if (it < a.n_per) set_prev(un);
    }
}

template <int C, bool HAS_U>
__global__ void __launch_bounds__(TN_THREADS, 1) temporalnet_kernel(const __grid_constant__ TnArgs args) {
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int KC = C / 8, HALF = C / 2;
    constexpr uint32_t LBO_B = HALF * 16, TAPB = KC * LBO_B;
    const distb200_temporalnet_desc& d = args.d;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rank = (int)ptx::cluster_ctarank();
    const uint32_t bar0 = ptx::smem_u32(smem + args.off_bar);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem + args.off_bar + 8 * TN_NBAR);

    if (warp == TN_MMA_WARP) {
        if (lane == 0) {
            ptx::mbar_init(bar0 + 8u * B_LN_FULL, 2 * TN_LN_WARPS);
            ptx::mbar_init(bar0 + 8u * B_EPI1, 2 * TN_EPI_WARPS);
            ptx::mbar_init(bar0 + 8u * B_ACC2_FREE, 2 * TN_EPI_WARPS);
            ptx::mbar_init(bar0 + 8u * (B_ACC2_FREE + 1), 2 * TN_EPI_WARPS);
            for (int s = B_C311; s < B_W_READY; ++s) ptx::mbar_init(bar0 + 8u * s, 1);
            ptx::mbar_init(bar0 + 8u * B_W_READY, 2 * (TN_EPI_WARPS + 1));
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc_2sm(ptx::smem_u32(tmem_slot_ptr), TN_TMEM_COLS);
        ptx::tmem_relinquish_2sm();
    }
    if (warp >= TN_EPI_WARPS && warp < TN_MMA_WARP) {
        // LayerNorm parameters (the other warps fill the rest of the parameter block below)
        float* par = reinterpret_cast<float*>(smem + args.off_par);
        for (int idx = threadIdx.x - TN_EPI_WARPS * 32; idx < C; idx += TN_LN_WARPS * 32) {
            par[idx] = d.ln_g[idx];
            par[C + idx] = d.ln_b[idx];
        }
    }
    ptx::tc_fence_before();
    ptx::cluster_sync();      // barriers initialised, TMEM allocated, LayerNorm parameters in place - in both CTAs
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    grid_dep_sync();          // PDL: nothing above read the streams
    const int first = (int)blockIdx.x * args.n_per;

    if (warp >= TN_EPI_WARPS && warp < TN_MMA_WARP) {
        ln_role<C, HAS_U>(args, smem, warp - TN_EPI_WARPS, lane, first);        // starts on the first window right away
    } else {
        // ---- resident operands, loaded by the nine warps that have nothing to do until the first window is normalised: this
        //      CTA's half of the output channels of all twelve taps, chunked K-major.  Eight consecutive lanes take the eight rows
        //      of a core matrix (128 contiguous bytes of shared memory), the next lanes the next k-chunks (64 contiguous bytes
        //      per weight row and instruction).
        const int tid = warp == TN_MMA_WARP ? TN_EPI_WARPS * 32 + lane : (int)threadIdx.x;
        constexpr int NT = (TN_EPI_WARPS + 1) * 32;
        const bf16* w1 = reinterpret_cast<const bf16*>(d.w1);
        const bf16* w2 = reinterpret_cast<const bf16*>(d.w2);
        constexpr int TOTAL = 12 * HALF * KC;
        for (int idx = tid; idx < TOTAL; idx += NT) {
            const int r8 = idx & 7;
            int rest = idx >> 3;
            const int kc = rest % KC;
            rest /= KC;
            const int grp = rest % (HALF / 8), t = rest / (HALF / 8);
            const int n = grp * 8 + r8;
            const bf16* src = (t < 3 ? w1 + (long long)t * C * C : w2 + (long long)(t - 3) * C * C) + (long long)(rank * HALF + n) * C + kc * 8;
            *reinterpret_cast<uint4*>(smem + (uint32_t)t * TAPB + (uint32_t)kc * LBO_B + (uint32_t)n * 16u) = __ldg(reinterpret_cast<const uint4*>(src));
        }
        float* par = reinterpret_cast<float*>(smem + args.off_par);
        for (int idx = tid; idx < C; idx += NT) {
            par[2 * C + idx] = d.b1[idx];
            par[3 * C + idx] = d.b2[idx];
        }
        uint4* z4 = reinterpret_cast<uint4*>(smem + args.off_z);
        for (int idx = tid; idx < KC * args.z_rows; idx += NT) z4[idx] = make_uint4(0u, 0u, 0u, 0u);
        ptx::fence_proxy_async();
        __syncwarp();
        if (lane == 0) arrive_leader(ptx::mapa(bar0 + 8u * B_W_READY, 0));
        if (warp < TN_EPI_WARPS) {
            asm volatile("bar.sync 1, %0;" ::"n"(TN_EPI_WARPS * 32) : "memory");       // z is zeroed and the biases are in place before any epilogue warp goes on
            epi_role<C, HAS_U>(args, smem, tmem_base, warp, lane, first);
        } else if (rank == 0) {
            mma_role<C>(args, smem, tmem_base);
        }
    }

    ptx::tc_fence_before();
    ptx::cluster_sync();      // nobody leaves while the peer may still arrive on this CTA's barriers
    if (warp == TN_MMA_WARP) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc_2sm(tmem_base, TN_TMEM_COLS);
    }
}

template <int C, bool HAS_U>
int launch_instance(const TnArgs& args, int grid, int smem, cudaStream_t stream) {
    static bool attr_done[DISTB200_MAX_DEVICES] = {};
    const int dev = current_device();
    if (!attr_done[dev]) {
        cudaError_t e = cudaFuncSetAttribute(temporalnet_kernel<C, HAS_U>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        DISTB200_REQUIRE(e == cudaSuccess, "temporalnet: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        attr_done[dev] = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(TN_THREADS);
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    cudaError_t e = cudaLaunchKernelEx(&cfg, temporalnet_kernel<C, HAS_U>, args);
    DISTB200_REQUIRE(e == cudaSuccess, "temporalnet: launch failed: %s", cudaGetErrorString(e));
    return check_launch("temporalnet");
}

}  // namespace

int temporalnet_launch(const distb200_temporalnet_desc& d, cudaStream_t stream) {
    const int C = d.channels, g = d.grid;
    DISTB200_REQUIRE(C == 32 || C == 64 || C == 96, "temporalnet: channels=%d (supported: 32, 64, 96)", C);
    DISTB200_REQUIRE(g >= 1 && d.frames >= 1 && d.clips >= 0, "temporalnet: bad sizes");
    if (d.clips == 0) return 0;
    int br = TN_M / g - 2;
    if (TN_M / (g + 2) < br) br = TN_M / (g + 2);
    if (g < br) br = g;
    DISTB200_REQUIRE(br >= 1, "temporalnet: grid=%d is too wide for a 128-row tile", g);
    DISTB200_REQUIRE(!d.u || (d.alpha >= 1 && d.frames % d.alpha == 0), "temporalnet: frames=%d is not a multiple of alpha=%d", d.frames, d.alpha);
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    DISTB200_REQUIRE(al16(d.x) && al16(d.u) && al16(d.w1) && al16(d.w2) && al16(d.out) && al16(d.out2), "temporalnet: pointers must be 16-byte aligned");
    DISTB200_REQUIRE(!d.out2 || (d.ld_out2 % 8 == 0 && d.out2_cstep % 8 == 0 && d.out2_gdiv >= 0), "temporalnet: ld_out2 / out2_cstep must be multiples of 8");
    DISTB200_REQUIRE(d.out != d.x, "temporalnet: out must not alias x (neighbouring frames are read while a frame is written)");

    TnArgs a;
    a.d = d;
    if (!d.u) a.d.alpha = 1;
    a.nb = (g + br - 1) / br;
    a.band_base = g / a.nb;
    a.band_rem = g - a.band_base * a.nb;
    a.br = (g + a.nb - 1) / a.nb;
    a.W = g + 2;
    a.z_rows = (a.br + 2) * a.W;      // no guard rows: only pad-column outputs (never stored) read one row before / behind the tile
    a.P = g * g;
    a.ts = d.u ? d.frames / d.alpha : d.frames;
    const long long units = (long long)d.clips * a.nb * d.frames;
    DISTB200_REQUIRE(units < (1ll << 30) && (long long)d.clips * d.frames < (1ll << 30), "temporalnet: too many frames");
    a.units = (int)units;
    const int KC = C / 8;
    a.off_ln = (uint32_t)(12 * KC * (C / 2) * 16);
    a.off_z = a.off_ln + (uint32_t)(3 * KC) * TN_LBO_A;
    a.off_stg = a.off_z + (uint32_t)(KC * a.z_rows * 16);
    a.off_stg = (a.off_stg + 127u) & ~127u;
    a.off_par = a.off_stg + TN_EPI_WARPS * TN_STAGE_BYTES;
    a.off_bar = a.off_par + (uint32_t)(4 * C * 4);
    const int smem = (int)a.off_bar + 8 * TN_NBAR + 16;
    DISTB200_REQUIRE(smem <= 227 * 1024, "temporalnet: needs %d bytes of shared memory", smem);

    int ctas = sm_count() & ~1;
    if (d.max_ctas >= 2 && d.max_ctas < ctas) ctas = d.max_ctas & ~1;
    if ((long long)ctas > ((units + 1) & ~1ll)) ctas = (int)((units + 1) & ~1ll);
    a.n_per = (int)((units + ctas - 1) / ctas);
    ctas = (int)(((units + a.n_per - 1) / a.n_per + 1) & ~1ll);
    if (C == 32) return d.u ? launch_instance<32, true>(a, ctas, smem, stream) : launch_instance<32, false>(a, ctas, smem, stream);
    if (C == 64) return d.u ? launch_instance<64, true>(a, ctas, smem, stream) : launch_instance<64, false>(a, ctas, smem, stream);
    return d.u ? launch_instance<96, true>(a, ctas, smem, stream) : launch_instance<96, false>(a, ctas, smem, stream);
}

}  // namespace distb200

using namespace distb200;

#ifdef DISTB200_TN_TRACE
extern "C" int distb200_debug_tn_trace(long long* buf, int dbg) {
    cudaMemcpyToSymbol(g_tn_trace, &buf, sizeof(buf));
    cudaMemcpyToSymbol(g_tn_dbg, &dbg, sizeof(dbg));
    return 0;
}
#endif

extern "C" int distb200_temporalnet(const distb200_temporalnet_desc* desc, void* stream) {
    DISTB200_REQUIRE(desc != nullptr, "temporalnet: null descriptor");
    const distb200_temporalnet_desc& d = *desc;
    DISTB200_REQUIRE(d.x && d.ln_g && d.ln_b && d.w1 && d.b1 && d.w2 && d.b2, "temporalnet: null pointer");
    DISTB200_REQUIRE(d.out || d.out2, "temporalnet: no output");
    DISTB200_REQUIRE(d.dtype == DISTB200_BF16, "temporalnet: the fused kernel takes bf16 operands (compose the fp32 path from layernorm + gemm)");
    return temporalnet_launch(d, (cudaStream_t)stream);
}
