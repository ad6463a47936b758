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
//                                               band (lr = 0 is the halo row above) at row 1 + lr * (g+2) + c + 1.  Pad columns
//                                               and rows outside the frame hold zeros, so spatial tap (i, j) of the (1,3,3)
//                                               convolution is the SAME tile with its start address moved by (i*(g+2) + j) rows.
//   staging   8 x 2 KB                          transposes of the output epilogue
// TMEM: conv(3,1,1) accumulators at columns [0, 2C), conv(1,3,3) accumulators at [2C, 4C) (two stages each).
//
// Warps: 0-7 epilogue (TMEM lane quadrant = warp & 3, channel half = warp >> 2), 8-15 LayerNorm producers, 16 MMA issuer
// (leader CTA only).  Per iteration `it` the issuer runs conv311(it) then conv133(it-1); the epilogue warps turn the conv311
// accumulator into z (bias, QuickGELU, bf16) and then finish unit it-1 (bias, residual, QuickGELU, fp32 + bf16 stores) while the
// tensor core is busy with the next unit; the LayerNorm warps fill the ring one frame ahead.
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"

namespace distb200 {

namespace {

constexpr int TN_EPI_WARPS = 8;
constexpr int TN_LN_WARPS = 8;
constexpr int TN_MMA_WARP = TN_EPI_WARPS + TN_LN_WARPS;
constexpr int TN_THREADS = (TN_MMA_WARP + 1) * 32;
constexpr int TN_M = 128;                     // accumulator rows per CTA
constexpr int TN_STAGE_BYTES = 2048;          // per epilogue warp: 32 rows x 16 fp32 columns
constexpr int TN_TMEM_COLS = 512;
// mbarriers (8 bytes each, same offsets in both CTAs)
constexpr int B_LN_FULL = 0;                  // leader: LayerNorm warps of both CTAs filled the ring for iteration `it`
constexpr int B_EPI1 = 1;                     // leader: z of iteration j written (and its conv311 accumulator read) in both CTAs
constexpr int B_ACC2_FREE = 2;                // leader [2]: conv133 accumulator stage read by both CTAs
constexpr int B_C311 = 4;                     // each CTA [2]: conv311(it) completed (tcgen05.commit, multicast)
constexpr int B_C133 = 6;                     // each CTA [2]: conv133(j) completed
constexpr int TN_NBAR = 8;

struct alignas(16) TnArgs {
    distb200_temporalnet_desc d;
    int nb;            // bands per frame
    int br;            // rows of the largest band
    int W;             // pitch of the padded z layout = g + 2
    int z_rows;        // (br + 2) * W + 2
    int P;             // g * g
    int ts;            // frames / alpha
    int n_per;         // iterations (units per CTA)
    int units;         // clips * nb * frames
    uint32_t off_ln, off_z, off_stg, off_par, off_bar;
};

struct Unit {
    bool valid;
    int clip, tau, r0, nrows, frame;
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
    const int base = a.d.grid / a.nb, rem = a.d.grid - base * a.nb;
    un.nrows = base + (band < rem ? 1 : 0);
    un.r0 = band * base + (band < rem ? band : rem);
    un.frame = un.clip * (int)T + un.tau;
    return un;
}

__device__ __forceinline__ float gelu_fast(float x) {        // bf16 destinations: one MUFU.TANH (see gemm_tcgen05.cu)
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * x));
    return x * fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ float gelu_precise(float x) { return __fdividef(x, 1.0f + __expf(-1.702f * x)); }

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// ---- LayerNorm producers -------------------------------------------------------------------------------------------
// Eight lanes per row (lane & 7 = 16-byte piece inside every 128-byte third of the row), four rows per warp and pass, the
// eight warps cover the 128 slot rows in four passes; a job (one frame) is done in two halves of two passes so that
// x and u of a half (up to 12 x 16 bytes per lane) are in flight together.
template <int C, bool HAS_U>
__device__ __forceinline__ void ln_role(const TnArgs& a, uint8_t* smem, int wl, int lane, int first) {
    constexpr int NJ = C / 32;
    constexpr int KC = C / 8;
    constexpr uint32_t LBO_A = TN_M * 16;
    constexpr uint32_t SLOT = KC * LBO_A;
    const distb200_temporalnet_desc& d = a.d;
    const int g = d.grid, T = d.frames, P = a.P;
    const int sub = lane >> 3, cl = lane & 7;
    const float* par = reinterpret_cast<const float*>(smem + a.off_par);
    const uint32_t bar0 = ptx::smem_u32(smem + a.off_bar);
    const uint32_t ln_full = ptx::mapa(bar0 + 8u * B_LN_FULL, 0);
    const float inv_c = 1.0f / (float)C;

    for (int it = 0; it < a.n_per; ++it) {
        const Unit un = decode_unit(a, first + it);
        int js[3], jf[3], nj = 0;
        if (un.valid) {
            if (it == 0 || un.tau == 0) {                 // a new (clip, band): the whole window
                js[nj] = (it + 2) % 3; jf[nj++] = un.tau - 1;
                js[nj] = it % 3;       jf[nj++] = un.tau;
            }
            js[nj] = (it + 1) % 3; jf[nj++] = un.tau + 1 < T ? un.tau + 1 : -1;
        }
        // every slot written in iteration `it` was last read by conv311(it - 1)
        bool waited = it == 0;
        const uint32_t wbar = bar0 + 8u * (uint32_t)(B_C311 + ((it - 1) & 1)), wpar = (uint32_t)((it - 1) >> 1) & 1u;
        const int rows_valid = (un.nrows + 2) * g, pos0 = (un.r0 - 1) * g;
        for (int j = 0; j < nj; ++j) {
            uint8_t* slot = smem + a.off_ln + (uint32_t)js[j] * SLOT;
            const int sigma = jf[j];
            const bool zero = sigma < 0;                    // frame outside the clip: the convolution pads LayerNorm's OUTPUT with zeros
            const float* xf = d.x + (long long)(un.clip * T + (zero ? 0 : sigma)) * P * C + 4 * cl;
            const float* uf = HAS_U ? d.u + (long long)(un.clip * a.ts + (zero ? 0 : sigma) / d.alpha) * P * C + 4 * cl : nullptr;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float4 v[2][NJ], w[2][NJ];
                bool ok[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int m = (2 * half + q) * 32 + wl * 4 + sub, pos = pos0 + m;
                    ok[q] = !zero && m < rows_valid && pos >= 0 && pos < P;
#pragma unroll
                    for (int jj = 0; jj < NJ; ++jj) {
                        v[q][jj] = ok[q] ? ldg4(xf + (long long)pos * C + 32 * jj) : make_float4(0.f, 0.f, 0.f, 0.f);
                        if (HAS_U) w[q][jj] = ok[q] ? ldg4(uf + (long long)pos * C + 32 * jj) : make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                }
                if (!waited) {
                    ptx::mbar_wait(wbar, wpar);
                    waited = true;
                }
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    const int m = (2 * half + q) * 32 + wl * 4 + sub;
                    float s = 0.f;
#pragma unroll
                    for (int jj = 0; jj < NJ; ++jj) {
                        if (HAS_U) { v[q][jj].x += w[q][jj].x; v[q][jj].y += w[q][jj].y; v[q][jj].z += w[q][jj].z; v[q][jj].w += w[q][jj].w; }
                        s += (v[q][jj].x + v[q][jj].y) + (v[q][jj].z + v[q][jj].w);
                    }
                    s += __shfl_xor_sync(0xffffffffu, s, 1);
                    s += __shfl_xor_sync(0xffffffffu, s, 2);
                    s += __shfl_xor_sync(0xffffffffu, s, 4);
                    const float mean = s * inv_c;
                    float qq = 0.f;
#pragma unroll
                    for (int jj = 0; jj < NJ; ++jj) {
                        v[q][jj].x -= mean; v[q][jj].y -= mean; v[q][jj].z -= mean; v[q][jj].w -= mean;
                        qq += fmaf(v[q][jj].x, v[q][jj].x, v[q][jj].y * v[q][jj].y) + fmaf(v[q][jj].z, v[q][jj].z, v[q][jj].w * v[q][jj].w);
                    }
                    qq += __shfl_xor_sync(0xffffffffu, qq, 1);
                    qq += __shfl_xor_sync(0xffffffffu, qq, 2);
                    qq += __shfl_xor_sync(0xffffffffu, qq, 4);
                    const float rstd = rsqrtf(qq * inv_c + d.eps);
                    if (m < rows_valid && (zero || ok[q])) {
#pragma unroll
                        for (int jj = 0; jj < NJ; ++jj) {
                            const float4 gm = *reinterpret_cast<const float4*>(par + 32 * jj + 4 * cl);
                            const float4 bt = *reinterpret_cast<const float4*>(par + C + 32 * jj + 4 * cl);
                            uint2 pk;
                            if (zero) pk = make_uint2(0u, 0u);
                            else pk = make_uint2(pack_bf16x2(fmaf(v[q][jj].x * rstd, gm.x, bt.x), fmaf(v[q][jj].y * rstd, gm.y, bt.y)),
                                                 pack_bf16x2(fmaf(v[q][jj].z * rstd, gm.z, bt.z), fmaf(v[q][jj].w * rstd, gm.w, bt.w)));
                            // channels 32 jj + 4 cl .. +3  ->  k-chunk 4 jj + cl / 2, half (cl & 1)
                            *reinterpret_cast<uint2*>(slot + (uint32_t)(4 * jj + (cl >> 1)) * LBO_A + (uint32_t)m * 16u + (uint32_t)(cl & 1) * 8u) = pk;
                        }
                    }
                }
            }
        }
        if (!waited) ptx::mbar_wait(wbar, wpar);       // keep every warp within one phase of the others (the barrier counts arrivals per phase)
        ptx::fence_proxy_async();                      // generic-proxy stores -> visible to the tensor core's reads
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_release_cluster(ln_full);
    }
}

// ---- MMA issuer (leader CTA) -----------------------------------------------------------------------------------------
template <int C>
__device__ __forceinline__ void mma_role(const TnArgs& a, uint8_t* smem, uint32_t tmem_base) {
    constexpr int KC = C / 8, KS = C / 16;
    constexpr uint32_t LBO_A = TN_M * 16, SLOT = KC * LBO_A;
    constexpr uint32_t LBO_B = (C / 2) * 16, TAPB = KC * LBO_B;
    const uint32_t LBO_Z = (uint32_t)a.z_rows * 16u;
    const uint32_t s0 = ptx::smem_u32(smem);
    const uint32_t bar0 = s0 + a.off_bar;
    const uint32_t idesc = ptx::umma_idesc_bf16(2 * TN_M, C);
    for (int it = 0; it <= a.n_per; ++it) {
        if (it < a.n_per) {
            // ---- conv(3,1,1) of iteration it: taps = ring slots (it+2, it, it+1) % 3 = frames tau-1, tau, tau+1
            ptx::mbar_wait_acquire_cluster(bar0 + 8u * B_LN_FULL, (uint32_t)it & 1u);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint32_t acc = tmem_base + (uint32_t)((it & 1) * C);
#pragma unroll
                for (int k = 0; k < 3; ++k) {
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
            ptx::mbar_wait_acquire_cluster(bar0 + 8u * B_EPI1, (uint32_t)j & 1u);
            ptx::mbar_wait_acquire_cluster(bar0 + 8u * (uint32_t)(B_ACC2_FREE + (j & 1)), ((uint32_t)(j >> 1) & 1u) ^ 1u);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
                const uint32_t acc = tmem_base + (uint32_t)(2 * C + (j & 1) * C);
#pragma unroll
                for (int t = 0; t < 9; ++t) {
                    const uint32_t sa = s0 + a.off_z + (uint32_t)((t / 3) * a.W + (t % 3)) * 16u;
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
template <int C, bool HAS_U>
__device__ __forceinline__ void epi_role(const TnArgs& a, uint8_t* smem, uint32_t tmem_base, int warp, int lane, int first) {
    constexpr int NCH = C / 32;             // 16-column chunks per warp (its half of the channels)
    const distb200_temporalnet_desc& d = a.d;
    const int g = d.grid, T = d.frames, P = a.P, W = a.W;
    const int quad = warp & 3, hf = warp >> 2;
    const uint32_t LBO_Z = (uint32_t)a.z_rows * 16u;
    const float* par = reinterpret_cast<const float*>(smem + a.off_par);
    const float* b1s = par + 2 * C;
    const float* b2s = par + 3 * C;
    uint8_t* zb = smem + a.off_z;
    float4* stg = reinterpret_cast<float4*>(smem + a.off_stg + warp * TN_STAGE_BYTES);
    const uint32_t bar0 = ptx::smem_u32(smem + a.off_bar);
    const uint32_t epi1_done = ptx::mapa(bar0 + 8u * B_EPI1, 0);
    const uint32_t acc2_free = ptx::mapa(bar0 + 8u * B_ACC2_FREE, 0);
    const uint32_t tq = tmem_base + ((uint32_t)(quad * 32) << 16);
    // thread = row domain (conv311 accumulator -> z): accumulator row m1 = position (lr1, c1) of the haloed band
    const int m1 = quad * 32 + lane;
    const int lr1 = m1 / g, c1 = m1 - lr1 * g;
    const bool in_z = m1 < (a.br + 2) * g;
    const uint32_t zoff = (uint32_t)(1 + lr1 * W + c1 + 1) * 16u;
    // column domain (conv133 accumulator -> global): four lanes per row, eight rows per pass; accumulator row m2 = padded
    // position (lo, cc) of the band, cc = 0 and cc = g + 1 being pad columns
    const int rs = lane >> 2, l4 = lane & 3;
    int rel[4], lo_[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const int m2 = quad * 32 + 8 * p + rs;
        const int lo = m2 / W, cc = m2 - lo * W;
        lo_[p] = lo;
        rel[p] = (cc >= 1 && cc <= g) ? lo * g + cc - 1 : -1;
    }

    auto load_res = [&](const Unit& un, int ch, float4* rx, float4* ru) {
        const int col = hf * (C / 2) + 16 * ch + 4 * l4;
        const float* xf = d.x + (long long)un.frame * P * C + col;
        const float* uf = HAS_U ? d.u + (long long)(un.clip * a.ts + un.tau / d.alpha) * P * C + col : nullptr;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            const bool ok = un.valid && rel[p] >= 0 && lo_[p] < un.nrows;
            const long long pos = un.r0 * g + rel[p];
            rx[p] = ok ? ldg4(xf + pos * C) : make_float4(0.f, 0.f, 0.f, 0.f);
            if constexpr (HAS_U) ru[p] = ok ? ldg4(uf + pos * C) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };

    Unit prev;
    prev.valid = false;
    prev.clip = prev.tau = prev.r0 = prev.nrows = prev.frame = 0;
    float4 rx[4], ru[HAS_U ? 4 : 1];
    for (int it = 0; it <= a.n_per; ++it) {
        Unit un = prev;
        if (it < a.n_per) un = decode_unit(a, first + it);
        if (it >= 1) load_res(prev, 0, rx, ru);        // residual of the unit about to be finished: in flight during the z epilogue

        if (it < a.n_per) {
            ptx::mbar_wait(bar0 + 8u * (uint32_t)(B_C311 + (it & 1)), (uint32_t)(it >> 1) & 1u);
            if (it >= 1) ptx::mbar_wait(bar0 + 8u * (uint32_t)(B_C133 + ((it - 1) & 1)), (uint32_t)((it - 1) >> 1) & 1u);   // z is free again
            ptx::tc_fence_after();
            // ---- conv311 accumulator -> z (bf16, padded layout); rows outside the frame are zeros, not q(b1)
            const int pos = (un.r0 - 1) * g + m1;
            const bool ok = un.valid && m1 < (un.nrows + 2) * g && pos >= 0 && pos < P;
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const int col0 = hf * (C / 2) + 16 * ch;
                uint32_t acc[16];
                ptx::tmem_ld16(tq + (uint32_t)((it & 1) * C + col0), acc);
                ptx::tmem_ld_wait();
                if (in_z) {
#pragma unroll
                    for (int h8 = 0; h8 < 2; ++h8) {
                        uint32_t pk[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int c = col0 + 8 * h8 + 2 * e;
                            const float v0 = gelu_fast(__uint_as_float(acc[8 * h8 + 2 * e]) + b1s[c]);
                            const float v1 = gelu_fast(__uint_as_float(acc[8 * h8 + 2 * e + 1]) + b1s[c + 1]);
                            pk[e] = ok ? pack_bf16x2(v0, v1) : 0u;
                        }
                        *reinterpret_cast<uint4*>(zb + (uint32_t)(col0 / 8 + h8) * LBO_Z + zoff) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
            }
            ptx::fence_proxy_async();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive_release_cluster(epi1_done);
        } else {
            ptx::mbar_wait(bar0 + 8u * (uint32_t)(B_C133 + ((it - 1) & 1)), (uint32_t)((it - 1) >> 1) & 1u);
            ptx::tc_fence_after();
        }

        if (it >= 1) {
            // ---- conv133 accumulator of unit it-1 -> out / out2
            const int j = it - 1;
            long long row2 = 0;
            int col2 = 0;
            if (d.out2_gdiv > 0) {
                const int qd = prev.frame / d.out2_gdiv;
                row2 = (long long)qd * d.out2_gstride + d.out2_roff;
                col2 = (prev.frame - qd * d.out2_gdiv) * d.out2_cstep;
            } else {
                row2 = (long long)prev.frame * P;
            }
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                const int col0 = hf * (C / 2) + 16 * ch;
                uint32_t acc[16];
                ptx::tmem_ld16(tq + (uint32_t)(2 * C + (j & 1) * C + col0), acc);
                ptx::tmem_ld_wait();
                if (ch == NCH - 1) {                     // the accumulator stage is in registers: release it
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive_release_cluster(acc2_free + 8u * (uint32_t)(j & 1));
                }
                // transpose 32 rows x 16 columns through shared memory: 16-byte piece jj of row r at slot jj ^ ((r >> 1) & 3)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                    stg[lane * 4 + (jj ^ ((lane >> 1) & 3))] = make_float4(__uint_as_float(acc[4 * jj]), __uint_as_float(acc[4 * jj + 1]),
                                                                           __uint_as_float(acc[4 * jj + 2]), __uint_as_float(acc[4 * jj + 3]));
                __syncwarp();
                const float4 bias = *reinterpret_cast<const float4*>(b2s + col0 + 4 * l4);
                float4 v[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const int rr = 8 * p + rs;
                    v[p] = stg[rr * 4 + (l4 ^ ((rr >> 1) & 3))];
                    float4 r = rx[p];
                    if constexpr (HAS_U) { r.x += ru[p].x; r.y += ru[p].y; r.z += ru[p].z; r.w += ru[p].w; }
                    v[p].x += bias.x + r.x; v[p].y += bias.y + r.y; v[p].z += bias.z + r.z; v[p].w += bias.w + r.w;
                }
                __syncwarp();                            // the staging tile is rewritten by the next chunk
                if (ch + 1 < NCH) load_res(prev, ch + 1, rx, ru);      // next chunk's residual: in flight during the activation and the stores
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const bool ok = prev.valid && rel[p] >= 0 && lo_[p] < prev.nrows;
                    if (ok) {
                        float4 y;
                        y.x = gelu_precise(v[p].x);
                        y.y = gelu_precise(v[p].y);
                        y.z = gelu_precise(v[p].z);
                        y.w = gelu_precise(v[p].w);
                        const long long pos = prev.r0 * g + rel[p];
                        const int col = col0 + 4 * l4;
                        if (d.out) *reinterpret_cast<float4*>(d.out + ((long long)prev.frame * P + pos) * C + col) = y;
                        if (d.out2)
                            *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(d.out2) + (row2 + pos) * d.ld_out2 + col2 + col) =
                                make_uint2(pack_bf16x2(y.x, y.y), pack_bf16x2(y.z, y.w));
                    }
                }
            }
        }
        prev = un;
    }
    (void)T;
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
            for (int s = B_C311; s < TN_NBAR; ++s) ptx::mbar_init(bar0 + 8u * s, 1);
            ptx::fence_barrier_init();
        }
        __syncwarp();
        ptx::tmem_alloc_2sm(ptx::smem_u32(tmem_slot_ptr), TN_TMEM_COLS);
        ptx::tmem_relinquish_2sm();
    }
    // ---- resident operands: this CTA's half of the output channels of all twelve taps, chunked K-major.  Eight consecutive
    //      lanes take the eight rows of a core matrix (128 contiguous bytes of shared memory), the next lanes the next k-chunks
    //      (64 contiguous bytes per weight row and instruction).
    {
        const bf16* w1 = reinterpret_cast<const bf16*>(d.w1);
        const bf16* w2 = reinterpret_cast<const bf16*>(d.w2);
        constexpr int TOTAL = 12 * HALF * KC;
        for (int idx = threadIdx.x; idx < TOTAL; idx += TN_THREADS) {
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
        for (int idx = threadIdx.x; idx < C; idx += TN_THREADS) {
            par[idx] = d.ln_g[idx];
            par[C + idx] = d.ln_b[idx];
            par[2 * C + idx] = d.b1[idx];
            par[3 * C + idx] = d.b2[idx];
        }
        uint4* z4 = reinterpret_cast<uint4*>(smem + args.off_z);
        for (int idx = threadIdx.x; idx < KC * args.z_rows; idx += TN_THREADS) z4[idx] = make_uint4(0u, 0u, 0u, 0u);
    }
    ptx::fence_proxy_async();
    ptx::tc_fence_before();
    ptx::cluster_sync();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    grid_dep_sync();          // PDL: the weights are constants; the streams are touched below

    const int first = (int)blockIdx.x * args.n_per;
    if (warp < TN_EPI_WARPS) epi_role<C, HAS_U>(args, smem, tmem_base, warp, lane, first);
    else if (warp < TN_MMA_WARP) ln_role<C, HAS_U>(args, smem, warp - TN_EPI_WARPS, lane, first);
    else if (rank == 0) mma_role<C>(args, smem, tmem_base);

    ptx::tc_fence_before();
    ptx::cluster_sync();      // nobody leaves while the peer may still arrive on this CTA's barriers
    if (warp == TN_MMA_WARP) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc_2sm(tmem_base, TN_TMEM_COLS);
    }
}

template <int C, bool HAS_U>
int launch_instance(const TnArgs& args, int grid, int smem, cudaStream_t stream) {
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_done[dev]) {
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
    a.br = (g + a.nb - 1) / a.nb;
    a.W = g + 2;
    a.z_rows = (a.br + 2) * a.W + 2;
    a.P = g * g;
    a.ts = d.u ? d.frames / d.alpha : d.frames;
    const long long units = (long long)d.clips * a.nb * d.frames;
    DISTB200_REQUIRE(units < (1ll << 30) && (long long)d.clips * d.frames < (1ll << 30), "temporalnet: too many frames");
    a.units = (int)units;
    const int KC = C / 8;
    a.off_ln = (uint32_t)(12 * KC * (C / 2) * 16);
    a.off_z = a.off_ln + (uint32_t)(3 * KC * TN_M * 16);
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

extern "C" int distb200_temporalnet(const distb200_temporalnet_desc* desc, void* stream) {
    DISTB200_REQUIRE(desc != nullptr, "temporalnet: null descriptor");
    const distb200_temporalnet_desc& d = *desc;
    DISTB200_REQUIRE(d.x && d.ln_g && d.ln_b && d.w1 && d.b1 && d.w2 && d.b2, "temporalnet: null pointer");
    DISTB200_REQUIRE(d.out || d.out2, "temporalnet: no output");
    DISTB200_REQUIRE(d.dtype == DISTB200_BF16, "temporalnet: the fused kernel takes bf16 operands (compose the fp32 path from layernorm + gemm)");
    return temporalnet_launch(d, (cudaStream_t)stream);
}
