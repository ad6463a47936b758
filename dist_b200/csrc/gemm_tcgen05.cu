// distb200_gemm on the 5th-generation tensor cores (sm_100a): persistent, warp-specialised.
//
//   warp 0        TMA producer: cp.async.bulk.tensor tiles of A (4-D map, zero fill outside the tensor gives the
//                 convolution halos / ragged edges for free) and B (3-D map: k, n, tap) into a ring of
//                 128B-swizzled shared-memory stages, completion on mbarriers
//   warp 1        MMA issuer: one elected lane issues tcgen05.mma (M=128, N=block_n, K=16, bf16 x bf16 -> fp32)
//                 into one of two TMEM accumulator stages; tcgen05.commit releases smem stages and publishes
//                 the finished accumulator
//   warps 2..9    epilogue (two warps per TMEM lane quadrant, alternating 32-column chunks): tcgen05.ld gives one
//                 accumulator row per thread; the 32x32 block is transposed through a swizzled shared-memory
//                 staging buffer so that lanes run along columns: bias, fp32 residual loads, QuickGELU and the
//                 fp32 / bf16 stores are then fully coalesced (128 B per row), with the row remapping of
//                 include/distb200.h.  The next tile's MMAs overlap the epilogue (two TMEM accumulator stages).
//
// Tile = rows_per_tile (<=128) output rows of one group x block_n columns; the reduction runs over
// num_taps x ceil(K/64) k-blocks.  All rows of the 128-row MMA that lie outside the tile are computed on whatever
// the smem holds and never stored.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace distb200 {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;             // 64 bf16 = one 128-byte swizzle row
constexpr int MAX_STAGES = 8;
// Epilogue warps: 4 TMEM lane quadrants x EW/4 interleaved 32-column chunks.  8 warps leave room for the deepest
// operand ring (long-K GEMMs), 12 warps hide more of the epilogue's latency (narrow / short-K GEMMs).
__host__ __device__ constexpr int num_threads(int ew) { return 64 + ew * 32; }
__host__ __device__ constexpr int staging_bytes(int ew) { return ew * 32 * 32 * 4; }
__host__ __device__ constexpr int smem_budget(int ew) { return 227 * 1024 - staging_bytes(ew) - 2048; }
constexpr uint32_t A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr uint32_t TMEM_COLS = 512;
constexpr uint32_t ACC_STAGE_COLS = 256;

struct alignas(64) TcArgs {
    CUtensorMap tm_a;
    CUtensorMap tm_b;
    distb200_gemm_desc d;
    int block_n;
    int stages;
    int k_blocks;
    int kpack;                // 64-wide k-blocks per pipeline stage: 2 when 64 < K <= 128 (one stage per tap: half the barrier round trips)
    int rows_per_tile;
    int tiles_per_group;
    int n_tiles;
    int group_dim;
    int dbg;
    int ctas;                 // 1, or 2 = CTA pair (cta_group::2): 256 rows x block_n per pair, each CTA holds half of B
    uint32_t b_stage_bytes;
    uint32_t tx_bytes;
    long long total_tiles;
};

// Bottleneck probes, compiled in only with -DDISTB200_GEMM_PROBES and driven by the environment variable
// DISTB200_GEMM_DBG: 1 = no epilogue memory traffic, 2 = no MMA issue, 4 = no A loads, 8 = no B loads,
// 16 = no TMEM reads in the epilogue, 32 = one k-iteration per tile.
#ifdef DISTB200_GEMM_PROBES
#define PROBE(bit) (args.dbg & (bit))
#else
#define PROBE(bit) false
#endif

struct TileCoord {
    long long gi;
    int r0;
    int n0;
};

// tile = index of a unit of work of one CTA (ctas == 1) or of one CTA pair (ctas == 2, tiles_per_group then counts
// pairs and the CTA of rank r takes the (2 * pair + r)-th row tile of the group, which may lie past the group's end)
__device__ __forceinline__ TileCoord decode_tile(const TcArgs& a, long long tile, int rank) {
    // 32-bit arithmetic: the host guarantees total_tiles < 2^31 (64-bit divisions cost hundreds of cycles on the
    // single-warp producer / issuer paths)
    TileCoord t;
    const uint32_t tl = (uint32_t)tile;
    const uint32_t m_idx = tl / (uint32_t)a.n_tiles;
    const uint32_t n_idx = tl - m_idx * (uint32_t)a.n_tiles;
    const uint32_t gi = m_idx / (uint32_t)a.tiles_per_group;
    t.gi = gi;
    t.r0 = (int)((m_idx - gi * (uint32_t)a.tiles_per_group) * (uint32_t)a.ctas + (uint32_t)rank) * a.rows_per_tile;
    t.n0 = (int)n_idx * a.block_n;
    return t;
}

// sigmoid(1.702 x) = 0.5 + 0.5 tanh(0.851 x): one MUFU.TANH per value (tanh.approx.f32, |rel err| < 2^-11), used for bf16 outputs
// where the rounding of the result dominates that error.  (tanh.approx.f16x2 compiles to two MUFU.TANH.F16 plus pack / unpack
// instructions - it saves no MUFU work and costs six instead of four instructions per value.)
__device__ __forceinline__ void quick_gelu_pair_fast(float& x0, float& x1) {
    float t0, t1;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(0.851f * x0));
    asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(0.851f * x1));
    x0 = x0 * fmaf(0.5f, t0, 0.5f);
    x1 = x1 * fmaf(0.5f, t1, 0.5f);
}

__device__ __forceinline__ float quick_gelu_precise(float x) { return __fdividef(x, 1.0f + __expf(-1.702f * x)); }

// registers (thread = row) -> swizzled smem: 16-byte chunk j of row r lives at chunk slot (j ^ (r & 7))
__device__ __forceinline__ void stage_block(const uint32_t* acc, float* stage, int lane) {
    float4* st4 = reinterpret_cast<float4*>(stage);
#pragma unroll
    for (int j = 0; j < 8; ++j)
        st4[lane * 8 + (j ^ (lane & 7))] = make_float4(__uint_as_float(acc[4 * j]), __uint_as_float(acc[4 * j + 1]),
                                                       __uint_as_float(acc[4 * j + 2]), __uint_as_float(acc[4 * j + 3]));
}

// Epilogue variants (compile-time so that the inner loop carries no flag tests):
//   OUT:  0 = fp32 `out`, 1 = bf16 `out`, 2 = fp32 `out` + bf16 `out2`, 3 = anything (runtime flags)
//   ACT:  0 none, 1 QuickGELU via tanh.approx.f16x2, 2 QuickGELU via ex2 + rcp
// LNF: LayerNorm folded into the GEMM (see distb200_gemm_desc.ln_stats): acc -> rstd * (acc - mean * wsum[n]) + bias[n]
// STATS: per-row sum / sum of squares of the bf16 values written to out2 (distb200_gemm_desc.stat_partials), accumulated in st[]
template <int OUT, bool RES, int ACT, bool LNF = false, bool STATS = false>
struct Epi {
    // Column domain: eight lanes cover the 32 columns of a row (four adjacent columns = 16 bytes each), the four
    // lane groups take four consecutive rows.  One loop iteration = four rows: 512 B of fp32 (256 B of bf16) per warp
    // instruction, every row a full 128-byte line.
    static __device__ __forceinline__ void prefetch(const distb200_gemm_desc& d, float4* rv, float4& bias, float4& ws, long long res_row0,
                                                    long long srow0, int n, int sub, int rows_here, bool col_ok) {
        bias = (d.bias && col_ok) ? __ldg(reinterpret_cast<const float4*>(d.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (LNF) {
            // per-row (mean, rstd) travel in the (otherwise unused) residual registers; the eight lanes of a row read one address
            ws = col_ok ? __ldg(reinterpret_cast<const float4*>(d.ln_wsum + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
            const float2* sp = reinterpret_cast<const float2*>(d.ln_stats) + srow0 + sub;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float2 st = 4 * i + sub < rows_here ? __ldg(sp + 4 * i) : make_float2(0.f, 0.f);
                rv[i] = make_float4(st.x, st.y, 0.f, 0.f);
            }
        }
        if (RES) {
            const float* rp = d.res + (res_row0 + sub) * d.ld_res + n;
            const long long step = 4 * d.ld_res;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                rv[i] = (col_ok && 4 * i + sub < rows_here) ? *reinterpret_cast<const float4*>(rp) : make_float4(0.f, 0.f, 0.f, 0.f);
                rp += step;
            }
        }
    }

    template <bool FULL>
    static __device__ __forceinline__ void finish_rows(const distb200_gemm_desc& d, const float* stage, const float4* rv, const float4 bias,
                                                       const float4 ws, int lane, int n, int rows_here, long long dst_row0, float2* st,
                                                       long long dst2_row0, int col2) {
        const int sub = lane >> 3, chunk = lane & 7;
        char* o1 = nullptr;
        char* o2 = nullptr;
        long long s1 = 0, s2 = 0;
        const bool f32_1 = OUT == 0 || OUT == 2 || (OUT == 3 && d.out_dtype == DISTB200_F32);
        if (OUT != 3 || d.out) {
            const int es = f32_1 ? 4 : 2;
            o1 = reinterpret_cast<char*>(d.out) + ((dst_row0 + sub) * d.ld_out + n) * es;
            s1 = 4 * d.ld_out * es;
        }
        const bool f32_2 = OUT == 3 && d.out2_dtype == DISTB200_F32;
        if (OUT == 2 || (OUT == 3 && d.out2)) {
            const int es = f32_2 ? 4 : 2;
            o2 = reinterpret_cast<char*>(d.out2) + ((dst2_row0 + sub) * d.ld_out2 + n + col2) * es;
            s2 = 4 * d.ld_out2 * es;
        }
        const bool act_on = ACT != 0 && n >= d.act_from && (d.act_to == 0 || n < d.act_to);      // the activation may cover a column range only
        const float4* st4 = reinterpret_cast<const float4*>(stage);
        float4 vv[8];                      // all shared-memory reads first: their latency overlaps instead of serialising
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = 4 * i + sub;
            vv[i] = st4[row * 8 + (chunk ^ (row & 7))];
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int row = 4 * i + sub;
            float4 v = vv[i];
            if (LNF) {
                const float mu = rv[i].x, rs = rv[i].y;
                v.x = rs * (v.x - mu * ws.x); v.y = rs * (v.y - mu * ws.y); v.z = rs * (v.z - mu * ws.z); v.w = rs * (v.w - mu * ws.w);
            }
            v.x += bias.x; v.y += bias.y; v.z += bias.z; v.w += bias.w;
            if (RES) {
                v.x += rv[i].x; v.y += rv[i].y; v.z += rv[i].z; v.w += rv[i].w;
            }
            if (ACT == 1 && act_on) {
                quick_gelu_pair_fast(v.x, v.y);
                quick_gelu_pair_fast(v.z, v.w);
            }
            if (ACT == 2 && act_on) {
                v.x = quick_gelu_precise(v.x); v.y = quick_gelu_precise(v.y);
                v.z = quick_gelu_precise(v.z); v.w = quick_gelu_precise(v.w);
            }
            if (FULL || row < rows_here) {
                if (o1) {
                    if (f32_1) *reinterpret_cast<float4*>(o1) = v;
                    else *reinterpret_cast<uint2*>(o1) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
                }
                if (o2) {
                    if (f32_2) *reinterpret_cast<float4*>(o2) = v;
                    else {
                        const uint32_t p0 = pack_bf16x2(v.x, v.y), p1 = pack_bf16x2(v.z, v.w);
                        *reinterpret_cast<uint2*>(o2) = make_uint2(p0, p1);
                        if (STATS) {            // statistics of exactly the values the consuming GEMM will read
                            const float a = __uint_as_float(p0 << 16), b = __uint_as_float(p0 & 0xffff0000u);
                            const float c = __uint_as_float(p1 << 16), e = __uint_as_float(p1 & 0xffff0000u);
                            st[i].x += (a + b) + (c + e);
                            st[i].y += fmaf(a, a, b * b) + fmaf(c, c, e * e);
                        }
                    }
                }
            }
            o1 += s1;
            o2 += s2;
        }
    }

    static __device__ __forceinline__ void finish(const distb200_gemm_desc& d, const float* stage, const float4* rv, const float4 bias,
                                                  const float4 ws, int lane, int n, int rows_here, long long dst_row0, float2* st,
                                                  long long dst2_row0, int col2) {
        if (rows_here >= 32) finish_rows<true>(d, stage, rv, bias, ws, lane, n, rows_here, dst_row0, st, dst2_row0, col2);
        else finish_rows<false>(d, stage, rv, bias, ws, lane, n, rows_here, dst_row0, st, dst2_row0, col2);
    }
};

// Per-tile quantities of one epilogue warp.
struct EpiTile {
    long long dst0, res0, srow0, dst2_0;
    int n0, ncols, rows_valid, col2;
};

__device__ __forceinline__ EpiTile epi_tile(const TcArgs& args, long long tile, int rank, int quad) {
    const distb200_gemm_desc& d = args.d;
    const TileCoord tc = decode_tile(args, tile, rank);
    EpiTile t;
    const long long rows_left = d.rows_per_group - tc.r0;
    int rows_valid = (int)(rows_left < args.rows_per_tile ? rows_left : (long long)args.rows_per_tile) - quad * 32;
    t.rows_valid = rows_valid > 32 ? 32 : rows_valid;                   // rows of this quadrant that exist
    const long long r = (long long)tc.r0 + quad * 32;
    t.dst0 = tc.gi * d.out_gstride + d.out_roff + r;
    t.res0 = tc.gi * d.res_gstride + d.res_roff + r;
    t.srow0 = tc.gi * d.rows_per_group + r;            // row of the logical A matrix (index of ln_stats)
    t.dst2_0 = t.dst0;                                 // out2 follows out unless it is placed independently (out2_gdiv)
    t.col2 = 0;
    if (d.out2_gdiv > 0) {
        const uint32_t q = (uint32_t)tc.gi / (uint32_t)d.out2_gdiv;
        t.dst2_0 = (long long)q * d.out2_gstride + d.out2_roff + r;
        t.col2 = (int)((uint32_t)tc.gi - q * (uint32_t)d.out2_gdiv) * d.out2_cstep;
    }
    t.n0 = tc.n0;
    t.ncols = min(args.block_n, d.n - tc.n0);
    return t;
}

// The chunks of a warp form one sequence across its tiles (tile, c0 = half*32, half*32+64, ...).  The residual of
// chunk i+1 is requested before chunk i is processed, so that a warp always has one chunk of loads (4 KB) in
// flight while it transposes, activates and stores another: the epilogue of the narrow, short-K GEMMs is bound by
// the latency of these loads, not by their bandwidth.
template <int OUT, bool RES, int ACT, int EW, bool LNF = false, bool STATS = false>
__device__ __forceinline__ void epilogue_role(const TcArgs& args, uint32_t tmem_base, float* stage, uint32_t tfull0, uint32_t tempty0,
                                              int warp, int lane, long long tile0, long long tile_step, int rank) {
    typedef Epi<OUT, RES, ACT, LNF, STATS> E;
    const distb200_gemm_desc& d = args.d;
    float2 st[STATS ? 8 : 1];                // rows 4*i + sub of this warp's quadrant: (sum, sum of squares) over the warp's chunks of the tile
    if (STATS) {
#pragma unroll
        for (int i = 0; i < 8; ++i) st[i] = make_float2(0.f, 0.f);
    }
    const int quad = warp & 3;              // TMEM lanes [32*quad, 32*quad+32) are the ones this warp may read
    const int half = (warp - 2) >> 2;       // which of the alternating 32-column chunks
    const int sub = lane >> 3, cl = (lane & 7) * 4;
    int acc_stage = 0;
    uint32_t acc_phase = 0;
    if (tile0 >= args.total_tiles) return;
    long long tile = tile0;
    EpiTile cur = epi_tile(args, tile, rank, quad);
    int c0 = half * 32;
    constexpr int CSTEP = 32 * (EW / 4);
    constexpr bool EPI_PIPE_RES = EW <= 8 && !LNF;   // double-buffered residual prefetch only when registers allow
    float4 rv[8], bias = make_float4(0.f, 0.f, 0.f, 0.f), ws = bias;
    if (!PROBE(1) && c0 < cur.ncols) E::prefetch(d, rv, bias, ws, cur.res0, cur.srow0, cur.n0 + c0 + cl, sub, cur.rows_valid, c0 + cl < cur.ncols);
    bool waited = false;
    while (true) {
        // ---- next chunk of this warp (possibly in its next tile): request its residual now
        long long next_tile = tile;
        int next_c0 = c0 + CSTEP;
        EpiTile nxt = cur;
        if (next_c0 >= cur.ncols) {
            next_tile = tile + tile_step;
            next_c0 = half * 32;
            if (next_tile < args.total_tiles) nxt = epi_tile(args, next_tile, rank, quad);
        }
        float4 rv2[EPI_PIPE_RES ? 8 : 1], bias2 = make_float4(0.f, 0.f, 0.f, 0.f), ws2 = bias2;
        const bool have_next = !PROBE(1) && next_tile < args.total_tiles && next_c0 < nxt.ncols;
        if (EPI_PIPE_RES && have_next)
            E::prefetch(d, rv2, bias2, ws2, nxt.res0, nxt.srow0, nxt.n0 + next_c0 + cl, sub, nxt.rows_valid, next_c0 + cl < nxt.ncols);

        // ---- current chunk
        if (!waited) {                       // first chunk of the tile (or no chunk at all: still observe the phase)
            ptx::mbar_wait(tfull0 + 8u * acc_stage, acc_phase);
            ptx::tc_fence_after();
            waited = true;
        }
        if (c0 < cur.ncols) {
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc_stage * ACC_STAGE_COLS;
            const int n = cur.n0 + c0 + cl;
            const bool col_ok = c0 + cl < cur.ncols;
            uint32_t acc[32];
            const bool two = c0 + 16 < cur.ncols;
            if (!PROBE(16)) {
                ptx::tmem_ld16(taddr + (uint32_t)c0, acc);
                if (two) ptx::tmem_ld16(taddr + (uint32_t)c0 + 16u, acc + 16);
                ptx::tmem_ld_wait();
            }
            if (cur.rows_valid > 0 && !PROBE(1)) {
                stage_block(acc, stage, lane);
                __syncwarp();
                for (int rep = 0; rep < d.out_rep; ++rep) {
                    if (rep > 0) E::prefetch(d, rv, bias, ws, cur.res0 + (long long)rep * d.res_rep_stride, cur.srow0, n, sub, cur.rows_valid, col_ok);
                    if (col_ok) E::finish(d, stage, rv, bias, ws, lane, n, cur.rows_valid, cur.dst0 + (long long)rep * d.out_rep_stride, st,
                                          cur.dst2_0 + (long long)rep * d.out_rep_stride, cur.col2);
                }
                __syncwarp();
            }
        }
        if (next_tile != tile) {
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                // the accumulator stage is released to the MMA issuer, which lives in the leader CTA of a pair
                if (rank == 0) ptx::mbar_arrive(tempty0 + 8u * acc_stage);
                else ptx::mbar_arrive_cluster(ptx::mapa(tempty0 + 8u * acc_stage, 0));
            }
            if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1u; }
            waited = false;
            if (STATS) {
                // The tile is done: fold the eight lanes of every row together and publish this warp's slot of the row statistics.
                const int slot = (cur.n0 / args.block_n) * (EW / 4) + half;
                float2* dst = reinterpret_cast<float2*>(d.stat_partials) + (cur.srow0 + sub) * DISTB200_STAT_SLOTS + slot;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    float a = st[i].x, q = st[i].y;
#pragma unroll
                    for (int o = 1; o < 8; o <<= 1) {
                        a += __shfl_xor_sync(0xffffffffu, a, o);
                        q += __shfl_xor_sync(0xffffffffu, q, o);
                    }
                    if ((lane & 7) == 0 && 4 * i + sub < cur.rows_valid) dst[4 * i * DISTB200_STAT_SLOTS] = make_float2(a, q);
                    st[i] = make_float2(0.f, 0.f);
                }
            }
            if (next_tile >= args.total_tiles) break;
        }
        tile = next_tile;
        c0 = next_c0;
        cur = nxt;
        if (EPI_PIPE_RES) {
            bias = bias2;
            ws = ws2;
            if (RES) {
#pragma unroll
                for (int i = 0; i < 8; ++i) rv[i] = rv2[i];
            }
        } else if (have_next) {
            E::prefetch(d, rv, bias, ws, cur.res0, cur.srow0, cur.n0 + c0 + cl, sub, cur.rows_valid, c0 + cl < cur.ncols);
        }
    }
}


template <int CTAS, int EW>
__global__ void __launch_bounds__(num_threads(EW), 1) gemm_tcgen05_kernel(const __grid_constant__ TcArgs args) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const distb200_gemm_desc& d = args.d;

    // carve shared memory: [stages x (A | B)] [epilogue staging] [barriers]; swizzle-128B needs 1024-byte aligned
    // stage bases.  Both CTAs of a pair use the same offsets (the pair MMA addresses the peer's operands by offset).
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sub_bytes = A_STAGE_BYTES + args.b_stage_bytes;            // one 64-wide k-block of A and B
    const uint32_t stage_bytes = sub_bytes * (uint32_t)args.kpack;
    const uint32_t staging_base = smem_base + (uint32_t)args.stages * stage_bytes;
    const uint32_t bar_base = staging_base + staging_bytes(EW);
    // barriers (8 bytes each): full[MAX_STAGES], empty[MAX_STAGES], tmem_full[2], tmem_empty[2], then the TMEM base
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (MAX_STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * MAX_STAGES + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * MAX_STAGES + 4);
    uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - ptx::smem_u32(smem_raw)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = CTAS == 2 ? (int)ptx::cluster_ctarank() : 0;
    const long long tile0 = blockIdx.x / CTAS, tile_step = gridDim.x / CTAS;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&args.tm_a);
        ptx::prefetch_tensormap(&args.tm_b);
        for (int s = 0; s < args.stages; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(tfull_bar(s), 1);
            ptx::mbar_init(tempty_bar(s), EW * CTAS);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        if (CTAS == 2) { ptx::tmem_alloc_2sm(tmem_slot, TMEM_COLS); ptx::tmem_relinquish_2sm(); }
        else { ptx::tmem_alloc(tmem_slot, TMEM_COLS); ptx::tmem_relinquish(); }
    }
    ptx::tc_fence_before();
    if (CTAS == 2) ptx::cluster_sync(); else __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    grid_dep_sync();          // PDL: everything above overlapped the previous kernel's tail; global memory is touched below

    const int iters = PROBE(32) ? 1 : d.num_taps * (args.k_blocks / args.kpack);

    if (warp == 0) {
        // ===================== TMA producer =====================
        // The whole warp runs the loop (warp-uniform control flow keeps addresses and coordinates in uniform
        // registers); one elected lane issues the asynchronous operations.
        {
            int stage = 0;
            uint32_t phase = 0;
            const int b_rows = args.block_n / CTAS;     // a pair splits the B tile: rank r holds rows [r * block_n / 2, ...)
            for (long long tile = tile0; tile < args.total_tiles; tile += tile_step) {
                const TileCoord tc = decode_tile(args, tile, rank);
                const int img_row0 = d.img_w > 0 ? tc.r0 / d.img_w : 0;
                for (int tap = 0; tap < (PROBE(32) ? 1 : d.num_taps); ++tap) {
                    int c1, c2, c3;
                    if (d.img_w > 0) {
                        c1 = d.tap_off[tap][0];
                        c2 = img_row0 + d.tap_off[tap][1];
                        c3 = (int)tc.gi + d.tap_off[tap][2];
                    } else {
                        c1 = tc.r0 + d.tap_off[tap][0];
                        c2 = d.tap_off[tap][1] + (args.group_dim == 3 ? 0 : (int)tc.gi);
                        c3 = d.tap_off[tap][2] + (args.group_dim == 3 ? (int)tc.gi : 0);
                    }
                    for (int kb0 = 0; kb0 < (PROBE(32) ? 1 : args.k_blocks); kb0 += args.kpack) {
                        ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
                        const uint32_t s0 = smem_base + (uint32_t)stage * stage_bytes;
                        if (CTAS == 2) {
                            // All bytes of the pair are counted on the leader's barrier, where the MMA issuer waits.  The peer
                            // needs no arrive of its own: it can only refill a stage after the leader's MMAs released it
                            // (multicast commit below), i.e. after the barrier's previous phase completed.
                            const uint32_t lead_full = ptx::mapa(full_bar(stage), 0);
                            if (ptx::elect_one()) {
                                if (rank == 0) ptx::mbar_arrive_expect_tx(full_bar(stage), 2 * args.tx_bytes * (uint32_t)args.kpack);
                                for (int j = 0; j < args.kpack; ++j) {
                                    const uint32_t sa = s0 + (uint32_t)j * sub_bytes, sb = sa + A_STAGE_BYTES;
                                    ptx::tma_load_4d_2sm(sa, &args.tm_a, lead_full, (kb0 + j) * BLOCK_K, c1, c2, c3);
                                    ptx::tma_load_3d_2sm(sb, &args.tm_b, lead_full, (kb0 + j) * BLOCK_K, tc.n0 + rank * b_rows, tap);
                                }
                            }
                        } else if (ptx::elect_one()) {
#ifdef DISTB200_GEMM_PROBES
                            uint32_t tx = args.tx_bytes;
                            const uint32_t b_bytes = (uint32_t)(BLOCK_K * args.block_n * 2);
                            if (PROBE(4)) tx -= args.tx_bytes - b_bytes;
                            if (PROBE(8)) tx -= b_bytes;
                            tx *= (uint32_t)args.kpack;
                            if (tx) ptx::mbar_arrive_expect_tx(full_bar(stage), tx); else ptx::mbar_arrive(full_bar(stage));
                            for (int j = 0; j < args.kpack; ++j) {
                                const uint32_t sa = s0 + (uint32_t)j * sub_bytes, sb = sa + A_STAGE_BYTES;
                                if (!PROBE(4)) ptx::tma_load_4d(sa, &args.tm_a, full_bar(stage), (kb0 + j) * BLOCK_K, c1, c2, c3);
                                if (!PROBE(8)) ptx::tma_load_3d(sb, &args.tm_b, full_bar(stage), (kb0 + j) * BLOCK_K, tc.n0, tap);
                            }
#else
                            ptx::mbar_arrive_expect_tx(full_bar(stage), args.tx_bytes * (uint32_t)args.kpack);
                            for (int j = 0; j < args.kpack; ++j) {
                                const uint32_t sa = s0 + (uint32_t)j * sub_bytes, sb = sa + A_STAGE_BYTES;
                                ptx::tma_load_4d(sa, &args.tm_a, full_bar(stage), (kb0 + j) * BLOCK_K, c1, c2, c3);
                                ptx::tma_load_3d(sb, &args.tm_b, full_bar(stage), (kb0 + j) * BLOCK_K, tc.n0, tap);
                            }
#endif
                        }
                        __syncwarp();
                        if (++stage == args.stages) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA of a pair only) =====================
        // Whole warp in the loop, one elected lane issues tcgen05.mma / tcgen05.commit (see the producer).
        if (rank == 0) {
            const uint32_t idesc = ptx::umma_idesc_bf16(BLOCK_M * CTAS, args.block_n);
            const uint64_t desc0 = ptx::umma_desc_k_sw128(smem_base);          // stage bases are 1024-byte multiples: no carry into other fields
            const int ksteps_last = (d.k - (args.k_blocks - 1) * BLOCK_K + 15) / 16;
            int stage = 0;
            uint32_t phase = 0;
            int acc_stage = 0;
            uint32_t acc_phase = 0;
            for (long long tile = tile0; tile < args.total_tiles; tile += tile_step) {
                ptx::mbar_wait(tempty_bar(acc_stage), acc_phase ^ 1u);
                ptx::tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)acc_stage * ACC_STAGE_COLS;
                int kb = 0;
                for (int it = 0; it < iters; ++it) {
                    ptx::mbar_wait(full_bar(stage), phase);
                    ptx::tc_fence_after();
                    if (ptx::elect_one()) {
                        for (int j = 0; j < args.kpack; ++j) {
                            const uint64_t da = desc0 + (uint64_t)(((uint32_t)stage * stage_bytes + (uint32_t)j * sub_bytes) >> 4);
                            const uint64_t db = da + (uint64_t)(A_STAGE_BYTES >> 4);
                            const int ksteps = kb + j == args.k_blocks - 1 ? ksteps_last : BLOCK_K / 16;
                            for (int ks = 0; ks < (PROBE(2) ? 0 : ksteps); ++ks) {
                                // advancing 16 bf16 (32 bytes) along K inside the swizzle row: +2 in the (addr >> 4) field
                                if (CTAS == 2) ptx::mma_f16_ss_2sm(tmem_d, da + (uint64_t)(2 * ks), db + (uint64_t)(2 * ks), idesc, (it | j | ks) != 0);
                                else ptx::mma_f16_ss(tmem_d, da + (uint64_t)(2 * ks), db + (uint64_t)(2 * ks), idesc, (it | j | ks) != 0);
                            }
                        }
                        if (CTAS == 2) ptx::mma_commit_2sm(empty_bar(stage), 3);    // frees the stage in both CTAs
                        else ptx::mma_commit(empty_bar(stage));
                    }
                    __syncwarp();
                    if (++stage == args.stages) { stage = 0; phase ^= 1u; }
                    kb += args.kpack;
                    if (kb >= args.k_blocks) kb = 0;
                }
                if (ptx::elect_one()) {
                    if (CTAS == 2) ptx::mma_commit_2sm(tfull_bar(acc_stage), 3);    // both CTAs' epilogues may start
                    else ptx::mma_commit(tfull_bar(acc_stage));
                }
                __syncwarp();
                if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1u; }
            }
        }
    } else {
        // ===================== epilogue (8 warps: 4 lane quadrants x 2 column halves) =====================
        float* stage = reinterpret_cast<float*>(smem_raw + (staging_base - ptx::smem_u32(smem_raw))) + (warp - 2) * 32 * 32;
        const uint32_t tf = tfull_bar(0), te = tempty_bar(0);
        const bool gelu = d.act == DISTB200_ACT_QUICKGELU;
        const bool bf_only = d.out && d.out_dtype == DISTB200_BF16 && !d.out2;
        const bool f_only = d.out && d.out_dtype == DISTB200_F32 && !d.out2;
        const bool f_and_bf = d.out && d.out_dtype == DISTB200_F32 && d.out2 && d.out2_dtype == DISTB200_BF16;
#define DISTB200_EPI(OUT, RES, ACT) epilogue_role<OUT, RES, ACT, EW>(args, tmem_base, stage, tf, te, warp, lane, tile0, tile_step, rank)
        if (d.ln_stats) {                                  // host side guarantees: bf16 out only, no residual
            if (gelu) epilogue_role<1, false, 1, EW, true>(args, tmem_base, stage, tf, te, warp, lane, tile0, tile_step, rank);
            else epilogue_role<1, false, 0, EW, true>(args, tmem_base, stage, tf, te, warp, lane, tile0, tile_step, rank);
        } else if (bf_only && !d.res && !gelu) DISTB200_EPI(1, false, 0);
        else if (bf_only && !d.res && gelu) DISTB200_EPI(1, false, 1);
        else if (f_only && d.res && !gelu) DISTB200_EPI(0, true, 0);
        else if (f_and_bf && d.res && !gelu && d.stat_partials) epilogue_role<2, true, 0, EW, false, true>(args, tmem_base, stage, tf, te, warp, lane, tile0, tile_step, rank);
        else if (f_and_bf && d.res && !gelu) DISTB200_EPI(2, true, 0);
        else if (f_and_bf && d.res && gelu) DISTB200_EPI(2, true, 2);
        else if (d.res) {
            if (gelu) DISTB200_EPI(3, true, 2);
            else DISTB200_EPI(3, true, 0);
        } else {
            if (gelu) DISTB200_EPI(3, false, 2);
            else DISTB200_EPI(3, false, 0);
        }
#undef DISTB200_EPI
    }

    ptx::tc_fence_before();
    if (CTAS == 2) ptx::cluster_sync(); else __syncthreads();      // nobody leaves while the peer may still address this CTA
    if (warp == 1) {
        ptx::tc_fence_after();
        if (CTAS == 2) ptx::tmem_dealloc_2sm(tmem_base, TMEM_COLS);
        else ptx::tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ---- host side ---------------------------------------------------------------------------------------

int pick_block_n(int n) {
    if (n <= 256) return n;
    for (int bn = 256; bn >= 64; bn -= 16)
        if (n % bn == 0) return bn;
    return 256;
}

}  // namespace

int gemm_tcgen05_launch(const distb200_gemm_desc& d, cudaStream_t stream) {
    const long long total_rows = d.groups * d.rows_per_group;
    if (total_rows == 0 || d.n == 0) return 0;
    DISTB200_REQUIRE(d.n % 16 == 0, "gemm(tcgen05): n=%d must be a multiple of 16", d.n);
    DISTB200_REQUIRE(d.num_taps >= 1 && d.num_taps <= DISTB200_MAX_TAPS, "gemm(tcgen05): num_taps=%d", d.num_taps);
    DISTB200_REQUIRE(d.a_stride[0] == 1, "gemm(tcgen05): a_stride[0] must be 1");
    DISTB200_REQUIRE(d.ldb % 8 == 0 && d.b_tap_stride % 8 == 0, "gemm(tcgen05): ldb / b_tap_stride must be multiples of 8");
    DISTB200_REQUIRE(!d.bias || (reinterpret_cast<uintptr_t>(d.bias) & 15) == 0, "gemm(tcgen05): bias must be 16-byte aligned");
    DISTB200_REQUIRE(!d.res || ((reinterpret_cast<uintptr_t>(d.res) & 15) == 0 && d.ld_res % 4 == 0),
                    "gemm(tcgen05): res must be 16-byte aligned with ld_res %% 4 == 0");
    DISTB200_REQUIRE(!d.out || ((reinterpret_cast<uintptr_t>(d.out) & 15) == 0 && d.ld_out % 8 == 0),
                    "gemm(tcgen05): out must be 16-byte aligned with ld_out %% 8 == 0");
    DISTB200_REQUIRE(!d.out2 || ((reinterpret_cast<uintptr_t>(d.out2) & 15) == 0 && d.ld_out2 % 8 == 0),
                    "gemm(tcgen05): out2 must be 16-byte aligned with ld_out2 %% 8 == 0");
    DISTB200_REQUIRE(d.out_rep >= 1, "gemm(tcgen05): out_rep must be >= 1");
    if (d.ln_stats) {
        DISTB200_REQUIRE(d.ln_wsum && d.out && d.out_dtype == DISTB200_BF16 && !d.out2 && !d.res && d.out_rep == 1 && d.num_taps == 1,
                        "gemm(tcgen05): the folded LayerNorm needs ln_wsum, a bf16 `out` only, no residual, one tap");
        DISTB200_REQUIRE((reinterpret_cast<uintptr_t>(d.ln_stats) & 7) == 0 && (reinterpret_cast<uintptr_t>(d.ln_wsum) & 15) == 0,
                        "gemm(tcgen05): ln_stats / ln_wsum alignment");
    }

    DISTB200_REQUIRE(d.out2_gdiv >= 0 && (d.out2_gdiv == 0 || (d.out2 && d.out_rep == 1 && d.out2_cstep % 8 == 0 && d.groups < (1ll << 31))),
                    "gemm(tcgen05): out2_gdiv needs out2, out_rep == 1 and out2_cstep %% 8 == 0");
    DISTB200_REQUIRE(d.act_from >= 0 && d.act_from % 4 == 0 && d.act_to >= 0 && d.act_to % 4 == 0,
                    "gemm(tcgen05): act_from=%d / act_to=%d must be non-negative multiples of 4", d.act_from, d.act_to);
    if (d.stat_partials) {
        DISTB200_REQUIRE(d.out && d.out_dtype == DISTB200_F32 && d.out2 && d.out2_dtype == DISTB200_BF16 && d.res && d.act == DISTB200_ACT_NONE &&
                        d.out_rep == 1 && !d.ln_stats, "gemm(tcgen05): stat_partials needs fp32 out + bf16 out2 + res, no activation, out_rep 1");
        DISTB200_REQUIRE((reinterpret_cast<uintptr_t>(d.stat_partials) & 7) == 0, "gemm(tcgen05): stat_partials alignment");
    }

    TcArgs args;
    args.d = d;
    args.group_dim = d.group_dim == 3 ? 3 : 2;
    args.block_n = d.block_n > 0 ? d.block_n : pick_block_n(d.n);
    if (d.block_n == 0 && d.img_w == 0 && d.impl != DISTB200_IMPL_TCGEN05_2CTA && !d.stat_partials) {
        // Few-row GEMMs (the ada-pooling head: 32 ... 256 rows) would occupy a handful of SMs, each walking the whole reduction
        // alone: narrower column tiles spread the weight matrix over more SMs (the operand re-reads stay in L2).
        const long long row_tiles0 = d.groups * ((d.rows_per_group + BLOCK_M - 1) / BLOCK_M);
        static const int narrow_env = getenv("DISTB200_GEMM_NARROW") ? atoi(getenv("DISTB200_GEMM_NARROW")) : 1;
        if (narrow_env && row_tiles0 * ((d.n + args.block_n - 1) / args.block_n) * 4 <= sm_count()) {
            for (int bn = args.block_n - 16; bn >= 16; bn -= 16) {
                if (d.n % bn) continue;
                args.block_n = bn;
                if (row_tiles0 * (d.n / bn) >= sm_count()) break;
            }
        }
    }
    DISTB200_REQUIRE(args.block_n % 16 == 0 && args.block_n >= 16 && args.block_n <= 256, "gemm(tcgen05): block_n=%d", args.block_n);
    args.k_blocks = (d.k + BLOCK_K - 1) / BLOCK_K;
    if (d.img_w > 0) {
        DISTB200_REQUIRE(d.img_w <= BLOCK_M, "gemm(tcgen05): img_w=%d exceeds %d", d.img_w, BLOCK_M);
        DISTB200_REQUIRE(d.rows_per_group % d.img_w == 0, "gemm(tcgen05): rows_per_group must be a multiple of img_w");
        const int img_h = (int)(d.rows_per_group / d.img_w);
        int rows_h = BLOCK_M / d.img_w;
        if (rows_h > img_h) rows_h = img_h;
        // balance the tiles of one image: e.g. 14 rows -> 2 tiles of 7 rather than 9 + 5
        const int tiles = (img_h + rows_h - 1) / rows_h;
        rows_h = (img_h + tiles - 1) / tiles;
        args.rows_per_tile = rows_h * d.img_w;
    } else {
        args.rows_per_tile = BLOCK_M;
    }
    const int row_tiles = (int)((d.rows_per_group + args.rows_per_tile - 1) / args.rows_per_tile);
    args.n_tiles = (d.n + args.block_n - 1) / args.block_n;
    // CTA pairs (cta_group::2) halve the L2 -> SM traffic of B; they need an even split of the B tile and enough tiles
    static const int ctas_env = getenv("DISTB200_GEMM_CTAS") ? atoi(getenv("DISTB200_GEMM_CTAS")) : 2;
    args.ctas = (ctas_env == 2 && args.block_n % 32 == 0 && d.groups * row_tiles * args.n_tiles >= 2 * sm_count()) ? 2 : 1;
    if (args.ctas == 2 && d.impl == DISTB200_IMPL_AUTO) {
        // Wave quantisation: a pair takes two row tiles of ONE group, so an odd tile count per group leaves half-empty pairs
        // (IntegrationNetwork's temporal convolution: 13 tiles per clip -> 224 pairs = 3.03 waves of 74).  Fall back to single
        // CTAs when they fill the machine clearly better.
        const long long t1 = d.groups * row_tiles * args.n_tiles, t2 = d.groups * ((row_tiles + 1) / 2) * args.n_tiles;
        const long long u1 = sm_count(), u2 = sm_count() / 2;
        const double eff1 = (double)t1 / (double)(((t1 + u1 - 1) / u1) * u1);
        const double eff2 = (double)t1 / (double)(((t2 + u2 - 1) / u2) * u2 * 2);
        if (eff1 > 1.1 * eff2) args.ctas = 1;
    }
    if (d.impl == DISTB200_IMPL_TCGEN05_1CTA) args.ctas = 1;
    if (d.impl == DISTB200_IMPL_TCGEN05_2CTA) {
        DISTB200_REQUIRE(args.block_n % 32 == 0, "gemm(tcgen05): CTA pairs need block_n %% 32 == 0 (block_n=%d)", args.block_n);
        args.ctas = 2;
    }
    static const int dbg_env = getenv("DISTB200_GEMM_DBG") ? atoi(getenv("DISTB200_GEMM_DBG")) : 0;
    args.dbg = dbg_env;
    args.tiles_per_group = (row_tiles + args.ctas - 1) / args.ctas;
    args.total_tiles = d.groups * args.tiles_per_group * args.n_tiles;
    DISTB200_REQUIRE(args.total_tiles < (1ll << 31) && d.groups < (1ll << 31), "gemm(tcgen05): too many tiles (%lld)", args.total_tiles);
    args.b_stage_bytes = (uint32_t)(args.block_n / args.ctas) * BLOCK_K * 2;
    // Several 64-wide k-blocks per pipeline stage when the MMA work of one k-block (proportional to block_n) is smaller than
    // the barrier round trip of a stage: narrow tiles pack up to 4 k-blocks, wide ones up to 2; at least 3 stages must remain.
    static const int kpack_env = getenv("DISTB200_GEMM_KPACK") ? atoi(getenv("DISTB200_GEMM_KPACK")) : 4;
    args.kpack = 1;
    {
        const int limit = args.block_n <= 128 ? 4 : (args.block_n <= 192 ? 2 : 1);
        // epilogue width is decided below from K; use the larger staging footprint for the budget check
        for (int p = 2; p <= limit && p <= kpack_env; ++p)
            if (args.k_blocks % p == 0 && 3 * p * (int)(A_STAGE_BYTES + args.b_stage_bytes) <= smem_budget(12)) args.kpack = p;
    }
    const uint32_t stage_bytes = (A_STAGE_BYTES + args.b_stage_bytes) * (uint32_t)args.kpack;
    // epilogue width: long reductions want the deepest operand ring, everything else the better latency hiding
    static const int ew_env = getenv("DISTB200_GEMM_EPI_WARPS") ? atoi(getenv("DISTB200_GEMM_EPI_WARPS")) : 0;
    // Measured per call site (B/16, 32 clips): 8 warps (double-buffered residual prefetch, deeper operand ring) also win for the
    // plain single-tap GEMMs from K = 768 up (out_proj 1.40 -> 1.34 ms, qkv 1.64 -> 1.62, DiST input GEMM 0.92 -> 0.89), while
    // QuickGELU epilogues (FC1 2.41 vs 2.53) and the tapped convolutions ((1,3,3): 0.62 vs 0.86) need the 12.
    const bool plain = d.num_taps == 1 && d.act == DISTB200_ACT_NONE && d.k >= 768;
    const int ew = ew_env == 8 || ew_env == 12 ? ew_env : (((long long)d.k * d.num_taps >= 2048 || plain) ? 8 : 12);
    DISTB200_REQUIRE(!d.stat_partials || args.n_tiles * (ew / 4) <= DISTB200_STAT_SLOTS,
                    "gemm(tcgen05): %d column tiles x %d epilogue phases exceed the %d statistic slots", args.n_tiles, ew / 4, DISTB200_STAT_SLOTS);
    args.stages = smem_budget(ew) / (int)stage_bytes;
    if (args.stages > MAX_STAGES) args.stages = MAX_STAGES;
    if (args.stages < 2) args.stages = 2;

    // A: (k, c1, c2, c3)
    {
        long long dims[4] = {d.a_dim[0], d.a_dim[1], d.a_dim[2], d.a_dim[3]};
        long long str[4] = {1, d.a_stride[1], d.a_stride[2], d.a_stride[3]};
        int box[4] = {BLOCK_K, 1, 1, 1};
        if (d.img_w > 0) {
            box[1] = d.img_w;
            box[2] = args.rows_per_tile / d.img_w;
        } else {
            box[1] = args.rows_per_tile;
        }
        if (make_map(&args.tm_a, d.a, 4, dims, str, box, "A")) return 1;
        args.tx_bytes = (uint32_t)(box[0] * box[1] * box[2] * box[3] * 2);
    }
    // B: (k, n, tap)
    {
        long long dims[3] = {d.k, d.n, d.num_taps};
        long long str[3] = {1, d.ldb, d.num_taps > 1 ? d.b_tap_stride : d.ldb * d.n};
        int box[3] = {BLOCK_K, args.block_n / args.ctas, 1};
        if (make_map(&args.tm_b, d.b, 3, dims, str, box, "B")) return 1;
        args.tx_bytes += (uint32_t)(BLOCK_K * (args.block_n / args.ctas) * 2);
    }

    const int smem = args.stages * (int)stage_bytes + staging_bytes(ew) + 1024 + 8 * (2 * MAX_STAGES + 4) + 16;
    static bool attr_done_dev[DISTB200_MAX_DEVICES] = {};
    bool& attr_done = attr_done_dev[current_device()];
    if (!attr_done) {
        const void* fns[4] = {(const void*)gemm_tcgen05_kernel<1, 8>, (const void*)gemm_tcgen05_kernel<2, 8>,
                              (const void*)gemm_tcgen05_kernel<1, 12>, (const void*)gemm_tcgen05_kernel<2, 12>};
        for (int i = 0; i < 4; ++i) {
            cudaError_t e = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            DISTB200_REQUIRE(e == cudaSuccess, "gemm(tcgen05): cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        }
        attr_done = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.blockDim = dim3((unsigned)num_threads(ew));
    cfg.dynamicSmemBytes = (size_t)smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    cfg.attrs = attr;
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.numAttrs = 1;
    cudaError_t e;
    if (args.ctas == 2) {
        long long clusters = args.total_tiles < sm_count() / 2 ? args.total_tiles : sm_count() / 2;
        cfg.gridDim = dim3((unsigned)(2 * clusters));
        attr[1].id = cudaLaunchAttributeClusterDimension;
        attr[1].val.clusterDim.x = 2;
        attr[1].val.clusterDim.y = 1;
        attr[1].val.clusterDim.z = 1;
        cfg.numAttrs = 2;
        e = ew == 8 ? cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<2, 8>, args) : cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<2, 12>, args);
    } else {
        long long grid = args.total_tiles < sm_count() ? args.total_tiles : sm_count();
        cfg.gridDim = dim3((unsigned)grid);
        e = ew == 8 ? cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<1, 8>, args) : cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<1, 12>, args);
    }
    DISTB200_REQUIRE(e == cudaSuccess, "gemm(tcgen05): launch failed: %s", cudaGetErrorString(e));
    return check_launch("gemm_tcgen05");
}

}  // namespace distb200
