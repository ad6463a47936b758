// distb200_gemm_wgrad on the 5th-generation tensor cores (sm_100a).
//
//   dw[tap][n][k] += sum_rows dy[row][n] * X_tap[row][k]
//
// is a GEMM whose reduction runs over the ROWS of two row-major activations, i.e. both MMA operands are "MN-major":
// D[M' = n][N' = k] = A[M'][r] B[N'][r]^T with A = dy^T, B = X^T.  TMA drops 64-row x 64-column (128-byte) slabs of dy
// and of the shifted / zero-filled X (the same 4-D tensor map as the forward A operand, so convolution taps and
// group boundaries come for free) into 128B-swizzled shared memory; tcgen05.mma reads them through MN-major
// descriptors (leading-dimension byte offset = slab stride), 16 rows of reduction per instruction.
//
// Work item = (tap, 128-wide tile of n, <=256-wide tile of k, split of the row blocks).  Persistent CTAs:
//   warp 0      TMA producer (ring of 4 stages: 2 dy slabs + 4 X slabs = 48 KB each)
//   warp 1      MMA issuer, fp32 accumulator in one of two TMEM stages (2 x 256 columns)
//   warps 2..5  epilogue: tcgen05.ld one accumulator row per thread, fp32 vector reductions (red.global.add.v4.f32)
//               into dw - the splits of a tile meet there, so the result is independent of the split count up to
//               fp32 summation order.
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "ptx.cuh"
#include "tma_host.cuh"

namespace distb200 {

namespace {

constexpr int W_ROWS = 64;                      // reduction rows per stage (4 MMAs of K = 16)
constexpr int W_TM = 128;                       // tile of n (MMA M)
constexpr int W_TN = 256;                       // tile of k (MMA N)
constexpr uint32_t SLAB_BYTES = W_ROWS * 128;   // 64 rows x 64 bf16
constexpr int W_A_SLABS = W_TM / 64, W_B_SLABS = W_TN / 64;
constexpr uint32_t W_STAGE_BYTES = (W_A_SLABS + W_B_SLABS) * SLAB_BYTES;
constexpr int W_STAGES = 4;
constexpr int W_THREADS = 64 + 128;

struct alignas(64) WgArgs {
    CUtensorMap tm_x;
    CUtensorMap tm_dy;
    distb200_wgrad_desc d;
    int m_tiles, n_tiles, units, splits;
    int box_rows;                 // rows a TMA box covers (64, or rows_h * img_w in image mode)
    int rows_h;                   // image rows per block (image mode)
    int blocks_per_group;
    int ksteps;                   // MMAs per block = ceil(box_rows / 16)
    long long total_blocks, blocks_per_split, items;
    int vec_ok;                   // dw rows are 16-byte aligned: vector reductions
};

// MN-major operand tile, 128-byte swizzle: 64 MN-elements (128 B) x 8 reduction rows per atom, atoms 1024 B apart along the
// reduction and `lbo` bytes apart along MN (cute::UMMA canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

struct Item {
    int tap, n0, k0, ncols;       // ncols = width of the k tile (multiple of 16)
    long long blk0, blk1;
};

__device__ __forceinline__ Item decode_item(const WgArgs& a, long long item) {
    Item it;
    const int unit = (int)(item % a.units);
    const long long split = item / a.units;
    const int nt = unit % a.n_tiles;
    const int rest = unit / a.n_tiles;
    const int mt = rest % a.m_tiles;
    it.tap = rest / a.m_tiles;
    it.n0 = mt * W_TM;
    it.k0 = nt * W_TN;
    int nc = a.d.k - it.k0;
    nc = nc > W_TN ? W_TN : nc;
    it.ncols = (nc + 15) & ~15;
    it.blk0 = split * a.blocks_per_split;
    it.blk1 = it.blk0 + a.blocks_per_split;
    if (it.blk1 > a.total_blocks) it.blk1 = a.total_blocks;
    return it;
}

__global__ void __launch_bounds__(W_THREADS, 1) wgrad_tcgen05_kernel(const __grid_constant__ WgArgs args) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[2 * W_STAGES + 4];
    __shared__ uint32_t tmem_slot;
    const distb200_wgrad_desc& d = args.d;
    const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bar0 = ptx::smem_u32(bars);
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (W_STAGES + s); };
    auto tfull_bar = [&](int s) { return bar0 + 8u * (2 * W_STAGES + s); };
    auto tempty_bar = [&](int s) { return bar0 + 8u * (2 * W_STAGES + 2 + s); };
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // rows of a stage that no TMA box covers (box_rows < 16 * ksteps) must read as zero: clear the ring once
    {
        uint4* z = reinterpret_cast<uint4*>(smem_raw + (smem_base - ptx::smem_u32(smem_raw)));
        for (uint32_t i = threadIdx.x; i < W_STAGES * W_STAGE_BYTES / 16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
        ptx::fence_proxy_async();
    }
    if (warp == 0 && lane == 0) {
        ptx::prefetch_tensormap(&args.tm_x);
        ptx::prefetch_tensormap(&args.tm_dy);
        for (int s = 0; s < W_STAGES; ++s) {
            ptx::mbar_init(full_bar(s), 1);
            ptx::mbar_init(empty_bar(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            ptx::mbar_init(tfull_bar(s), 1);
            ptx::mbar_init(tempty_bar(s), 4);
        }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(ptx::smem_u32(&tmem_slot), 512);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    grid_dep_sync();          // PDL: ring clearing, barrier init and TMEM allocation overlapped the previous kernel's tail

    const int a_slabs = (min(d.n, W_TM) + 63) / 64;      // dy slabs actually loaded (n < 64 needs one)
    if (warp == 0) {
        // ===================== TMA producer =====================
        int stage = 0;
        uint32_t phase = 0;
        for (long long item = blockIdx.x; item < args.items; item += gridDim.x) {
            const Item it = decode_item(args, item);
            const int b_slabs = (it.ncols + 63) / 64;
            const int a_here = min(a_slabs, (d.n - it.n0 + 63) / 64);
            const uint32_t tx = (uint32_t)(a_here + b_slabs) * (uint32_t)args.box_rows * 128u;
            long long gi = it.blk0 / args.blocks_per_group;
            int rb = (int)(it.blk0 - gi * args.blocks_per_group);
            for (long long blk = it.blk0; blk < it.blk1; ++blk) {
                ptx::mbar_wait(empty_bar(stage), phase ^ 1u);
                const uint32_t sa = smem_base + (uint32_t)stage * W_STAGE_BYTES;
                const uint32_t sb = sa + W_A_SLABS * SLAB_BYTES;
                int c1, c2, c3, r0;
                if (d.img_w > 0) {
                    r0 = rb * args.rows_h * d.img_w;
                    c1 = d.tap_off[it.tap][0];
                    c2 = rb * args.rows_h + d.tap_off[it.tap][1];
                    c3 = (int)gi + d.tap_off[it.tap][2];
                } else {
                    r0 = rb * W_ROWS;
                    c1 = r0 + d.tap_off[it.tap][0];
                    c2 = d.tap_off[it.tap][1] + (d.group_dim == 3 ? 0 : (int)gi);
                    c3 = d.tap_off[it.tap][2] + (d.group_dim == 3 ? (int)gi : 0);
                }
                if (ptx::elect_one()) {
                    ptx::mbar_arrive_expect_tx(full_bar(stage), tx);
                    for (int j = 0; j < a_here; ++j)
                        ptx::tma_load_3d(sa + j * SLAB_BYTES, &args.tm_dy, full_bar(stage), it.n0 + 64 * j, r0, (int)gi);
                    for (int j = 0; j < b_slabs; ++j)
                        ptx::tma_load_4d(sb + j * SLAB_BYTES, &args.tm_x, full_bar(stage), it.k0 + 64 * j, c1, c2, c3);
                }
                __syncwarp();
                if (++stage == W_STAGES) { stage = 0; phase ^= 1u; }
                if (++rb == args.blocks_per_group) { rb = 0; ++gi; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        int stage = 0, acc_stage = 0;
        uint32_t phase = 0, acc_phase = 0;
        for (long long item = blockIdx.x; item < args.items; item += gridDim.x) {
            const Item it = decode_item(args, item);
            // kind::f16, D = f32, A = B = bf16, both MN-major (bits 15 / 16), M = 128, N = ncols
            const uint32_t idesc = ptx::umma_idesc_bf16(W_TM, it.ncols) | (1u << 15) | (1u << 16);
            ptx::mbar_wait(tempty_bar(acc_stage), acc_phase ^ 1u);
            ptx::tc_fence_after();
            const uint32_t tmem_d = tmem_base + (uint32_t)acc_stage * 256u;
            bool first = true;
            for (long long blk = it.blk0; blk < it.blk1; ++blk) {
                ptx::mbar_wait(full_bar(stage), phase);
                ptx::tc_fence_after();
                const uint32_t sa = smem_base + (uint32_t)stage * W_STAGE_BYTES;
                const uint32_t sb = sa + W_A_SLABS * SLAB_BYTES;
                if (ptx::elect_one()) {
                    for (int ks = 0; ks < args.ksteps; ++ks) {
                        const uint64_t da = umma_desc_mn_sw128(sa + (uint32_t)ks * 2048u, SLAB_BYTES);
                        const uint64_t db = umma_desc_mn_sw128(sb + (uint32_t)ks * 2048u, SLAB_BYTES);
                        ptx::mma_f16_ss(tmem_d, da, db, idesc, (first && ks == 0) ? 0u : 1u);
                    }
                    ptx::mma_commit(empty_bar(stage));
                }
                __syncwarp();
                first = false;
                if (++stage == W_STAGES) { stage = 0; phase ^= 1u; }
            }
            if (ptx::elect_one()) ptx::mma_commit(tfull_bar(acc_stage));
            __syncwarp();
            if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1u; }
        }
    } else {
        // ===================== epilogue: reduce the tile into dw =====================
        const int quad = warp & 3;               // warps 2..5 -> TMEM lane quadrants 2, 3, 0, 1
        int acc_stage = 0;
        uint32_t acc_phase = 0;
        for (long long item = blockIdx.x; item < args.items; item += gridDim.x) {
            const Item it = decode_item(args, item);
            ptx::mbar_wait(tfull_bar(acc_stage), acc_phase);
            ptx::tc_fence_after();
            const int n = it.n0 + quad * 32 + lane;
            const bool row_ok = n < d.n && it.blk1 > it.blk0;
            float* row = d.dw + (long long)it.tap * d.dw_tap_stride + (long long)n * d.ld_dw + it.k0;
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)acc_stage * 256u;
            const int kmax = d.k - it.k0;        // valid columns of this tile
            for (int c0 = 0; c0 < it.ncols; c0 += 32) {
                uint32_t acc[32];
                const bool two = c0 + 16 < it.ncols;
                ptx::tmem_ld16(taddr + (uint32_t)c0, acc);
                if (two) ptx::tmem_ld16(taddr + (uint32_t)c0 + 16u, acc + 16);
                ptx::tmem_ld_wait();
                if (!row_ok) continue;
                const int lim = two ? 32 : 16;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (j >= lim) break;
                    const int c = c0 + j;
                    if (args.vec_ok && c + 4 <= kmax) {
                        red_add_v4(row + c, __uint_as_float(acc[j]), __uint_as_float(acc[j + 1]), __uint_as_float(acc[j + 2]),
                                   __uint_as_float(acc[j + 3]));
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            if (c + e < kmax) atomicAdd(row + c + e, __uint_as_float(acc[j + e]));
                    }
                }
            }
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(tempty_bar(acc_stage));
            if (++acc_stage == 2) { acc_stage = 0; acc_phase ^= 1u; }
        }
    }
    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace

int wgrad_tcgen05_launch(const distb200_wgrad_desc& d, cudaStream_t stream) {
    const long long total_rows = d.groups * d.rows_per_group;
    if (total_rows == 0) return 0;
    // shapes the TMA path cannot address fall back to the FFMA kernel (same definition)
    const bool aligned = (reinterpret_cast<uintptr_t>(d.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d.dy) & 15) == 0 && d.ld_dy % 8 == 0 &&
                         d.a_stride[0] == 1 && d.a_stride[1] % 8 == 0 && (d.a_dim[2] == 1 || d.a_stride[2] % 8 == 0) &&
                         (d.a_dim[3] == 1 || d.a_stride[3] % 8 == 0) && (d.groups == 1 || (d.dy_gstride * d.ld_dy) % 8 == 0) &&
                         (d.dy_roff * d.ld_dy) % 8 == 0 && (d.img_w == 0 || (d.img_w <= W_ROWS && d.rows_per_group % d.img_w == 0));
    if (!aligned) return wgrad_simt_launch(d, stream);

    WgArgs args;
    args.d = d;
    args.d.group_dim = d.group_dim == 3 ? 3 : 2;
    args.m_tiles = (d.n + W_TM - 1) / W_TM;
    args.n_tiles = (d.k + W_TN - 1) / W_TN;
    args.units = d.num_taps * args.m_tiles * args.n_tiles;
    if (d.img_w > 0) {
        const int img_h = (int)(d.rows_per_group / d.img_w);
        args.rows_h = W_ROWS / d.img_w;
        if (args.rows_h > img_h) args.rows_h = img_h;
        args.box_rows = args.rows_h * d.img_w;
        args.blocks_per_group = (img_h + args.rows_h - 1) / args.rows_h;
    } else {
        args.rows_h = 0;
        args.box_rows = d.rows_per_group < W_ROWS ? (int)d.rows_per_group : W_ROWS;
        args.blocks_per_group = (int)((d.rows_per_group + W_ROWS - 1) / W_ROWS);
    }
    args.ksteps = (args.box_rows + 15) / 16;
    args.total_blocks = d.groups * args.blocks_per_group;
    // enough items to fill the GPU about twice, at least 4 row blocks per item
    long long splits = (2LL * sm_count() + args.units - 1) / args.units;
    const long long max_splits = (args.total_blocks + 3) / 4;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    args.blocks_per_split = (args.total_blocks + splits - 1) / splits;
    args.splits = (int)((args.total_blocks + args.blocks_per_split - 1) / args.blocks_per_split);
    args.items = (long long)args.units * args.splits;
    args.vec_ok = (reinterpret_cast<uintptr_t>(d.dw) & 15) == 0 && d.ld_dw % 4 == 0 && d.dw_tap_stride % 4 == 0;

    {   // X: (k, c1, c2, c3) exactly as the forward A operand
        long long dims[4] = {d.a_dim[0], d.a_dim[1], d.a_dim[2], d.a_dim[3]};
        long long str[4] = {1, d.a_stride[1], d.a_stride[2], d.a_stride[3]};
        int box[4] = {64, args.box_rows, 1, 1};
        if (d.img_w > 0) {
            box[1] = d.img_w;
            box[2] = args.rows_h;
        }
        if (make_map(&args.tm_x, d.x, 4, dims, str, box, "wgrad X")) return 1;
    }
    {   // dy: (n, row in group, group)
        long long dims[3] = {d.n, d.rows_per_group, d.groups};
        long long str[3] = {1, d.ld_dy, d.dy_gstride * d.ld_dy};
        int box[3] = {64, args.box_rows, 1};
        const bf16* base = reinterpret_cast<const bf16*>(d.dy) + d.dy_roff * d.ld_dy;
        if (make_map(&args.tm_dy, base, 3, dims, str, box, "wgrad dY")) return 1;
    }
    const int smem = W_STAGES * (int)W_STAGE_BYTES + 1024;
    static bool attr_done_dev[DISTB200_MAX_DEVICES] = {};
    bool& attr_done = attr_done_dev[current_device()];
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        DISTB200_REQUIRE(e == cudaSuccess, "gemm_wgrad(tcgen05): cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        attr_done = true;
    }
    const long long grid = args.items < sm_count() ? args.items : sm_count();
    DISTB200_LAUNCH(wgrad_tcgen05_kernel, (unsigned)grid, W_THREADS, smem, stream, args);
    return check_launch("wgrad_tcgen05");
}

}  // namespace distb200
