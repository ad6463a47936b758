// SIMT (FFMA) implementation of distb200_gemm: the fp32 parity path (fp32 operands, fp32 accumulation)
// and a cross-check for the tcgen05 kernel (bf16 operands).  Same operand addressing and epilogue
// definition as include/distb200.h; 64x64x16 shared-memory tiles, 4x4 outputs per thread.
#include "common.cuh"

namespace distb200 {

namespace {

constexpr int TM = 64, TN = 64, TK = 16, NT = 256;

struct SimtArgs {
    distb200_gemm_desc d;
    long long total_rows;
};

template <typename T>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(const SimtArgs args) {
    grid_dep_sync();
    const distb200_gemm_desc& d = args.d;
    __shared__ float As[TK][TM + 4];
    __shared__ float Bs[TK][TN + 4];

    const int tid = threadIdx.x;
    const long long row0 = (long long)blockIdx.x * TM;
    const int n0 = blockIdx.y * TN;
    const T* __restrict__ A = reinterpret_cast<const T*>(d.a);
    const T* __restrict__ B = reinterpret_cast<const T*>(d.b);

    // loader role: one A row / one B row, 4 consecutive k
    const int lrow = tid >> 2, lk = (tid & 3) * 4;
    const long long grow = row0 + lrow;
    const bool row_ok = grow < args.total_rows;
    const long long gi = row_ok ? grow / d.rows_per_group : 0;
    const long long r = row_ok ? grow % d.rows_per_group : 0;
    const int ncol = n0 + lrow;
    const bool n_ok = ncol < d.n;

    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int K = d.k;
    for (int tap = 0; tap < d.num_taps; ++tap) {
        long long c1, c2, c3;
        if (d.img_w > 0) {
            c1 = r % d.img_w + d.tap_off[tap][0];
            c2 = r / d.img_w + d.tap_off[tap][1];
            c3 = gi + d.tap_off[tap][2];
        } else {
            c1 = r + d.tap_off[tap][0];
            c2 = d.tap_off[tap][1] + (d.group_dim == 3 ? 0 : gi);
            c3 = d.tap_off[tap][2] + (d.group_dim == 3 ? gi : 0);
        }
        const bool a_ok = row_ok && c1 >= 0 && c1 < d.a_dim[1] && c2 >= 0 && c2 < d.a_dim[2] && c3 >= 0 && c3 < d.a_dim[3];
        const T* arow = A + (a_ok ? c1 * d.a_stride[1] + c2 * d.a_stride[2] + c3 * d.a_stride[3] : 0);
        const T* brow = B + (long long)tap * d.b_tap_stride + (long long)(n_ok ? ncol : 0) * d.ldb;
        const int ka = (int)(d.a_dim[0] < K ? d.a_dim[0] : K);
        for (int k0 = 0; k0 < K; k0 += TK) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int k = k0 + lk + e;
                As[lk + e][lrow] = (a_ok && k < ka) ? to_float(arow[k]) : 0.f;
                Bs[lk + e][lrow] = (n_ok && k < K) ? to_float(brow[k]) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < TK; ++k) {
                float a4[4], b4[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a4[i] = As[k][ty * 4 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) b4[j] = Bs[k][tx * 4 + j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
            }
            __syncthreads();
        }
    }

    // epilogue
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long long orow = row0 + ty * 4 + i;
        if (orow >= args.total_rows) continue;
        const long long ogi = orow / d.rows_per_group, orr = orow % d.rows_per_group;
        for (int rep = 0; rep < d.out_rep; ++rep) {
            const long long dst = ogi * d.out_gstride + d.out_roff + orr + (long long)rep * d.out_rep_stride;
            const long long rsrc = ogi * d.res_gstride + d.res_roff + orr + (long long)rep * d.res_rep_stride;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + tx * 4 + j;
                if (n >= d.n) continue;
                float v = acc[i][j];
                if (d.ln_stats) v = d.ln_stats[2 * orow + 1] * (v - d.ln_stats[2 * orow] * d.ln_wsum[n]);      // folded LayerNorm
                if (d.bias) v += d.bias[n];
                if (d.res) v += d.res[rsrc * d.ld_res + n];
                if (d.act == DISTB200_ACT_QUICKGELU && n >= d.act_from && (d.act_to == 0 || n < d.act_to)) v = quick_gelu(v);
                if (d.out) {
                    if (d.out_dtype == DISTB200_F32) reinterpret_cast<float*>(d.out)[dst * d.ld_out + n] = v;
                    else reinterpret_cast<bf16*>(d.out)[dst * d.ld_out + n] = __float2bfloat16_rn(v);
                }
                if (d.out2) {
                    long long dst2 = dst;
                    int n2 = n;
                    if (d.out2_gdiv > 0) {
                        dst2 = (ogi / d.out2_gdiv) * d.out2_gstride + d.out2_roff + orr;
                        n2 = n + (int)(ogi % d.out2_gdiv) * d.out2_cstep;
                    }
                    if (d.out2_dtype == DISTB200_F32) reinterpret_cast<float*>(d.out2)[dst2 * d.ld_out2 + n2] = v;
                    else reinterpret_cast<bf16*>(d.out2)[dst2 * d.ld_out2 + n2] = __float2bfloat16_rn(v);
                }
            }
        }
    }
}


// ---- weight gradient (distb200_gemm_wgrad), FFMA version: the fp32 parity path and the cross-check of the tensor-core
// kernel.  Block = 16 (n) x 16 (k) outputs of one tap; the rows are split over blockIdx.z and reduced with fp32 atomics.
constexpr int WT = 16, WR = 32;

struct WgradSimtArgs {
    distb200_wgrad_desc d;
    long long total_rows;
    long long rows_per_split;
    int splits;
    int n_tiles;
};

template <typename T>
__global__ void __launch_bounds__(WT * WT) wgrad_simt_kernel(const WgradSimtArgs args) {
    grid_dep_sync();
    const distb200_wgrad_desc& d = args.d;
    __shared__ float Ys[WR][WT + 1];
    __shared__ float Xs[WR][WT + 1];
    const int tid = threadIdx.x, tn = tid / WT, tk = tid % WT;
    const int n0 = (blockIdx.x % args.n_tiles) * WT, tap = blockIdx.x / args.n_tiles;
    const int k0 = blockIdx.y * WT;
    const int split = blockIdx.z;
    const T* __restrict__ X = reinterpret_cast<const T*>(d.x);
    const T* __restrict__ DY = reinterpret_cast<const T*>(d.dy);
    const long long r_begin = (long long)split * args.rows_per_split;
    long long r_end = r_begin + args.rows_per_split;
    if (r_end > args.total_rows) r_end = args.total_rows;
    const int ka = (int)(d.a_dim[0] < d.k ? d.a_dim[0] : d.k);
    float acc = 0.f;
    for (long long base = r_begin; base < r_end; base += WR) {
        // loader: thread (lr, lc) fetches rows lr and lr + 16 of the chunk, column lc of each operand tile
        for (int h = 0; h < WR / WT; ++h) {
            const int lr = tn + h * WT, lc = tk;
            const long long row = base + lr;
            float yv = 0.f, xv = 0.f;
            if (row < r_end) {
                const long long gi = row / d.rows_per_group, r = row - gi * d.rows_per_group;
                if (n0 + lc < d.n) yv = to_float(DY[(gi * d.dy_gstride + d.dy_roff + r) * d.ld_dy + n0 + lc]);
                long long c1, c2, c3;
                if (d.img_w > 0) {
                    c1 = r % d.img_w + d.tap_off[tap][0];
                    c2 = r / d.img_w + d.tap_off[tap][1];
                    c3 = gi + d.tap_off[tap][2];
                } else {
                    c1 = r + d.tap_off[tap][0];
                    c2 = d.tap_off[tap][1] + (d.group_dim == 3 ? 0 : gi);
                    c3 = d.tap_off[tap][2] + (d.group_dim == 3 ? gi : 0);
                }
                const bool ok = c1 >= 0 && c1 < d.a_dim[1] && c2 >= 0 && c2 < d.a_dim[2] && c3 >= 0 && c3 < d.a_dim[3] && k0 + lc < ka;
                if (ok) xv = to_float(X[c1 * d.a_stride[1] + c2 * d.a_stride[2] + c3 * d.a_stride[3] + k0 + lc]);
            }
            Ys[lr][lc] = yv;
            Xs[lr][lc] = xv;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < WR; ++r) acc = fmaf(Ys[r][tn], Xs[r][tk], acc);
        __syncthreads();
    }
    if (n0 + tn < d.n && k0 + tk < d.k)
        atomicAdd(d.dw + (long long)tap * d.dw_tap_stride + (long long)(n0 + tn) * d.ld_dw + k0 + tk, acc);
}

}  // namespace

int gemm_simt_launch(const distb200_gemm_desc& d, cudaStream_t stream) {
    SimtArgs args;
    args.d = d;
    args.total_rows = (long long)d.groups * d.rows_per_group;
    if (args.total_rows == 0 || d.n == 0) return 0;
    const long long mt = (args.total_rows + TM - 1) / TM;
    DISTB200_REQUIRE(mt < 2147483647LL, "gemm(simt): too many row tiles (%lld)", mt);
    dim3 grid((unsigned)mt, (unsigned)((d.n + TN - 1) / TN));
    if (d.dtype == DISTB200_F32) DISTB200_LAUNCH(gemm_simt_kernel<float>, grid, NT, 0, stream, args);
    else DISTB200_LAUNCH(gemm_simt_kernel<bf16>, grid, NT, 0, stream, args);
    return check_launch("gemm_simt");
}

int wgrad_simt_launch(const distb200_wgrad_desc& d, cudaStream_t stream) {
    WgradSimtArgs args;
    args.d = d;
    args.total_rows = (long long)d.groups * d.rows_per_group;
    if (args.total_rows == 0 || d.n == 0 || d.k == 0) return 0;
    args.n_tiles = (d.n + WT - 1) / WT;
    const int k_tiles = (d.k + WT - 1) / WT;
    const long long blocks_xy = (long long)args.n_tiles * d.num_taps * k_tiles;
    long long splits = ((long long)sm_count() * 8 + blocks_xy - 1) / blocks_xy;
    const long long max_splits = (args.total_rows + 4 * WR - 1) / (4 * WR);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    args.rows_per_split = ((args.total_rows + splits - 1) / splits + WR - 1) / WR * WR;
    args.splits = (int)((args.total_rows + args.rows_per_split - 1) / args.rows_per_split);
    dim3 grid((unsigned)(args.n_tiles * d.num_taps), (unsigned)k_tiles, (unsigned)args.splits);
    if (d.dtype == DISTB200_F32) DISTB200_LAUNCH(wgrad_simt_kernel<float>, grid, WT * WT, 0, stream, args);
    else DISTB200_LAUNCH(wgrad_simt_kernel<bf16>, grid, WT * WT, 0, stream, args);
    return check_launch("wgrad_simt");
}

}  // namespace distb200
