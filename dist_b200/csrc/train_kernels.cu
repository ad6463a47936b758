// Backward-pass and optimizer kernels of the fine-tuning step other than the GEMMs:
// QuickGELU forward / backward on saved pre-activations, casts, frame-group sums, column sums (bias-like gradients),
// LayerNorm backward, single-query cross-attention backward, the soft-target cross-entropy head, AdamW and the
// operand copies of the master weights.  All HBM-bound: vectorised accesses, warp-shuffle reductions, fp32 math.
#include "common.cuh"

namespace distb200 {

namespace {

inline unsigned grid_cap(long long work, int block, int per_sm = 16) {
    long long g = (work + block - 1) / block;
    const long long cap = (long long)sm_count() * per_sm;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

__device__ __forceinline__ float ld_any(const void* p, int dtype, long long i) {
    return dtype == DISTB200_F32 ? reinterpret_cast<const float*>(p)[i] : __bfloat162float(reinterpret_cast<const bf16*>(p)[i]);
}
__device__ __forceinline__ void st_any(void* p, int dtype, long long i, float v) {
    if (dtype == DISTB200_F32) reinterpret_cast<float*>(p)[i] = v;
    else reinterpret_cast<bf16*>(p)[i] = __float2bfloat16_rn(v);
}

// four consecutive elements (i % 4 == 0, 16 / 8-byte aligned)
__device__ __forceinline__ float4 ld4_any(const void* p, int dtype, long long i) {
    if (dtype == DISTB200_F32) return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i);
    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(p) + i);
    const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
    const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st4_any(void* p, int dtype, long long i, float4 v) {
    if (dtype == DISTB200_F32) *reinterpret_cast<float4*>(reinterpret_cast<float*>(p) + i) = v;
    else *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(p) + i) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

__device__ __forceinline__ float qgelu_grad(float z) {
    const float s = 1.0f / (1.0f + __expf(-1.702f * z));
    return s * (1.0f + 1.702f * z * (1.0f - s));
}

// ---------------------------------------------------------------------------------------------------
// elementwise (n % 4 == 0 fast path, scalar tail)
// ---------------------------------------------------------------------------------------------------
enum { EW_GELU = 0, EW_GELU_BWD = 1, EW_CAST = 2 };

template <int OP>
__global__ void __launch_bounds__(256) elementwise_kernel(const void* a, int a_dtype, const void* b, int b_dtype, float* o32, void* olp,
                                                          int lp_dtype, long long n) {
    grid_dep_sync();
    const long long n4 = n >> 2;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 x = ld4_any(a, a_dtype, 4 * i);
        float4 r;
        if (OP == EW_GELU) {
            r = make_float4(quick_gelu(x.x), quick_gelu(x.y), quick_gelu(x.z), quick_gelu(x.w));
        } else if (OP == EW_GELU_BWD) {
            const float4 z = ld4_any(b, b_dtype, 4 * i);
            r = make_float4(x.x * qgelu_grad(z.x), x.y * qgelu_grad(z.y), x.z * qgelu_grad(z.z), x.w * qgelu_grad(z.w));
        } else {
            r = x;
        }
        if (o32) *reinterpret_cast<float4*>(o32 + 4 * i) = r;
        if (olp) st4_any(olp, lp_dtype, 4 * i, r);
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const long long i = (n4 << 2) + threadIdx.x;
        const float x = ld_any(a, a_dtype, i);
        float r = x;
        if (OP == EW_GELU) r = quick_gelu(x);
        if (OP == EW_GELU_BWD) r = x * qgelu_grad(ld_any(b, b_dtype, i));
        if (o32) o32[i] = r;
        if (olp) st_any(olp, lp_dtype, i, r);
    }
}

// bf16 operands, 16-byte vectors, U independent vectors in flight per thread (the generic kernel above moves 8 bytes per thread and
// iteration behind run-time dtype tests: 2.7-3.7 TB/s on the activation passes of the fine-tuning step).
// sigmoid(t) = 0.5 + 0.5 tanh(t / 2) with ONE MUFU (tanh.approx, relative error ~2^-11: below bf16 resolution); used when only a bf16
// result is written - two MUFU per element (ex2 + rcp) made the activation passes MUFU-bound, not HBM-bound.
__device__ __forceinline__ float sigmoid_fast(float t) {
    float th;
    asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * t));
    return fmaf(0.5f, th, 0.5f);
}

template <int OP, int U, bool FAST>
__global__ void __launch_bounds__(256) elementwise_bf16_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b, float* o32, uint4* olp, long long n8) {
    grid_dep_sync();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x; i0 < n8; i0 += U * stride) {
        uint4 x[U], z[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i < n8) {
                x[u] = __ldcs(a + i);
                if (OP == EW_GELU_BWD) z[u] = __ldcs(b + i);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i >= n8) continue;
            const uint32_t xw[4] = {x[u].x, x[u].y, x[u].z, x[u].w}, zw[4] = {z[u].x, z[u].y, z[u].z, z[u].w};
            float r[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float lo = __uint_as_float(xw[e] << 16), hi = __uint_as_float(xw[e] & 0xffff0000u);
                if (OP == EW_GELU) {
                    r[2 * e] = FAST ? lo * sigmoid_fast(1.702f * lo) : quick_gelu(lo);
                    r[2 * e + 1] = FAST ? hi * sigmoid_fast(1.702f * hi) : quick_gelu(hi);
                } else {
                    const float z0 = __uint_as_float(zw[e] << 16), z1 = __uint_as_float(zw[e] & 0xffff0000u);
                    if (FAST) {
                        const float s0 = sigmoid_fast(1.702f * z0), s1 = sigmoid_fast(1.702f * z1);
                        r[2 * e] = lo * (s0 * fmaf(1.702f * z0, 1.0f - s0, 1.0f));
                        r[2 * e + 1] = hi * (s1 * fmaf(1.702f * z1, 1.0f - s1, 1.0f));
                    } else {
                        r[2 * e] = lo * qgelu_grad(z0);
                        r[2 * e + 1] = hi * qgelu_grad(z1);
                    }
                }
            }
            if (o32) {
                reinterpret_cast<float4*>(o32)[2 * i] = make_float4(r[0], r[1], r[2], r[3]);
                reinterpret_cast<float4*>(o32)[2 * i + 1] = make_float4(r[4], r[5], r[6], r[7]);
            }
            if (olp) olp[i] = make_uint4(pack_bf16x2(r[0], r[1]), pack_bf16x2(r[2], r[3]), pack_bf16x2(r[4], r[5]), pack_bf16x2(r[6], r[7]));
        }
    }
}

__global__ void __launch_bounds__(256) group_sum_kernel(const void* src, int src_dtype, long long groups, int alpha, long long inner, void* dst,
                                                        int dst_dtype) {
    grid_dep_sync();
    const long long inner4 = inner >> 2, total = groups * inner4;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long g = idx / inner4, i = (idx - g * inner4) * 4;
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < alpha; ++k) {
            const float4 v = ld4_any(src, src_dtype, (g * alpha + k) * inner + i);
            s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
        }
        st4_any(dst, dst_dtype, g * inner + i, s);
    }
}

// ---------------------------------------------------------------------------------------------------
// column sums: block = 32 columns x 8 row lanes; rows of all groups are split over blockIdx.y; one atomicAdd per
// (block, column, period slot).  period > 1 (per-frame class tokens, positional embeddings) uses rows_per_group == 1.
// ---------------------------------------------------------------------------------------------------
// bf16, contiguous rows, period 1 (every bias gradient of the fine-tuning step): thread = 8 consecutive columns (one 16-byte load), cols / 8
// threads per row and 256 / (cols / 8) rows per block pass, U independent loads in flight per thread.  The 4-column version below ran the
// 96-wide gradients at 0.8-1.1 TB/s (24 of 32 lanes, 8-byte loads).
template <int U, bool DENSE>
__global__ void __launch_bounds__(256) colsum_bf16_dense_kernel(const uint4* __restrict__ src, long long ld8, unsigned total, long long roff, unsigned rpg,
                                                                long long gstride, int tpr, int rpb, float* out, float* out2, unsigned rows_per_block) {
    grid_dep_sync();
    __shared__ float red[256][9];
    const int tr = threadIdx.x / tpr, tc = threadIdx.x - tr * tpr;
    const unsigned r_begin = blockIdx.x * rows_per_block;
    unsigned r_end = r_begin + rows_per_block;
    if (r_end > total) r_end = total;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    // DENSE: rows are contiguous from roff; otherwise row i of the walk is row (i / rpg) * gstride + roff + i % rpg of the matrix
    auto addr = [&](unsigned i) -> const uint4* {
        const long long srow = DENSE ? (long long)i + roff : (long long)(i / rpg) * gstride + roff + (long long)(i % rpg);
        return src + srow * ld8 + tc;
    };
    auto add8 = [&](const uint4& v) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            acc[2 * e] += __uint_as_float(w[e] << 16);
            acc[2 * e + 1] += __uint_as_float(w[e] & 0xffff0000u);
        }
    };
    if (tr < rpb) {
        unsigned i = r_begin + tr;
        for (; i + (unsigned)((U - 1) * rpb) < r_end; i += (unsigned)(U * rpb)) {
            uint4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = __ldg(addr(i + (unsigned)(u * rpb)));
#pragma unroll
            for (int u = 0; u < U; ++u) add8(v[u]);
        }
        for (; i < r_end; i += (unsigned)rpb) add8(__ldg(addr(i)));
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) red[threadIdx.x][e] = acc[e];
    __syncthreads();
    // thread (e, column vector): sums the row lanes of one column
    for (int o = threadIdx.x; o < tpr * 8; o += 256) {
        const int cv = o >> 3, e = o & 7;
        float t = 0.f;
        for (int j = 0; j < rpb; ++j) t += red[j * tpr + cv][e];
        atomicAdd(out + o, t);
        if (out2) atomicAdd(out2 + o, t);
    }
}

__global__ void __launch_bounds__(256) colsum_kernel(const void* src, int dtype, long long ld, long long groups, long long rows_per_group,
                                                     long long gstride, long long roff, long long period, int cols, float* out, float* out2,
                                                     long long rows_per_block) {
    grid_dep_sync();
    // thread = 4 consecutive columns (16-byte / 8-byte loads), 32 column vectors x 8 row lanes per block
    __shared__ float4 red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = (blockIdx.x * 32 + tx) * 4;
    const bool col_ok = c < cols;
    if (period == 1) {
        const unsigned total = (unsigned)(groups * rows_per_group), rpg = (unsigned)rows_per_group;
        const unsigned r_begin = (unsigned)(blockIdx.y * rows_per_block);
        unsigned r_end = r_begin + (unsigned)rows_per_block;
        if (r_end > total) r_end = total;
        const bool dense = gstride == rows_per_group;          // rows of all groups are contiguous (plus roff)
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col_ok) {
            unsigned i = r_begin + ty;
            for (; i + 24 < r_end; i += 32) {                   // four independent loads in flight
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const unsigned row = i + 8 * u;
                    const long long srow = dense ? (long long)row + roff : (long long)(row / rpg) * gstride + roff + row % rpg;
                    v[u] = ld4_any(src, dtype, srow * ld + c);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
            }
            for (; i < r_end; i += 8) {
                const long long srow = dense ? (long long)i + roff : (long long)(i / rpg) * gstride + roff + i % rpg;
                const float4 v = ld4_any(src, dtype, srow * ld + c);
                s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
            }
        }
        red[ty][tx] = s;
        __syncthreads();
        if (ty == 0 && col_ok) {
            float4 t = red[0][tx];
#pragma unroll
            for (int j = 1; j < 8; ++j) { t.x += red[j][tx].x; t.y += red[j][tx].y; t.z += red[j][tx].z; t.w += red[j][tx].w; }
            atomicAdd(out + c, t.x); atomicAdd(out + c + 1, t.y); atomicAdd(out + c + 2, t.z); atomicAdd(out + c + 3, t.w);
            if (out2) { atomicAdd(out2 + c, t.x); atomicAdd(out2 + c + 1, t.y); atomicAdd(out2 + c + 2, t.z); atomicAdd(out2 + c + 3, t.w); }
        }
    } else {
        // slot p collects the groups g with g % period == p (rows_per_group rows each); blockIdx.y strides over p
        for (long long p = blockIdx.y; p < period; p += gridDim.y) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
            if (col_ok)
                for (long long g = p + (long long)ty * period; g < groups; g += 8 * period)
                    for (long long r = 0; r < rows_per_group; ++r) {
                        const float4 v = ld4_any(src, dtype, (g * gstride + roff + r) * ld + c);
                        s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
                    }
            red[ty][tx] = s;
            __syncthreads();
            if (ty == 0 && col_ok) {
                float4 t = red[0][tx];
#pragma unroll
                for (int j = 1; j < 8; ++j) { t.x += red[j][tx].x; t.y += red[j][tx].y; t.z += red[j][tx].z; t.w += red[j][tx].w; }
                float* o = out + p * cols + c;
                atomicAdd(o, t.x); atomicAdd(o + 1, t.y); atomicAdd(o + 2, t.z); atomicAdd(o + 3, t.w);
            }
            __syncthreads();
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// LayerNorm backward: one warp per row (row in registers, cols <= 1024), statistics recomputed; the parameter
// gradients are reduced per block in shared memory (fp32 atomics within the block, then one global atomic per
// column and block).
// ---------------------------------------------------------------------------------------------------
constexpr int LNB_MAXV = 8;      // widest row: 128 * LNB_MAXV columns
constexpr int LNB_WARPS = 4;

// gradient rows arrive in bf16 (tensor-core path) or fp32 (parity path): keep them packed in registers until they are used
template <bool BF> struct DyRaw;
template <> struct DyRaw<true> {
    typedef uint2 T;
    static __device__ __forceinline__ T load(const void* p, long long i) { return *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(p) + i); }
    static __device__ __forceinline__ float4 get(const T& u) {
        return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
    }
};
template <> struct DyRaw<false> {
    typedef float4 T;
    static __device__ __forceinline__ T load(const void* p, long long i) { return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p) + i); }
    static __device__ __forceinline__ float4 get(const T& u) { return u; }
};

// Every operand of a row (x, the periodic addend, both gradient rows, the skip gradient, the old dx when accumulating) is requested
// at the top of the row's iteration: round 1 issued them in three dependent phases (x -> statistics -> gradients -> skip), which
// left ~2 KB per warp in flight and 1.5 TB/s (23 % of HBM) on the IntegrationNetwork's double LayerNorm.
// LPR lanes share a row (32, or 8 for rows of <= 128 columns: four rows per warp pass, three shuffle steps per reduction, every lane busy).
template <int LPR>
__device__ __forceinline__ float lanes_sum(float v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <int LNB_V, bool DYBF, int LPR>
__global__ void __launch_bounds__(LNB_WARPS * 32) layernorm_bwd_kernel(
    const float* __restrict__ in1, long long ld_in1, const float* __restrict__ in2, long long ld_in2, long long in2_period, long long rows, int cols,
    float eps, const float* __restrict__ g1, const void* dy1, long long ld_dy1, const float* __restrict__ g2, const void* dy2, long long ld_dy2,
    const float* add, long long ld_add, float* dx, long long ld_dx, int accumulate, void* dx_lp, long long ld_dx_lp, int lp_dtype,
    float* dg1, float* db1, float* dg2, float* db2, int rows_per_warp) {
    grid_dep_sync();
    typedef DyRaw<DYBF> Dy;
    extern __shared__ float sh[];            // [4][cols]: dg1, db1, dg2, db2 partials of this block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int RPW = 32 / LPR;                 // rows per warp pass
    const int lc = lane % LPR, sub = lane / LPR;
    for (int i = threadIdx.x; i < 4 * cols; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
    float a_dg1[LNB_V][4], a_db1[LNB_V][4], a_dg2[LNB_V][4], a_db2[LNB_V][4];
#pragma unroll
    for (int i = 0; i < LNB_V; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) a_dg1[i][j] = a_db1[i][j] = a_dg2[i][j] = a_db2[i][j] = 0.f;
    const bool rmw = dx && accumulate;

    const long long row_begin = ((long long)blockIdx.x * LNB_WARPS + warp) * rows_per_warp;
    for (int rr = 0; rr < rows_per_warp; rr += RPW) {
        if (row_begin + rr >= rows) break;
        const long long row = row_begin + rr + sub;
        const bool live = rr + sub < rows_per_warp && row < rows;      // lanes of a missing row still take part in the shuffles
        const float* x = in1 + row * ld_in1;
        const float* x2 = in2 ? in2 + ((live ? row : 0) % in2_period) * ld_in2 : nullptr;
        float4 v[LNB_V], ad[LNB_V], old[LNB_V];
        typename Dy::T r1[LNB_V], r2[LNB_V];
        // ---- all loads of the row
#pragma unroll
        for (int i = 0; i < LNB_V; ++i) {
            const int c = (i * LPR + lc) * 4;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c < cols && live) {
                v[i] = *reinterpret_cast<const float4*>(x + c);
                r1[i] = Dy::load(dy1, row * ld_dy1 + c);
                if (dy2) r2[i] = Dy::load(dy2, row * ld_dy2 + c);
                if (add) ad[i] = *reinterpret_cast<const float4*>(add + row * ld_add + c);
                if (rmw) old[i] = *reinterpret_cast<const float4*>(dx + row * ld_dx + c);
            }
        }
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < LNB_V; ++i) {
            const int c = (i * LPR + lc) * 4;
            if (c < cols && live) {
                if (x2) {
                    const float4 w = *reinterpret_cast<const float4*>(x2 + c);
                    v[i].x += w.x; v[i].y += w.y; v[i].z += w.z; v[i].w += w.w;
                }
                sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            }
        }
        const float mean = lanes_sum<LPR>(sum) / (float)cols;
        float sq = 0.f;
#pragma unroll
        for (int i = 0; i < LNB_V; ++i) {
            const int c = (i * LPR + lc) * 4;
            if (c < cols && live) {
                v[i].x -= mean; v[i].y -= mean; v[i].z -= mean; v[i].w -= mean;
                sq += (v[i].x * v[i].x + v[i].y * v[i].y) + (v[i].z * v[i].z + v[i].w * v[i].w);
            }
        }
        const float rstd = rsqrtf(lanes_sum<LPR>(sq) / (float)cols + eps);
        // g = dy1*g1 + dy2*g2; accumulate parameter gradients; row means of g and g*xhat
        float4 g[LNB_V];
        float sg = 0.f, sgx = 0.f;
#pragma unroll
        for (int i = 0; i < LNB_V; ++i) {
            const int c = (i * LPR + lc) * 4;
            if (c < cols && live) {
                v[i].x *= rstd; v[i].y *= rstd; v[i].z *= rstd; v[i].w *= rstd;          // xhat
                const float4 d1 = Dy::get(r1[i]);
                const float4 ga = *reinterpret_cast<const float4*>(g1 + c);
                g[i] = make_float4(d1.x * ga.x, d1.y * ga.y, d1.z * ga.z, d1.w * ga.w);
                a_dg1[i][0] += d1.x * v[i].x; a_dg1[i][1] += d1.y * v[i].y; a_dg1[i][2] += d1.z * v[i].z; a_dg1[i][3] += d1.w * v[i].w;
                a_db1[i][0] += d1.x; a_db1[i][1] += d1.y; a_db1[i][2] += d1.z; a_db1[i][3] += d1.w;
                if (dy2) {
                    const float4 d2 = Dy::get(r2[i]);
                    const float4 gb = *reinterpret_cast<const float4*>(g2 + c);
                    g[i].x += d2.x * gb.x; g[i].y += d2.y * gb.y; g[i].z += d2.z * gb.z; g[i].w += d2.w * gb.w;
                    a_dg2[i][0] += d2.x * v[i].x; a_dg2[i][1] += d2.y * v[i].y; a_dg2[i][2] += d2.z * v[i].z; a_dg2[i][3] += d2.w * v[i].w;
                    a_db2[i][0] += d2.x; a_db2[i][1] += d2.y; a_db2[i][2] += d2.z; a_db2[i][3] += d2.w;
                }
                sg += (g[i].x + g[i].y) + (g[i].z + g[i].w);
                sgx += (g[i].x * v[i].x + g[i].y * v[i].y) + (g[i].z * v[i].z + g[i].w * v[i].w);
            }
        }
        const float mg = lanes_sum<LPR>(sg) / (float)cols, mgx = lanes_sum<LPR>(sgx) / (float)cols;
#pragma unroll
        for (int i = 0; i < LNB_V; ++i) {
            const int c = (i * LPR + lc) * 4;
            if (c < cols && live) {
                float4 o = make_float4(rstd * (g[i].x - mg - v[i].x * mgx), rstd * (g[i].y - mg - v[i].y * mgx),
                                       rstd * (g[i].z - mg - v[i].z * mgx), rstd * (g[i].w - mg - v[i].w * mgx));
                if (add) { o.x += ad[i].x; o.y += ad[i].y; o.z += ad[i].z; o.w += ad[i].w; }
                if (rmw) { o.x += old[i].x; o.y += old[i].y; o.z += old[i].z; o.w += old[i].w; }
                if (dx) *reinterpret_cast<float4*>(dx + row * ld_dx + c) = o;
                if (dx_lp) st4_any(dx_lp, lp_dtype, row * ld_dx_lp + c, o);
            }
        }
    }
    // block reduction of the parameter gradients
#pragma unroll
    for (int i = 0; i < LNB_V; ++i) {
        const int c = (i * LPR + lc) * 4;
        if (c < cols) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (dg1) atomicAdd(&sh[c + j], a_dg1[i][j]);
                if (db1) atomicAdd(&sh[cols + c + j], a_db1[i][j]);
                if (dg2) atomicAdd(&sh[2 * cols + c + j], a_dg2[i][j]);
                if (db2) atomicAdd(&sh[3 * cols + c + j], a_db2[i][j]);
            }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < cols; c += blockDim.x) {
        if (dg1) atomicAdd(dg1 + c, sh[c]);
        if (db1) atomicAdd(db1 + c, sh[cols + c]);
        if (dg2) atomicAdd(dg2 + c, sh[2 * cols + c]);
        if (db2) atomicAdd(db2 + c, sh[3 * cols + c]);
    }
}

// ---------------------------------------------------------------------------------------------------
// single-query cross-attention backward: one warp per (batch element, head), head dim 64, lanes own two dims.
//   p = softmax(q.k / 8); dp_j = dO.v_j; ds_j = p_j (dp_j - sum_i p_i dp_i);
//   dq = sum_j ds_j k_j / 8; dk_j = ds_j q / 8; dv_j = p_j dO
// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) cross_attention_bwd_kernel(const T* __restrict__ q, const T* __restrict__ kv, const T* __restrict__ d_out,
                                                                  T* __restrict__ dq, T* __restrict__ dkv, int batch, int keys, int heads) {
    grid_dep_sync();
    extern __shared__ float sc_all[];          // per warp: [keys] probabilities, [keys] dp
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const long long item = (long long)blockIdx.x * wpb + warp;
    if (item >= (long long)batch * heads) return;
    float* pr = sc_all + (size_t)warp * 2 * keys;
    float* dp = pr + keys;
    const int b = (int)(item / heads), h = (int)(item % heads), C = heads * 64;
    const long long qoff = (long long)b * C + h * 64 + lane * 2;
    const float q0 = to_float(q[qoff]), q1 = to_float(q[qoff + 1]);
    const float o0 = to_float(d_out[qoff]), o1 = to_float(d_out[qoff + 1]);
    const T* kbase = kv + (long long)b * keys * 2 * C + h * 64 + lane * 2;
    float mx = -INFINITY;
    for (int j = 0; j < keys; ++j) {
        const T* kp = kbase + (long long)j * 2 * C;
        const float s = warp_sum(q0 * to_float(kp[0]) + q1 * to_float(kp[1])) * 0.125f;
        const float d = warp_sum(o0 * to_float(kp[C]) + o1 * to_float(kp[C + 1]));
        if (lane == 0) { pr[j] = s; dp[j] = d; }
        mx = fmaxf(mx, s);
    }
    __syncwarp();
    float den = 0.f;
    for (int j = lane; j < keys; j += 32) {
        const float e = __expf(pr[j] - mx);
        pr[j] = e;
        den += e;
    }
    den = warp_sum(den);
    __syncwarp();
    const float inv = 1.f / den;
    float dot = 0.f;
    for (int j = lane; j < keys; j += 32) {
        pr[j] *= inv;
        dot += pr[j] * dp[j];
    }
    dot = warp_sum(dot);
    __syncwarp();
    float dq0 = 0.f, dq1 = 0.f;
    T* dbase = dkv + (long long)b * keys * 2 * C + h * 64 + lane * 2;
    for (int j = 0; j < keys; ++j) {
        const float p = pr[j];
        const float ds = p * (dp[j] - dot) * 0.125f;
        const T* kp = kbase + (long long)j * 2 * C;
        dq0 = fmaf(ds, to_float(kp[0]), dq0);
        dq1 = fmaf(ds, to_float(kp[1]), dq1);
        T* dk = dbase + (long long)j * 2 * C;
        dk[0] = from_float<T>(ds * q0);
        dk[1] = from_float<T>(ds * q1);
        dk[C] = from_float<T>(p * o0);
        dk[C + 1] = from_float<T>(p * o1);
    }
    dq[qoff] = from_float<T>(dq0);
    dq[qoff + 1] = from_float<T>(dq1);
}

// ---------------------------------------------------------------------------------------------------
// train-mode head + soft-target cross entropy: one block per clip.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* red, int tid) {
    const int lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += red[w];
    return t;
}
__device__ __forceinline__ float block_max(float v, float* red, int tid) {
    const int lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float t = red[0];
    for (int w = 1; w < nw; ++w) t = fmaxf(t, red[w]);
    return t;
}

__global__ void __launch_bounds__(256) softce_head_kernel(const float* __restrict__ emb, const float* __restrict__ text_n, float scale,
                                                          const float* __restrict__ target, int batch, int E, int C, float* logits, float* loss,
                                                          float* d_emb) {
    grid_dep_sync();
    extern __shared__ float sh[];          // [E] unit embedding, [C] logits -> d_logits, [E] u = d_logits . text_n, [32] scratch
    float* se = sh;
    float* sl = sh + E;
    float* su = sl + C;
    float* red = su + E;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    float ss = 0.f;
    for (int e = tid; e < E; e += blockDim.x) {
        const float v = emb[(long long)b * E + e];
        se[e] = v;
        ss += v * v;
    }
    const float inv_norm = rsqrtf(block_sum(ss, red, tid));
    for (int e = tid; e < E; e += blockDim.x) se[e] *= inv_norm;
    __syncthreads();
    for (int c = warp; c < C; c += nw) {
        float dot = 0.f;
        for (int e = lane; e < E; e += 32) dot = fmaf(se[e], text_n[(long long)c * E + e], dot);
        dot = warp_sum(dot);
        if (lane == 0) {
            sl[c] = dot * scale;
            if (logits) logits[(long long)b * C + c] = dot * scale;
        }
    }
    __syncthreads();
    float mx = -INFINITY;
    for (int c = tid; c < C; c += blockDim.x) mx = fmaxf(mx, sl[c]);
    mx = block_max(mx, red, tid);
    float den = 0.f, tsum = 0.f, tl = 0.f;
    for (int c = tid; c < C; c += blockDim.x) {
        den += __expf(sl[c] - mx);
        const float t = target[(long long)b * C + c];
        tsum += t;
        tl += t * sl[c];
    }
    den = block_sum(den, red, tid);
    tsum = block_sum(tsum, red, tid);
    tl = block_sum(tl, red, tid);
    const float lse = mx + __logf(den);
    if (tid == 0) atomicAdd(loss, (tsum * lse - tl) / (float)batch);       // sum_c -t_c (l_c - lse)
    __syncthreads();
    // d loss / d logits = (softmax * sum(t) - t) / batch
    for (int c = tid; c < C; c += blockDim.x)
        sl[c] = (__expf(sl[c] - lse) * tsum - target[(long long)b * C + c]) / (float)batch;
    __syncthreads();
    // u = scale * d_logits . text_n ; d_emb = (u - ehat (ehat . u)) / |emb|
    float eu = 0.f;
    for (int e = tid; e < E; e += blockDim.x) {
        float u = 0.f;
        for (int c = 0; c < C; ++c) u = fmaf(sl[c], text_n[(long long)c * E + e], u);
        u *= scale;
        su[e] = u;
        eu += u * se[e];
    }
    eu = block_sum(eu, red, tid);
    for (int e = tid; e < E; e += blockDim.x) d_emb[(long long)b * E + e] = (su[e] - se[e] * eu) * inv_norm;
}

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                    long long n, float lr, float beta1, float beta2, float eps, float decay, float inv_bc1,
                                                    float inv_sqrt_bc2, float grad_scale) {
    grid_dep_sync();
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float gi = g[i] * grad_scale;
        const float mi = beta1 * m[i] + (1.f - beta1) * gi;
        const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
        p[i] = p[i] * decay - lr * inv_bc1 * mi / denom;
    }
}

// operand copies: 32x32 tiles through shared memory so that both the straight and the transposed copy are coalesced
template <typename T>
__global__ void __launch_bounds__(256) pack_weight_kernel(const float* __restrict__ w, int n, int k, T* out, long long ld_out, T* out_t,
                                                          long long ld_out_t) {
    grid_dep_sync();
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const long long bz = blockIdx.z;
    const float* src = w + bz * (long long)n * k;
    const int n0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
    for (int i = ty; i < 32; i += 8) {
        const int nn = n0 + i, kk = k0 + tx;
        const float v = (nn < n && kk < k) ? src[(long long)nn * k + kk] : 0.f;
        tile[i][tx] = v;
        if (out && nn < n && kk < ld_out) out[bz * n * ld_out + (long long)nn * ld_out + kk] = from_float<T>(v);
    }
    __syncthreads();
    if (out_t)
        for (int i = ty; i < 32; i += 8) {
            const int kk = k0 + i, nn = n0 + tx;
            if (kk < k && nn < ld_out_t) out_t[bz * k * ld_out_t + (long long)kk * ld_out_t + nn] = from_float<T>(nn < n ? tile[tx][i] : 0.f);
        }
}

}  // namespace

}  // namespace distb200

using namespace distb200;

#define DISTB200_DTYPE_OK(dt) ((dt) == DISTB200_F32 || (dt) == DISTB200_BF16)

extern "C" int distb200_quickgelu(const void* z, int32_t z_dtype, float* y_f32, void* y_lp, int32_t lp_dtype, int64_t n, void* stream) {
    if (n == 0) return 0;
    DISTB200_REQUIRE(z && (y_f32 || y_lp) && DISTB200_DTYPE_OK(z_dtype) && DISTB200_DTYPE_OK(lp_dtype), "quickgelu: bad arguments");
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (z_dtype == DISTB200_BF16 && (!y_lp || lp_dtype == DISTB200_BF16) && n % 8 == 0 && al16(z) && al16(y_f32) && al16(y_lp)) {
        const long long n8 = n / 8;
        long long blocks = (n8 + 256 * 4 - 1) / (256 * 4);
        if (blocks > (long long)sm_count() * 8) blocks = (long long)sm_count() * 8;
        if (y_f32)
            DISTB200_LAUNCH((elementwise_bf16_kernel<EW_GELU, 4, false>), (unsigned)blocks, 256, 0, (cudaStream_t)stream, reinterpret_cast<const uint4*>(z),
                            (const uint4*)nullptr, y_f32, reinterpret_cast<uint4*>(y_lp), n8);
        else
            DISTB200_LAUNCH((elementwise_bf16_kernel<EW_GELU, 4, true>), (unsigned)blocks, 256, 0, (cudaStream_t)stream, reinterpret_cast<const uint4*>(z),
                            (const uint4*)nullptr, y_f32, reinterpret_cast<uint4*>(y_lp), n8);
        return check_launch("quickgelu");
    }
    DISTB200_LAUNCH(elementwise_kernel<EW_GELU>, grid_cap(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream, z, z_dtype, nullptr, 0, y_f32, y_lp, lp_dtype, n);
    return check_launch("quickgelu");
}

extern "C" int distb200_quickgelu_bwd(const void* dy, int32_t dy_dtype, const void* z, int32_t z_dtype, float* dz_f32, void* dz_lp,
                                      int32_t lp_dtype, int64_t n, void* stream) {
    if (n == 0) return 0;
    DISTB200_REQUIRE(dy && z && (dz_f32 || dz_lp) && DISTB200_DTYPE_OK(dy_dtype) && DISTB200_DTYPE_OK(z_dtype) && DISTB200_DTYPE_OK(lp_dtype),
                    "quickgelu_bwd: bad arguments");
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (dy_dtype == DISTB200_BF16 && z_dtype == DISTB200_BF16 && (!dz_lp || lp_dtype == DISTB200_BF16) && n % 8 == 0 && al16(dy) && al16(z) && al16(dz_f32) &&
        al16(dz_lp)) {
        const long long n8 = n / 8;
        long long blocks = (n8 + 256 * 4 - 1) / (256 * 4);
        if (blocks > (long long)sm_count() * 8) blocks = (long long)sm_count() * 8;
        if (dz_f32)
            DISTB200_LAUNCH((elementwise_bf16_kernel<EW_GELU_BWD, 4, false>), (unsigned)blocks, 256, 0, (cudaStream_t)stream, reinterpret_cast<const uint4*>(dy),
                            reinterpret_cast<const uint4*>(z), dz_f32, reinterpret_cast<uint4*>(dz_lp), n8);
        else
            DISTB200_LAUNCH((elementwise_bf16_kernel<EW_GELU_BWD, 4, true>), (unsigned)blocks, 256, 0, (cudaStream_t)stream, reinterpret_cast<const uint4*>(dy),
                            reinterpret_cast<const uint4*>(z), dz_f32, reinterpret_cast<uint4*>(dz_lp), n8);
        return check_launch("quickgelu_bwd");
    }
    DISTB200_LAUNCH(elementwise_kernel<EW_GELU_BWD>, grid_cap(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream, dy, dy_dtype, z, z_dtype, dz_f32, dz_lp, lp_dtype, n);
    return check_launch("quickgelu_bwd");
}

extern "C" int distb200_cast(const float* src, void* dst, int32_t dst_dtype, int64_t n, void* stream) {
    if (n == 0) return 0;
    DISTB200_REQUIRE(src && dst && DISTB200_DTYPE_OK(dst_dtype), "cast: bad arguments");
    DISTB200_LAUNCH(elementwise_kernel<EW_CAST>, grid_cap(n / 4 + 1, 256), 256, 0, (cudaStream_t)stream, src, DISTB200_F32, nullptr, 0, nullptr, dst, dst_dtype, n);
    return check_launch("cast");
}

extern "C" int distb200_group_sum(const void* src, int32_t src_dtype, int64_t groups, int32_t alpha, int64_t inner, void* dst,
                                  int32_t dst_dtype, void* stream) {
    if (groups == 0 || inner == 0) return 0;
    DISTB200_REQUIRE(src && dst && alpha >= 1 && inner % 4 == 0 && DISTB200_DTYPE_OK(src_dtype) && DISTB200_DTYPE_OK(dst_dtype),
                    "group_sum: bad arguments (inner must be a multiple of 4)");
    DISTB200_LAUNCH(group_sum_kernel, grid_cap(groups * inner / 4, 256), 256, 0, (cudaStream_t)stream, src, src_dtype, groups, alpha, inner, dst, dst_dtype);
    return check_launch("group_sum");
}

extern "C" int distb200_colsum(const void* src, int32_t src_dtype, int64_t ld, int64_t groups, int64_t rows_per_group, int64_t gstride,
                               int64_t roff, int64_t period, int32_t cols, float* out, float* out2, void* stream) {
    if (groups == 0 || rows_per_group == 0 || cols == 0) return 0;
    DISTB200_REQUIRE(src && out && period >= 1 && DISTB200_DTYPE_OK(src_dtype), "colsum: bad arguments");
    DISTB200_REQUIRE(cols % 4 == 0 && ld % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0, "colsum: cols and ld must be multiples of 4, src 16-byte aligned");
    DISTB200_REQUIRE(!out2 || period == 1, "colsum: the second output needs period == 1");
    const long long total = groups * rows_per_group;
    DISTB200_REQUIRE(total < (1ll << 31), "colsum: too many rows");
    if (src_dtype == DISTB200_BF16 && period == 1 && cols % 8 == 0 && ld % 8 == 0 && cols <= 2048) {
        const int tpr = cols / 8, rpb = 256 / tpr;
        long long blocks = (long long)sm_count() * 6;
        long long rpblk = (total + blocks - 1) / blocks;
        if (rpblk < 8 * rpb) rpblk = 8 * rpb;
        blocks = (total + rpblk - 1) / rpblk;
        if (gstride == rows_per_group)
            DISTB200_LAUNCH((colsum_bf16_dense_kernel<8, true>), (unsigned)blocks, 256, 0, (cudaStream_t)stream, reinterpret_cast<const uint4*>(src),
                            (long long)(ld / 8), (unsigned)total, (long long)roff, (unsigned)rows_per_group, (long long)gstride, tpr, rpb, out, out2, (unsigned)rpblk);
        else
            DISTB200_LAUNCH((colsum_bf16_dense_kernel<8, false>), (unsigned)blocks, 256, 0, (cudaStream_t)stream, reinterpret_cast<const uint4*>(src),
                            (long long)(ld / 8), (unsigned)total, (long long)roff, (unsigned)rows_per_group, (long long)gstride, tpr, rpb, out, out2, (unsigned)rpblk);
        return check_launch("colsum");
    }
    dim3 grid((unsigned)((cols / 4 + 31) / 32), 1);
    long long rows_per_block = total;
    if (period == 1) {
        long long splits = (long long)sm_count() * 8 / grid.x;
        if (splits < 1) splits = 1;
        rows_per_block = (total + splits - 1) / splits;
        if (rows_per_block < 128) rows_per_block = 128;
        grid.y = (unsigned)((total + rows_per_block - 1) / rows_per_block);
    } else {
        grid.y = (unsigned)(period < 65535 ? period : 65535);
    }
    DISTB200_LAUNCH(colsum_kernel, grid, 256, 0, (cudaStream_t)stream, src, src_dtype, ld, groups, rows_per_group, gstride, roff, period, cols, out, out2,
                                                          rows_per_block);
    return check_launch("colsum");
}

extern "C" int distb200_layernorm_bwd(const float* in1, int64_t ld_in1, const float* in2, int64_t ld_in2, int64_t in2_period, int64_t rows,
                                      int32_t cols, float eps, const float* g1, const void* dy1, int64_t ld_dy1, const float* g2,
                                      const void* dy2, int64_t ld_dy2, int32_t dy_dtype, const float* add, int64_t ld_add, float* dx,
                                      int64_t ld_dx, int32_t accumulate, void* dx_lp, int64_t ld_dx_lp, int32_t lp_dtype, float* dg1,
                                      float* db1, float* dg2, float* db2, void* stream) {
    if (rows == 0) return 0;
    DISTB200_REQUIRE(cols % 4 == 0 && cols <= 128 * LNB_MAXV, "layernorm_bwd: cols=%d must be a multiple of 4 and <= %d", cols, 128 * LNB_MAXV);
    DISTB200_REQUIRE(in1 && g1 && dy1 && (dx || dx_lp), "layernorm_bwd: null pointer");
    DISTB200_REQUIRE(!dy2 || g2, "layernorm_bwd: second affine parameters missing");
    DISTB200_REQUIRE(ld_in1 % 4 == 0 && ld_dy1 % 4 == 0 && (!dy2 || ld_dy2 % 4 == 0) && (!in2 || ld_in2 % 4 == 0) && (!add || ld_add % 4 == 0) &&
                        (!dx || ld_dx % 4 == 0) && (!dx_lp || ld_dx_lp % 4 == 0),
                    "layernorm_bwd: row pitches must be multiples of 4");
    DISTB200_REQUIRE(DISTB200_DTYPE_OK(dy_dtype) && DISTB200_DTYPE_OK(lp_dtype), "layernorm_bwd: bad dtype");
    if (!in2) in2_period = 1;
    DISTB200_REQUIRE(in2_period >= 1, "layernorm_bwd: in2_period must be >= 1");
    const size_t smem = (size_t)4 * cols * sizeof(float);
    // one wave of resident blocks (the register footprint, and with it the residency, follows the row width): every warp walks
    // rows_per_warp consecutive rows, which amortises the block-level reduction of the parameter gradients
#define DISTB200_LNB2(V, BF, LPR)                                                                                                          \
    do {                                                                                                                                \
        int per_sm = 1;                                                                                                                 \
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, layernorm_bwd_kernel<V, BF, LPR>, LNB_WARPS * 32, smem);                     \
        if (per_sm < 1) per_sm = 1;                                                                                                     \
        const long long warps = (long long)sm_count() * per_sm * LNB_WARPS;                                                            \
        int rows_per_warp = (int)((rows + warps - 1) / warps);                                                                          \
        if (rows_per_warp < 1) rows_per_warp = 1;                                                                                       \
        rows_per_warp = (rows_per_warp + 32 / LPR - 1) / (32 / LPR) * (32 / LPR);                                                       \
        const long long blocks = (rows + (long long)LNB_WARPS * rows_per_warp - 1) / ((long long)LNB_WARPS * rows_per_warp);            \
        DISTB200_LAUNCH((layernorm_bwd_kernel<V, BF, LPR>), (unsigned)blocks, LNB_WARPS * 32, smem, (cudaStream_t)stream,                                  \
            in1, ld_in1, in2, ld_in2, in2_period, rows, cols, eps, g1, dy1, ld_dy1, g2, dy2, ld_dy2, add, ld_add, dx, ld_dx, accumulate, dx_lp,  \
            ld_dx_lp, lp_dtype, dg1, db1, dg2, db2, rows_per_warp);                                                                     \
    } while (0)
#define DISTB200_LNB(V, LPR) do { if (dy_dtype == DISTB200_BF16) DISTB200_LNB2(V, true, LPR); else DISTB200_LNB2(V, false, LPR); } while (0)
    // float4 per lane: the register footprint (and with it the number of resident warps) follows the row width
    if (cols <= 32) DISTB200_LNB(1, 8);
    else if (cols <= 64) DISTB200_LNB(2, 8);
    else if (cols <= 96) DISTB200_LNB(3, 8);
    else if (cols <= 128) DISTB200_LNB(4, 8);
    else if (cols <= 256) DISTB200_LNB(2, 32);
    else if (cols <= 384) DISTB200_LNB(3, 32);
    else if (cols <= 512) DISTB200_LNB(4, 32);
    else if (cols <= 768) DISTB200_LNB(6, 32);
    else DISTB200_LNB(8, 32);
#undef DISTB200_LNB2
#undef DISTB200_LNB
    return check_launch("layernorm_bwd");
}

extern "C" int distb200_cross_attention_bwd(const void* q, const void* kv, const void* d_out, void* dq, void* dkv, int32_t batch, int32_t keys,
                                            int32_t heads, int32_t dtype, void* stream) {
    if (batch == 0) return 0;
    DISTB200_REQUIRE(q && kv && d_out && dq && dkv, "cross_attention_bwd: null pointer");
    DISTB200_REQUIRE(keys >= 1 && keys <= 2048, "cross_attention_bwd: keys=%d out of range", keys);
    const int wpb = 4;
    const long long items = (long long)batch * heads;
    const unsigned grid = (unsigned)((items + wpb - 1) / wpb);
    const size_t smem = (size_t)wpb * 2 * keys * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DISTB200_F32)
        DISTB200_LAUNCH(cross_attention_bwd_kernel<float>, grid, wpb * 32, smem, st, (const float*)q, (const float*)kv, (const float*)d_out, (float*)dq, (float*)dkv, batch, keys, heads);
    else
        DISTB200_LAUNCH(cross_attention_bwd_kernel<bf16>, grid, wpb * 32, smem, st, (const bf16*)q, (const bf16*)kv, (const bf16*)d_out, (bf16*)dq, (bf16*)dkv, batch, keys, heads);
    return check_launch("cross_attention_bwd");
}

extern "C" int distb200_softce_head(const float* emb, const float* text_n, float scale, const float* target, int32_t batch, int32_t embed_dim,
                                    int32_t classes, float* logits, float* loss, float* d_emb, void* stream) {
    if (batch == 0) return 0;
    DISTB200_REQUIRE(emb && text_n && target && loss && d_emb, "softce_head: null pointer");
    const size_t smem = (size_t)(2 * embed_dim + classes + 32) * sizeof(float);
    DISTB200_REQUIRE(smem <= 48 * 1024, "softce_head: E + C too large for one block (%zu bytes)", smem);
    DISTB200_LAUNCH(softce_head_kernel, batch, 256, smem, (cudaStream_t)stream, emb, text_n, scale, target, batch, embed_dim, classes, logits, loss, d_emb);
    return check_launch("softce_head");
}

extern "C" int distb200_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                              float weight_decay, int32_t step, float grad_scale, void* stream) {
    if (n == 0) return 0;
    DISTB200_REQUIRE(p && g && m && v && step >= 1, "adamw: bad arguments");
    const double bc1 = 1.0 - pow((double)beta1, (double)step), bc2 = 1.0 - pow((double)beta2, (double)step);
    DISTB200_LAUNCH(adamw_kernel, grid_cap(n, 256), 256, 0, (cudaStream_t)stream, p, g, m, v, n, lr, beta1, beta2, eps, 1.0f - lr * weight_decay, (float)(1.0 / bc1),
                                                                    (float)(1.0 / sqrt(bc2)), grad_scale);
    return check_launch("adamw");
}

extern "C" int distb200_pack_weight(const float* w, int64_t batch, int32_t n, int32_t k, void* out, int64_t ld_out, void* out_t,
                                    int64_t ld_out_t, int32_t dtype, void* stream) {
    if (batch == 0 || n == 0 || k == 0) return 0;
    DISTB200_REQUIRE(w && (out || out_t) && DISTB200_DTYPE_OK(dtype), "pack_weight: bad arguments");
    DISTB200_REQUIRE((!out || ld_out >= k) && (!out_t || ld_out_t >= n), "pack_weight: row pitch too small");
    const long long kc = out ? (ld_out > k ? ld_out : k) : k, nc = out_t ? (ld_out_t > n ? ld_out_t : n) : n;
    dim3 grid((unsigned)((kc + 31) / 32), (unsigned)((nc + 31) / 32), (unsigned)batch);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DISTB200_F32) DISTB200_LAUNCH(pack_weight_kernel<float>, grid, 256, 0, st, w, n, k, (float*)out, ld_out, (float*)out_t, ld_out_t);
    else DISTB200_LAUNCH(pack_weight_kernel<bf16>, grid, 256, 0, st, w, n, k, (bf16*)out, ld_out, (bf16*)out_t, ld_out_t);
    return check_launch("pack_weight");
}
