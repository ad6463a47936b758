// HBM-bound kernels of the path: LayerNorm(+cast), patchify, row broadcasts, frame means, single-query
// cross attention and the class-score head.  One warp per row / per (batch, head); vectorised, coalesced
// accesses; warp-shuffle reductions; fp32 statistics everywhere.
#include "common.cuh"

namespace distb200 {

namespace {

// ---------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, the row lives in registers (cols <= 32*4*MAXV).
// ---------------------------------------------------------------------------------------------------
constexpr int LN_MAXV = 8;   // float4 per lane -> cols <= 1024

template <typename OutT>
__device__ __forceinline__ void ln_store4(OutT* y, int c, float4 v);
template <>
__device__ __forceinline__ void ln_store4<float>(float* y, int c, float4 v) {
    *reinterpret_cast<float4*>(y + c) = v;
}
template <>
__device__ __forceinline__ void ln_store4<bf16>(bf16* y, int c, float4 v) {
    *reinterpret_cast<uint2*>(y + c) = make_uint2(pack_bf16x2(v.x, v.y), pack_bf16x2(v.z, v.w));
}

// LPR lanes cooperate on one row (32 for wide rows; 8 for rows of <= 128 floats such as the Ct = 96 temporal
// stream, so that a warp covers 4 rows and no lane idles); each lane holds up to LN_MAXV float4 of its row.
// Persistent: a warp walks over its rows with a grid stride and requests row i+1 (all of its 16-byte loads) before it
// normalises and stores row i, so every warp keeps a full row of loads in flight at all times instead of one row per
// (short-lived) block - isolated, 32 clips of B/16: 4.4 -> 5.2 TB/s (67 -> 79 % of the measured copy bandwidth) for the two-output 384-wide rows.
template <int LPR, int V>
__device__ __forceinline__ void ln_load(float4* v, const float* __restrict__ x, int sub, int cols) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int c = (i * LPR + sub) * 4;
        if (c < cols) v[i] = __ldcs(reinterpret_cast<const float4*>(x + c));
    }
}

template <typename OutT, int LPR, int V>
__device__ __forceinline__ void ln_finish(const float4* v, int sub, int cols, float eps, const float* __restrict__ g1, const float* __restrict__ b1,
                                          OutT* y1, const float* __restrict__ g2, const float* __restrict__ b2, OutT* y2, bool row_ok) {
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i)
        if ((i * LPR + sub) * 4 < cols) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)cols;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i)
        if ((i * LPR + sub) * 4 < cols) {
            const float a = v[i].x - mean, b = v[i].y - mean, cc = v[i].z - mean, dd = v[i].w - mean;
            sq += (a * a + b * b) + (cc * cc + dd * dd);
        }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / (float)cols + eps);
    if (!row_ok) return;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int c = (i * LPR + sub) * 4;
        if (c < cols) {
            float4 n;
            n.x = (v[i].x - mean) * rstd; n.y = (v[i].y - mean) * rstd;
            n.z = (v[i].z - mean) * rstd; n.w = (v[i].w - mean) * rstd;
            const float4 ga = __ldg(reinterpret_cast<const float4*>(g1 + c)), be = __ldg(reinterpret_cast<const float4*>(b1 + c));
            ln_store4<OutT>(y1, c, make_float4(n.x * ga.x + be.x, n.y * ga.y + be.y, n.z * ga.z + be.z, n.w * ga.w + be.w));
            if (y2) {
                const float4 gb = __ldg(reinterpret_cast<const float4*>(g2 + c)), bb = __ldg(reinterpret_cast<const float4*>(b2 + c));
                ln_store4<OutT>(y2, c, make_float4(n.x * gb.x + bb.x, n.y * gb.y + bb.y, n.z * gb.z + bb.z, n.w * gb.w + bb.w));
            }
        }
    }
}

template <typename OutT, int LPR, int V>
__global__ void __launch_bounds__(256, V <= 4 ? 3 : 2) layernorm_kernel(const float* __restrict__ in1, long long ld_in1,
                                                        const float* __restrict__ in2, long long ld_in2, long long in2_period,
                                                        long long rows, int cols, float eps,
                                                        const float* __restrict__ g1, const float* __restrict__ b1, OutT* y1, long long ld_y1,
                                                        const float* __restrict__ g2, const float* __restrict__ b2, OutT* y2, long long ld_y2) {
    grid_dep_sync();
    constexpr int RPW = 32 / LPR;                       // rows per warp and step
    const int lane = threadIdx.x & 31, sub = lane % LPR;
    const long long stride = (long long)gridDim.x * (blockDim.x >> 5) * RPW;
    long long row = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LPR;
    // warp-uniform trip count (the shuffles inside ln_finish need every lane): rows of lanes past the end are clamped
    const long long warp_row0 = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW;
    if (warp_row0 >= rows) return;
    const long long last = rows - 1;
    float4 va[V], vb[V];
    if (in2) {
        // periodic addend (ada-pooling inputs): few rows, no prefetch
        for (long long w0 = warp_row0; w0 < rows; w0 += stride, row += stride) {
            const long long r = row < rows ? row : last;
            ln_load<LPR, V>(va, in1 + r * ld_in1, sub, cols);
            ln_load<LPR, V>(vb, in2 + (r % in2_period) * ld_in2, sub, cols);
#pragma unroll
            for (int i = 0; i < V; ++i)
                if ((i * LPR + sub) * 4 < cols) { va[i].x += vb[i].x; va[i].y += vb[i].y; va[i].z += vb[i].z; va[i].w += vb[i].w; }
            ln_finish<OutT, LPR, V>(va, sub, cols, eps, g1, b1, y1 + r * ld_y1, g2, b2, y2 ? y2 + r * ld_y2 : nullptr, row < rows);
        }
        return;
    }
    ln_load<LPR, V>(va, in1 + (row < rows ? row : last) * ld_in1, sub, cols);
    for (long long w0 = warp_row0; w0 < rows; w0 += 2 * stride, row += 2 * stride) {
        const long long r1 = row + stride, r2 = row + 2 * stride;
        const bool more1 = w0 + stride < rows, more2 = w0 + 2 * stride < rows;
        if (more1) ln_load<LPR, V>(vb, in1 + (r1 < rows ? r1 : last) * ld_in1, sub, cols);
        {
            const long long r = row < rows ? row : last;
            ln_finish<OutT, LPR, V>(va, sub, cols, eps, g1, b1, y1 + r * ld_y1, g2, b2, y2 ? y2 + r * ld_y2 : nullptr, row < rows);
        }
        if (!more1) break;
        if (more2) ln_load<LPR, V>(va, in1 + (r2 < rows ? r2 : last) * ld_in1, sub, cols);
        {
            const long long r = r1 < rows ? r1 : last;
            ln_finish<OutT, LPR, V>(vb, sub, cols, eps, g1, b1, y1 + r * ld_y1, g2, b2, y2 ? y2 + r * ld_y2 : nullptr, r1 < rows);
        }
    }
}


__device__ __forceinline__ void load_row8(const float* p, float* v) {
    *reinterpret_cast<float4*>(v) = *reinterpret_cast<const float4*>(p);
    *reinterpret_cast<float4*>(v + 4) = *reinterpret_cast<const float4*>(p + 4);
}
__device__ __forceinline__ void load_row8(const bf16* p, float* v) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
        v[2 * j] = f.x;
        v[2 * j + 1] = f.y;
    }
}

// ---------------------------------------------------------------------------------------------------
// row statistics for the LayerNorm folded into a GEMM: one warp per row, the row in registers, two-pass variance
// ---------------------------------------------------------------------------------------------------
// LPR lanes per row (32, or 16 for rows of <= 512 values so that a warp covers two rows), V 16-byte pieces per lane.
template <typename T, int V, int LPR>
__global__ void __launch_bounds__(256) row_stats_kernel(const T* __restrict__ x, long long ld, long long rows, int cols, float eps, float2* stats) {
    grid_dep_sync();
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31, sub = lane % LPR;
    const long long row = ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * RPW + lane / LPR;
    const bool ok = row < rows;
    const T* p = x + (ok ? row : rows - 1) * ld;
    float v[8 * V];                      // 8 consecutive values (16 bytes of bf16) per lane and step
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int c = (i * LPR + sub) * 8;
        if (c < cols) {
            load_row8(p + c, v + 8 * i);
#pragma unroll
            for (int j = 0; j < 8; ++j) sum += v[8 * i + j];
        }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)cols;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < V; ++i) {
        const int c = (i * LPR + sub) * 8;
        if (c < cols)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float dlt = v[8 * i + j] - mean;
                sq += dlt * dlt;
            }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if (ok && sub == 0) stats[row] = make_float2(mean, rsqrtf(sq / (float)cols + eps));
}

// ---------------------------------------------------------------------------------------------------
// patchify: one group of TPR threads per image row (clip, frame, channel, y); VEC pixels per thread.  Reads are
// contiguous along the row, writes are VEC*esize contiguous pieces of the patch rows; index arithmetic once per row.
// ---------------------------------------------------------------------------------------------------
template <typename OutT, int VEC>
__device__ __forceinline__ void store_pixels(OutT* o, const float* px);
template <> __device__ __forceinline__ void store_pixels<float, 4>(float* o, const float* px) { *reinterpret_cast<float4*>(o) = make_float4(px[0], px[1], px[2], px[3]); }
template <> __device__ __forceinline__ void store_pixels<float, 2>(float* o, const float* px) { *reinterpret_cast<float2*>(o) = make_float2(px[0], px[1]); }
template <> __device__ __forceinline__ void store_pixels<bf16, 4>(bf16* o, const float* px) { *reinterpret_cast<uint2*>(o) = make_uint2(pack_bf16x2(px[0], px[1]), pack_bf16x2(px[2], px[3])); }
template <> __device__ __forceinline__ void store_pixels<bf16, 2>(bf16* o, const float* px) { *reinterpret_cast<uint32_t*>(o) = pack_bf16x2(px[0], px[1]); }

template <typename OutT, int VEC, int TPR>
__global__ void __launch_bounds__(256) patchify_kernel(const float* __restrict__ video, OutT* __restrict__ out, int clips, int T, int H, int W,
                                                       int p, int first, int step, int n_sel, long long ld_out) {
    grid_dep_sync();
    constexpr int RPB = 256 / TPR;                      // image rows per block
    const int g = W / p;
    const long long n_rows = (long long)clips * n_sel * 3 * H;
    const int tx = threadIdx.x % TPR;
    for (long long row = (long long)blockIdx.x * RPB + threadIdx.x / TPR; row < n_rows; row += (long long)gridDim.x * RPB) {
        long long rest = row;
        const int y = (int)(rest % H); rest /= H;
        const int c = (int)(rest % 3); rest /= 3;
        const int i = (int)(rest % n_sel);
        const int clip = (int)(rest / n_sel);
        const int frame = first + i * step;
        const float* src = video + ((((long long)clip * 3 + c) * T + frame) * H + y) * W;
        const int pr = y / p, iy = y - pr * p;
        OutT* dst = out + (((long long)clip * n_sel + i) * (g * g) + (long long)pr * g) * ld_out + (c * p + iy) * p;
        for (int x = tx * VEC; x < W; x += TPR * VEC) {
            float px[VEC];
            if (VEC == 4) *reinterpret_cast<float4*>(px) = *reinterpret_cast<const float4*>(src + x);
            else *reinterpret_cast<float2*>(px) = *reinterpret_cast<const float2*>(src + x);
            const int pc = x / p, ix = x - pc * p;
            store_pixels<OutT, VEC>(dst + (long long)pc * ld_out + ix, px);
        }
    }
}


// bf16 patch rows, staged: one block per (clip, selected frame, patch-row band).  The 3*p image rows of the band are read
// with coalesced 16-byte streaming loads (all of a thread's loads requested before any is used), re-ordered into patch rows
// in shared memory (row pitch padded by 32 bytes against bank conflicts) and written out as whole patch rows - the g patch
// rows of a band are contiguous in `out`, so the write is a flat stream of 16-byte stores, pad columns (zero) included.
template <bool P4>
__global__ void __launch_bounds__(256) patchify_band_kernel(const float* __restrict__ video, bf16* __restrict__ out, int T, int H, int W, int p,
                                                            int first, int step, int n_sel, int ld_out, int pitch) {
    grid_dep_sync();
    extern __shared__ __align__(16) unsigned char band_smem[];
    bf16* tile = reinterpret_cast<bf16*>(band_smem);                 // [g][pitch]
    const int g = W / p, gh = H / p, kk = 3 * p * p, pad = ld_out - kk;
    const long long band = blockIdx.x;
    const int pr = (int)(band % gh);
    const long long ci = band / gh;
    const int i = (int)(ci % n_sel);
    const long long clip = ci / n_sel;
    const int frame = first + i * step;
    for (int idx = threadIdx.x; idx < g * pad; idx += 256) tile[(idx / pad) * pitch + kk + idx % pad] = __float2bfloat16_rn(0.f);
    const int w4 = W >> 2, n4 = 3 * p * w4;
    const float* src0 = video + ((clip * 3 * T + frame) * H + (long long)pr * p) * W;       // channel 0, first row of the band
    const long long cstride = (long long)T * H * W;
    constexpr int U = 6;
    for (int base = threadIdx.x; base < n4; base += 256 * U) {
        float4 v[U];
        int r[U], x4[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int idx = base + u * 256;
            r[u] = idx / w4;                                          // c * p + iy
            x4[u] = idx - r[u] * w4;
            if (idx < n4) {
                const int c = r[u] / p, iy = r[u] - c * p;
                v[u] = __ldcs(reinterpret_cast<const float4*>(src0 + c * cstride + (long long)iy * W) + x4[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (base + u * 256 >= n4) break;
            const int x = 4 * x4[u];
            if (P4) {                                                 // p % 4 == 0: the four pixels share a patch
                const int pc = x / p, ix = x - pc * p;
                *reinterpret_cast<uint2*>(tile + pc * pitch + r[u] * p + ix) = make_uint2(pack_bf16x2(v[u].x, v[u].y), pack_bf16x2(v[u].z, v[u].w));
            } else {
                const float px[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int pc = (x + e) / p, ix = x + e - pc * p;
                    tile[pc * pitch + r[u] * p + ix] = __float2bfloat16_rn(px[e]);
                }
            }
        }
    }
    __syncthreads();
    const int l8 = ld_out >> 3;                                       // 16-byte pieces per patch row
    uint4* dst = reinterpret_cast<uint4*>(out + band * g * (long long)ld_out);
    for (int idx = threadIdx.x; idx < g * l8; idx += 256) {
        const int row = idx / l8, c8 = idx - row * l8;
        dst[idx] = *reinterpret_cast<const uint4*>(tile + row * pitch + 8 * c8);
    }
}


// patchify straight from decoded frames: uint8 [clips, T, H, W, 3] (the decoder's layout, before ToTensorVideo) with the
// normalisation (x / 255 - mean) / std fused in - same operation order as torchvision's to_tensor + normalize, so the fp32
// result is bit-identical to patchifying the normalised float clip.  Thread = VEC pixels (3 * VEC contiguous bytes).
template <typename OutT, int VEC, int TPR>
__global__ void __launch_bounds__(256) patchify_u8_kernel(const uint8_t* __restrict__ frames, OutT* __restrict__ out, int clips, int T, int H, int W,
                                                          int p, int first, int step, int n_sel, long long ld_out, float m0, float m1, float m2,
                                                          float s0, float s1, float s2) {
    grid_dep_sync();
    constexpr int RPB = 256 / TPR;
    const int g = W / p;
    const long long n_rows = (long long)clips * n_sel * H;
    const int tx = threadIdx.x % TPR;
    const float mean[3] = {m0, m1, m2}, sd[3] = {s0, s1, s2};
    for (long long row = (long long)blockIdx.x * RPB + threadIdx.x / TPR; row < n_rows; row += (long long)gridDim.x * RPB) {
        long long rest = row;
        const int y = (int)(rest % H); rest /= H;
        const int i = (int)(rest % n_sel);
        const int clip = (int)(rest / n_sel);
        const int frame = first + i * step;
        const uint8_t* src = frames + (((long long)clip * T + frame) * H + y) * W * 3;
        const int pr = y / p, iy = y - pr * p;
        OutT* dst = out + (((long long)clip * n_sel + i) * (g * g) + (long long)pr * g) * ld_out + iy * p;
        for (int x = tx * VEC; x < W; x += TPR * VEC) {
            uint8_t b[3 * VEC];
            if (VEC == 4) {
                const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src + 3 * x);
                uint32_t w3[3] = {s32[0], s32[1], s32[2]};
#pragma unroll
                for (int j = 0; j < 12; ++j) b[j] = (uint8_t)(w3[j >> 2] >> (8 * (j & 3)));
            } else {
                const uint16_t* s16 = reinterpret_cast<const uint16_t*>(src + 3 * x);
                uint16_t w3[3] = {s16[0], s16[1], s16[2]};
#pragma unroll
                for (int j = 0; j < 6; ++j) b[j] = (uint8_t)(w3[j >> 1] >> (8 * (j & 1)));
            }
            const int pc = x / p, ix = x - pc * p;
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float px[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) px[v] = ((float)b[3 * v + c] / 255.0f - mean[c]) / sd[c];
                store_pixels<OutT, VEC>(dst + (long long)pc * ld_out + c * p * p + ix, px);
            }
        }
    }
}

template <typename OutT>
__global__ void zero_pad_cols_kernel(OutT* out, long long rows, int c0, long long ld) {
    grid_dep_sync();
    const int pad = (int)ld - c0;
    const long long total = rows * pad;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
        out[(idx / pad) * ld + c0 + (idx % pad)] = from_float<OutT>(0.f);
}

// ---------------------------------------------------------------------------------------------------
__global__ void rows_bcast_kernel(float* dst, long long row_stride, long long n_rows, int cols, const float* __restrict__ table,
                                  long long period, int accumulate, bf16* dst2, long long row_stride2) {
    grid_dep_sync();
    const long long total = n_rows * cols;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long i = idx / cols;
        const int c = (int)(idx % cols);
        const float t = table[(i % period) * cols + c];
        float* p = dst + i * row_stride + c;
        const float v = accumulate ? *p + t : t;
        *p = v;
        if (dst2) dst2[i * row_stride2 + c] = __float2bfloat16_rn(v);
    }
}

template <typename OutT>
__global__ void mean_rows_kernel(const float* __restrict__ src, long long row_stride, int count, long long batch, int cols, OutT* out) {
    grid_dep_sync();
    const long long total = batch * cols;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long b = idx / cols;
        const int c = (int)(idx % cols);
        float s = 0.f;
        for (int i = 0; i < count; ++i) s += src[(b * count + i) * row_stride + c];
        out[idx] = from_float<OutT>(s / (float)count);
    }
}

// ---------------------------------------------------------------------------------------------------
// single-query cross attention: one warp per (batch element, head), head dim 64.
//   scores: one key per lane (the lane reads its whole 64-wide key row: full 128 B lines, 64 independent FMAs),
//   softmax over the lanes' scores, then P.V with two dims per lane and four keys in flight.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load_row64(const float* p, float* v) {
#pragma unroll
    for (int i = 0; i < 16; ++i) *reinterpret_cast<float4*>(v + 4 * i) = *reinterpret_cast<const float4*>(p + 4 * i);
}
__device__ __forceinline__ void load_row64(const bf16* p, float* v) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint4 u = *reinterpret_cast<const uint4*>(p + 8 * i);
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[j]));
            v[8 * i + 2 * j] = f.x;
            v[8 * i + 2 * j + 1] = f.y;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(128) cross_attention_kernel(const T* __restrict__ q, const T* __restrict__ kv, T* __restrict__ out,
                                                              int batch, int keys, int heads) {
    grid_dep_sync();
    extern __shared__ float sc_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const long long item = (long long)blockIdx.x * wpb + warp;
    if (item >= (long long)batch * heads) return;
    float* sc = sc_all + (size_t)warp * keys;
    const int b = (int)(item / heads), h = (int)(item % heads);
    const int C = heads * 64;
    float qv[64];
    load_row64(q + (long long)b * C + h * 64, qv);
    const T* kbase = kv + (long long)b * keys * 2 * C + h * 64;
    float mx = -INFINITY;
    for (int j = lane; j < keys; j += 32) {
        float kr[64];
        load_row64(kbase + (long long)j * 2 * C, kr);
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < 64; i += 2) {
            s0 = fmaf(qv[i], kr[i], s0);
            s1 = fmaf(qv[i + 1], kr[i + 1], s1);
        }
        const float s = (s0 + s1) * 0.125f;
        sc[j] = s;
        mx = fmaxf(mx, s);
    }
    mx = warp_max(mx);
    float den = 0.f;
    for (int j = lane; j < keys; j += 32) {
        const float e = __expf(sc[j] - mx);
        sc[j] = e;
        den += e;
    }
    den = warp_sum(den);
    __syncwarp();
    float o0[4] = {0.f, 0.f, 0.f, 0.f}, o1[4] = {0.f, 0.f, 0.f, 0.f};
    const T* vbase = kbase + C + lane * 2;
    int j = 0;
    for (; j + 4 <= keys; j += 4) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const T* vp = vbase + (long long)(j + u) * 2 * C;
            const float pj = sc[j + u];
            o0[u] = fmaf(pj, to_float(vp[0]), o0[u]);
            o1[u] = fmaf(pj, to_float(vp[1]), o1[u]);
        }
    }
    for (; j < keys; ++j) {
        const T* vp = vbase + (long long)j * 2 * C;
        o0[0] = fmaf(sc[j], to_float(vp[0]), o0[0]);
        o1[0] = fmaf(sc[j], to_float(vp[1]), o1[0]);
    }
    const float inv = 1.f / den;
    out[(long long)b * C + h * 64 + lane * 2] = from_float<T>(((o0[0] + o0[1]) + (o0[2] + o0[3])) * inv);
    out[(long long)b * C + h * 64 + lane * 2 + 1] = from_float<T>(((o1[0] + o1[1]) + (o1[2] + o1[3])) * inv);
}

// ---------------------------------------------------------------------------------------------------
// class head: one block per clip.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) class_head_kernel(const float* __restrict__ emb, const float* __restrict__ text_n, float scale,
                                                         int E, int C, float* logits, float* probs) {
    grid_dep_sync();
    extern __shared__ float sh[];          // [E] embedding, [C] logits, [32] scratch
    float* se = sh;
    float* sl = sh + E;
    float* red = sl + C;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    float ss = 0.f;
    for (int e = tid; e < E; e += blockDim.x) {
        const float v = emb[(long long)b * E + e];
        se[e] = v;
        ss += v * v;
    }
    ss = warp_sum(ss);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    float tot = 0.f;
    for (int w = 0; w < nw; ++w) tot += red[w];
    const float inv = scale * rsqrtf(tot);
    __syncthreads();
    for (int c = warp; c < C; c += nw) {
        // four independent loads in flight per lane: the loop is bound by the latency of the label-embedding reads
        const float* tr = text_n + (long long)c * E;
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
        int e = lane;
        for (; e + 96 < E; e += 128) {
            const float t0 = __ldg(tr + e), t1 = __ldg(tr + e + 32), t2 = __ldg(tr + e + 64), t3 = __ldg(tr + e + 96);
            d0 = fmaf(se[e], t0, d0); d1 = fmaf(se[e + 32], t1, d1); d2 = fmaf(se[e + 64], t2, d2); d3 = fmaf(se[e + 96], t3, d3);
        }
        for (; e < E; e += 32) d0 = fmaf(se[e], __ldg(tr + e), d0);
        float dot = warp_sum((d0 + d1) + (d2 + d3));
        if (lane == 0) {
            sl[c] = dot * inv;
            if (logits) logits[(long long)b * C + c] = dot * inv;
        }
    }
    __syncthreads();
    if (!probs) return;
    float mx = -INFINITY;
    for (int c = tid; c < C; c += blockDim.x) mx = fmaxf(mx, sl[c]);
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
    for (int w = 1; w < nw; ++w) mx = fmaxf(mx, red[w]);
    __syncthreads();
    float den = 0.f;
    for (int c = tid; c < C; c += blockDim.x) den += __expf(sl[c] - mx);
    den = warp_sum(den);
    if (lane == 0) red[warp] = den;
    __syncthreads();
    den = 0.f;
    for (int w = 0; w < nw; ++w) den += red[w];
    for (int c = tid; c < C; c += blockDim.x) probs[(long long)b * C + c] = __expf(sl[c] - mx) / den;
}

// class_head with the zero-shot / prediction-fusion branch of clip.py:519-527 fused in: the clip's video embedding and its t
// per-frame CLIP image embeddings are scored against the label matrix in one pass; logits = w * video + (1 - w) * mean_t frame.
// Also writes the L2-normalised frame embeddings (the `img_logits` the reference returns on this branch).
__global__ void __launch_bounds__(1024) class_head_fused_kernel(const float* __restrict__ emb, const float* __restrict__ img, int t, float w,
                                                               const float* __restrict__ text_n, float scale, int E, int C, float* logits,
                                                               float* probs, float* img_n) {
    grid_dep_sync();
    extern __shared__ float sh[];          // [E] current vector, [C] blended logits, [32] scratch
    float* se = sh;
    float* sl = sh + E;
    float* red = sl + C;
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    for (int c = tid; c < C; c += blockDim.x) sl[c] = 0.f;
    for (int f = -1; f < t; ++f) {
        const float* src = f < 0 ? emb + (long long)b * E : img + ((long long)b * t + f) * E;
        const float weight = f < 0 ? w : (1.0f - w) / (float)t;
        __syncthreads();
        float ss = 0.f;
        for (int e = tid; e < E; e += blockDim.x) {
            const float v = src[e];
            se[e] = v;
            ss += v * v;
        }
        ss = warp_sum(ss);
        if (lane == 0) red[warp] = ss;
        __syncthreads();
        float tot = 0.f;
        for (int i = 0; i < nw; ++i) tot += red[i];
        const float rn = rsqrtf(tot);
        if (f >= 0 && img_n)
            for (int e = tid; e < E; e += blockDim.x) img_n[((long long)b * t + f) * E + e] = se[e] * rn;
        const float inv = scale * rn * weight;
        for (int c = warp; c < C; c += nw) {
            const float* tr = text_n + (long long)c * E;
            float d0 = 0.f, d1 = 0.f;
            int e = lane;
            for (; e + 32 < E; e += 64) {
                const float t0 = __ldg(tr + e), t1 = __ldg(tr + e + 32);
                d0 = fmaf(se[e], t0, d0); d1 = fmaf(se[e + 32], t1, d1);
            }
            for (; e < E; e += 32) d0 = fmaf(se[e], __ldg(tr + e), d0);
            const float dot = warp_sum(d0 + d1);
            if (lane == 0) sl[c] += dot * inv;           // class c belongs to this warp in every round: no race
        }
    }
    __syncthreads();
    if (logits)
        for (int c = tid; c < C; c += blockDim.x) logits[(long long)b * C + c] = sl[c];
    if (!probs) return;
    float mx = -INFINITY;
    for (int c = tid; c < C; c += blockDim.x) mx = fmaxf(mx, sl[c]);
    mx = warp_max(mx);
    if (lane == 0) red[warp] = mx;
    __syncthreads();
    mx = red[0];
    for (int i = 1; i < nw; ++i) mx = fmaxf(mx, red[i]);
    __syncthreads();
    float den = 0.f;
    for (int c = tid; c < C; c += blockDim.x) den += __expf(sl[c] - mx);
    den = warp_sum(den);
    if (lane == 0) red[warp] = den;
    __syncthreads();
    den = 0.f;
    for (int i = 0; i < nw; ++i) den += red[i];
    for (int c = tid; c < C; c += blockDim.x) probs[(long long)b * C + c] = __expf(sl[c] - mx) / den;
}


// ---------------------------------------------------------------------------------------------------
// multi-view ensemble (utils/meters.py:83-115): every clip adds (or max-es) its class scores into the row of its video
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) view_ensemble_kernel(const float* __restrict__ preds, const long long* __restrict__ labels,
                                                            const long long* __restrict__ clip_ids, int n, int C, int num_clips, int method,
                                                            float* video_preds, long long* video_labels, long long* clip_count,
                                                            long long num_videos) {
    grid_dep_sync();
    const long long total = (long long)n * C;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx / C), c = (int)(idx % C);
        const long long vid = clip_ids[i] / num_clips;
        if (vid < 0 || vid >= num_videos) continue;
        const float v = preds[idx];
        float* dst = video_preds + vid * C + c;
        if (method == 0) atomicAdd(dst, v);
        else atomicMax(reinterpret_cast<int*>(dst), __float_as_int(fmaxf(v, 0.f)));      // scores are probabilities (>= 0): int order = float order
        if (c == 0) {
            video_labels[vid] = labels[i];
            atomicAdd(reinterpret_cast<unsigned long long*>(clip_count + vid), 1ull);
        }
    }
}

// top-k hits (utils/metrics.py topks_correct): a video counts for k when fewer than k classes score strictly higher than its label
__global__ void __launch_bounds__(256) topk_correct_kernel(const float* __restrict__ video_preds, const long long* __restrict__ video_labels,
                                                           long long num_videos, int C, const int* __restrict__ ks, int nk, long long* correct) {
    grid_dep_sync();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long vid = (long long)blockIdx.x * (blockDim.x >> 5) + warp;
    if (vid >= num_videos) return;
    const long long lab = video_labels[vid];
    if (lab < 0 || lab >= C) return;
    const float ref = video_preds[vid * C + lab];
    int higher = 0;
    for (int c = lane; c < C; c += 32) higher += video_preds[vid * C + c] > ref ? 1 : 0;
    for (int o = 16; o > 0; o >>= 1) higher += __shfl_xor_sync(0xffffffffu, higher, o);
    if (lane == 0)
        for (int j = 0; j < nk; ++j)
            if (higher < ks[j]) atomicAdd(reinterpret_cast<unsigned long long*>(correct + j), 1ull);
}

// ---------------------------------------------------------------------------------------------------
// CLIP text tower input / output rows (clip.py:419-434): token embedding lookup + positional embedding, and the row of
// the end-of-text token (the largest id of a sequence; first occurrence, as torch.argmax returns it).
__global__ void embed_tokens_kernel(const long long* __restrict__ ids, const float* __restrict__ table, const float* __restrict__ pos,
                                    long long rows, int ctx, int width, float* __restrict__ out) {
    grid_dep_sync();
    const int w4 = width >> 2;
    const long long total = rows * w4;
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const long long r = idx / w4;
        const int c = (int)(idx % w4);
        const float4 e = reinterpret_cast<const float4*>(table + ids[r] * width)[c];
        const float4 p = reinterpret_cast<const float4*>(pos + (r % ctx) * width)[c];
        reinterpret_cast<float4*>(out + r * width)[c] = make_float4(e.x + p.x, e.y + p.y, e.z + p.z, e.w + p.w);
    }
}

__global__ void gather_eot_kernel(const float* __restrict__ x, const long long* __restrict__ ids, int ctx, int width, float* __restrict__ out) {
    grid_dep_sync();
    __shared__ int s_best;
    const long long s = blockIdx.x;
    if (threadIdx.x < 32) {
        long long best = ids[s * ctx];
        int at = 0;
        for (int i = threadIdx.x; i < ctx; i += 32) {
            const long long v = ids[s * ctx + i];
            if (v > best) { best = v; at = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const long long ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oa = __shfl_xor_sync(0xffffffffu, at, o);
            if (ob > best || (ob == best && oa < at)) { best = ob; at = oa; }
        }
        if (threadIdx.x == 0) s_best = at;
    }
    __syncthreads();
    const float* src = x + (s * ctx + s_best) * width;
    for (int c = threadIdx.x; c < width; c += blockDim.x) out[s * width + c] = src[c];
}

// (mean, rstd) from the statistic slots a GEMM epilogue emitted (distb200_gemm_desc.stat_partials): 8 lanes per row, two slots each
__global__ void __launch_bounds__(256) row_stats_finalize_kernel(const float4* __restrict__ partials, long long rows, float inv_cols, float eps,
                                                                 float2* __restrict__ stats) {
    grid_dep_sync();
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long row = idx >> 3;
    const int part = (int)(idx & 7);
    float s = 0.f, q = 0.f;
    if (row < rows) {
        const float4 v = __ldcs(partials + row * (DISTB200_STAT_SLOTS / 2) + part);      // slots 2*part, 2*part + 1
        s = v.x + v.z;
        q = v.y + v.w;
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (row < rows && part == 0) {
        const float mean = s * inv_cols;
        stats[row] = make_float2(mean, rsqrtf(fmaxf(q * inv_cols - mean * mean, 0.f) + eps));
    }
}

// persistent LayerNorm grid: every resident block slot of the device, fewer when the rows do not fill them
inline unsigned ln_grid(long long rows, int rows_per_block, int blocks_per_sm) {
    const long long want = (rows + rows_per_block - 1) / rows_per_block, cap = (long long)sm_count() * blocks_per_sm;
    return (unsigned)(want < cap ? want : cap);
}

inline unsigned grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    const long long cap = (long long)sm_count() * 16;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

}  // namespace distb200

using namespace distb200;

extern "C" int distb200_layernorm(const float* in1, int64_t ld_in1, const float* in2, int64_t ld_in2, int64_t in2_period,
                                  int64_t rows, int32_t cols, float eps, const float* g1, const float* b1, void* y1, int64_t ld_y1,
                                  const float* g2, const float* b2, void* y2, int64_t ld_y2, int32_t out_dtype, void* stream) {
    if (rows == 0) return 0;
    DISTB200_REQUIRE(cols % 4 == 0 && cols <= 128 * LN_MAXV, "layernorm: cols=%d must be a multiple of 4 and <= %d", cols, 128 * LN_MAXV);
    DISTB200_REQUIRE(ld_in1 % 4 == 0 && ld_y1 % 4 == 0 && (!in2 || ld_in2 % 4 == 0) && (!y2 || ld_y2 % 4 == 0), "layernorm: row pitches must be multiples of 4");
    DISTB200_REQUIRE(in1 && g1 && b1 && y1, "layernorm: null pointer");
    DISTB200_REQUIRE(!y2 || (g2 && b2), "layernorm: second affine parameters missing");
    if (!in2) in2_period = 1;
    DISTB200_REQUIRE(in2_period >= 1, "layernorm: in2_period must be >= 1");
    const int wpb = 8;
    cudaStream_t st = (cudaStream_t)stream;
#define DISTB200_LN(T, LPR, V)                                                                                                      \
    DISTB200_LAUNCH((layernorm_kernel<T, LPR, V>), ln_grid(rows, wpb * (32 / LPR), (V) <= 4 ? 3 : 2), wpb * 32, 0, st,                              \
        in1, ld_in1, in2, ld_in2, in2_period, rows, cols, eps, g1, b1, (T*)y1, ld_y1, g2, b2, (T*)y2, ld_y2)
#define DISTB200_LN_T(T)                                                                                                             \
    do {                                                                                                                             \
        if (cols <= 128) DISTB200_LN(T, 8, 4);                                                                                        \
        else if (cols <= 256) DISTB200_LN(T, 32, 2);                                                                                  \
        else if (cols <= 384) DISTB200_LN(T, 32, 3);                                                                                  \
        else if (cols <= 512) DISTB200_LN(T, 32, 4);                                                                                  \
        else if (cols <= 768) DISTB200_LN(T, 32, 6);                                                                                  \
        else DISTB200_LN(T, 32, 8);                                                                                                   \
    } while (0)
    // the register footprint (and with it the number of resident warps) follows the row width
    if (out_dtype == DISTB200_F32) DISTB200_LN_T(float); else DISTB200_LN_T(bf16);
#undef DISTB200_LN_T
#undef DISTB200_LN
    return check_launch("layernorm");
}


extern "C" int distb200_row_stats(const void* x, int32_t dtype, int64_t ld, int64_t rows, int32_t cols, float eps, float* stats, void* stream) {
    if (rows == 0) return 0;
    DISTB200_REQUIRE(x && stats, "row_stats: null pointer");
    DISTB200_REQUIRE(cols % 8 == 0 && cols <= 1024 && ld % 8 == 0, "row_stats: cols=%d must be a multiple of 8 and <= 1024 (ld %% 8 == 0)", cols);
    DISTB200_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(stats) & 7) == 0, "row_stats: alignment");
#define DISTB200_RS(T, V, LPR) DISTB200_LAUNCH((row_stats_kernel<T, V, LPR>), (unsigned)((rows + 8 * (32 / LPR) - 1) / (8 * (32 / LPR))), 256, 0, stream, \
                                               (const T*)x, ld, rows, cols, eps, (float2*)stats)
    if (dtype == DISTB200_F32) {
        if (cols <= 256) DISTB200_RS(float, 2, 16); else if (cols <= 512) DISTB200_RS(float, 4, 16); else if (cols <= 768) DISTB200_RS(float, 3, 32); else DISTB200_RS(float, 4, 32);
    } else {
        if (cols <= 256) DISTB200_RS(bf16, 2, 16); else if (cols <= 512) DISTB200_RS(bf16, 4, 16); else if (cols <= 768) DISTB200_RS(bf16, 3, 32); else DISTB200_RS(bf16, 4, 32);
    }
#undef DISTB200_RS
    return check_launch("row_stats");
}

extern "C" int distb200_patchify(const float* video, void* out, int32_t clips, int32_t T, int32_t H, int32_t W, int32_t p,
                                 int32_t first_frame, int32_t frame_step, int32_t n_sel, int64_t ld_out, int32_t out_dtype, void* stream) {
    DISTB200_REQUIRE(H % p == 0 && W % p == 0 && W % 2 == 0 && p % 2 == 0, "patchify: H=%d W=%d p=%d", H, W, p);
    DISTB200_REQUIRE(first_frame >= 0 && frame_step >= 1 && first_frame + (n_sel - 1) * frame_step < T, "patchify: frame selection out of range");
    DISTB200_REQUIRE(ld_out >= 3 * p * p, "patchify: ld_out too small");
    const long long total = (long long)clips * n_sel * 3 * H * (W / 2);
    if (total == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const long long rows = (long long)clips * n_sel * (H / p) * (W / p);
    const long long img_rows = (long long)clips * n_sel * 3 * H;
    const bool v4 = p % 4 == 0 && W % 4 == 0 && ld_out % 4 == 0;
#define DISTB200_PATCHIFY(T, VEC, TPR)                                                                                             \
    DISTB200_LAUNCH((patchify_kernel<T, VEC, TPR>), grid_for(img_rows * TPR, 256), 256, 0, st, video, (T*)out, clips, T_, H, W, p, first_frame,   \
                                                                                 frame_step, n_sel, ld_out)
    const int T_ = T;
    if (out_dtype == DISTB200_F32) {
        if (v4) DISTB200_PATCHIFY(float, 4, 64); else DISTB200_PATCHIFY(float, 2, 128);
        if (ld_out > 3 * p * p) DISTB200_LAUNCH(zero_pad_cols_kernel<float>, grid_for(rows * (ld_out - 3 * p * p), 256), 256, 0, st, (float*)out, rows, 3 * p * p, ld_out);
    } else {
        const int pitch = (int)ld_out + 16;                              // +32 bytes per patch row: conflict-free transposing stores
        const size_t band_smem = (size_t)(W / p) * pitch * sizeof(bf16);
        if (W % 4 == 0 && ld_out % 8 == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && (reinterpret_cast<uintptr_t>(video) & 15) == 0 &&
            band_smem <= 48 * 1024 && ld_out < (1 << 20)) {
            const long long bands = (long long)clips * n_sel * (H / p);
            DISTB200_REQUIRE(bands < (1ll << 31), "patchify: too many bands");
            if (p % 4 == 0) DISTB200_LAUNCH(patchify_band_kernel<true>, (unsigned)bands, 256, band_smem, st, video, (bf16*)out, T, H, W, p, first_frame, frame_step, n_sel, (int)ld_out, pitch);
            else DISTB200_LAUNCH(patchify_band_kernel<false>, (unsigned)bands, 256, band_smem, st, video, (bf16*)out, T, H, W, p, first_frame, frame_step, n_sel, (int)ld_out, pitch);
        } else {
            if (v4) DISTB200_PATCHIFY(bf16, 4, 64); else DISTB200_PATCHIFY(bf16, 2, 128);
            if (ld_out > 3 * p * p) DISTB200_LAUNCH(zero_pad_cols_kernel<bf16>, grid_for(rows * (ld_out - 3 * p * p), 256), 256, 0, st, (bf16*)out, rows, 3 * p * p, ld_out);
        }
    }
#undef DISTB200_PATCHIFY
    return check_launch("patchify");
}


extern "C" int distb200_patchify_u8(const uint8_t* frames, void* out, int32_t clips, int32_t T, int32_t H, int32_t W, int32_t p,
                                    int32_t first_frame, int32_t frame_step, int32_t n_sel, int64_t ld_out, int32_t out_dtype,
                                    const float* mean3, const float* std3, void* stream) {
    DISTB200_REQUIRE(frames && out && mean3 && std3, "patchify_u8: null pointer");
    DISTB200_REQUIRE(H % p == 0 && W % p == 0 && W % 2 == 0 && p % 2 == 0, "patchify_u8: H=%d W=%d p=%d", H, W, p);
    DISTB200_REQUIRE(first_frame >= 0 && frame_step >= 1 && first_frame + (n_sel - 1) * frame_step < T, "patchify_u8: frame selection out of range");
    DISTB200_REQUIRE(ld_out >= 3 * p * p, "patchify_u8: ld_out too small");
    DISTB200_REQUIRE((reinterpret_cast<uintptr_t>(frames) & 3) == 0, "patchify_u8: frames must be 4-byte aligned");
    const long long img_rows = (long long)clips * n_sel * H;
    if (img_rows == 0) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    const long long rows = (long long)clips * n_sel * (H / p) * (W / p);
    const bool v4 = p % 4 == 0 && W % 4 == 0 && ld_out % 4 == 0;
#define DISTB200_PATCHIFY_U8(T_, VEC, TPR)                                                                                          \
    DISTB200_LAUNCH((patchify_u8_kernel<T_, VEC, TPR>), grid_for(img_rows * TPR, 256), 256, 0, st, frames, (T_*)out, clips, T, H, W, p, first_frame, \
                                                                                   frame_step, n_sel, ld_out, mean3[0], mean3[1],  \
                                                                                   mean3[2], std3[0], std3[1], std3[2])
    if (out_dtype == DISTB200_F32) {
        if (v4) DISTB200_PATCHIFY_U8(float, 4, 64); else DISTB200_PATCHIFY_U8(float, 2, 128);
        if (ld_out > 3 * p * p) DISTB200_LAUNCH(zero_pad_cols_kernel<float>, grid_for(rows * (ld_out - 3 * p * p), 256), 256, 0, st, (float*)out, rows, 3 * p * p, ld_out);
    } else {
        if (v4) DISTB200_PATCHIFY_U8(bf16, 4, 64); else DISTB200_PATCHIFY_U8(bf16, 2, 128);
        if (ld_out > 3 * p * p) DISTB200_LAUNCH(zero_pad_cols_kernel<bf16>, grid_for(rows * (ld_out - 3 * p * p), 256), 256, 0, st, (bf16*)out, rows, 3 * p * p, ld_out);
    }
#undef DISTB200_PATCHIFY_U8
    return check_launch("patchify_u8");
}


extern "C" int distb200_view_ensemble(const float* preds, const int64_t* labels, const int64_t* clip_ids, int32_t n, int32_t classes,
                                      int32_t num_clips, int32_t method, float* video_preds, int64_t* video_labels, int64_t* clip_count,
                                      int64_t num_videos, void* stream) {
    if (n == 0) return 0;
    DISTB200_REQUIRE(preds && labels && clip_ids && video_preds && video_labels && clip_count, "view_ensemble: null pointer");
    DISTB200_REQUIRE(num_clips >= 1 && (method == 0 || method == 1), "view_ensemble: bad arguments");
    DISTB200_LAUNCH(view_ensemble_kernel, grid_for((long long)n * classes, 256), 256, 0, (cudaStream_t)stream, 
        preds, (const long long*)labels, (const long long*)clip_ids, n, classes, num_clips, method, video_preds, (long long*)video_labels,
        (long long*)clip_count, num_videos);
    return check_launch("view_ensemble");
}

extern "C" int distb200_topk_correct(const float* video_preds, const int64_t* video_labels, int64_t num_videos, int32_t classes,
                                     const int32_t* ks, int32_t num_ks, int64_t* correct, void* stream) {
    if (num_videos == 0 || num_ks == 0) return 0;
    DISTB200_REQUIRE(video_preds && video_labels && ks && correct, "topk_correct: null pointer");
    DISTB200_LAUNCH(topk_correct_kernel, (unsigned)((num_videos + 7) / 8), 256, 0, (cudaStream_t)stream, video_preds, (const long long*)video_labels, num_videos,
                                                                                       classes, ks, num_ks, (long long*)correct);
    return check_launch("topk_correct");
}

extern "C" int distb200_rows_bcast(float* dst, int64_t row_stride, int64_t n_rows, int32_t cols, const float* table, int64_t period,
                                   int32_t accumulate, void* dst2_bf16, int64_t row_stride2, void* stream) {
    if (n_rows == 0 || cols == 0) return 0;
    DISTB200_REQUIRE(period >= 1, "rows_bcast: period must be >= 1");
    DISTB200_LAUNCH(rows_bcast_kernel, grid_for(n_rows * cols, 256), 256, 0, (cudaStream_t)stream, dst, row_stride, n_rows, cols, table, period, accumulate,
                    (bf16*)dst2_bf16, (long long)row_stride2);
    return check_launch("rows_bcast");
}

extern "C" int distb200_mean_rows(const float* src, int64_t row_stride, int32_t count, int64_t batch, int32_t cols, void* out,
                                  int32_t out_dtype, void* stream) {
    if (batch == 0 || cols == 0) return 0;
    DISTB200_REQUIRE(count >= 1, "mean_rows: count must be >= 1");
    cudaStream_t st = (cudaStream_t)stream;
    if (out_dtype == DISTB200_F32) DISTB200_LAUNCH(mean_rows_kernel<float>, grid_for(batch * cols, 256), 256, 0, st, src, row_stride, count, batch, cols, (float*)out);
    else DISTB200_LAUNCH(mean_rows_kernel<bf16>, grid_for(batch * cols, 256), 256, 0, st, src, row_stride, count, batch, cols, (bf16*)out);
    return check_launch("mean_rows");
}

extern "C" int distb200_cross_attention(const void* q, const void* kv, void* out, int32_t batch, int32_t keys, int32_t heads,
                                        int32_t dtype, void* stream) {
    if (batch == 0) return 0;
    DISTB200_REQUIRE(keys >= 1 && keys <= 2048, "cross_attention: keys=%d out of range", keys);
    const int wpb = 4;
    const long long items = (long long)batch * heads;
    const unsigned grid = (unsigned)((items + wpb - 1) / wpb);
    const size_t smem = (size_t)wpb * keys * sizeof(float);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == DISTB200_F32) DISTB200_LAUNCH(cross_attention_kernel<float>, grid, wpb * 32, smem, st, (const float*)q, (const float*)kv, (float*)out, batch, keys, heads);
    else DISTB200_LAUNCH(cross_attention_kernel<bf16>, grid, wpb * 32, smem, st, (const bf16*)q, (const bf16*)kv, (bf16*)out, batch, keys, heads);
    return check_launch("cross_attention");
}

extern "C" int distb200_class_head(const float* emb, const float* text_n, float scale, int32_t batch, int32_t embed_dim, int32_t classes,
                                   float* logits, float* probs, void* stream) {
    if (batch == 0) return 0;
    const size_t smem = (size_t)(embed_dim + classes + 32) * sizeof(float);
    DISTB200_REQUIRE(smem <= 48 * 1024, "class_head: E + C too large for one block (%zu bytes)", smem);
    DISTB200_LAUNCH(class_head_kernel, batch, 1024, smem, (cudaStream_t)stream, emb, text_n, scale, embed_dim, classes, logits, probs);
    return check_launch("class_head");
}

extern "C" int distb200_class_head_fused(const float* emb, const float* img, int32_t frames_per_clip, float w, const float* text_n, float scale,
                                         int32_t batch, int32_t embed_dim, int32_t classes, float* logits, float* probs, float* img_n,
                                         void* stream) {
    if (batch == 0) return 0;
    DISTB200_REQUIRE(emb && img && text_n && frames_per_clip >= 1, "class_head_fused: null pointer / no frames");
    const size_t smem = (size_t)(embed_dim + classes + 32) * sizeof(float);
    DISTB200_REQUIRE(smem <= 48 * 1024, "class_head_fused: E + C too large for one block (%zu bytes)", smem);
    DISTB200_LAUNCH(class_head_fused_kernel, batch, 1024, smem, (cudaStream_t)stream, emb, img, frames_per_clip, w, text_n, scale, embed_dim, classes,
                    logits, probs, img_n);
    return check_launch("class_head_fused");
}

extern "C" int distb200_embed_tokens(const int64_t* ids, const float* table, const float* pos, int64_t seqs, int32_t ctx, int32_t width,
                                     float* out, void* stream) {
    if (seqs == 0) return 0;
    DISTB200_REQUIRE(ids && table && pos && out, "embed_tokens: null pointer");
    DISTB200_REQUIRE(ctx >= 1 && width >= 4 && width % 4 == 0, "embed_tokens: ctx=%d width=%d (width must be a multiple of 4)", ctx, width);
    DISTB200_REQUIRE(((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(pos) | reinterpret_cast<uintptr_t>(out)) & 15) == 0, "embed_tokens: alignment");
    DISTB200_LAUNCH(embed_tokens_kernel, grid_for(seqs * ctx * (width / 4), 256), 256, 0, (cudaStream_t)stream, (const long long*)ids, table, pos,
                    (long long)seqs * ctx, ctx, width, out);
    return check_launch("embed_tokens");
}

extern "C" int distb200_gather_eot(const float* x, const int64_t* ids, int64_t seqs, int32_t ctx, int32_t width, float* out, void* stream) {
    if (seqs == 0) return 0;
    DISTB200_REQUIRE(x && ids && out, "gather_eot: null pointer");
    DISTB200_REQUIRE(ctx >= 1 && width >= 1, "gather_eot: bad sizes");
    DISTB200_LAUNCH(gather_eot_kernel, (unsigned)seqs, 128, 0, (cudaStream_t)stream, x, (const long long*)ids, ctx, width, out);
    return check_launch("gather_eot");
}

extern "C" int distb200_row_stats_finalize(const float* stat_partials, int64_t rows, int32_t cols, float eps, float* stats, void* stream) {
    if (rows == 0) return 0;
    DISTB200_REQUIRE(stat_partials && stats && cols >= 1, "row_stats_finalize: bad arguments");
    DISTB200_REQUIRE((reinterpret_cast<uintptr_t>(stat_partials) & 15) == 0 && (reinterpret_cast<uintptr_t>(stats) & 7) == 0, "row_stats_finalize: alignment");
    static_assert(DISTB200_STAT_SLOTS == 16, "the finalize kernel reads eight float4 per row");
    DISTB200_LAUNCH(row_stats_finalize_kernel, (unsigned)((rows * 8 + 255) / 256), 256, 0, (cudaStream_t)stream, reinterpret_cast<const float4*>(stat_partials),
                    (long long)rows, 1.0f / (float)cols, eps, reinterpret_cast<float2*>(stats));
    return check_launch("row_stats_finalize");
}
