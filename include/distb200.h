/*
 * distb200 C ABI - the CUDA side of the DiST video forward path for NVIDIA B200 (sm_100a).
 *
 * The reference (alibaba-mmai-research/DiST) is pure PyTorch: every operator on its forward path is
 * an ATen / cuBLASLt / cuDNN / SDPA library call and there is no native interface to mirror
 * (SURVEY.md section 0.1, 2.2).  This header therefore *defines* the boundary a maintainer of the
 * reference would bind (ctypes stub shown in INTEGRATION.md): each entry point below names the
 * reference call sites it replaces.
 *
 * Conventions
 *   - plain C: raw device pointers, sizes and strides in ELEMENTS, a CUDA stream handle passed as
 *     void* (cudaStream_t); no torch types;
 *   - every function returns 0 on success; on failure a non-zero code is returned, nothing is
 *     launched and distb200_last_error() describes the problem (thread-local string);
 *   - the caller owns all memory.  The library allocates nothing per call and never synchronises;
 *   - one calling thread per device (the reference is one process per GPU, utils/launcher.py:29-34).
 *   - tensors: "act" dtype is DISTB200_BF16 on the tensor-core path and DISTB200_F32 on the fp32 parity
 *     path; residual streams, biases, LayerNorm parameters and statistics are always fp32.
 */
#ifndef DISTB200_H_
#define DISTB200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DISTB200_VERSION 100
#define DISTB200_MAX_TAPS 9

enum { DISTB200_F32 = 0, DISTB200_BF16 = 1 };
enum { DISTB200_ACT_NONE = 0, DISTB200_ACT_QUICKGELU = 1 };   /* x * sigmoid(1.702 x), clip.py:199-201 */
enum { DISTB200_IMPL_AUTO = 0, DISTB200_IMPL_SIMT = 1, DISTB200_IMPL_TCGEN05 = 2,
       DISTB200_IMPL_TCGEN05_1CTA = 3, DISTB200_IMPL_TCGEN05_2CTA = 4 };   /* 3 / 4 force single-CTA / CTA-pair (cta_group::2) tiles */

int         distb200_version(void);
int         distb200_arch(void);          /* 100 : compiled for sm_100a only */
const char* distb200_last_error(void);

/* ------------------------------------------------------------------------------------------------
 * distb200_gemm - tapped, grouped GEMM with fused epilogue.  One entry point covers every dense
 * contraction of the path:
 *   nn.Linear            clip.py:157-161 (c_fc / c_proj), nn.MultiheadAttention in/out projections
 *                        (clip.py:155,168), dist.py:26-28 (ffn), :97 (linear_fuse), :183 (input_linears),
 *                        :125-137 (token MLPs), :192,:195 (proj_spatial_cls_token, proj)
 *   nn.Conv2d patch embed    clip.py:243,271            (A = patchified frames)
 *   nn.Conv3d temporal_stem  dist.py:178-181,225        (5 temporal taps over patchified frames)
 *   nn.Conv3d (kt,1,1)       dist.py:55, :31            (kt taps = row shifts inside a clip)
 *   nn.Conv3d (1,3,3)        dist.py:57                 (9 taps on the g x g frame grid, image mode)
 *   nn.Conv3d (alpha,1,1)/stride alpha   dist.py:75     (alpha taps over a frame group)
 *   nearest upsample + add   dist.py:105,231            (epilogue row replication)
 *
 * Definition.  Output rows are indexed by (group gi in [0,groups), row r in [0,rows_per_group)).
 * A is a logical 4-D tensor with element (c0,c1,c2,c3) at a + sum(ci * a_stride[i]), a_stride[0] == 1,
 * reading as zero outside [0,a_dim[i]).  For tap j with offsets (d1,d2,d3) = tap_off[j]:
 *     img_w == 0 :  A_j(gi,r,k) = A[k, r + d1, d2 (+ gi if group_dim == 2), d3 (+ gi if group_dim == 3)]
 *     img_w  > 0 :  A_j(gi,r,k) = A[k, r % img_w + d1, r / img_w + d2, gi + d3]
 *     acc(gi,r,n) = sum_j sum_k A_j(gi,r,k) * B[j*b_tap_stride + n*ldb + k]
 *     v           = acc + bias[n] + res[res_row(gi,r,rep)*ld_res + n]      (each term optional)
 *     out[out_row(gi,r,rep)*ld_out + n] = act(v)   for rep in [0,out_rep)
 *     out_row(gi,r,rep) = gi*out_gstride + out_roff + r + rep*out_rep_stride      (res_row alike)
 * out2, when non-null, receives the same values at the same rows in out2_dtype (the low-precision copy
 * that the next GEMM consumes as its A operand).  out may alias res.
 *
 * dtype selects the arithmetic: DISTB200_BF16 = bf16 operands, fp32 accumulation on tcgen05 tensor
 * cores (TMA-fed, TMEM accumulators); DISTB200_F32 = fp32 FFMA.  impl = DISTB200_IMPL_SIMT forces the
 * SIMT kernel for either dtype (debug / cross-check).
 * tcgen05 constraints: n % 16 == 0, a_stride[1..3] and ldb multiples of 8 elements, 16-byte aligned
 * bases, img_w <= 128.
 */
typedef struct distb200_gemm_desc {
    const void* a;
    const void* b;
    int32_t dtype;                 /* dtype of a and b */
    int32_t impl;
    int64_t a_dim[4];
    int64_t a_stride[4];
    int32_t img_w;
    int32_t num_taps;
    int32_t tap_off[DISTB200_MAX_TAPS][3];
    int64_t ldb;
    int64_t b_tap_stride;
    int32_t n;
    int32_t k;
    int64_t groups;
    int64_t rows_per_group;
    const float* bias;
    const float* res;
    int64_t ld_res, res_gstride, res_roff, res_rep_stride;
    void*   out;
    int32_t out_dtype;
    int32_t out_rep;
    int64_t ld_out, out_gstride, out_roff, out_rep_stride;
    void*   out2;
    int32_t out2_dtype;
    int32_t act;
    int64_t ld_out2;
    int32_t block_n;               /* tcgen05 N tile, 0 = choose */
    int32_t group_dim;             /* 2 (default when 0) or 3: which A coordinate the group index adds to */
} distb200_gemm_desc;

int distb200_gemm(const distb200_gemm_desc* desc, void* stream);

/* LayerNorm over the last dim (eps, biased variance; clip.py:181-187) of x = in1 (+ in2[row % in2_period]),
 * fp32 statistics.  Writes y1 = xhat*g1+b1 and, when y2 != NULL, y2 = xhat*g2+b2 (two affine views of the
 * same statistics: dist.py:43-45 ln / ln_temporal).  y may alias in1 when out_dtype is fp32.
 * Replaces: ln_pre/ln_1/ln_2 (clip.py:276,172,173), dist.py:65 (LN over Ct), :43-45, clip.py:147, dist.py:145,160,243. */
int distb200_layernorm(const float* in1, int64_t ld_in1, const float* in2, int64_t ld_in2, int64_t in2_period,
                       int64_t rows, int32_t cols, float eps,
                       const float* g1, const float* b1, void* y1, int64_t ld_y1,
                       const float* g2, const float* b2, void* y2, int64_t ld_y2,
                       int32_t out_dtype, void* stream);

/* Multi-head self attention over the N tokens of each frame, head dim 64, no mask
 * (nn.MultiheadAttention inside ResidualAttentionBlockMid, clip.py:155,166-168).
 * qkv [frames, tokens, 3*heads*64] (q | k | v column blocks, as in_proj produces them), out [frames, tokens, heads*64]. */
int distb200_attention(const void* qkv, void* out, int32_t frames, int32_t tokens, int32_t heads,
                       int32_t dtype, int32_t impl, void* stream);

/* Single-query cross attention (CrossAttentionBlockGenral inside the ada-pooling head, clip.py:139-147,
 * dist.py:144,158): q [batch, heads*64]; kv [batch, keys, 2*heads*64] (k | v); out [batch, heads*64]. */
int distb200_cross_attention(const void* q, const void* kv, void* out, int32_t batch, int32_t keys,
                             int32_t heads, int32_t dtype, void* stream);

/* Cut frames into patch rows: video fp32 [clips, 3, T, H, W] -> out[(clip*n_sel + i), patch, (c, y, x)] with row
 * pitch ld_out, for the frames first_frame + i*frame_step, i < n_sel; pad columns [3*p*p, ld_out) are zeroed.
 * Feeds the patch-embedding conv (clip.py:271, on the kept frames of :281-284) and the temporal stem (dist.py:225);
 * replaces the permute+reshape copies of backbone.py:232-233 and dist.py:225. */
int distb200_patchify(const float* video, void* out, int32_t clips, int32_t T, int32_t H, int32_t W, int32_t p,
                      int32_t first_frame, int32_t frame_step, int32_t n_sel, int64_t ld_out, int32_t out_dtype,
                      void* stream);

/* dst[i*row_stride + c] = (accumulate ? dst[...] : 0) + table[(i % period)*cols + c], i < n_rows (fp32).
 * Class-token rows (clip.py:274), per-frame cls tokens of dist.py:84, broadcast of the aggregated tokens (dist.py:237-238). */
int distb200_rows_bcast(float* dst, int64_t row_stride, int64_t n_rows, int32_t cols, const float* table,
                        int64_t period, int32_t accumulate, void* stream);

/* out[b, :] = mean_i src[(b*count + i)*row_stride + :], i < count  (dist.py:243 mean over the sparse frames). */
int distb200_mean_rows(const float* src, int64_t row_stride, int32_t count, int64_t batch, int32_t cols,
                       void* out, int32_t out_dtype, void* stream);

/* Cosine-similarity class scores and head: logits = scale * emb/|emb| . text_n^T (text_n pre-normalised rows),
 * probs = softmax(logits) (clip.py:511-518, base_blocks.py:579-585).  Either output may be NULL. */
int distb200_class_head(const float* emb, const float* text_n, float scale, int32_t batch, int32_t embed_dim,
                        int32_t classes, float* logits, float* probs, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DISTB200_H_ */
