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
#define DISTB200_STAT_SLOTS 16

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
    /* LayerNorm folded into the GEMM: with A = the raw (un-normalised) rows x, B = W * gamma (per input column) and
     * ln_stats[row] = (mean, rstd) of row gi*rows_per_group + r (distb200_row_stats), ln_wsum[n] = sum_k B[n][k],
     * bias[n] = sum_k W[n][k] beta[k] + b[n], the epilogue computes  v = rstd * (acc - mean * ln_wsum[n]) + bias[n],
     * which equals LayerNorm(x) W^T + b (clip.py:172-173 followed by in_proj / c_fc) without materialising LayerNorm(x).
     * Both NULL = off.  Needs a bf16 `out` only (no out2 / res / taps). */
    const float* ln_stats;
    const float* ln_wsum;
    /* Producer side of the folded LayerNorm: when non-NULL, the epilogue also emits, for every output row, the sum and the sum of
     * squares of the bf16 values it writes to out2 - as DISTB200_STAT_SLOTS float2 slots per row,
     *     stat_partials[((gi*rows_per_group + r) * DISTB200_STAT_SLOTS + slot) * 2 + {0, 1}],
     * one slot per (column tile, epilogue warp phase), each written by exactly one warp with plain stores (deterministic; slots
     * the launch does not own are left untouched, so the buffer starts zeroed and is used by launches of one shape).
     * distb200_row_stats_finalize reduces the slots to the (mean, rstd) pairs that ln_stats takes: the stand-alone pass over
     * the bf16 copy (distb200_row_stats) disappears.  Needs fp32 out + bf16 out2 + res, no activation, out_rep == 1, tcgen05. */
    float* stat_partials;
    /* The activation applies to the output columns act_from <= n < act_to only (act_to 0 = up to n): lets one GEMM produce an
     * activated and a linear block side by side (IntegrationNetwork: QuickGELU(ffn.c_fc) | temporal_ffn.c_fc1, dist.py:40-45). */
    int32_t act_from;
    /* Independent placement of out2 (0 = the rows of `out`): with out2_gdiv = g > 0 the copy of output row (gi, r) goes to row
     * (gi / g) * out2_gstride + out2_roff + r, column offset (gi % g) * out2_cstep.  Lets the (1,3,3) convolution of
     * TemporalNet drop the bf16 copy of frame alpha*ti + k next to token row 1 + r of sparse frame ti - the K-concatenated
     * operand [tap | temporal rows | one-hot] from which input_linear, the temporal->integration convolution and the cls
     * token (dist.py:229,80-86) come out of ONE GEMM.  Needs out_rep == 1. */
    int32_t out2_gdiv;
    int32_t out2_cstep;
    int32_t act_to;
    int64_t out2_gstride;
    int64_t out2_roff;
} distb200_gemm_desc;

int distb200_gemm(const distb200_gemm_desc* desc, void* stream);

/* ------------------------------------------------------------------------------------------------
 * distb200_temporalnet - one fused TemporalNet block (models/module_zoo/branches/dist.py:48-65) on the
 * channels-last temporal stream, with the integration->temporal add of the PREVIOUS DiST layer
 * (dist.py:231, nearest upsample over alpha dense frames, dist.py:105) folded into its input:
 *
 *     xe[b,tau,p,:] = x[b,tau,p,:] + (u ? u[b, tau / alpha, p, :] : 0)                    p = r*g + c, r,c < g
 *     y             = LayerNorm_C(xe) * ln_g + ln_b                     (eps, biased variance; dist.py:65, clip.py:181-187)
 *     z[b,tau,p,:]  = q( sum_{k<3}   y[b, tau+k-1, p, :] . w1[k]^T + b1 )                 zero outside 0 <= tau+k-1 < T
 *     o[b,tau,p,:]  = sum_{i,j<3} z[b, tau, (r+i-1)*g + (c+j-1), :] . w2[3i+j]^T + b2      zero outside the g x g frame
 *     out[b,tau,p,:] = q( xe + o )                                      q(v) = v * sigmoid(1.702 v), clip.py:199-201
 *
 * w1 [3][C][C] / w2 [9][C][C] are the per-tap K-major matrices distb200_gemm takes for the same convolutions
 * (w1[k][n][c] = temporal_net.c_fc1.weight[n,c,k,0,0], w2[3i+j][n][c] = c_fc2.weight[n,c,0,i,j]).
 * out (fp32, may be NULL; must not alias x) receives the new stream; out2 (bf16, may be NULL) the operand copy, placed
 * like distb200_gemm_desc.out2 with groups = frames: with out2_gdiv = a > 0 the copy of frame f = b*T + tau, position p
 * goes to row (f / a) * out2_gstride + out2_roff + p, column (f % a) * out2_cstep (+ channel), row pitch ld_out2;
 * out2_gdiv = 0 places it at row f * g*g + p.
 *
 * One launch replaces distb200_layernorm + two tapped distb200_gemm calls (+ the read-modify-write of the i2t add):
 * a CTA pair (cta_group::2, both weight sets resident in shared memory, split between the pair) walks bands of image rows
 * frame by frame; LayerNorm'd frames live in a three-slot ring of un-swizzled K-major shared-memory tiles, z stays in
 * shared memory in a zero-padded (g+2)-pitch layout so that the nine spatial taps are nine row-shifted descriptors of ONE
 * tile, both accumulators live in TMEM.  bf16 operands / fp32 accumulation only (dtype = DISTB200_BF16); the fp32 parity
 * path composes the same block from distb200_layernorm + distb200_gemm.
 * Constraints: C in {32, 64, 96}; g <= 60; 16-byte aligned pointers; ld_out2 and out2_cstep multiples of 8. */
typedef struct distb200_temporalnet_desc {
    const float* x;
    const void* u;                 /* optional addend [clips, frames / alpha, g*g, C] in `dtype` (bf16: the i2t GEMM's output) */
    int32_t alpha;                 /* >= 1 (ignored when u is NULL) */
    int32_t dtype;                 /* operand dtype of w1 / w2 / out2: DISTB200_BF16 */
    const float* ln_g;
    const float* ln_b;
    const void* w1;
    const float* b1;
    const void* w2;
    const float* b2;
    float* out;
    void*  out2;
    int64_t ld_out2;
    int32_t out2_gdiv;
    int32_t out2_cstep;
    int64_t out2_gstride;
    int64_t out2_roff;
    int32_t clips, frames, grid, channels;
    float   eps;
    int32_t max_ctas;              /* 0 = one CTA per SM; otherwise an upper bound on the (even) CTA count */
} distb200_temporalnet_desc;

int distb200_temporalnet(const distb200_temporalnet_desc* desc, void* stream);

/* LayerNorm over the last dim (eps, biased variance; clip.py:181-187) of x = in1 (+ in2[row % in2_period]),
 * fp32 statistics.  Writes y1 = xhat*g1+b1 and, when y2 != NULL, y2 = xhat*g2+b2 (two affine views of the
 * same statistics: dist.py:43-45 ln / ln_temporal).  y may alias in1 when out_dtype is fp32.
 * Replaces: ln_pre/ln_1/ln_2 (clip.py:276,172,173), dist.py:65 (LN over Ct), :43-45, clip.py:147, dist.py:145,160,243. */
int distb200_layernorm(const float* in1, int64_t ld_in1, const float* in2, int64_t ld_in2, int64_t in2_period,
                       int64_t rows, int32_t cols, float eps,
                       const float* g1, const float* b1, void* y1, int64_t ld_y1,
                       const float* g2, const float* b2, void* y2, int64_t ld_y2,
                       int32_t out_dtype, void* stream);

/* (mean, rstd) of every row of a [rows, cols] matrix (row pitch ld), biased variance + eps as LayerNorm uses them
 * (clip.py:181-187): stats[2*row] = mean, stats[2*row + 1] = 1 / sqrt(var + eps).  Feeds distb200_gemm_desc.ln_stats. */
int distb200_row_stats(const void* x, int32_t dtype, int64_t ld, int64_t rows, int32_t cols, float eps, float* stats, void* stream);

/* stats[row] = (mean, rstd) from the DISTB200_STAT_SLOTS (sum, sum of squares) slots per row that distb200_gemm emitted
 * (distb200_gemm_desc.stat_partials), summed in slot order: mean = S/cols, rstd = 1/sqrt(max(Q/cols - mean^2, 0) + eps). */
int distb200_row_stats_finalize(const float* stat_partials, int64_t rows, int32_t cols, float eps, float* stats, void* stream);

/* Multi-head self attention over the N tokens of each frame, head dim 64, no mask
 * (nn.MultiheadAttention inside ResidualAttentionBlockMid, clip.py:155,166-168).
 * qkv [frames, tokens, 3*heads*64] (q | k | v column blocks, as in_proj produces them), out [frames, tokens, heads*64]. */
int distb200_attention(const void* qkv, void* out, int32_t frames, int32_t tokens, int32_t heads,
                       int32_t dtype, int32_t impl, void* stream);

/* Causal multi-head self attention of the CLIP text transformer: token i attends to the tokens j <= i (the additive
 * upper-triangular -inf mask built at clip.py:404-410 and passed to nn.MultiheadAttention at clip.py:122-124).  Same layout as
 * distb200_attention.  FFMA kernel for both dtypes: the label set is encoded once and cached (clip.py:436-452). */
int distb200_attention_causal(const void* qkv, void* out, int32_t seqs, int32_t tokens, int32_t heads, int32_t dtype,
                              void* stream);

/* Text tower input rows (clip.py:420-421): out[s*ctx + i, :] = table[ids[s*ctx + i], :] + pos[i, :], fp32.  ids int64 on the
 * device, every id in [0, vocab) (checked by the caller: nn.Embedding raises on the host). */
int distb200_embed_tokens(const int64_t* ids, const float* table, const float* pos, int64_t seqs, int32_t ctx, int32_t width,
                          float* out, void* stream);

/* Text tower output rows (clip.py:429): out[s, :] = x[s*ctx + argmax_i ids[s, i], :] - the end-of-text token has the largest
 * id of its sequence; ties resolve to the first position like torch.argmax. */
int distb200_gather_eot(const float* x, const int64_t* ids, int64_t seqs, int32_t ctx, int32_t width, float* out, void* stream);

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

/* The same patch rows straight from decoded frames: frames uint8 [clips, T, H, W, 3] (the layout before
 * transforms.ToTensorVideo, dataset/base/ssv2.py:137), with ToTensorVideo + NormalizeVideo fused in:
 * value = ((float)u8 / 255 - mean[c]) / std[c]  (DATA.MEAN / DATA.STD, ssv2.py:139-143).  mean3 / std3 are HOST pointers to
 * three floats.  A quarter of the bytes of the float clip cross PCIe / HBM. */
int distb200_patchify_u8(const uint8_t* frames, void* out, int32_t clips, int32_t T, int32_t H, int32_t W, int32_t p,
                         int32_t first_frame, int32_t frame_step, int32_t n_sel, int64_t ld_out, int32_t out_dtype,
                         const float* mean3, const float* std3, void* stream);

/* dst[i*row_stride + c] = (accumulate ? dst[...] : 0) + table[(i % period)*cols + c], i < n_rows (fp32).
 * Class-token rows (clip.py:274), per-frame cls tokens of dist.py:84, broadcast of the aggregated tokens (dist.py:237-238).
 * dst2_bf16 (optional): a bf16 copy of the resulting rows at row pitch row_stride2 (the GEMM operand copy of the stream). */
int distb200_rows_bcast(float* dst, int64_t row_stride, int64_t n_rows, int32_t cols, const float* table,
                        int64_t period, int32_t accumulate, void* dst2_bf16, int64_t row_stride2, void* stream);

/* out[b, :] = mean_i src[(b*count + i)*row_stride + :], i < count  (dist.py:243 mean over the sparse frames). */
int distb200_mean_rows(const float* src, int64_t row_stride, int32_t count, int64_t batch, int32_t cols,
                       void* out, int32_t out_dtype, void* stream);

/* Cosine-similarity class scores and head: logits = scale * emb/|emb| . text_n^T (text_n pre-normalised rows),
 * probs = softmax(logits) (clip.py:511-518, base_blocks.py:579-585).  Either output may be NULL. */
int distb200_class_head(const float* emb, const float* text_n, float scale, int32_t batch, int32_t embed_dim,
                        int32_t classes, float* logits, float* probs, void* stream);

/* The same head with the zero-shot / prediction-fusion branch (clip.py:519-527, TEST.ZEROSHOT.ENABLE) fused in: img
 * [batch * frames_per_clip, embed_dim] are the per-frame CLIP image embeddings ln_post(class token) @ visual.proj (clip.py:291-298),
 *     logits[b] = w * scale * cos(emb[b], text) + (1 - w) * mean_f scale * cos(img[b, f], text)
 * (w = 0.5 in the reference unless its gating parameter is enabled), probs = softmax(logits).  img_n (optional) receives the
 * L2-normalised image embeddings, which is what the reference returns as img_logits on this branch (clip.py:520,532). */
int distb200_class_head_fused(const float* emb, const float* img, int32_t frames_per_clip, float w, const float* text_n, float scale,
                              int32_t batch, int32_t embed_dim, int32_t classes, float* logits, float* probs, float* img_n,
                              void* stream);


/* Multi-view test ensemble on the device (TestMeter.update_stats, utils/meters.py:83-115, without the per-clip Python loop
 * and the .cpu() synchronisation of runs/test.py:136-145): for every clip i, video = clip_ids[i] / num_clips;
 * video_preds[video] += preds[i] (method 0, "sum") or = max(video_preds[video], preds[i]) (method 1, "max"; scores >= 0);
 * video_labels[video] = labels[i]; clip_count[video] += 1.  labels / clip_ids / counts are int64 device arrays. */
int distb200_view_ensemble(const float* preds, const int64_t* labels, const int64_t* clip_ids, int32_t n, int32_t classes,
                           int32_t num_clips, int32_t method, float* video_preds, int64_t* video_labels, int64_t* clip_count,
                           int64_t num_videos, void* stream);

/* correct[j] += #videos whose label is among the ks[j] best scores (utils/metrics.py topks_correct; TestMeter.finalize_metrics,
 * utils/meters.py:135-163).  ks is a DEVICE array of num_ks ints. */
int distb200_topk_correct(const float* video_preds, const int64_t* video_labels, int64_t num_videos, int32_t classes,
                          const int32_t* ks, int32_t num_ks, int64_t* correct, void* stream);

/* ================================================================================================
 * Fine-tuning step (runs/train.py:97-112): the reference gets the backward of
 * the DiST branches from autograd (ATen kernels), the loss from models/utils/losses.py:20-31 and the update
 * from torch.optim.AdamW (models/utils/optimizer.py:67-73).  The entry points below are the operators that
 * replace those calls.  Gradients with respect to parameters are always ACCUMULATED (+=) into fp32 buffers
 * the caller zeroes once per step, so that tied uses and split reductions compose.
 * ================================================================================================ */

/* Input gradient of distb200_gemm needs no entry point of its own: it is distb200_gemm with a = dY, the tap
 * offsets negated and b = the per-tap transposed weights (distb200_pack_weight writes them). */

/* Weight gradient of distb200_gemm (nn.Linear / nn.Conv3d weight.grad):
 *     dw[j*dw_tap_stride + n*ld_dw + k] += sum_{gi, r} dy[(gi*dy_gstride + dy_roff + r)*ld_dy + n] * A_j(gi, r, k)
 * with A_j(gi, r, k) exactly as defined for distb200_gemm (same a_dim / a_stride / tap_off / img_w / group_dim),
 * j < num_taps, n < n, k < k.  x and dy have dtype `dtype` (bf16: tcgen05 tensor cores with both operands
 * MN-major, split over row blocks with fp32 reductions into dw; fp32: FFMA).  impl as for distb200_gemm. */
typedef struct distb200_wgrad_desc {
    const void* x;
    const void* dy;
    int32_t dtype;
    int32_t impl;
    int64_t a_dim[4];
    int64_t a_stride[4];
    int32_t img_w;
    int32_t num_taps;
    int32_t tap_off[DISTB200_MAX_TAPS][3];
    int32_t n;
    int32_t k;
    int64_t groups;
    int64_t rows_per_group;
    int32_t group_dim;
    int32_t reserved;
    int64_t ld_dy, dy_gstride, dy_roff;
    float*  dw;
    int64_t ld_dw, dw_tap_stride;
} distb200_wgrad_desc;

int distb200_gemm_wgrad(const distb200_wgrad_desc* desc, void* stream);

/* QuickGELU on a saved pre-activation (training keeps z for the backward): y = z * sigmoid(1.702 z) written as fp32
 * (y_f32) and / or in lp_dtype (y_lp); either may be NULL.  clip.py:199-201. */
int distb200_quickgelu(const void* z, int32_t z_dtype, float* y_f32, void* y_lp, int32_t lp_dtype, int64_t n, void* stream);

/* dz = dy * d/dz [z sigmoid(1.702 z)] = dy * s * (1 + 1.702 z (1 - s)), s = sigmoid(1.702 z). */
int distb200_quickgelu_bwd(const void* dy, int32_t dy_dtype, const void* z, int32_t z_dtype, float* dz_f32, void* dz_lp,
                           int32_t lp_dtype, int64_t n, void* stream);

/* dst = (dst_dtype) src, n contiguous elements (the low-precision operand copy of an fp32 gradient). */
int distb200_cast(const float* src, void* dst, int32_t dst_dtype, int64_t n, void* stream);

/* dst[g*inner + i] = sum_{k < alpha} src[(g*alpha + k)*inner + i]: gradient of the nearest upsample along time
 * (dist.py:105) - each sparse frame collects its alpha dense frames. */
int distb200_group_sum(const void* src, int32_t src_dtype, int64_t groups, int32_t alpha, int64_t inner, void* dst,
                       int32_t dst_dtype, void* stream);

/* out[(g % period)*cols + c] += sum_{r < rows_per_group} src[(g*gstride + roff + r)*ld + c]   (bias.grad, cls_token.grad,
 * positional_embedding.grad, aggregated token grads).  out2, when non-NULL (period == 1 only), receives the same sums (two
 * biases fed by the same gradient).  cols and ld must be multiples of 4. */
int distb200_colsum(const void* src, int32_t src_dtype, int64_t ld, int64_t groups, int64_t rows_per_group, int64_t gstride,
                    int64_t roff, int64_t period, int32_t cols, float* out, float* out2, void* stream);

/* Backward of distb200_layernorm for x = in1 (+ in2[row % in2_period]); statistics are recomputed from x.
 *   dx_row = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dy1 * g1 (+ dy2 * g2)
 *   dx (= or +=, `accumulate`) add[row] + dx_row;   dx_lp, when non-NULL, receives the final dx in lp_dtype
 *   dg1 += sum_rows dy1 * xhat, db1 += sum_rows dy1 (dg2 / db2 alike; all optional). */
int distb200_layernorm_bwd(const float* in1, int64_t ld_in1, const float* in2, int64_t ld_in2, int64_t in2_period,
                           int64_t rows, int32_t cols, float eps,
                           const float* g1, const void* dy1, int64_t ld_dy1,
                           const float* g2, const void* dy2, int64_t ld_dy2, int32_t dy_dtype,
                           const float* add, int64_t ld_add,
                           float* dx, int64_t ld_dx, int32_t accumulate,
                           void* dx_lp, int64_t ld_dx_lp, int32_t lp_dtype,
                           float* dg1, float* db1, float* dg2, float* db2, void* stream);

/* Backward of distb200_cross_attention (probabilities recomputed): dq [batch, heads*64], dkv [batch, keys, 2*heads*64]. */
int distb200_cross_attention_bwd(const void* q, const void* kv, const void* d_out, void* dq, void* dkv, int32_t batch,
                                 int32_t keys, int32_t heads, int32_t dtype, void* stream);

/* Train-mode head + SoftTargetCrossEntropy (base_blocks.py:579-585 without softmax, losses.py:20-31) and its gradient:
 * logits = scale * emb/|emb| . text_n^T;  loss[0] += mean_b sum_c -target * log_softmax(logits);
 * d_emb = d loss / d emb.  logits may be NULL. */
int distb200_softce_head(const float* emb, const float* text_n, float scale, const float* target, int32_t batch,
                         int32_t embed_dim, int32_t classes, float* logits, float* loss, float* d_emb, void* stream);

/* torch.optim.AdamW on n contiguous fp32 elements (decoupled decay, bias correction with `step` >= 1); the gradient
 * is multiplied by grad_scale first (1 / world_size after a SUM all-reduce). */
int distb200_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int32_t step, float grad_scale, void* stream);

/* Operand copies of a master weight w [batch][n][k] fp32: out[b][n][k] (row pitch ld_out, columns [k, ld_out) zeroed)
 * and out_t[b][k][n] (row pitch ld_out_t) in `dtype`; either may be NULL. */
int distb200_pack_weight(const float* w, int64_t batch, int32_t n, int32_t k, void* out, int64_t ld_out, void* out_t,
                         int64_t ld_out_t, int32_t dtype, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DISTB200_H_ */
