"""Forward engine of the DiST video path on the distb200 CUDA library.

``DistEngine`` packs a reference-format ``state_dict`` once (bf16 or fp32 operands, convolution weights
re-laid as per-tap K-major matrices), allocates every activation buffer for a fixed clip count, and
*plans* the whole forward as a flat list of prepared C calls (``ops.Call``).  Running the model is
replaying that list on a stream - directly, or as a captured CUDA graph.  No torch operator runs on
the hot path; PyTorch only provides device memory and streams.

Data layout (HBM), b clips, T dense / t = T/alpha sparse frames, N = P+1 tokens:
  ViT stream            h      [b*t*N, D]   fp32   frame-major tokens, frame = clip*t + ti, token 0 = class
  temporal stream       xT     [b*T*P, Ct]  fp32   channels-last (clip, frame, row, col)
  integration stream    mid    [b*t*N, Ci]  fp32   (becomes ``upd`` in place), ``res`` alike
  GEMM operands         bf16 (or fp32 on the parity path) copies written by the producing kernel's epilogue

Semantics follow SURVEY.md section 8(a'); the reference lines are cited next to each step.
"""

import math
import os

import torch

from . import ops
from .arch import DistArch


def _pad8(n):
    return (n + 7) // 8 * 8


def block_n_by_waves(rows, n, candidates, units=74):
    """Column tile of a narrow, epilogue-bound GEMM chosen by wave quantisation: `units` CTA pairs work through
    ceil(rows / 256) * (n / block_n) tiles; ties go to the wider tile."""
    best, best_eff = 0, 0.0
    for bn in sorted(candidates, reverse=True):
        if n % bn:
            continue
        tiles = (rows + 255) // 256 * (n // bn)
        eff = tiles / float((tiles + units - 1) // units * units)
        if eff > best_eff + 0.02:
            best, best_eff = bn, eff
    return best


def kcat_layout(arch):
    """Column layout of the K-concatenated DiST operand buffer (bf16 path), every block on a 128-byte boundary:
    [tap D | hidden activations of the PREVIOUS IntegrationNetwork (ffn Ih | c_fc1 Cm | temporal Cm) | alpha temporal rows |
     one-hot(ti) | bf16 upd Ci]."""
    r64 = lambda n: (n + 63) // 64 * 64
    c_h = r64(arch.width)
    c_x = r64(c_h + arch.integration_hidden + 2 * arch.integration_temporal_hidden)
    c_oh = r64(c_x + arch.alpha * arch.temporal_dim)
    c_u = r64(c_oh + arch.sparse_frames)
    return dict(c_h=c_h, c_x=c_x, c_oh=c_oh, c_u=c_u, tw=r64(c_u + arch.integration_dim))


def plan_branches(names, sel):
    """Split a planned forward (call names in plan order) into the ViT branch (0) and the DiST branch (1) of the CUDA graph.

    Returns ``(branch of every call, {call index: [indices of calls of the OTHER branch that must have completed]})``:
      * the first DiST call (the stem) follows the patch-row kernels;
      * ``dist.input_linear`` of DiST layer i follows ``vit.fc2`` of ViT block sel[i] (it reads that block's bf16 tap);
      * ``vit.fc2`` of block l follows ``dist.input_linear`` of the DiST layer fed by block l-2: the tap buffers alternate by
        layer, so block l overwrites the buffer block l-2 filled;
      * the tail follows ``dist.cls_mean``, which runs on the ViT branch right behind the last tapped block (it reads the
        fp32 stream that later blocks keep modifying).
    """
    branch, deps = [], {}
    fc2_of, inlin_of = {}, {}
    first_dist = last_patchify = cls_mean = None
    for k, n in enumerate(names):
        on_v = n.startswith("patchify") or n.startswith("vit.") or n == "dist.cls_mean"
        branch.append(0 if on_v else 1)
        if n.startswith("patchify"):
            last_patchify = k
        elif n == "vit.fc2":
            layer = len(fc2_of)
            fc2_of[layer] = k
            if layer - 2 in sel and sel.index(layer - 2) in inlin_of:
                deps.setdefault(k, []).append(inlin_of[sel.index(layer - 2)])
        elif n == "dist.input_linear":
            i = len(inlin_of)
            inlin_of[i] = k
            deps.setdefault(k, []).append(fc2_of[sel[i]])
        elif n == "dist.cls_mean":
            cls_mean = k
        elif n == "tail.proj_spatial_cls" and cls_mean is not None:
            deps.setdefault(k, []).append(cls_mean)
        if not on_v and first_dist is None:
            first_dist = k
            if last_patchify is not None:
                deps.setdefault(k, []).append(last_patchify)
    return branch, deps


class PackedWeights:
    """Device-resident, kernel-ready weights (reference key names in the comments)."""

    def __init__(self, sd, arch: DistArch, device, act_dtype):
        self.arch, self.device, self.adt = arch, device, act_dtype
        a = arch
        f32 = lambda x: x.detach().to(device=device, dtype=torch.float32).contiguous()
        op = lambda x: x.detach().to(device=device, dtype=torch.float32).contiguous().to(act_dtype)

        def padk(w2d, kp):
            n, k = w2d.shape
            if k == kp:
                return op(w2d)
            out = torch.zeros(n, kp, dtype=torch.float32)
            out[:, :k] = w2d.detach().float().cpu()
            return op(out)

        D, L, p = a.width, a.layers, a.patch
        self.kp = _pad8(3 * p * p)
        self.conv1_w = padk(sd["visual.conv1.weight"].reshape(D, -1), self.kp)                 # clip.py:243
        pos = sd["visual.positional_embedding"].float()
        self.pos = f32(pos)                                                                    # [N, D]
        self.cls_row = f32((sd["visual.class_embedding"].float() + pos[0]).reshape(1, D))      # clip.py:274-275
        self.ln_pre = (f32(sd["visual.ln_pre.weight"]), f32(sd["visual.ln_pre.bias"]))
        # per-frame CLIP image embeddings (clip.py:291-298): only the zero-shot / prediction-fusion branch and `img_logits` read them
        self.vis_ln_post = (f32(sd["visual.ln_post.weight"]), f32(sd["visual.ln_post.bias"])) if "visual.ln_post.weight" in sd else None
        self.vis_proj_w = op(sd["visual.proj"].float().t()) if "visual.proj" in sd else None          # [E, D]
        self.vit = []

        def folded(w_key, b_key, g_key, beta_key):
            """LayerNorm folded into the following linear (distb200_gemm_desc.ln_stats): operand W * gamma, its row sums
            (of the ROUNDED operand, so that the mean term cancels exactly) and the bias W beta + b."""
            W, bb = sd[w_key].detach().double(), sd[b_key].detach().double()
            gam, bet = sd[g_key].detach().double(), sd[beta_key].detach().double()
            wf = (W * gam[None, :]).float().to(device=device).to(act_dtype).contiguous()
            return wf, wf.float().sum(dim=1).contiguous(), f32(W @ bet + bb)

        for l in range(L):
            pre = "visual.transformer.resblocks.%d." % l
            qf = folded(pre + "attn.in_proj_weight", pre + "attn.in_proj_bias", pre + "ln_1.weight", pre + "ln_1.bias")
            ff = folded(pre + "mlp.c_fc.weight", pre + "mlp.c_fc.bias", pre + "ln_2.weight", pre + "ln_2.bias")
            self.vit.append(dict(
                qkv_wf=qf[0], qkv_ws=qf[1], qkv_bf=qf[2], fc1_wf=ff[0], fc1_ws=ff[1], fc1_bf=ff[2],
                ln1=(f32(sd[pre + "ln_1.weight"]), f32(sd[pre + "ln_1.bias"])),
                qkv_w=op(sd[pre + "attn.in_proj_weight"]), qkv_b=f32(sd[pre + "attn.in_proj_bias"]),
                proj_w=op(sd[pre + "attn.out_proj.weight"]), proj_b=f32(sd[pre + "attn.out_proj.bias"]),
                ln2=(f32(sd[pre + "ln_2.weight"]), f32(sd[pre + "ln_2.bias"])),
                fc1_w=op(sd[pre + "mlp.c_fc.weight"]), fc1_b=f32(sd[pre + "mlp.c_fc.bias"]),
                fc2_w=op(sd[pre + "mlp.c_proj.weight"]), fc2_b=f32(sd[pre + "mlp.c_proj.bias"]),
            ))

        # ---- DiST ----
        Ci, Ct, ps = a.integration_dim, a.temporal_dim, a.s_patch
        self.kps = _pad8(3 * ps * ps)
        w = sd["dist_net.temporal_stem.weight"].float()                                        # [Ct, 3, kt, ps, ps]
        w = w.permute(2, 0, 1, 3, 4).reshape(a.t_patch, Ct, 3 * ps * ps)
        self.stem_w = torch.stack([padk(w[k], self.kps) for k in range(a.t_patch)]).contiguous()   # [kt, Ct, Kp]
        self.stem_b = f32(sd["dist_net.temporal_stem.bias"])
        self.dist = []
        for i in range(len(a.selected_layers)):
            tn, it = "dist_net.temporal_nets.%d." % i, "dist_net.integration_nets.%d." % i
            t2i, i2t = "dist_net.temporal2integration_nets.%d." % i, "dist_net.integration2temporal_nets.%d." % i
            w1 = sd[tn + "temporal_net.c_fc1.weight"].float()[:, :, :, 0, 0].permute(2, 0, 1)     # [kt, Ch, Ct]
            w2 = sd[tn + "temporal_net.c_fc2.weight"].float()[:, :, 0].permute(2, 3, 0, 1)        # [3, 3, Ct, Ch]
            wt = sd[t2i + "linear_fuse.weight"].float()[:, :, :, 0, 0].permute(2, 0, 1)           # [alpha, Ci, Ct]
            wk = sd[it + "temporal_ffn.c_fc2.weight"].float()[:, :, :, 0, 0].permute(2, 0, 1)      # [kt, Cm, Cm]
            self.dist.append(dict(
                tn_ln=(f32(sd[tn + "ln.weight"]), f32(sd[tn + "ln.bias"])),
                tn_w1=op(w1), tn_b1=f32(sd[tn + "temporal_net.c_fc1.bias"]),
                tn_w2=op(w2.reshape(9, w2.shape[2], w2.shape[3])), tn_b2=f32(sd[tn + "temporal_net.c_fc2.bias"]),
                in_w=op(sd["dist_net.input_linears.%d.weight" % i]), in_b=f32(sd["dist_net.input_linears.%d.bias" % i]),
                i2t_w=op(sd[i2t + "linear_fuse.weight"]), i2t_b=f32(sd[i2t + "linear_fuse.bias"]),
                t2i_w=op(wt), t2i_b=f32(sd[t2i + "linear_fuse.bias"]),
                t2i_cls=f32(sd[t2i + "cls_token"].float().reshape(a.sparse_frames, Ci)),
                ln=(f32(sd[it + "ln.weight"]), f32(sd[it + "ln.bias"])),
                ln_t=(f32(sd[it + "ln_temporal.weight"]), f32(sd[it + "ln_temporal.bias"])),
                fc_w=op(sd[it + "ffn.c_fc.weight"]), fc_b=f32(sd[it + "ffn.c_fc.bias"]),
                # ffn.c_proj and temporal_ffn.c_proj are applied as ONE GEMM over the K-concatenated hidden activations
                # [ffn hidden | temporal hidden]: res = [h_f | h_t] . [W_proj | W_tproj]^T + (b_proj + b_tproj)   (dist.py:45)
                prj_w=op(torch.cat([sd[it + "ffn.c_proj.weight"].float(), sd[it + "temporal_ffn.c_proj.weight"].float()[:, :, 0, 0, 0]], dim=1)),
                prj_b=f32(sd[it + "ffn.c_proj.bias"].float() + sd[it + "temporal_ffn.c_proj.bias"].float()),
                tf1_w=op(sd[it + "temporal_ffn.c_fc1.weight"].float()[:, :, 0, 0, 0]), tf1_b=f32(sd[it + "temporal_ffn.c_fc1.bias"]),
                tf2_w=op(wk), tf2_b=f32(sd[it + "temporal_ffn.c_fc2.bias"]),
            ))
            # IntegrationNetwork with both LayerNorms folded (bf16 path): ONE GEMM on the raw rows of `upd` produces
            # [QuickGELU(ffn.c_fc(LN x)) | temporal_ffn.c_fc1(LN_t x)] - operand rows [W_fc * gamma ; W_tf1 * gamma_t], their row sums
            # (of the rounded operand) and the biases W beta + b; the activation covers the ffn columns only (dist.py:40-45)
            wt = sd[it + "temporal_ffn.c_fc1.weight"].detach().double()[:, :, 0, 0, 0]
            wf = sd[it + "ffn.c_fc.weight"].detach().double()
            g1, b1 = sd[it + "ln.weight"].detach().double(), sd[it + "ln.bias"].detach().double()
            g2, b2 = sd[it + "ln_temporal.weight"].detach().double(), sd[it + "ln_temporal.bias"].detach().double()
            # K-concatenated operands (bf16 path, DistEngine.kcat; column layout = kcat_layout):
            #   upd_i = [tap | h_{i-1} | xT pair | e_ti] . [W_in | W_proj,i-1 | W_t2i,0..alpha-1 | cls_ti - b_t2i]^T + (b_in + b_t2i + b_proj,i-1)
            #           = input_linear(tap) + res_{i-1} + t2i(xT) + cls token (dist.py:229,80-86,232) with the previous
            #           IntegrationNetwork's output projection res_{i-1} = h_{i-1} W_proj^T + b (dist.py:45) riding in the same reduction
            #   i2t   = [xT pair | e_ti | upd] . [-W_i2t W_t2i | 0 | W_i2t]^T + (b_i2t - W_i2t b_t2i)   on the patch rows,
            #           i.e. linear_fuse(mid) with mid = upd - t2i(xT) (dist.py:99-105,231 reads the PRE-fusion mid)
            KL = kcat_layout(a)
            Ih, Cm = a.integration_hidden, a.integration_temporal_hidden
            cpu64 = lambda x: x.detach().double().cpu()
            w_in = cpu64(sd["dist_net.input_linears.%d.weight" % i])
            wt2 = cpu64(sd[t2i + "linear_fuse.weight"])[:, :, :, 0, 0].permute(2, 0, 1)                     # [alpha, Ci, Ct]
            b_t2 = cpu64(sd[t2i + "linear_fuse.bias"])
            cls = cpu64(sd[t2i + "cls_token"]).reshape(a.sparse_frames, Ci)
            w_t2_flat = torch.cat([wt2[k] for k in range(a.alpha)], dim=1)                                   # [Ci, alpha*Ct]
            cat_w = torch.zeros(Ci, KL["c_u"], dtype=torch.float64)
            cat_w[:, :D] = w_in
            cat_b = cpu64(sd["dist_net.input_linears.%d.bias" % i]) + b_t2
            if i > 0:
                pit = "dist_net.integration_nets.%d." % (i - 1)
                cat_w[:, KL["c_h"]:KL["c_h"] + Ih] = cpu64(sd[pit + "ffn.c_proj.weight"])
                cat_w[:, KL["c_h"] + Ih + Cm:KL["c_h"] + Ih + 2 * Cm] = cpu64(sd[pit + "temporal_ffn.c_proj.weight"])[:, :, 0, 0, 0]
                cat_b = cat_b + cpu64(sd[pit + "ffn.c_proj.bias"]) + cpu64(sd[pit + "temporal_ffn.c_proj.bias"])
            cat_w[:, KL["c_x"]:KL["c_x"] + a.alpha * Ct] = w_t2_flat
            cat_w[:, KL["c_oh"]:KL["c_oh"] + a.sparse_frames] = (cls - b_t2[None, :]).t()
            w_i2 = cpu64(sd[i2t + "linear_fuse.weight"])                                                     # [Ct, Ci]
            i2_cat = torch.zeros(Ct, KL["c_u"] + Ci - KL["c_x"], dtype=torch.float64)
            i2_cat[:, :a.alpha * Ct] = -(w_i2 @ w_t2_flat)
            i2_cat[:, KL["c_u"] - KL["c_x"]:] = w_i2
            # IntegrationNetwork hidden block [QuickGELU(ffn.c_fc) Ih | temporal c_fc1 Cm | temporal hidden Cm]: folded c_fc / c_fc1 rows
            # in that order, and the last layer's output projection over the same block (zero weights on the c_fc1 columns)
            cat2 = torch.cat([wf * g1[None, :], wt * g2[None, :]], dim=0).float().to(device=device).to(act_dtype).contiguous()
            prj = torch.zeros(Ci, Ih + 2 * Cm, dtype=torch.float64)
            prj[:, :Ih] = cpu64(sd[it + "ffn.c_proj.weight"])
            prj[:, Ih + Cm:] = cpu64(sd[it + "temporal_ffn.c_proj.weight"])[:, :, 0, 0, 0]
            self.dist[-1].update(
                cat_w=op(cat_w.float()), cat_b=f32(cat_b), i2t_cat_w=op(i2_cat.float()),
                i2t_cat_b=f32(cpu64(sd[i2t + "linear_fuse.bias"]) - w_i2 @ b_t2),
                int_wf2=cat2, int_ws2=cat2.float().sum(dim=1).contiguous(),
                int_bf2=f32(torch.cat([wf @ b1 + sd[it + "ffn.c_fc.bias"].detach().double(),
                                       wt @ b2 + sd[it + "temporal_ffn.c_fc1.bias"].detach().double()])),
                prj_w_h=op(prj.float()))
        self.ada = []
        for j in range(a.ada_layers):
            pre = "dist_net.adapooling_nets.%d." % j
            entry = dict(pos=f32(sd[pre + "positional_embedding"].float().reshape(a.sparse_frames, Ci)))
            for tag, which, ln_out, mlp in (("sp", "spatial_transformer", "ln_out_spat_cls_token", "output_map_spatial_cls_token"),
                                            ("tp", "temporal_transformer", "ln_out_temp_cls_token", "output_map_cls_token")):
                wq = sd[pre + which + ".attn.in_proj_weight"].float()
                bq = sd[pre + which + ".attn.in_proj_bias"].float()
                entry[tag] = dict(
                    ln=(f32(sd[pre + which + ".ln_1.weight"]), f32(sd[pre + which + ".ln_1.bias"])),
                    q_w=op(wq[:Ci]), q_b=f32(bq[:Ci]), kv_w=op(wq[Ci:]), kv_b=f32(bq[Ci:]),
                    # K / V projection on the NORMALISED rows (unit affine): W diag(gamma), b + W beta.  Every ada layer normalises the same
                    # tensor (res + upd of the last DiST layer, dist.py:239-241), so one LayerNorm pass serves all of them.
                    kv_wn=op(wq[Ci:].detach().double().cpu() * sd[pre + which + ".ln_1.weight"].detach().double().cpu()[None, :]),
                    kv_bn=f32(bq[Ci:].detach().double().cpu() + wq[Ci:].detach().double().cpu() @ sd[pre + which + ".ln_1.bias"].detach().double().cpu()),
                    o_w=op(sd[pre + which + ".attn.out_proj.weight"]), o_b=f32(sd[pre + which + ".attn.out_proj.bias"]),
                    ln_out=(f32(sd[pre + ln_out + ".weight"]), f32(sd[pre + ln_out + ".bias"])),
                    fc_w=op(sd[pre + mlp + ".c_fc.weight"]), fc_b=f32(sd[pre + mlp + ".c_fc.bias"]),
                    pr_w=op(sd[pre + mlp + ".c_proj.weight"]), pr_b=f32(sd[pre + mlp + ".c_proj.bias"]),
                )
            self.ada.append(entry)
        self.unit_ln = (f32(torch.ones(Ci)), f32(torch.zeros(Ci)))
        self.agg_cls = f32(sd["dist_net.aggregated_cls_token"].float().reshape(1, Ci))
        self.agg_sp = f32(sd["dist_net.aggregated_spatial_cls_token"].float().reshape(1, Ci))
        self.pcls_w = op(sd["dist_net.proj_spatial_cls_token.weight"])
        self.pcls_b = f32(sd["dist_net.proj_spatial_cls_token.bias"])
        self.ln_post = (f32(sd["dist_net.ln_post.weight"]), f32(sd["dist_net.ln_post.bias"]))
        self.proj_w = op(sd["dist_net.proj"].float().t())                                       # [E, Ci]
        self.logit_scale = float(sd["logit_scale"].float().exp()) if "logit_scale" in sd else 1.0 / 0.07


class DistEngine:
    """Planned forward for a fixed number of clips per call."""

    # CLIP normalisation (DATA.MEAN / DATA.STD, configs/projects/dist/vit_base_16_ssv2.yaml:31-32)
    CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
    CLIP_STD = (0.26862954, 0.26130258, 0.27577711)

    def __init__(self, state_dict, arch: DistArch, batch, device="cuda", precision="bf16", text_features=None,
                 gemm_impl=ops.IMPL_AUTO, attn_impl=ops.IMPL_AUTO, input_format="float", mean=None, std=None, image_logits=None,
                 fusion_weight=0.5):
        """``image_logits``: ``None`` (DiST's own path: the ViT's ``ln_post`` / ``proj`` are dead code, clip.py:291-298), ``"raw"`` (also
        compute the per-frame CLIP image embeddings ``img_emb`` [b*t, E]) or ``"fused"`` (the zero-shot / prediction-fusion branch of
        clip.py:519-527: class scores = ``fusion_weight`` x video logits + the rest x mean per-frame logits; ``img_emb_n`` holds the
        normalised embeddings).  ``input_format="uint8"``: clips arrive as decoded frames ``[b, T, H, W, 3]`` uint8 and ToTensorVideo +
        NormalizeVideo (``dataset/base/ssv2.py:137-143``) are fused into the patch-row kernel."""
        assert precision in ("bf16", "fp32") and input_format in ("float", "uint8") and image_logits in (None, "raw", "fused")
        self.image_logits, self.fusion_weight = image_logits, float(fusion_weight)
        self.input_format = input_format
        self.mean, self.std = tuple(mean or self.CLIP_MEAN), tuple(std or self.CLIP_STD)
        arch.validate()
        ops.lib()     # fail loudly before touching the GPU if the extension is missing
        self.arch, self.batch, self.device, self.precision = arch, int(batch), torch.device(device), precision
        self.adt = torch.bfloat16 if precision == "bf16" else torch.float32
        self.gemm_impl, self.attn_impl = gemm_impl, attn_impl
        self.w = PackedWeights(state_dict, arch, self.device, self.adt)
        self.text_n = None
        if text_features is not None:
            self.set_text_features(text_features)
        self._alloc()
        self.calls = []
        self.static_calls = []
        self._plan()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        for c in self.static_calls:                       # weight-only work of the plan (see _plan_head): once per engine
            c.launch(stream)
        if self.static_calls:
            torch.cuda.current_stream(self.device).synchronize()       # later forwards may run on another stream
        self.graph = None

    # ------------------------------------------------------------------------------------------
    def set_text_features(self, feats):
        """Cached label embeddings [C, E] (clip.py:437-452); normalised once (clip.py:514)."""
        f = feats.detach().to(device=self.device, dtype=torch.float32)
        f = (f / f.norm(dim=1, keepdim=True)).contiguous()
        if hasattr(self, "logits") and f.shape[0] != self.logits.shape[1]:
            raise ValueError("text features changed the number of classes; rebuild the engine")
        if getattr(self, "text_n", None) is not None and self.text_n.shape == f.shape:
            self.text_n.copy_(f)            # in place: the planned head call (and a captured graph) keep pointing at this buffer
        else:
            self.text_n = f

    def _alloc(self):
        a, b, dev, adt = self.arch, self.batch, self.device, self.adt
        F, N, P, D = b * a.sparse_frames, a.tokens, a.patches, a.width
        T, Ci, Ct = a.frames, a.integration_dim, a.temporal_dim
        Mv, Mt = F * N, b * T * P
        z = lambda *s, dtype=adt: torch.zeros(*s, device=dev, dtype=dtype)
        f32 = torch.float32
        if self.input_format == "uint8":
            self.video = z(b, T, a.resolution, a.resolution, 3, dtype=torch.uint8)
        else:
            self.video = z(b, 3, T, a.resolution, a.resolution, dtype=f32)
        self.patches_s = z(F * P, self.w.kp) if a.patch != a.s_patch else None
        self.patches_d = z(b * T * P, self.w.kps)
        self.h = z(Mv, D, dtype=f32)
        self.ln_buf = z(Mv, D)
        self.qkv = z(Mv, 3 * D)
        self.attn_out = z(Mv, D)
        self.fc1 = z(Mv, 4 * D)
        # bf16 copies of the ViT block outputs (the taps; with the folded LayerNorm also the next block's GEMM operand).  Two
        # buffers, alternating by layer, so that the DiST layer reading tap l may still run while ViT block l+1 writes its own.
        self.ln_fold = self.precision == "bf16" and os.environ.get("DISTB200_LN_FOLD", "1") != "0"
        # K-concatenated DiST operands (see PackedWeights): the tap buffers grow the columns [alpha temporal rows | one-hot(ti) |
        # bf16 upd]; input_linear + the temporal->integration convolution + the cls token become ONE GEMM, `mid` is written once
        self.kcat = self.ln_fold and os.environ.get("DISTB200_KCAT", "1") != "0"
        if self.kcat:
            self.kl = kcat_layout(a)
            self.tw = self.kl["tw"]
            self.tapw = [z(Mv, self.tw), z(Mv, self.tw)]
            frames = torch.arange(F, device=dev)
            for buf in self.tapw:                      # static one-hot(ti) on the class-token row of every frame
                buf[frames * N, self.kl["c_oh"] + frames % a.sparse_frames] = 1.0
            self.taps = [buf[:, :D] for buf in self.tapw]
        else:
            self.taps = [z(Mv, D), z(Mv, D)] if self.precision == "bf16" else [self.h, self.h]
        self.tap = self.taps[0]
        # LayerNorm folded into the QKV / FC1 GEMMs (bf16 path): bf16 copies of the residual stream + per-row statistics
        # ... and the statistics themselves come out of the producing GEMM's epilogue (distb200_gemm_desc.stat_partials) instead of a
        # separate pass over the bf16 copy: one slot buffer per producer shape (FC2 -> ln_1, out_proj -> ln_2)
        # Measured (B/16 8+16f, 32 clips, per step): FC2 is tensor-bound and its epilogue has slack - emitting the statistics costs it
        # +0.03 ms and removes 0.18 ms of row_stats; out_proj is bound by its epilogue / the fp32 stream and loses +0.18 ms for the
        # 0.19 ms it saves.  Hence mode 2 by default (0: stand-alone row_stats, 1: both producers, 2: FC2 only).
        mode = os.environ.get("DISTB200_LN_STATS_FUSED", "2")
        self.stats_fused = self.ln_fold and mode != "0" and self.gemm_impl != ops.IMPL_SIMT       # emitted by the tcgen05 epilogue only
        self.stats_fused_out_proj = self.stats_fused and mode != "2"
        if self.ln_fold:
            self.hb_mid = z(Mv, D)
            self.row_st = z(Mv, 2, dtype=f32)
            if self.stats_fused:
                self.stat_p = [z(Mv, ops.STAT_SLOTS, 2, dtype=f32), z(Mv, ops.STAT_SLOTS, 2, dtype=f32)]
        self._hb_prev = None
        self.xT = z(Mt, Ct, dtype=f32)
        # Fused TemporalNet (distb200_temporalnet, bf16 path): LayerNorm + both convolutions + residual in one launch, the
        # integration->temporal add of the previous layer folded into its input.  The stream ping-pongs between two buffers (a frame's
        # neighbours are read while it is written) and the i2t GEMM only writes its [b*t*P, Ct] result `u`.
        self.tn_fused = (self.kcat and os.environ.get("DISTB200_TN_FUSED", "1") != "0" and Ct in (32, 64, 96)
                         and a.temporal_hidden == Ct and a.t_kernel == 3 and 128 // a.grid - 2 >= 1)
        if self.tn_fused:
            self.xTs = [self.xT, z(Mt, Ct, dtype=f32)]
            self.u_i2t = z(F * P, Ct)
        self.xT_a = z(Mt, Ct)
        self.xln = z(Mt, Ct)
        self.y1 = z(Mt, a.temporal_hidden)
        self.mid = z(Mv, Ci, dtype=f32)
        self.mid_a = z(Mv, Ci)
        self.res = z(Mv, Ci, dtype=f32)
        self.int_a1 = z(Mv, Ci)
        self.int_a2 = z(Mv, Ci)
        self.int_h = z(Mv, a.integration_hidden + a.integration_temporal_hidden)     # [ffn hidden | temporal hidden]
        if self.kcat:
            self.int_st = z(Mv, 2, dtype=f32)                                           # row statistics of upd (both LayerNorms folded)
            self.int_w = z(Mv, 2 * a.integration_temporal_hidden + a.integration_hidden)     # the LAST layer's hidden block
        self.tf1 = z(Mv, a.integration_temporal_hidden)
        self.kv_s = z(Mv, 2 * Ci)
        self.sp = z(F, Ci, dtype=f32)
        self.top = z(b, Ci, dtype=f32)
        self.sp_ln, self.q_s, self.o_s, self.mlp_s = z(F, Ci), z(F, Ci), z(F, Ci), z(F, 4 * Ci)
        self.kv_t = z(F, 2 * Ci)
        self.top_ln, self.q_t, self.o_t, self.mlp_t = z(b, Ci), z(b, Ci), z(b, Ci), z(b, 4 * Ci)
        self.clsmean = z(b, D)
        self.zbuf = z(b, Ci, dtype=f32)
        self.z_ln = z(b, Ci)
        self.emb = z(b, a.embed_dim, dtype=f32)
        C = self.text_n.shape[0] if self.text_n is not None else max(a.num_classes, 1)
        self.logits = z(b, C, dtype=f32)
        self.probs = z(b, C, dtype=f32)
        if self.image_logits:
            self.img_ln = z(F, D)
            self.img_emb = z(F, a.embed_dim, dtype=f32)
            self.img_emb_n = z(F, a.embed_dim, dtype=f32)

    # ------------------------------------------------------------------------------------------
    def _gemm(self, *args, **kw):
        kw.setdefault("impl", self.gemm_impl)
        self.calls.append(ops.gemm(*args, **kw))

    def _lin(self, a, w, bias, out, *, res=None, out2=None, act=ops.ACT_NONE, ld_out=None, name="linear", **kw):
        """out[M, n] = act(a[M, k] @ w[n, k]^T + bias (+ res)); ``ld_out`` > n writes into a column slice of a wider buffer"""
        n, k = w.shape
        ld2 = out2.stride(0) if (out2 is not None and out2.dim() == 2) else n
        self._gemm(a, w, n, k, bias=bias, res=res, ld_res=n, out=out, ld_out=ld_out or n, out2=out2, ld_out2=ld2, act=act, name=name, **kw)

    def _ln(self, x, gb, y, **kw):
        self.calls.append(ops.layernorm(x, gb[0], gb[1], y, **kw))

    def _plan_inputs(self):
        """Patch rows, ViT embedding and temporal stem (everything that reads the clip)."""
        a, b, w = self.arch, self.batch, self.w
        t, T, N, P, D = a.sparse_frames, a.frames, a.tokens, a.patches, a.width
        F, Ct, al = b * t, a.temporal_dim, a.alpha
        R = a.resolution
        add = self.calls.append

        # ---- patch rows: sparse frames for the ViT (clip.py:271 restricted to the frames kept at :281-284),
        #      all frames for the temporal stem (dist.py:225)
        self.sections = {"vit": [], "dist": []}
        mark = len(self.calls)
        if self.input_format == "uint8":
            cut = lambda *args, **kw: ops.patchify_u8(*args, self.mean, self.std, **kw)
        else:
            cut = ops.patchify
        add(cut(self.video, self.patches_d, b, T, R, R, a.s_patch, 0, 1, T, w.kps, name="patchify.dense"))
        shared = a.patch == a.s_patch          # then the ViT's patches are the rows of every alpha-th dense frame
        if not shared:
            add(cut(self.video, self.patches_s, b, T, R, R, a.patch, 0, al, t, w.kp, name="patchify.sparse"))

        self.sections["patchify"] = (mark, len(self.calls))
        mark = len(self.calls)
        # ---- ViT embedding: conv1 as GEMM, + positional embedding, class row, ln_pre (clip.py:271-276)
        k1 = 3 * a.patch * a.patch
        src, fstride = (self.patches_d, al * P * w.kp) if shared else (self.patches_s, P * w.kp)
        self._gemm(src, w.conv1_w, D, k1, a_dim=(k1, P, F, 1), a_stride=(1, w.kp, fstride, F * fstride),
                   groups=F, rows_per_group=P, ldb=w.kp, res=w.pos, ld_res=D, res_gstride=0, res_roff=1,
                   out=self.h, ld_out=D, out_gstride=N, out_roff=1, name="vit.patch_embed")
        add(ops.rows_bcast(self.h, N * D, F, D, w.cls_row, 1, False, name="vit.cls_rows"))   # class embedding + pos[0]
        self._ln(self.h, w.ln_pre, self.h, name="vit.ln_pre")

        self.sections["embed"] = (mark, len(self.calls))
        mark = len(self.calls)
        # ---- temporal stem: Conv3d(3->Ct,(kt,ps,ps)) = kt row-shifted GEMMs over patch rows (dist.py:178-181,225)
        self._plan_stem()
        self.sections["stem"] = (mark, len(self.calls))

    def _stem_operand(self):
        """(a_dim, a_stride, taps) of the stem's A operand: the dense patch rows of a clip, kt row-shifted taps."""
        a, b, w = self.arch, self.batch, self.w
        T, P = a.frames, a.patches
        ks = 3 * a.s_patch * a.s_patch
        half = a.t_patch // 2
        return ((ks, T * P, b, 1), (1, w.kps, T * P * w.kps, b * T * P * w.kps), [((k - half) * P, 0, 0) for k in range(a.t_patch)])

    def _plan_stem(self):
        a, b, w = self.arch, self.batch, self.w
        Ct, ks = a.temporal_dim, 3 * a.s_patch * a.s_patch
        a_dim, a_stride, taps = self._stem_operand()
        self._gemm(self.patches_d, w.stem_w, Ct, ks, a_dim=a_dim, a_stride=a_stride, taps=taps, b_tap_stride=Ct * w.kps, ldb=w.kps,
                   groups=b, rows_per_group=a.frames * a.patches, bias=w.stem_b, out=self.xT, ld_out=Ct, name="dist.stem")

    def _plan(self):
        a, b, w = self.arch, self.batch, self.w
        t, N, D = a.sparse_frames, a.tokens, a.width
        add = self.calls.append
        bf = self.precision == "bf16"
        self._plan_inputs()

        sel = list(a.selected_layers)
        monotonic = all(x < y for x, y in zip(sel, sel[1:]))
        assert monotonic, "SELECTED_LAYERS must be increasing (taps are consumed as the ViT produces them)"
        last_sel = sel[-1]
        for l in range(a.layers):
            want_tap = l in sel
            mark = len(self.calls)
            self._plan_vit_layer(l, self.taps[l % 2] if (want_tap and bf) else None)
            self.sections["vit"].append((mark, len(self.calls)))
            if want_tap:
                mark = len(self.calls)
                self._plan_dist_layer(sel.index(l))
                self.sections["dist"].append((mark, len(self.calls)))
            if l == last_sel:
                # mean over the sparse frames of the last tapped class token (dist.py:243)
                mark = len(self.calls)
                add(ops.mean_rows(self.h, N * D, t, b, D, self.clsmean, name="dist.cls_mean"))
                self.sections["cls_mean"] = (mark, len(self.calls))
        mark = len(self.calls)
        self._plan_head()
        self.sections["head"] = (mark, len(self.calls))

    def _plan_vit_layer(self, l, tap_out):
        """ResidualAttentionBlockMid (clip.py:170-178); ``tap_out`` receives a copy of the block output (the tap).

        bf16 path: ln_1 / ln_2 are folded into the QKV / FC1 GEMMs.  The producing GEMM (FC2 of the previous block,
        out_proj of this one) writes a bf16 copy of the residual stream next to the fp32 one; ``row_stats`` reads that copy
        (half the bytes of the fp32 stream, nothing written back) and the consuming GEMM normalises in its epilogue.
        The first block's ln_1 follows ln_pre, which is no GEMM: it keeps the stand-alone kernel."""
        a, v = self.arch, self.w.vit[l]
        F, N = self.batch * a.sparse_frames, a.tokens
        add = self.calls.append
        fold = self.ln_fold
        fused = fold and getattr(self, "stats_fused", False)
        D = a.width
        if fold and self._hb_prev is not None:
            if fused:
                add(ops.row_stats_finalize(self.stat_p[0], D, self.row_st, name="vit.ln_1.stats"))
            else:
                add(ops.row_stats(self._hb_prev, self.row_st, name="vit.ln_1.stats"))
            self._lin(self._hb_prev, v["qkv_wf"], v["qkv_bf"], self.qkv, ln_stats=self.row_st, ln_wsum=v["qkv_ws"], name="vit.qkv")
        else:
            self._ln(self.h, v["ln1"], self.ln_buf, name="vit.ln_1")
            self._lin(self.ln_buf, v["qkv_w"], v["qkv_b"], self.qkv, name="vit.qkv")
        add(ops.attention(self.qkv, self.attn_out, F, N, a.heads, impl=self.attn_impl, name="vit.attention"))
        if fold:
            self._lin(self.attn_out, v["proj_w"], v["proj_b"], self.h, res=self.h, out2=self.hb_mid, name="vit.out_proj",
                      stat_partials=self.stat_p[1] if (fused and self.stats_fused_out_proj) else None)
            if fused and self.stats_fused_out_proj:
                add(ops.row_stats_finalize(self.stat_p[1], D, self.row_st, name="vit.ln_2.stats"))
            else:
                add(ops.row_stats(self.hb_mid, self.row_st, name="vit.ln_2.stats"))
            self._lin(self.hb_mid, v["fc1_wf"], v["fc1_bf"], self.fc1, act=ops.ACT_QUICKGELU, ln_stats=self.row_st, ln_wsum=v["fc1_ws"],
                      name="vit.fc1")
            hb = tap_out if tap_out is not None else self.taps[l % 2]   # bf16 copy of the block output: the tap AND the next block's operand
            self._lin(self.fc1, v["fc2_w"], v["fc2_b"], self.h, res=self.h, out2=hb, name="vit.fc2", stat_partials=self.stat_p[0] if fused else None)
            self._hb_prev = hb
        else:
            self._lin(self.attn_out, v["proj_w"], v["proj_b"], self.h, res=self.h, name="vit.out_proj")
            self._ln(self.h, v["ln2"], self.ln_buf, name="vit.ln_2")
            self._lin(self.ln_buf, v["fc1_w"], v["fc1_b"], self.fc1, act=ops.ACT_QUICKGELU, name="vit.fc1")
            self._lin(self.fc1, v["fc2_w"], v["fc2_b"], self.h, res=self.h, out2=tap_out, name="vit.fc2")

    def _plan_dist_layer(self, i):
        a, b, w = self.arch, self.batch, self.w
        d = w.dist[i]
        t, T, N, P, g = a.sparse_frames, a.frames, a.tokens, a.patches, a.grid
        F, Ci, Ct, al, Ch, Cm = b * t, a.integration_dim, a.temporal_dim, a.alpha, a.temporal_hidden, a.integration_temporal_hidden
        Mv, Mt = F * N, b * T * P
        add = self.calls.append
        half = a.t_kernel // 2

        if self.kcat and self.tn_fused:
            self._plan_dist_layer_kcat(i, None)
            return
        # ---- TemporalNet (dist.py:48-65) ----
        self._ln(self.xT, d["tn_ln"], self.xln, name="dist.tn.ln")
        self._gemm(self.xln, d["tn_w1"], Ch, Ct, a_dim=(Ct, T * P, b, 1), a_stride=(1, Ct, T * P * Ct, Mt * Ct),
                   taps=[((k - half) * P, 0, 0) for k in range(a.t_kernel)], b_tap_stride=Ch * Ct, ldb=Ct,
                   groups=b, rows_per_group=T * P, bias=d["tn_b1"], out=self.y1, ld_out=Ch, act=ops.ACT_QUICKGELU,
                   name="dist.tn.conv_t")
        conv_s = dict(a_dim=(Ch, g, g, b * T), a_stride=(1, Ch, g * Ch, P * Ch), img_w=g,
                      taps=[(j - 1, ii - 1, 0) for ii in range(3) for j in range(3)], b_tap_stride=Ct * Ch, ldb=Ch,
                      groups=b * T, rows_per_group=P, bias=d["tn_b2"], res=self.xT, ld_res=Ct, res_gstride=P,
                      out=self.xT, ld_out=Ct, act=ops.ACT_QUICKGELU, name="dist.tn.conv_s")
        if self.kcat:
            self._plan_dist_layer_kcat(i, conv_s)
            return
        self._gemm(self.y1, d["tn_w2"], Ct, Ch, out2=self.xT_a, ld_out2=Ct, **conv_s)

        # ---- input linear + previous integration output (dist.py:229) ----
        self._lin(self.taps[a.selected_layers[i] % 2], d["in_w"], d["in_b"], self.mid, res=self.res if i > 0 else None, out2=self.mid_a,
                  name="dist.input_linear")

        # ---- temporal -> integration (dist.py:68-86,232): alpha-tap GEMM over frame groups, result added onto the patch rows
        self._gemm(self.xT_a, d["t2i_w"], Ci, Ct, a_dim=(Ct, P, al, F), a_stride=(1, Ct, P * Ct, al * P * Ct), group_dim=3,
                   taps=[(0, k, 0) for k in range(al)], b_tap_stride=Ci * Ct, ldb=Ct, groups=F, rows_per_group=P,
                   bias=d["t2i_b"], res=self.mid, ld_res=Ci, res_gstride=N, res_roff=1,
                   out=self.mid, ld_out=Ci, out_gstride=N, out_roff=1, name="dist.t2i")
        add(ops.rows_bcast(self.mid, N * Ci, F, Ci, d["t2i_cls"], t, True, name="dist.t2i.cls"))

        # ---- integration -> temporal (dist.py:90-105,231): patch tokens only, nearest upsample = row replication
        # (the fused temporal stream of the LAST layer is read by nothing, dist.py:231-235: its branch is skipped)
        if i < len(a.selected_layers) - 1:
            self._gemm(self.mid_a, d["i2t_w"], Ct, Ci, a_dim=(Ci, N, F, 1), a_stride=(1, Ci, N * Ci, Mv * Ci), taps=[(1, 0, 0)],
                       groups=F, rows_per_group=P, ldb=Ci, bias=d["i2t_b"], res=self.xT, ld_res=Ct, res_gstride=al * P,
                       res_rep_stride=P, out=self.xT, ld_out=Ct, out_gstride=al * P, out_rep=al, out_rep_stride=P, name="dist.i2t")

        # ---- IntegrationNetwork (dist.py:16-45) on upd = mid ----
        add(ops.layernorm(self.mid, d["ln"][0], d["ln"][1], self.int_a1, g2=d["ln_t"][0], b2=d["ln_t"][1], y2=self.int_a2,
                          name="dist.int.ln"))
        Ih = a.integration_hidden
        wide = Ih + Cm
        self._lin(self.int_a1, d["fc_w"], d["fc_b"], self.int_h, act=ops.ACT_QUICKGELU, ld_out=wide, name="dist.int.ffn_fc")
        self._lin(self.int_a2, d["tf1_w"], d["tf1_b"], self.tf1, name="dist.int.t_fc1")
        self._gemm(self.tf1, d["tf2_w"], Cm, Cm, a_dim=(Cm, t * N, b, 1), a_stride=(1, Cm, t * N * Cm, Mv * Cm),
                   taps=[((k - half) * N, 0, 0) for k in range(a.t_kernel)], b_tap_stride=Cm * Cm, ldb=Cm,
                   groups=b, rows_per_group=t * N, bias=d["tf2_b"], out=self.int_h[:, Ih:], ld_out=wide, act=ops.ACT_QUICKGELU,
                   name="dist.int.t_conv")
        self._lin(self.int_h, d["prj_w"], d["prj_b"], self.res, name="dist.int.proj")

    def _plan_dist_layer_kcat(self, i, conv_s):
        """DiST layer i behind TemporalNet's first convolution, on the K-concatenated operand buffers (bf16 path, kcat_layout)."""
        a, b, w = self.arch, self.batch, self.w
        d = w.dist[i]
        sel = list(a.selected_layers)
        t, T, N, P = a.sparse_frames, a.frames, a.tokens, a.patches
        F, Ci, Ct, al, Cm, Ih = b * t, a.integration_dim, a.temporal_dim, a.alpha, a.integration_temporal_hidden, a.integration_hidden
        Mv = F * N
        half = a.t_kernel // 2
        last = i == len(sel) - 1
        L, tw = self.kl, self.tw
        buf = self.tapw[sel[i] % 2]
        upd_a = buf[:, L["c_u"]:L["c_u"] + Ci]
        # TemporalNet drops the bf16 copy of dense frame alpha*ti + k next to token row 1 + r of sparse frame ti
        if self.tn_fused:
            # dist.py:48-65 in one launch; x = stream of the previous layer + its integration->temporal term (dist.py:231).  The last
            # layer's fp32 stream is read by nothing (only its bf16 copy feeds the temporal->integration columns).
            self.calls.append(ops.temporalnet(
                self.xTs[i % 2], d["tn_ln"][0], d["tn_ln"][1], d["tn_w1"], d["tn_b1"], d["tn_w2"], d["tn_b2"], clips=b, frames=T, grid=a.grid,
                u=self.u_i2t if i > 0 else None, alpha=al, out=None if last else self.xTs[(i + 1) % 2], out2=buf[:, L["c_x"]:], ld_out2=tw,
                out2_gdiv=al, out2_cstep=Ct, out2_gstride=N, out2_roff=1, name="dist.tn"))
        else:
            self._gemm(self.y1, d["tn_w2"], Ct, a.temporal_hidden, out2=buf[:, L["c_x"]:], ld_out2=tw, out2_gdiv=al, out2_cstep=Ct, out2_gstride=N,
                       out2_roff=1, **conv_s)
        # upd in ONE pass: input_linear + previous IntegrationNetwork's output projection + temporal->integration + cls token
        self._gemm(buf, d["cat_w"], Ci, L["c_u"], bias=d["cat_b"], out=self.mid, ld_out=Ci, out2=upd_a, ld_out2=tw, name="dist.input_linear")
        # integration -> temporal on the pre-fusion stream mid = upd - t2i(xT) (dist.py:99-105,231); dead in the last layer
        if not last:
            k2 = L["c_u"] + Ci - L["c_x"]
            i2t = dict(a_dim=(k2, N, F, 1), a_stride=(1, tw, N * tw, Mv * tw), taps=[(1, 0, 0)], groups=F, rows_per_group=P, ldb=k2,
                       bias=d["i2t_cat_b"], name="dist.i2t")
            if self.tn_fused:      # the nearest upsample + add happens in the next layer's TemporalNet
                self._gemm(buf[:, L["c_x"]:], d["i2t_cat_w"], Ct, k2, out=self.u_i2t, ld_out=Ct, **i2t)
            else:
                self._gemm(buf[:, L["c_x"]:], d["i2t_cat_w"], Ct, k2, res=self.xT, ld_res=Ct, res_gstride=al * P, res_rep_stride=P,
                           out=self.xT, ld_out=Ct, out_gstride=al * P, out_rep=al, out_rep_stride=P, **i2t)
        # IntegrationNetwork (dist.py:16-45), both LayerNorms folded; its hidden block lands next to the NEXT layer's tap, where that
        # layer's first GEMM applies the output projection (the last layer projects here: the ada-pooling head reads `res`)
        hbuf = self.int_w if last else self.tapw[sel[i + 1] % 2][:, L["c_h"]:]
        hp = hbuf.stride(0)
        self.calls.append(ops.row_stats(upd_a, self.int_st, name="dist.int.stats"))
        self._gemm(upd_a, d["int_wf2"], Ih + Cm, Ci, bias=d["int_bf2"], out=hbuf, ld_out=hp, act=ops.ACT_QUICKGELU, act_to=Ih,
                   ln_stats=self.int_st, ln_wsum=d["int_ws2"], name="dist.int.fc",
                   block_n=block_n_by_waves(Mv, Ih + Cm, [bn for bn in (256, 240, 192, 160, 128) if (Ih + Cm) % bn == 0]))
        self._gemm(hbuf[:, Ih:], d["tf2_w"], Cm, Cm, a_dim=(Cm, t * N, b, 1), a_stride=(1, hp, t * N * hp, Mv * hp),
                   taps=[((k - half) * N, 0, 0) for k in range(a.t_kernel)], b_tap_stride=Cm * Cm, ldb=Cm,
                   groups=b, rows_per_group=t * N, bias=d["tf2_b"], out=hbuf[:, Ih + Cm:], ld_out=hp, act=ops.ACT_QUICKGELU,
                   name="dist.int.t_conv")
        if last:
            self._lin(self.int_w, d["prj_w_h"], d["prj_b"], self.res, name="dist.int.proj")

    def _plan_head(self):
        a, b, w = self.arch, self.batch, self.w
        t, N, Ci = a.sparse_frames, a.tokens, a.integration_dim
        F, Mv, H = b * t, b * t * N, a.integration_heads
        add = self.calls.append
        # ---- ada-pooling (dist.py:237-241, 139-162) ----
        # The two pooled streams START as the broadcast aggregation tokens (dist.py:237-238), i.e. as constants of the weights: the
        # broadcasts, and the LayerNorm + query projection of the FIRST ada layer, run once here (`static_calls`) instead of in every
        # step (six launches); the first layer's output projections take the constant streams as their residual.
        dev, f32 = self.device, torch.float32
        self.sp0, self.top0 = torch.zeros(F, Ci, device=dev, dtype=f32), torch.zeros(b, Ci, device=dev, dtype=f32)
        self.q_s0, self.q_t0 = torch.zeros(F, Ci, device=dev, dtype=self.adt), torch.zeros(b, Ci, device=dev, dtype=self.adt)
        self.static_calls = [ops.rows_bcast(self.sp0, Ci, F, Ci, w.agg_sp, 1, False, name="ada.init_sp"),
                             ops.rows_bcast(self.top0, Ci, b, Ci, w.agg_cls, 1, False, name="ada.init_top")]
        main = self.calls
        for j in range(a.ada_layers):
            e = w.ada[j]
            s, tp = e["sp"], e["tp"]
            first = j == 0
            # spatial: query = per-frame token, key = value = LN(res + upd)  (clip.py:146-147)
            if first:                                   # x_hat of res + upd, once for all ada layers (their LayerNorm affines are folded into kv_wn / kv_bn)
                self._ln(self.res, w.unit_ln, self.int_a1, in2=self.mid, in2_period=Mv, name="ada.sp.ln_kv")
            self._lin(self.int_a1, s["kv_wn"], s["kv_bn"], self.kv_s, name="ada.sp.kv")
            if first:
                self.calls = self.static_calls
            self._ln(self.sp0 if first else self.sp, s["ln"], self.sp_ln, name="ada.sp.ln_q")
            self._lin(self.sp_ln, s["q_w"], s["q_b"], self.q_s0 if first else self.q_s, name="ada.sp.q")
            self.calls = main
            add(ops.cross_attention(self.q_s0 if first else self.q_s, self.kv_s, self.o_s, F, N, H, name="ada.sp.attn"))
            self._lin(self.o_s, s["o_w"], s["o_b"], self.sp, res=self.sp0 if first else self.sp, name="ada.sp.out_proj")
            self._ln(self.sp, s["ln_out"], self.sp_ln, name="ada.sp.ln_out")
            self._lin(self.sp_ln, s["fc_w"], s["fc_b"], self.mlp_s, act=ops.ACT_QUICKGELU, name="ada.sp.fc")
            self._lin(self.mlp_s, s["pr_w"], s["pr_b"], self.sp, res=self.sp, name="ada.sp.proj")
            # temporal: query = clip token, keys = the t frame tokens + positional embedding (dist.py:155-160)
            self._ln(self.sp, tp["ln"], self.sp_ln, in2=e["pos"], in2_period=t, name="ada.tp.ln_kv")
            self._lin(self.sp_ln, tp["kv_w"], tp["kv_b"], self.kv_t, name="ada.tp.kv")
            if first:
                self.calls = self.static_calls
            self._ln(self.top0 if first else self.top, tp["ln"], self.top_ln, name="ada.tp.ln_q")
            self._lin(self.top_ln, tp["q_w"], tp["q_b"], self.q_t0 if first else self.q_t, name="ada.tp.q")
            self.calls = main
            add(ops.cross_attention(self.q_t0 if first else self.q_t, self.kv_t, self.o_t, b, t, H, name="ada.tp.attn"))
            self._lin(self.o_t, tp["o_w"], tp["o_b"], self.top, res=self.top0 if first else self.top, name="ada.tp.out_proj")
            self._ln(self.top, tp["ln_out"], self.top_ln, name="ada.tp.ln_out")
            self._lin(self.top_ln, tp["fc_w"], tp["fc_b"], self.mlp_t, act=ops.ACT_QUICKGELU, name="ada.tp.fc")
            self._lin(self.mlp_t, tp["pr_w"], tp["pr_b"], self.top, res=self.top, name="ada.tp.proj")
        if a.ada_layers == 0:                             # no pooling layer: the tail reads the constant clip token
            self.top = self.top0
        # ---- tail (dist.py:242-246) ----
        self._lin(self.clsmean, w.pcls_w, w.pcls_b, self.zbuf, res=self.top, name="tail.proj_spatial_cls")
        self._ln(self.zbuf, w.ln_post, self.z_ln, name="tail.ln_post")
        self._lin(self.z_ln, w.proj_w, None, self.emb, name="tail.proj")
        if self.image_logits:
            # ln_post of the last block's class token, times visual.proj (clip.py:291-298): rows N*D apart in the residual stream
            assert w.vis_ln_post is not None and w.vis_proj_w is not None, "the checkpoint has no visual.ln_post / visual.proj"
            add(ops.layernorm(self.h, w.vis_ln_post[0], w.vis_ln_post[1], self.img_ln, rows=F, cols=a.width, ld_in1=N * a.width, name="head.img.ln_post"))
            self._lin(self.img_ln, w.vis_proj_w, None, self.img_emb, name="head.img.proj")
        self.head_call = None
        if self.text_n is not None:
            # clip.py:511-518 + base_blocks.py:579-585 (+ clip.py:519-527 on the zero-shot branch)
            if self.image_logits == "fused":
                self.head_call = ops.class_head_fused(self.emb, self.img_emb, t, self.fusion_weight, self.text_n, w.logit_scale, b, a.embed_dim,
                                                      self.text_n.shape[0], self.logits, self.probs, img_n=self.img_emb_n, name="head.class_scores")
            else:
                self.head_call = ops.class_head(self.emb, self.text_n, w.logit_scale, b, a.embed_dim, self.text_n.shape[0],
                                                self.logits, self.probs, name="head.class_scores")
            add(self.head_call)

    # ------------------------------------------------------------------------------------------
    @property
    def num_launches(self):
        # patchify issues a second (padding) kernel when the patch row is padded - except on the staged bf16 path of the float clip
        staged = self.input_format == "float" and self.precision == "bf16"
        extra = 0 if staged else sum(1 for c in self.calls if c.name.startswith("patchify") and c.args[10] > 3 * c.args[6] * c.args[6])
        return len(self.calls) + extra

    def run(self, stream=None):
        """Replay the planned forward on ``stream`` (default: torch's current stream)."""
        s = (stream or torch.cuda.current_stream(self.device)).cuda_stream
        for c in self.calls:
            c.launch(s)

    def run_section(self, rng, stream=None):
        s = (stream or torch.cuda.current_stream(self.device)).cuda_stream
        for c in self.calls[rng[0]:rng[1]]:
            c.launch(s)

    def forward_vit(self, video):
        """The frozen ViT alone (``VisionTransformer.forward``, clip.py:263-300): returns the fp32 residual stream after every
        block, frame-major ``[b*t, N, D]`` (the reference's ``others["mid_feat"]["img"][l]`` is its ``[N, b*t, D]`` transpose)."""
        a = self.arch
        self.video.copy_(video, non_blocking=True)
        for name in ("patchify", "embed"):
            self.run_section(self.sections[name])
        taps = []
        for rng in self.sections["vit"]:
            self.run_section(rng)
            taps.append(self.h.view(self.batch * a.sparse_frames, a.tokens, a.width).clone())
        return taps

    def forward_dist(self, video, taps):
        """The DiST side alone (``DiSTNetwork.forward``, dist.py:222-247) on externally supplied taps: ``taps[l]`` is the
        frame-major ``[b*t, N, D]`` residual stream after ViT block ``l`` for every selected layer.  Returns ``[b, E]``."""
        a = self.arch
        self.video.copy_(video, non_blocking=True)
        self.run_section(self.sections["patchify"])
        self.run_section(self.sections["stem"])
        sel = list(a.selected_layers)
        for idx, l in enumerate(sel):
            tap = taps[l].reshape(-1, a.width)
            self.taps[l % 2].copy_(tap)                           # bf16 operand copy (fp32 path: the tap buffer is the stream itself)
            self.run_section(self.sections["dist"][idx])
        self.h.copy_(taps[sel[-1]].reshape(-1, a.width))           # the class-token mean reads the fp32 stream (dist.py:243)
        self.run_section(self.sections["cls_mean"])
        self.run_section(self.sections["head"])
        return self.emb

    # ------------------------------------------------------------------------------------------
    # Two-branch CUDA graph.  The ViT chain and the DiST chain only meet at the taps: DiST layer i reads the bf16 copy of
    # ViT block sel[i] (written by that block's FC2) and nothing of the DiST side flows back.  Captured on two streams the
    # chains become parallel branches of the graph: every kernel here is a persistent one-CTA-per-SM grid, so two kernels
    # never share an SM, but a kernel of one branch may fill the SMs that the other branch's kernel frees while it drains.
    # Optional (see capture()): it measured no faster than the linear graph.
    def _branches(self):
        return plan_branches([c.name for c in self.calls], list(self.arch.selected_layers))

    def run_branches(self, main, side):
        """Replay the plan with the ViT chain on ``main`` and the DiST chain on ``side`` (used under graph capture)."""
        branch, deps = self._branches()
        needed = {d for ds in deps.values() for d in ds}
        streams = (main, side)
        events = {}
        side.wait_stream(main)
        for k, c in enumerate(self.calls):
            st = streams[branch[k]]
            for d in deps.get(k, ()):
                if branch[d] != branch[k]:
                    st.wait_event(events[d])
            c.launch(st.cuda_stream)
            if k in needed:
                ev = torch.cuda.Event()
                ev.record(st)
                events[k] = ev
        main.wait_stream(side)

    def capture(self, branches=None):
        """Capture the plan into a CUDA graph (all buffers are static).  ``branches=2`` (or DISTB200_GRAPH_BRANCHES=2) captures
        the ViT and DiST chains as parallel branches; measured on B200 (B/16 8+16f, 32 clips): 15.01 ms vs 15.00 ms for the
        linear graph - the hardware already back-fills the drain of one persistent grid with the next launch - so the default
        stays the linear graph."""
        side = torch.cuda.Stream(self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            self.run(side)          # warm-up outside capture (function attributes, module load)
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        if branches is None:
            branches = int(os.environ.get("DISTB200_GRAPH_BRANCHES", "1"))
        two = self.precision == "bf16" and branches == 2
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            if two:
                self.run_branches(torch.cuda.current_stream(self.device), side)
            else:
                self.run(torch.cuda.current_stream(self.device))
        self.graph = graph
        self.graph_branches = 2 if two else 1
        return graph

    def forward(self, video=None, use_graph=True):
        """video [b, 3, T, H, W] fp32 - or [b, T, H, W, 3] uint8 for ``input_format="uint8"`` - (host or device)
        -> embedding [b, E] fp32 (view of an internal buffer)."""
        if video is not None:
            assert tuple(video.shape) == tuple(self.video.shape), (video.shape, self.video.shape)
            self.video.copy_(video, non_blocking=True)
        if use_graph and self.graph is not None:
            self.graph.replay()
        else:
            self.run()
        return self.emb

    def flops(self):
        return sum(c.flops for c in self.calls)
