"""CPU restatement of the DiST video forward path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this file.  The product (``dist_b200/``) never does: it fails loudly when its CUDA
extension is missing instead of falling back to anything in here.

What it is: the reference's algorithm for the path, written from its equations in a
frame-major token layout ``[b*t, N, D]`` and a channels-last temporal stream ``[b, T, g, g, Ct]``,
with plain torch CPU tensor algebra (``matmul``/``einsum``/``softmax``), in whatever dtype the
caller passes (float64 for parity references, float32 for the timed CPU baseline).  The reference
itself is PyTorch modules (``nn.Conv3d``, ``nn.MultiheadAttention`` ...); nothing here calls those.

Parity pin: the reference ships no tests or golden vectors.  The pin is
created by ``oracle/make_golden.py``, which imports the *unmodified* reference from
``/root/reference`` in the build container, loads the same synthetic weights, runs
``CLIP.forward_without_text`` / ``CLIP.forward`` (``models/base/clip.py:460-533``) on the same
synthetic clips and stores the outputs under ``tests/golden/``.  ``tests/test_oracle.py`` checks
this restatement against those fixtures (rel-L2 <= 2e-6 in float64 against the reference's fp32).

Reference lines each function follows are cited in its docstring (paths relative to
``/root/reference``).
"""

import math

import torch


def _ln(x, w, b, eps=1e-5):
    """LayerNorm over the last dim, biased variance (models/base/clip.py:181-187)."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def _qgelu(x):
    """QuickGELU (models/base/clip.py:199-201)."""
    return x * torch.sigmoid(1.702 * x)


def _lin(x, sd, prefix):
    return x @ sd[prefix + ".weight"].t() + sd[prefix + ".bias"]


def _mha(q_in, kv_in, sd, prefix, heads):
    """nn.MultiheadAttention forward without mask/dropout, batch-first inputs.

    q_in [B, Lq, C], kv_in [B, Lk, C].  Packed in-projection rows [0:C], [C:2C], [2C:3C] are
    q, k, v (torch.nn.functional.multi_head_attention_forward, as called at clip.py:155,168 and
    clip.py:139-147).
    """
    C = q_in.shape[-1]
    w, bias = sd[prefix + ".in_proj_weight"], sd[prefix + ".in_proj_bias"]
    q = q_in @ w[:C].t() + bias[:C]
    k = kv_in @ w[C:2 * C].t() + bias[C:2 * C]
    v = kv_in @ w[2 * C:].t() + bias[2 * C:]
    B, Lq, Lk, hd = q.shape[0], q.shape[1], k.shape[1], C // heads
    # contiguous per-head operands: torch's CPU bmm otherwise clones every (frame, head) slice separately (26 k small copies per clip),
    # which made this port 1.5 x slower than the reference's fused attention call on the same cores
    q = q.view(B, Lq, heads, hd).transpose(1, 2).contiguous()
    k = k.view(B, Lk, heads, hd).transpose(1, 2).contiguous()
    v = v.view(B, Lk, heads, hd).transpose(1, 2).contiguous()
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, Lq, C)
    return _lin(o, sd, prefix + ".out_proj")


def dims_from_state_dict(sd):
    D = sd["visual.conv1.weight"].shape[0]
    p = sd["visual.conv1.weight"].shape[-1]
    L = len([k for k in sd if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])
    N = sd["visual.positional_embedding"].shape[0]
    return dict(D=D, p=p, L=L, N=N, g=round((N - 1) ** 0.5))


def patchify(frames, p):
    """[F, 3, H, W] -> [F, P, 3*p*p] with the K order (c, i, j) of a conv weight [D, 3, p, p]."""
    F_, C, H, W = frames.shape
    g = H // p
    x = frames.view(F_, C, g, p, g, p).permute(0, 2, 4, 1, 3, 5)
    return x.reshape(F_, g * g, C * p * p)


def vit_forward(sd, video, alpha):
    """Frozen CLIP ViT over the sparse frames; returns the per-layer residual streams.

    video [b, 3, T, H, W].  Follows VisionTransformer.forward (models/base/clip.py:263-300):
    conv1 patch embedding (:271), class token + positional embedding (:274-275), ln_pre (:276),
    sparse frame pick ``[::alpha]`` (:281-284) - done here *before* the patch embedding, which
    gives the same kept frames -, then L x ResidualAttentionBlockMid (:170-178) whose outputs are
    the taps ``others["mid_feat"]["img"][l]`` (:177).  Taps are frame-major [b*t, N, D] with
    frame index b*t + ti (backbone.py:233, clip.py:284).
    """
    dm = dims_from_state_dict(sd)
    D, p, L = dm["D"], dm["p"], dm["L"]
    b, _, T, H, W = video.shape
    fs = video[:, :, ::alpha].permute(0, 2, 1, 3, 4).reshape(-1, 3, H, W)           # [b*t, 3, H, W]
    tok = patchify(fs, p) @ sd["visual.conv1.weight"].reshape(D, -1).t()             # [F, P, D]
    cls = sd["visual.class_embedding"].expand(tok.shape[0], 1, D)
    h = torch.cat([cls, tok], dim=1) + sd["visual.positional_embedding"]
    h = _ln(h, sd["visual.ln_pre.weight"], sd["visual.ln_pre.bias"])
    taps = []
    for l in range(L):
        pre = "visual.transformer.resblocks.%d" % l
        y = _ln(h, sd[pre + ".ln_1.weight"], sd[pre + ".ln_1.bias"])
        h = h + _mha(y, y, sd, pre + ".attn", D // 64)
        y = _ln(h, sd[pre + ".ln_2.weight"], sd[pre + ".ln_2.bias"])
        h = h + _lin(_qgelu(_lin(y, sd, pre + ".mlp.c_fc")), sd, pre + ".mlp.c_proj")
        taps.append(h)
    return taps


def temporal_stem(sd, video, ps):
    """Tubelet embedding of the dense clip (models/module_zoo/branches/dist.py:178-181,225).

    Conv3d(3 -> Ct, kernel (kt, ps, ps), stride (1, ps, ps), zero pad (kt//2, 0, 0)) written as a
    sum over temporal taps of patch GEMMs.  Returns [b, T, g, g, Ct].
    """
    w, bias = sd["dist_net.temporal_stem.weight"], sd["dist_net.temporal_stem.bias"]
    Ct, _, kt = w.shape[0], w.shape[1], w.shape[2]
    b, _, T, H, W = video.shape
    g = H // ps
    pat = patchify(video.permute(0, 2, 1, 3, 4).reshape(-1, 3, H, W), ps).view(b, T, g * g, -1)
    out = torch.zeros(b, T, g * g, Ct, dtype=video.dtype, device=video.device)
    for k in range(kt):
        shift = k - kt // 2                      # output frame tau reads input frame tau + shift
        lo, hi = max(0, -shift), min(T, T - shift)
        wk = w[:, :, k].reshape(Ct, -1)          # [Ct, 3*ps*ps]
        out[:, lo:hi] += _mm(pat[:, lo + shift:hi + shift], wk.t())
    return (out + bias).view(b, T, g, g, Ct)


def _mm(x, w_t):
    """``x @ w_t`` for a (possibly strided) view ``x [..., K]`` and a 2-D ``w_t [K, N]``: one copy of the view + one GEMM.  torch's CPU
    ``matmul`` on a non-contiguous N-d view expands the weight into a batched product and clones it per batch entry (25 k clones of a
    96 x 96 matrix per clip in the tap loops below) - same result, 2 x the time of the whole forward."""
    return (x.reshape(-1, x.shape[-1]) @ w_t).view(*x.shape[:-1], w_t.shape[1])


def temporal_net(sd, prefix, x):
    """TemporalNet.forward (dist.py:48-65): q(x + conv(1,3,3)(q(conv(kt,1,1)(LN_C(x))))).

    x [b, T, g, g, Ct] channels-last; both convolutions are dense, zero padded, with bias.
    """
    b, T, g, _, Ct = x.shape
    y = _ln(x, sd[prefix + ".ln.weight"], sd[prefix + ".ln.bias"])
    w1, b1 = sd[prefix + ".temporal_net.c_fc1.weight"], sd[prefix + ".temporal_net.c_fc1.bias"]
    kt = w1.shape[2]
    z = torch.zeros(b, T, g, g, w1.shape[0], dtype=x.dtype, device=x.device)
    for k in range(kt):
        s = k - kt // 2
        lo, hi = max(0, -s), min(T, T - s)
        z[:, lo:hi] += _mm(y[:, lo + s:hi + s], w1[:, :, k, 0, 0].t())
    z = _qgelu(z + b1)
    w2, b2 = sd[prefix + ".temporal_net.c_fc2.weight"], sd[prefix + ".temporal_net.c_fc2.bias"]
    o = torch.zeros(b, T, g, g, w2.shape[0], dtype=x.dtype, device=x.device)
    for i in range(3):
        for j in range(3):
            di, dj = i - 1, j - 1
            r0, r1 = max(0, -di), min(g, g - di)
            c0, c1 = max(0, -dj), min(g, g - dj)
            o[:, :, r0:r1, c0:c1] += _mm(z[:, :, r0 + di:r1 + di, c0 + dj:c1 + dj], w2[:, :, 0, i, j].t())
    return _qgelu(x + o + b2)


def integration_to_temporal(sd, prefix, mid, b, t, g, alpha):
    """Integration2TemporalNetwork.forward (dist.py:90-105): linear on the patch tokens,
    reshape to the frame grid, nearest upsample x alpha along time.  Returns [b, T, g, g, Ct]."""
    u = _lin(mid[:, 1:], sd, prefix + ".linear_fuse")                       # [b*t, P, Ct]
    u = u.view(b, t, g, g, -1)
    return u.repeat_interleave(alpha, dim=1)                                # out[tau] = in[tau // alpha]


def temporal_to_integration(sd, prefix, x, b, t, alpha):
    """Temporal2IntegrationNetwork.forward (dist.py:68-86): Conv3d(Ct -> Ci, (alpha,1,1), stride
    (alpha,1,1)) over frame groups, learned per-frame class token prepended.  Returns [b*t, N, Ci]."""
    w, bias = sd[prefix + ".linear_fuse.weight"], sd[prefix + ".linear_fuse.bias"]
    Ci = w.shape[0]
    _, T, g, _, Ct = x.shape
    xs = x.view(b, t, alpha, g * g, Ct)
    v = bias.expand(b, t, g * g, Ci).clone()
    for k in range(alpha):
        v = v + _mm(xs[:, :, k], w[:, :, k, 0, 0].t())
    cls = sd[prefix + ".cls_token"][0, 0].to(x.dtype)                       # [t, Ci]
    cls = cls.view(1, t, 1, Ci).expand(b, t, 1, Ci)
    return torch.cat([cls, v], dim=2).reshape(b * t, g * g + 1, Ci)


def integration_net(sd, prefix, x, b, t):
    """IntegrationNetwork.forward (dist.py:16-45): ffn(LN(x)) + temporal_ffn(LN_t(x)); the
    temporal branch is 1x1x1 conv Ci->Cm, (kt,1,1) conv along the sparse-frame axis for the same
    clip and token position, QuickGELU, 1x1x1 conv Cm->Ci.  x [b*t, N, Ci]; no skip from x."""
    N, Ci = x.shape[1], x.shape[2]
    a = _ln(x, sd[prefix + ".ln.weight"], sd[prefix + ".ln.bias"])
    f = _lin(_qgelu(_lin(a, sd, prefix + ".ffn.c_fc")), sd, prefix + ".ffn.c_proj")
    a2 = _ln(x, sd[prefix + ".ln_temporal.weight"], sd[prefix + ".ln_temporal.bias"])
    w1, b1 = sd[prefix + ".temporal_ffn.c_fc1.weight"], sd[prefix + ".temporal_ffn.c_fc1.bias"]
    z = (a2 @ w1[:, :, 0, 0, 0].t() + b1).view(b, t, N, -1)
    w2, b2 = sd[prefix + ".temporal_ffn.c_fc2.weight"], sd[prefix + ".temporal_ffn.c_fc2.bias"]
    kt = w2.shape[2]
    y = torch.zeros_like(z)
    for k in range(kt):
        s = k - kt // 2
        lo, hi = max(0, -s), min(t, t - s)
        y[:, lo:hi] += _mm(z[:, lo + s:hi + s], w2[:, :, k, 0, 0].t())
    y = _qgelu(y + b2)
    w3, b3 = sd[prefix + ".temporal_ffn.c_proj.weight"], sd[prefix + ".temporal_ffn.c_proj.bias"]
    return f + (y @ w3[:, :, 0, 0, 0].t() + b3).view(b * t, N, Ci)


def ada_pool(sd, prefix, cur, top, sp, b, t, heads):
    """SpatialTemporalAdaPoolingNetwork.forward (dist.py:139-162) with CrossAttentionBlockGenral
    (clip.py:139-147: the same ln_1 normalises query, key and value).

    cur [b*t, N, Ci]; top [b, 1, Ci]; sp [b*t, 1, Ci]."""
    pre_s, pre_t = prefix + ".spatial_transformer", prefix + ".temporal_transformer"
    ln_s = lambda z: _ln(z, sd[pre_s + ".ln_1.weight"], sd[pre_s + ".ln_1.bias"])
    ln_t = lambda z: _ln(z, sd[pre_t + ".ln_1.weight"], sd[pre_t + ".ln_1.bias"])
    sp = sp + _mha(ln_s(sp), ln_s(cur), sd, pre_s + ".attn", heads)
    y = _ln(sp, sd[prefix + ".ln_out_spat_cls_token.weight"], sd[prefix + ".ln_out_spat_cls_token.bias"])
    sp = sp + _lin(_qgelu(_lin(y, sd, prefix + ".output_map_spatial_cls_token.c_fc")),
                   sd, prefix + ".output_map_spatial_cls_token.c_proj")
    fr = sp[:, 0].view(b, t, -1) + sd[prefix + ".positional_embedding"].to(cur.dtype)   # [b, t, Ci]
    top = top + _mha(ln_t(top), ln_t(fr), sd, pre_t + ".attn", heads)
    y = _ln(top, sd[prefix + ".ln_out_temp_cls_token.weight"], sd[prefix + ".ln_out_temp_cls_token.bias"])
    top = top + _lin(_qgelu(_lin(y, sd, prefix + ".output_map_cls_token.c_fc")),
                     sd, prefix + ".output_map_cls_token.c_proj")
    return top, sp


def dist_forward(sd, video, taps, alpha, selected_layers, s_patch, ada_layers, return_parts=False):
    """DiSTNetwork.forward (dist.py:222-247).  Returns the video embedding [b, E]."""
    b, _, T, H, W = video.shape
    t = T // alpha
    g = H // s_patch
    Ci = sd["dist_net.proj"].shape[0]
    heads = Ci // 64
    xT = temporal_stem(sd, video, s_patch)
    res = None
    parts = {"stem": xT}
    for idx, l in enumerate(selected_layers):
        xT = temporal_net(sd, "dist_net.temporal_nets.%d" % idx, xT)
        mid = _lin(taps[l], sd, "dist_net.input_linears.%d" % idx)
        if res is not None:
            mid = mid + res
        xT_new = xT + integration_to_temporal(sd, "dist_net.integration2temporal_nets.%d" % idx, mid, b, t, g, alpha)
        upd = mid + temporal_to_integration(sd, "dist_net.temporal2integration_nets.%d" % idx, xT, b, t, alpha)
        res = integration_net(sd, "dist_net.integration_nets.%d" % idx, upd, b, t)
        xT = xT_new
        if return_parts:
            parts["xT.%d" % idx], parts["res.%d" % idx], parts["upd.%d" % idx] = xT, res, upd
    cur = res + upd
    top = sd["dist_net.aggregated_cls_token"].to(video.dtype).expand(b, 1, Ci)
    sp = sd["dist_net.aggregated_spatial_cls_token"].to(video.dtype).expand(b * t, 1, Ci)
    for j in range(ada_layers):
        top, sp = ada_pool(sd, "dist_net.adapooling_nets.%d" % j, cur, top, sp, b, t, heads)
        if return_parts:
            parts["top.%d" % j], parts["sp.%d" % j] = top, sp
    last = taps[selected_layers[-1]]
    cls_mean = last[:, 0].view(b, t, -1).mean(dim=1)
    z = top[:, 0] + _lin(cls_mean, sd, "dist_net.proj_spatial_cls_token")
    z = _ln(z, sd["dist_net.ln_post.weight"], sd["dist_net.ln_post.bias"])
    emb = z @ sd["dist_net.proj"]
    return (emb, parts) if return_parts else emb


def class_scores(sd, emb, text_features, softmax=True):
    """Cosine logits and head (clip.py:511-518, backbone.py:238-241, base_blocks.py:573-585)."""
    v = emb / emb.norm(dim=1, keepdim=True)
    tx = text_features / text_features.norm(dim=1, keepdim=True)
    logits = sd["logit_scale"].to(emb.dtype).exp() * v @ tx.t()
    return torch.softmax(logits, dim=-1) if softmax else logits


def image_embeddings(sd, last_tap):
    """Per-frame CLIP image embeddings (``VisionTransformer.forward``, clip.py:291-298): ``ln_post`` of the class token of the LAST
    ViT block, times ``visual.proj``.  last_tap [b*t, N, D] -> [b*t, E].  Returned as ``img_logits`` by ``CLIP.forward``
    (clip.py:532); DiST itself does not use them, the zero-shot / prediction-fusion branch does."""
    x = _ln(last_tap[:, 0], sd["visual.ln_post.weight"], sd["visual.ln_post.bias"])
    return x @ sd["visual.proj"]


def fused_class_scores(sd, emb, img, text_features, t, w=0.5):
    """Zero-shot / prediction fusion (clip.py:519-527): ``w`` x the video logits + ``(1 - w)`` x the mean over a clip's t sparse
    frames of the per-frame CLIP logits; both cosine logits share ``exp(logit_scale)``.  ``w`` = 0.5 unless the (never defined)
    gating parameter is enabled.  emb [b, E], img [b*t, E] -> logits [b, C]."""
    vid = class_scores(sd, emb, text_features, softmax=False)
    fr = class_scores(sd, img, text_features, softmax=False)
    return vid * w + fr.reshape(emb.shape[0], t, -1).mean(dim=1) * (1.0 - w)


def encode_text(sd, ids):
    """CLIP text tower (``CLIP.encode_text``, clip.py:419-434; blocks = ``ResidualAttentionBlock``, clip.py:112-136, with the
    causal additive mask of clip.py:404-410).  ids int64 [C, ctx] -> (features [C, E], eot rows [C, W])."""
    W = sd["ln_final.weight"].shape[0]
    heads = W // 64                                                                 # clip.py:590
    layers = len({k.split(".")[2] for k in sd if k.startswith("transformer.resblocks.")})
    x = sd["token_embedding.weight"][ids] + sd["positional_embedding"]              # [C, ctx, W]
    C, ctx, _ = x.shape
    mask = torch.full((ctx, ctx), float("-inf"), dtype=x.dtype).triu_(1)
    for l in range(layers):
        pre = "transformer.resblocks.%d." % l
        y = _ln(x, sd[pre + "ln_1.weight"], sd[pre + "ln_1.bias"])
        w, bias = sd[pre + "attn.in_proj_weight"], sd[pre + "attn.in_proj_bias"]
        q, k, v = (y @ w.t() + bias).split(W, dim=-1)
        sp = lambda z: z.view(C, ctx, heads, 64).transpose(1, 2)
        att = torch.softmax(sp(q) @ sp(k).transpose(-1, -2) / 8.0 + mask, dim=-1)
        o = (att @ sp(v)).transpose(1, 2).reshape(C, ctx, W)
        x = x + _lin(o, sd, pre + "attn.out_proj")
        y = _ln(x, sd[pre + "ln_2.weight"], sd[pre + "ln_2.bias"])
        x = x + _lin(_qgelu(_lin(y, sd, pre + "mlp.c_fc")), sd, pre + "mlp.c_proj")
    eot = x[torch.arange(C), ids.argmax(dim=-1)]                                     # clip.py:429
    feats = _ln(eot, sd["ln_final.weight"], sd["ln_final.bias"]) @ sd["text_projection"]
    return feats, eot


def forward(sd, video, alpha, selected_layers, s_patch, ada_layers, dtype=torch.float64, return_parts=False):
    """Whole path: [b, 3, T, H, W] -> [b, E] (= CLIP.forward_without_text(...)[:, 0], clip.py:466-480)."""
    sd = {k: v.to(dtype) for k, v in sd.items() if k.startswith(("visual.", "dist_net.", "logit_scale"))}
    video = video.to(dtype)
    taps = vit_forward(sd, video, alpha)
    out = dist_forward(sd, video, taps, alpha, selected_layers, s_patch, ada_layers, return_parts=return_parts)
    if return_parts:
        emb, parts = out
        for i, tp in enumerate(taps):
            parts["tap.%d" % i] = tp
        return emb, parts
    return out


def forward_arch(sd, video, arch, dtype=torch.float64, return_parts=False):
    """Convenience wrapper taking an object with the attributes of ``dist_b200.arch.DistArch``."""
    return forward(sd, video, arch.alpha, list(arch.selected_layers), arch.s_patch, arch.ada_layers,
                   dtype=dtype, return_parts=return_parts)
