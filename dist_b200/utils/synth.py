"""Deterministic synthetic weights, clips and label embeddings.

There is no network for checkpoints or datasets, so parity tests and the benchmark run on
random-init weights of the reference's architecture and on synthetic 224x224 clips.  Everything
here is generated from ``torch.Generator`` on the CPU with a fixed seed: the same call gives the
same tensors in the build container (where the golden fixtures are produced by running the
reference on them, ``oracle/make_golden.py``) and on the GPU box (where the CUDA path is checked
against those fixtures).  :func:`checksum` lets a test verify that claim before comparing outputs.

State-dict key names and shapes are the reference's weights contract (SURVEY.md section 8b;
``models/base/clip.py:218-261,303-372``, ``models/module_zoo/branches/dist.py:16-202``).

Two weight distributions:

``reference``  the distributions of the reference's own initialisers: truncated normal
               (std 0.02) linear/conv weights with zero biases, unit LayerNorms, Xavier-uniform
               attention in-projections, ``width**-0.5`` scaled embeddings
               (``clip.py:245-261``, ``dist.py:193-220``, ``torch.nn.MultiheadAttention``).
``scaled``     fan-in scaled normal weights, non-zero biases and non-trivial LayerNorm affine
               parameters, so that every branch of the network contributes O(1) to the output and
               a defect anywhere is visible in the final embedding.
"""

import math

import torch

from ..arch import DistArch


class _Gen:
    def __init__(self, seed, init):
        self.g = torch.Generator(device="cpu")
        self.g.manual_seed(int(seed))
        self.init = init

    def randn(self, *shape):
        return torch.randn(*shape, generator=self.g, dtype=torch.float32)

    def trunc(self, shape, std=0.02):
        # truncated normal on [-2, 2] (absolute), as timm's trunc_normal_ does for std << 2
        t = torch.empty(*shape, dtype=torch.float32)
        torch.nn.init.trunc_normal_(t, mean=0.0, std=std, a=-2.0, b=2.0, generator=self.g)
        return t

    def uniform(self, shape, bound):
        return (torch.rand(*shape, generator=self.g, dtype=torch.float32) * 2 - 1) * bound

    # ---- layer kinds ---------------------------------------------------------------------
    def weight(self, shape, fan_in):
        if self.init == "reference":
            return self.trunc(shape, 0.02)
        return self.randn(*shape) / math.sqrt(fan_in)

    def bias(self, n):
        if self.init == "reference":
            return torch.zeros(n)
        return 0.1 * self.randn(n)

    def ln(self, n):
        if self.init == "reference":
            return torch.ones(n), torch.zeros(n)
        return 1.0 + 0.1 * self.randn(n), 0.1 * self.randn(n)

    def token(self, shape, std=0.02):
        if self.init == "reference":
            return self.trunc(shape, std)
        return 0.5 * self.randn(*shape)

    def in_proj(self, dim):
        if self.init == "reference":
            bound = math.sqrt(6.0 / (dim + 3 * dim))  # xavier_uniform_ on [3*dim, dim]
            return self.uniform((3 * dim, dim), bound)
        return self.randn(3 * dim, dim) / math.sqrt(dim)


def _put_linear(sd, g, prefix, n_out, n_in):
    sd[prefix + ".weight"] = g.weight((n_out, n_in), n_in)
    sd[prefix + ".bias"] = g.bias(n_out)


def _put_ln(sd, g, prefix, n):
    sd[prefix + ".weight"], sd[prefix + ".bias"] = g.ln(n)


def _put_mha(sd, g, prefix, dim):
    sd[prefix + ".in_proj_weight"] = g.in_proj(dim)
    sd[prefix + ".in_proj_bias"] = g.bias(3 * dim)
    _put_linear(sd, g, prefix + ".out_proj", dim, dim)


def synth_state_dict(arch: DistArch, seed=0, init="reference"):
    """All ``visual.*`` and ``dist_net.*`` tensors plus ``logit_scale`` (fp32, CPU)."""
    assert init in ("reference", "scaled")
    g = _Gen(seed, init)
    sd = {}
    D, L, p, N, E = arch.width, arch.layers, arch.patch, arch.tokens, arch.embed_dim
    Ci, Ct, t = arch.integration_dim, arch.temporal_dim, arch.sparse_frames
    scale = D ** -0.5

    # ---- CLIP ViT (clip.py:218-261) ----
    sd["visual.class_embedding"] = scale * g.randn(D)
    sd["visual.positional_embedding"] = scale * g.randn(N, D)
    sd["visual.proj"] = scale * g.randn(D, E)
    kfan = 3 * p * p
    sd["visual.conv1.weight"] = (g.uniform((D, 3, p, p), 1.0 / math.sqrt(kfan)) if init == "reference"
                                 else g.randn(D, 3, p, p) / math.sqrt(kfan))
    _put_ln(sd, g, "visual.ln_pre", D)
    _put_ln(sd, g, "visual.ln_post", D)
    for i in range(L):
        pre = "visual.transformer.resblocks.%d" % i
        _put_mha(sd, g, pre + ".attn", D)
        _put_ln(sd, g, pre + ".ln_1", D)
        _put_linear(sd, g, pre + ".mlp.c_fc", 4 * D, D)
        _put_linear(sd, g, pre + ".mlp.c_proj", D, 4 * D)
        _put_ln(sd, g, pre + ".ln_2", D)

    # ---- DiST (dist.py:165-202) ----
    ps, pt, kt = arch.s_patch, arch.t_patch, arch.t_kernel
    Ch, Ih, Cm = arch.temporal_hidden, arch.integration_hidden, arch.integration_temporal_hidden
    sd["dist_net.proj"] = (Ci ** -0.5) * g.randn(Ci, E)
    sd["dist_net.aggregated_cls_token"] = g.token((1, 1, Ci))
    sd["dist_net.aggregated_spatial_cls_token"] = g.token((1, 1, Ci))
    sd["dist_net.temporal_stem.weight"] = g.weight((Ct, 3, pt, ps, ps), 3 * pt * ps * ps)
    sd["dist_net.temporal_stem.bias"] = g.bias(Ct)
    _put_linear(sd, g, "dist_net.proj_spatial_cls_token", Ci, D)
    _put_ln(sd, g, "dist_net.ln_post", Ci)
    for i in range(len(arch.selected_layers)):
        _put_linear(sd, g, "dist_net.input_linears.%d" % i, Ci, D)
        _put_linear(sd, g, "dist_net.integration2temporal_nets.%d.linear_fuse" % i, Ct, Ci)
        pre = "dist_net.temporal2integration_nets.%d" % i
        sd[pre + ".cls_token"] = g.token((1, 1, t, Ci))
        sd[pre + ".linear_fuse.weight"] = g.weight((Ci, Ct, arch.alpha, 1, 1), Ct * arch.alpha)
        sd[pre + ".linear_fuse.bias"] = g.bias(Ci)
        pre = "dist_net.temporal_nets.%d" % i
        _put_ln(sd, g, pre + ".ln", Ct)
        sd[pre + ".temporal_net.c_fc1.weight"] = g.weight((Ch, Ct, kt, 1, 1), Ct * kt)
        sd[pre + ".temporal_net.c_fc1.bias"] = g.bias(Ch)
        sd[pre + ".temporal_net.c_fc2.weight"] = g.weight((Ct, Ch, 1, 3, 3), Ch * 9)
        sd[pre + ".temporal_net.c_fc2.bias"] = g.bias(Ct)
        pre = "dist_net.integration_nets.%d" % i
        _put_ln(sd, g, pre + ".ln", Ci)
        _put_ln(sd, g, pre + ".ln_temporal", Ci)
        _put_linear(sd, g, pre + ".ffn.c_fc", Ih, Ci)
        _put_linear(sd, g, pre + ".ffn.c_proj", Ci, Ih)
        sd[pre + ".temporal_ffn.c_fc1.weight"] = g.weight((Cm, Ci, 1, 1, 1), Ci)
        sd[pre + ".temporal_ffn.c_fc1.bias"] = g.bias(Cm)
        sd[pre + ".temporal_ffn.c_fc2.weight"] = g.weight((Cm, Cm, kt, 1, 1), Cm * kt)
        sd[pre + ".temporal_ffn.c_fc2.bias"] = g.bias(Cm)
        sd[pre + ".temporal_ffn.c_proj.weight"] = g.weight((Ci, Cm, 1, 1, 1), Cm)
        sd[pre + ".temporal_ffn.c_proj.bias"] = g.bias(Ci)
    for j in range(arch.ada_layers):
        pre = "dist_net.adapooling_nets.%d" % j
        sd[pre + ".positional_embedding"] = g.token((1, t, Ci))
        for which in ("temporal_transformer", "spatial_transformer"):
            _put_mha(sd, g, pre + "." + which + ".attn", Ci)
            _put_ln(sd, g, pre + "." + which + ".ln_1", Ci)
        for which in ("output_map_cls_token", "output_map_spatial_cls_token"):
            _put_linear(sd, g, pre + "." + which + ".c_fc", 4 * Ci, Ci)
            _put_linear(sd, g, pre + "." + which + ".c_proj", Ci, 4 * Ci)
        _put_ln(sd, g, pre + ".ln_out_temp_cls_token", Ci)
        _put_ln(sd, g, pre + ".ln_out_spat_cls_token", Ci)
    sd["logit_scale"] = torch.tensor(math.log(1 / 0.07), dtype=torch.float32)  # clip.py:371
    return sd


def synth_clips(batch, arch: DistArch, seed=1234, kind="structured"):
    """``[b, 3, T, H, W]`` fp32 clips.

    ``iid``         unit normal noise.
    ``structured``  a per-clip low-frequency image (8x8 noise, bilinear) that translates with a
                    per-clip velocity, plus a little noise, normalised to unit variance; clips
                    then differ enough for top-1 to be a meaningful comparison.
    """
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    T, R = arch.frames, arch.resolution
    if kind == "iid":
        return torch.randn(batch, 3, T, R, R, generator=g, dtype=torch.float32)
    assert kind == "structured"
    coarse = torch.randn(batch, 3, 8, 8, generator=g, dtype=torch.float32)
    canvas = torch.nn.functional.interpolate(coarse, size=(2 * R, 2 * R), mode="bilinear", align_corners=False)
    vel = (torch.rand(batch, 2, generator=g) * 2 - 1) * (R / max(T, 1)) * 0.8
    noise = 0.1 * torch.randn(batch, 3, T, R, R, generator=g, dtype=torch.float32)
    out = torch.empty(batch, 3, T, R, R, dtype=torch.float32)
    for b in range(batch):
        for ti in range(T):
            oy = int(round(R / 2 + float(vel[b, 0]) * (ti - T / 2)))
            ox = int(round(R / 2 + float(vel[b, 1]) * (ti - T / 2)))
            oy, ox = max(0, min(R, oy)), max(0, min(R, ox))
            out[b, :, ti] = canvas[b, :, oy:oy + R, ox:ox + R]
    out = out + noise
    out = out - out.mean(dim=(1, 2, 3, 4), keepdim=True)
    out = out / out.std(dim=(1, 2, 3, 4), keepdim=True)
    return out.contiguous()


def synth_text_features(num_classes, embed_dim, seed=77):
    """Stand-in for the cached CLIP label embeddings ``[C, E]`` (``clip.py:437-452``)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return torch.randn(num_classes, embed_dim, generator=g, dtype=torch.float32)


def synth_text_tower(embed_dim, width=512, layers=12, context=77, vocab=49408, seed=5, init="reference"):
    """The CLIP text tower's tensors under the reference's key names (``clip.py:362-372``), drawn from the distributions of
    ``CLIP.initialize_parameters`` (``clip.py:375-402``) for ``init="reference"``: token embedding N(0, 0.02), positional
    embedding N(0, 0.01), attention / MLP weights N(0, width^-0.5 ...), ``text_projection`` N(0, width^-0.5)."""
    assert width % 64 == 0 and init in ("reference", "scaled")
    g = _Gen(seed, init)
    sd = {}
    ref = init == "reference"
    sd["token_embedding.weight"] = (0.02 if ref else 0.5) * g.randn(vocab, width)
    sd["positional_embedding"] = (0.01 if ref else 0.5) * g.randn(context, width)
    proj_std = (width ** -0.5) * ((2 * layers) ** -0.5)
    attn_std, fc_std = width ** -0.5, (2 * width) ** -0.5
    for i in range(layers):
        pre = "transformer.resblocks.%d" % i
        sd[pre + ".attn.in_proj_weight"] = attn_std * g.randn(3 * width, width)
        sd[pre + ".attn.in_proj_bias"] = g.bias(3 * width)
        sd[pre + ".attn.out_proj.weight"] = (proj_std if ref else attn_std) * g.randn(width, width)
        sd[pre + ".attn.out_proj.bias"] = g.bias(width)
        _put_ln(sd, g, pre + ".ln_1", width)
        sd[pre + ".mlp.c_fc.weight"] = (fc_std if ref else attn_std) * g.randn(4 * width, width)
        sd[pre + ".mlp.c_fc.bias"] = g.bias(4 * width)
        sd[pre + ".mlp.c_proj.weight"] = (proj_std if ref else (4 * width) ** -0.5) * g.randn(width, 4 * width)
        sd[pre + ".mlp.c_proj.bias"] = g.bias(width)
        _put_ln(sd, g, pre + ".ln_2", width)
    _put_ln(sd, g, "ln_final", width)
    sd["text_projection"] = (width ** -0.5) * g.randn(width, embed_dim)
    return sd


def synth_token_ids(num_prompts, context=77, vocab=49408, seed=11):
    """Token ids shaped like the CLIP tokenizer's output (``dataset/utils/simple_tokenizer.py``): start-of-text = vocab-2,
    1..context-2 word tokens, end-of-text = vocab-1 (the largest id, which ``encode_text`` locates by argmax), zero padding."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    ids = torch.zeros(num_prompts, context, dtype=torch.long)
    for r in range(num_prompts):
        n = int(torch.randint(1, context - 1, (1,), generator=g))
        ids[r, 0] = vocab - 2
        ids[r, 1:1 + n] = torch.randint(1, vocab - 2, (n,), generator=g)
        ids[r, 1 + n] = vocab - 1
    return ids


def synth_soft_targets(batch, num_classes, seed=99, smoothing=0.1):
    """Soft labels ``[b, C]`` of the kind mixup / cutmix + label smoothing produce (``dataset/utils/mixup.py:103-319``):
    each row mixes two smoothed one-hot labels with a Beta(0.8, 0.8)-like weight and sums to one."""
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    a = torch.randint(0, num_classes, (batch,), generator=g)
    b = torch.randint(0, num_classes, (batch,), generator=g)
    lam = torch.rand(batch, generator=g, dtype=torch.float32).clamp(0.05, 0.95)
    off, on = smoothing / num_classes, 1.0 - smoothing + smoothing / num_classes

    def one_hot(idx):
        t = torch.full((batch, num_classes), off, dtype=torch.float32)
        t.scatter_(1, idx.view(-1, 1), on)
        return t

    return lam.view(-1, 1) * one_hot(a) + (1.0 - lam.view(-1, 1)) * one_hot(b)


def checksum(tensors):
    """Order-independent fingerprint of a tensor or dict of tensors (float64 sums)."""
    if isinstance(tensors, torch.Tensor):
        tensors = {"_": tensors}
    s1 = s2 = 0.0
    for k in sorted(tensors):
        v = tensors[k].detach().double()
        s1 += float(v.sum())
        s2 += float(v.abs().sum())
    return [s1, s2]
