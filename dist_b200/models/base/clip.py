"""CLIP vision tower + DiST wiring on the distb200 CUDA path (reference: ``models/base/clip.py``).

The modules below own parameters under exactly the reference's ``state_dict`` names
(SURVEY.md section 8b), so reference checkpoints load unchanged; their arithmetic runs in
``dist_b200.engine.DistEngine`` through the C ABI - torch layers such as ``nn.Linear`` or
``nn.MultiheadAttention`` appear here only as parameter containers and are never called.
Label embeddings enter either as token ids ``[C, ctx]`` - encoded ONCE by the CLIP text tower on the same CUDA library
(``dist_b200/text.py``; ``cache_text``, ``clip.py:436-452``) when the checkpoint carries the tower - or as a constant
``[C, E]`` matrix (``set_text_features`` / ``others['label_embeddings']``, ``clip.py:437-439``).
"""

from collections import OrderedDict
import math
import os

import numpy as np
import torch
from torch import nn

from ...arch import arch_from_cfg
from ...registry import Registry
from ...utils import synth

TEMPORALNET_REGISTRY = Registry("TemporalNet")
ATTEN_BLOCK_REGISTRY = Registry("AttentionBlock")


class LayerNorm(nn.LayerNorm):
    """Parameter holder; statistics are always fp32 on the CUDA path (``clip.py:181-187``)."""


class QuickGELU(nn.Module):
    """x * sigmoid(1.702 x) - fused into GEMM epilogues (``clip.py:199-201``)."""

    def forward(self, x):
        raise RuntimeError("QuickGELU is fused into the distb200 GEMM epilogue; it is not called as a module")


class CrossAttentionBlockGenral(nn.Module):
    """Single-query cross attention block of the ada-pooling head (``clip.py:139-147``)."""

    def __init__(self, d_model, n_head, attn_mask=None, cfg=None, layer_id=0):
        super().__init__()
        self.layer_id = layer_id
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)


class ResidualAttentionBlock(nn.Module):
    """Block of the text transformer (``clip.py:112-136``, built with ``cfg=None`` at ``clip.py:205-208``): parameter holder,
    evaluated by ``dist_b200.text.TextEngine`` with the causal mask of ``clip.py:404-410``."""

    def __init__(self, d_model, n_head, attn_mask=None, cfg=None, layer_id=0):
        super().__init__()
        self.layer_id = layer_id
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, d_model * 4)),
            ("gelu", QuickGELU()),
            ("c_proj", nn.Linear(d_model * 4, d_model)),
        ]))
        self.ln_2 = LayerNorm(d_model)
        self.is_image_transformer = attn_mask is None


@ATTEN_BLOCK_REGISTRY.register()
class ResidualAttentionBlockMid(nn.Module):
    """ViT block whose output is tapped for DiST (``clip.py:150-178``).

    Block protocol of the reference: ctor ``(d_model, n_head, attn_mask, cfg=, layer_id=)``; inside the fused
    engine the tap ``others["mid_feat"]["img"][layer_id]`` is the bf16 copy the FC2 epilogue writes.
    """

    def __init__(self, d_model, n_head, attn_mask=None, cfg=None, layer_id=0):
        super().__init__()
        assert attn_mask is None, "the image transformer runs without a mask"
        self.layer_id = layer_id
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([
            ("c_fc", nn.Linear(d_model, d_model * 4)),
            ("gelu", QuickGELU()),
            ("c_proj", nn.Linear(d_model * 4, d_model)),
        ]))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = None
        self.is_image_transformer = True
        self._packed = None

    def forward(self, xo):
        """Block protocol of the reference (``clip.py:170-178``): ``(x [N, b*t, D], others) -> (x, others)`` with
        ``x += MHA(ln_1 x)``, ``x += mlp(ln_2 x)`` and the tap ``others["mid_feat"]["img"][layer_id] = x.clone()``.
        Stand-alone (eager) use of the same kernels the planned engine launches; fp32 residual stream, bf16 operands."""
        from ... import ops
        x, others = xo
        if not x.is_cuda:
            raise RuntimeError("dist_b200 runs on a CUDA device only; there is no CPU path")
        n_tok, bt, d = x.shape
        heads = self.attn.num_heads
        version = sum(p._version for p in self.parameters())
        if self._packed is None or self._packed[0] != version or self._packed[1] != x.device:
            bf = lambda t: t.detach().to(device=x.device, dtype=torch.bfloat16).contiguous()
            f32 = lambda t: t.detach().to(device=x.device, dtype=torch.float32).contiguous()
            self._packed = (version, x.device, dict(
                qkv_w=bf(self.attn.in_proj_weight), qkv_b=f32(self.attn.in_proj_bias), proj_w=bf(self.attn.out_proj.weight),
                proj_b=f32(self.attn.out_proj.bias), fc1_w=bf(self.mlp.c_fc.weight), fc1_b=f32(self.mlp.c_fc.bias),
                fc2_w=bf(self.mlp.c_proj.weight), fc2_b=f32(self.mlp.c_proj.bias),
                ln1=(f32(self.ln_1.weight), f32(self.ln_1.bias)), ln2=(f32(self.ln_2.weight), f32(self.ln_2.bias))))
        w = self._packed[2]
        st = torch.cuda.current_stream(x.device).cuda_stream
        h = x.detach().float().permute(1, 0, 2).reshape(bt * n_tok, d).contiguous()          # frame-major rows
        rows = h.shape[0]
        z = lambda *shape, dtype=torch.bfloat16: torch.empty(*shape, device=x.device, dtype=dtype)
        ln_buf, qkv, att, fc1 = z(rows, d), z(rows, 3 * d), z(rows, d), z(rows, 4 * d)
        lin = lambda a, wt, b, out, **kw: ops.gemm(a, wt, wt.shape[0], wt.shape[1], bias=b, out=out, ld_out=wt.shape[0], ld_res=wt.shape[0], **kw).launch(st)
        ops.layernorm(h, w["ln1"][0], w["ln1"][1], ln_buf).launch(st)
        lin(ln_buf, w["qkv_w"], w["qkv_b"], qkv)
        ops.attention(qkv, att, bt, n_tok, heads).launch(st)
        lin(att, w["proj_w"], w["proj_b"], h, res=h)
        ops.layernorm(h, w["ln2"][0], w["ln2"][1], ln_buf).launch(st)
        lin(ln_buf, w["fc1_w"], w["fc1_b"], fc1, act=ops.ACT_QUICKGELU)
        lin(fc1, w["fc2_w"], w["fc2_b"], h, res=h)
        out = h.view(bt, n_tok, d).permute(1, 0, 2).contiguous()
        if others is not None and "mid_feat" in others and "img" in others["mid_feat"]:
            others["mid_feat"]["img"][self.layer_id] = out.clone()                           # clip.py:177
        return out, others


class Transformer(nn.Module):
    def __init__(self, width, layers, heads, attn_mask=None, cfg=None):
        super().__init__()
        self.width, self.layers = width, layers
        if cfg is None:                                                    # the text tower (clip.py:208-209)
            block = ResidualAttentionBlock
        else:
            name = cfg.VIDEO.BACKBONE.ATTEN_BLOCK
            block = ATTEN_BLOCK_REGISTRY.get(name)
            if block is None:
                raise KeyError("attention block {!r} is not registered (the DiST configs select ResidualAttentionBlockMid)".format(name))
        self.resblocks = nn.Sequential(*[block(width, heads, attn_mask, cfg=cfg, layer_id=i) for i in range(layers)])


class VisionTransformer(nn.Module):
    """CLIP ViT parameters (``clip.py:218-261``)."""

    def __init__(self, cfg, input_resolution, patch_size, width, layers, heads, output_dim):
        super().__init__()
        self.cfg = cfg
        self.input_resolution, self.output_dim = input_resolution, output_dim
        self.num_frames = cfg.DATA.NUM_INPUT_FRAMES
        self.sparse_sample_alpha = getattr(cfg.DATA, "SPARSE_SAMPLE_ALPHA", 1)
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads, cfg=cfg)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))
        self._engines = {}

    def forward(self, x, others=None):
        """``VisionTransformer.forward`` of the reference (``clip.py:263-300``) as a stand-alone module: frames ``[b*T,3,H,W]``
        -> ``(cls_x, x_logits, x[:, 1:], others)`` and, when ``others["mid_feat"]["img"]`` exists, the per-block taps
        ``[N, b*t, D]`` (``clip.py:177``).  The blocks run on the CUDA path; ``ln_post`` / ``proj`` on the class tokens of the
        kept frames (dead for DiST, ``clip.py:293-298``) are evaluated with the same kernels."""
        from ... import ops
        from ...engine import DistEngine
        if not x.is_cuda:
            raise RuntimeError("dist_b200 runs on a CUDA device only; there is no CPU path")
        bt, c, hh, ww = x.shape
        T = self.num_frames
        b = bt // T
        arch = arch_from_cfg(self.cfg, {"visual." + k: v for k, v in self.state_dict().items()})
        key = (b, str(x.device), sum(p._version for p in self.parameters()))
        if key not in self._engines:
            self._engines.clear()
            sd = synth.synth_state_dict(arch, seed=0)              # the DiST half is never run here; any weights will do
            sd.update({"visual." + k: v.detach() for k, v in self.state_dict().items()})
            b200 = getattr(self.cfg, "B200", None)
            self._engines[key] = DistEngine(sd, arch, b, device=x.device, precision=getattr(b200, "PRECISION", "bf16") if b200 else "bf16")
        eng = self._engines[key]
        taps = eng.forward_vit(x.view(b, T, c, hh, ww).permute(0, 2, 1, 3, 4).float().contiguous())
        if others is not None and "mid_feat" in others and "img" in others["mid_feat"]:
            for l, tp in enumerate(taps):
                others["mid_feat"]["img"][l] = tp.permute(1, 0, 2).contiguous()
        last = taps[-1]                                                                  # [b*t, N, D]
        cls = last[:, 0].contiguous()
        x_logits = torch.empty_like(cls)
        st = torch.cuda.current_stream(x.device).cuda_stream
        ops.layernorm(cls, self.ln_post.weight.detach().float(), self.ln_post.bias.detach().float(), x_logits).launch(st)
        pw = self.proj.detach().float().t().contiguous()                               # [E, D]
        cls_x = torch.empty(cls.shape[0], pw.shape[0], device=x.device)
        ops.gemm(x_logits, pw, pw.shape[0], pw.shape[1], out=cls_x, ld_out=pw.shape[0]).launch(st)
        return cls_x, x_logits, last[:, 1:, :], others


class CLIP(nn.Module):
    """``visual`` + ``dist_net`` + ``logit_scale``; forwards run the planned CUDA path."""

    def __init__(self, cfg, embed_dim, image_resolution, vision_layers, vision_width, vision_patch_size, arch=None,
                 context_length=None, vocab_size=None, transformer_width=None, transformer_heads=None, transformer_layers=None):
        """The five text arguments are those of the reference (``clip.py:303-320``); ``None`` builds a model without the text
        tower (label embeddings are then supplied as a matrix)."""
        super().__init__()
        from ..module_zoo.branches.dist import DiSTNetwork
        self.cfg = cfg
        self.num_frames = cfg.DATA.NUM_INPUT_FRAMES
        self.freeze_text = cfg.VIDEO.BACKBONE.FREEZE_TEXT
        self.freeze_visual = getattr(cfg.VIDEO.BACKBONE, "FREEZE_VISUAL", True)
        self.num_classes = cfg.VIDEO.HEAD.NUM_CLASSES
        self.visual = VisionTransformer(cfg, image_resolution, vision_patch_size, vision_width, vision_layers,
                                        vision_width // 64, embed_dim)
        self.dist_net = DiSTNetwork(cfg, d_model=vision_width, width=vision_width, output_dim=embed_dim)
        self.context_length, self.vocab_size = context_length, vocab_size
        if transformer_width is not None:                                   # clip.py:360-370
            self.transformer = Transformer(transformer_width, transformer_layers, transformer_heads, attn_mask="causal")
            self.token_embedding = nn.Embedding(vocab_size, transformer_width)
            self.positional_embedding = nn.Parameter(torch.empty(context_length, transformer_width).normal_(std=0.01))
            self.ln_final = LayerNorm(transformer_width)
            self.text_projection = nn.Parameter(torch.empty(transformer_width, embed_dim).normal_(std=transformer_width ** -0.5))
        self.logit_scale = nn.Parameter(torch.ones([]) * np.log(1 / 0.07))
        self.arch = arch
        self._text_engine = None
        self._text_cache = None
        b200 = getattr(cfg, "B200", None)
        self.precision = getattr(b200, "PRECISION", "bf16") if b200 is not None else "bf16"
        self.use_graph = bool(getattr(b200, "CUDA_GRAPH", True)) if b200 is not None else True
        self.text_features = None
        test_cfg = getattr(cfg, "TEST", None)
        zs = getattr(test_cfg, "ZEROSHOT", None) if test_cfg is not None else None
        self.zero_shot_test = bool(getattr(zs, "ENABLE", False)) if zs is not None else False              # clip.py:327
        self.want_img_logits = bool(getattr(b200, "IMG_LOGITS", False)) if b200 is not None else False
        self._engines = {}

    # ---- label embeddings ---------------------------------------------------------------------
    def set_text_features(self, feats):
        self.text_features = feats.detach().float()
        self._engines.clear()

    @property
    def has_text_tower(self):
        return hasattr(self, "token_embedding")

    def encode_text(self, text, others=None):
        """``CLIP.encode_text`` (``clip.py:419-434``): ids int64 ``[C, ctx]`` -> ``(features [C, E], eot rows [C, W], others)``."""
        from ...text import TextEngine
        if not self.has_text_tower:
            raise RuntimeError("this model was built without the CLIP text tower (its checkpoint has no transformer.* / "
                               "token_embedding keys); supply label embeddings with set_text_features([C, E])")
        dev = self.logit_scale.device
        if dev.type != "cuda":
            raise RuntimeError("dist_b200 runs on a CUDA device only; there is no CPU path (move the model with .cuda())")
        names = ("transformer.", "token_embedding.", "positional_embedding", "ln_final.", "text_projection")
        version = sum(p._version for n, p in self.named_parameters() if n.startswith(names))
        key = (text.shape[0], str(dev), self.precision, version)
        if self._text_engine is None or self._text_engine[0] != key:
            sd = {k: v for k, v in self.state_dict().items() if k.startswith(names)}
            self._text_engine = (key, TextEngine(sd, text.shape[0], device=dev, precision=self.precision))
        feats, eot = self._text_engine[1].encode(text)
        return feats.clone(), eot.clone(), others

    def cache_text(self, text, others=None):
        """``CLIP.cache_text`` (``clip.py:436-452``): the label set is encoded once and reused while its size is unchanged."""
        if others is not None and "label_embeddings" in others:            # clip.py:437-439
            return others["label_embeddings"], None, others
        names = ("transformer.", "token_embedding.", "positional_embedding", "ln_final.", "text_projection")
        version = sum(p._version for n, p in self.named_parameters() if n.startswith(names))     # a load_state_dict after the first call invalidates
        if self._text_cache is None or self._text_cache[0].shape[0] != text.shape[0] or self._text_cache[2] != version:
            feats, eot, others = self.encode_text(text, others)
            self._text_cache = (feats, eot, version)
        return self._text_cache[0], self._text_cache[1], others

    def _text_from(self, text, others):
        if others is not None and "label_embeddings" in others:            # clip.py:437-439
            return others["label_embeddings"]
        if text is not None and torch.is_floating_point(text):
            return text
        if self.text_features is not None:
            return self.text_features
        if text is not None and self.has_text_tower:
            if self.freeze_text:                                             # clip.py:483-486
                return self.cache_text(text, others)[0]
            return self.encode_text(text, others)[0]
        raise NotImplementedError(
            "token ids were passed, but this model has no CLIP text tower (its checkpoint carried none) and no label "
            "embeddings are cached - call set_text_features([C, E]) with embeddings computed once")

    # ---- engine cache -------------------------------------------------------------------------
    def _image_logits_mode(self):
        """``"fused"`` on the zero-shot test branch (``TEST.ZEROSHOT.ENABLE`` in eval, clip.py:327,519), ``"raw"`` when the config asks
        for the per-frame CLIP embeddings (``B200.IMG_LOGITS``), else ``None``: DiST never reads them (clip.py:291-298 is dead code on
        its path), so they are not computed unless wanted."""
        if self.zero_shot_test and not self.training:
            return "fused"
        return "raw" if self.want_img_logits else None

    def _engine(self, batch, device, text, input_format="float"):
        from ...engine import DistEngine
        version = sum(p._version for p in self.parameters())
        tkey = None if text is None else (text.data_ptr(), tuple(text.shape), text._version)
        mode = self._image_logits_mode() if text is not None else None
        key = (batch, str(device), self.precision, input_format, mode)
        hit = self._engines.get(key)
        if hit is not None and hit[1] == version and hit[2] == tkey:
            return hit[0]
        if hit is not None and hit[1] == version and text is not None and hit[0].text_n is not None and hit[0].text_n.shape == text.shape:
            # only the label matrix changed (FREEZE_TEXT false, or fresh others["label_embeddings"] per call): refresh it in place
            # instead of re-packing every weight and re-capturing the graph
            hit[0].set_text_features(text)
            self._engines[key] = (hit[0], version, tkey)
            return hit[0]
        sd = {k: v for k, v in self.state_dict().items()}
        eng = DistEngine(sd, self.arch, batch, device=device, precision=self.precision, text_features=text, input_format=input_format,
                         mean=getattr(self.cfg.DATA, "MEAN", None), std=getattr(self.cfg.DATA, "STD", None), image_logits=mode)
        if self.use_graph:
            eng.capture()
        self._engines[key] = (eng, version, tkey)
        return eng

    def refresh_engine(self):
        self._engines.clear()

    # ---- forwards (clip.py:460-533) -------------------------------------------------------------
    def forward(self, image, text, others=None):
        if self.training and torch.is_grad_enabled():
            raise RuntimeError(
                "dist_b200: this module's forward is the planned CUDA inference path and builds no autograd graph, so "
                "loss.backward() cannot train it.  Fine-tune with dist_b200.runs.train.Trainer (planned forward + backward + "
                "all-reduce + AdamW, runs/train.py:97-112 of the reference), or call model.eval() / torch.no_grad() for inference.")
        if text is not None or (others is not None and "label_embeddings" in others):
            return self.forward_with_text(image, text, others)
        return self.forward_without_text(image)

    def _as_clips(self, image):
        """Accept the reference's frame-major ``[B*T,3,H,W]`` (clip.py:460), the native ``[B,3,T,H,W]``, or decoded uint8
        frames ``[B,T,H,W,3]`` (normalised inside the patch-row kernel with DATA.MEAN / DATA.STD)."""
        if image.dtype == torch.uint8:
            if image.dim() != 5 or image.shape[-1] != 3:
                raise ValueError("uint8 clips must be [B, T, H, W, 3] (decoder layout), got {}".format(tuple(image.shape)))
            return image
        if image.dim() == 4:
            bt, c, h, w = image.shape
            image = image.view(bt // self.num_frames, self.num_frames, c, h, w).permute(0, 2, 1, 3, 4).contiguous()
        return image

    def _device_for(self, image):
        """Clips may live on the model's GPU or in pinned host memory (copied straight into the engine's buffer)."""
        dev = self.logit_scale.device
        if dev.type != "cuda":
            raise RuntimeError("dist_b200 runs on a CUDA device only; there is no CPU path (move the model with .cuda())")
        if not image.is_cuda and not image.is_pinned():
            raise RuntimeError("clips must be CUDA tensors or pinned host tensors")
        return dev

    def forward_without_text(self, image):
        clips = self._as_clips(image)
        u8 = clips.dtype == torch.uint8
        eng = self._engine(clips.shape[0], self._device_for(image), None, "uint8" if u8 else "float")
        emb = eng.forward(clips if u8 else clips.float(), use_graph=self.use_graph)
        return emb.clone()[:, None, :]                                       # clip.py:480

    def forward_with_text(self, image, text, others=None):
        clips = self._as_clips(image)
        feats = self._text_from(text, others)
        u8 = clips.dtype == torch.uint8
        eng = self._engine(clips.shape[0], self._device_for(image), feats, "uint8" if u8 else "float")
        emb = eng.forward(clips if u8 else clips.float(), use_graph=self.use_graph)
        logits = eng.logits.clone()
        vid = emb / emb.norm(dim=1, keepdim=True)                           # clip.py:513 (returned, not on the scored path)
        # img_logits: the per-frame CLIP embeddings ln_post(cls) @ proj [B*t, E] - L2-normalised on the fusion branch, raw otherwise
        # (clip.py:520,532) - or None when nothing asked for them (see _image_logits_mode)
        img = None
        if eng.image_logits == "fused":
            img = eng.img_emb_n.clone()
        elif eng.image_logits == "raw":
            img = eng.img_emb.clone()
        return {"logits_per_image": logits, "logits_per_text": logits.t(), "probs_per_image": eng.probs.clone(),
                "img_logits": img, "vid_logits": vid[:, None, :]}


def build_model(cfg, state_dict):
    """Infer the geometry from the checkpoint like ``clip.py:564-611`` and load it with ``load_state_dict(strict=False)``
    semantics: missing and unexpected keys are tolerated (and recorded on the model), a key whose SHAPE differs raises - e.g. a
    ``cls_token`` / ``positional_embedding`` trained for another frame count (the weights are frame-count specific).

    ``dist_net`` tensors absent from the checkpoint - the normal case when fine-tuning starts from an OpenAI CLIP archive - get the
    reference's own initialisation (``dist.py:77-79,120-121,195-220``: truncated normal, std 0.02, for every Linear / Conv weight and
    token, zero biases, LayerNorm at identity), not torch's module defaults."""
    import logging
    from ...text import has_text_tower, text_geometry
    arch = arch_from_cfg(cfg, state_dict)
    tk = {}
    if has_text_tower(state_dict):                                           # clip.py:586-591
        g = text_geometry(state_dict)
        tk = dict(context_length=g["context"], vocab_size=g["vocab"], transformer_width=g["width"], transformer_heads=g["heads"],
                  transformer_layers=g["layers"])
    model = CLIP(cfg, arch.embed_dim, arch.resolution, arch.layers, arch.width, arch.patch, arch=arch, **tk)
    own = model.state_dict()
    bad = ["%s: checkpoint %s vs model %s" % (k, tuple(v.shape), tuple(own[k].shape)) for k, v in state_dict.items()
           if k in own and torch.is_tensor(v) and tuple(own[k].shape) != tuple(v.shape)]
    if bad:
        raise RuntimeError("size mismatch for %d tensor(s) while loading the checkpoint (weights are frame-count / geometry specific):\n  %s"
                           % (len(bad), "\n  ".join(bad[:20])))
    usable = {k: v for k, v in state_dict.items() if k in own}
    missing = sorted(set(own) - set(usable))
    unexpected = sorted(k for k in state_dict if k not in own)
    fresh = [k for k in missing if k.startswith("dist_net.")]
    if fresh:
        seed = int(getattr(cfg, "RANDOM_SEED", 0))
        init = synth.synth_state_dict(arch, seed=seed, init="reference")
        usable.update({k: init[k] for k in fresh})
    model.load_state_dict(usable, strict=False)
    model.missing_keys, model.unexpected_keys, model.initialised_keys = missing, unexpected, fresh
    log = logging.getLogger(__name__)
    if fresh:
        log.info("dist_net: %d tensor(s) not in the checkpoint were initialised like the reference (trunc_normal 0.02 / zero bias)", len(fresh))
    if [k for k in missing if k not in fresh]:
        log.warning("missing keys (kept at module defaults): %s", [k for k in missing if k not in fresh][:20])
    if unexpected:
        log.info("unexpected checkpoint keys ignored: %d (e.g. %s)", len(unexpected), unexpected[:5])
    return model.eval()


def load(cfg):
    """``clip.load`` (``clip.py:614-629``): a TorchScript CLIP archive or a ``.pyth`` state_dict; random init when no path is set."""
    bb = cfg.VIDEO.BACKBONE
    path = getattr(bb, "PRETRAIN_WEIGHT_PATH", "") or ""
    local = getattr(bb, "LOCAL_PRETRAIN_WEIGHT_PATH", "") or ""
    if local and os.path.exists(local):
        path = local
    if path and os.path.exists(path):
        from ...utils.checkpoint import load_state_dict
        return build_model(cfg, load_state_dict(path))
    # "random initialization, only for debugging" (utils/checkpoint.py:524-527)
    arch = arch_from_cfg(cfg)
    seed = int(getattr(cfg, "RANDOM_SEED", 0))
    return build_model(cfg, synth.synth_state_dict(arch, seed=seed, init="reference"))
