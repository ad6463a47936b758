"""CLIP text tower on the distb200 CUDA library (reference: ``CLIP.encode_text`` / ``cache_text``, ``models/base/clip.py:419-452``).

SURVEY.md section 8(f) rank 4.  The tower runs ONCE per label set: with ``FREEZE_TEXT: true`` (every DiST config) the
reference caches its output (``clip.py:441-446``) and the video path only ever sees the resulting ``[C, E]`` matrix.  It is
therefore planned like the video path - packed weights, static buffers, a flat list of prepared C calls - but not tuned:
the GEMMs are the library's ``distb200_gemm`` (tcgen05 on bf16 operands, FFMA on the fp32 parity path), the causal
attention is the FFMA kernel ``distb200_attention_causal``.

Layout: token rows ``[C*ctx, W]`` fp32 residual stream, sequence-major (row = prompt*ctx + position).
"""

import torch

from . import ops


def text_geometry(sd):
    """(embed_dim, context, vocab, width, heads, layers) inferred as ``clip.build_model`` does (``clip.py:586-591``)."""
    width = sd["ln_final.weight"].shape[0]
    layers = len({k.split(".")[2] for k in sd if k.startswith("transformer.resblocks.")})
    return dict(embed_dim=sd["text_projection"].shape[1], context=sd["positional_embedding"].shape[0],
                vocab=sd["token_embedding.weight"].shape[0], width=width, heads=width // 64, layers=layers)


TEXT_KEYS = ("token_embedding.weight", "positional_embedding", "ln_final.weight", "ln_final.bias", "text_projection")


def has_text_tower(sd):
    return all(k in sd for k in TEXT_KEYS) and any(k.startswith("transformer.resblocks.") for k in sd)


class TextEngine:
    """Planned ``encode_text`` for a fixed number of prompts: ids int64 ``[C, ctx]`` -> features ``[C, E]`` fp32."""

    def __init__(self, state_dict, prompts, device="cuda", precision="bf16", gemm_impl=ops.IMPL_AUTO):
        assert precision in ("bf16", "fp32")
        ops.lib()
        sd = state_dict
        g = text_geometry(sd)
        assert g["width"] % 64 == 0 and g["layers"] >= 1, g
        self.geom, self.prompts, self.device, self.precision = g, int(prompts), torch.device(device), precision
        self.adt = torch.bfloat16 if precision == "bf16" else torch.float32
        self.gemm_impl = gemm_impl
        dev, adt = self.device, self.adt
        f32 = lambda x: x.detach().to(device=dev, dtype=torch.float32).contiguous()
        op = lambda x: x.detach().to(device=dev, dtype=torch.float32).contiguous().to(adt)
        self.table, self.pos = f32(sd["token_embedding.weight"]), f32(sd["positional_embedding"])
        self.layers = []
        for l in range(g["layers"]):
            pre = "transformer.resblocks.%d." % l
            self.layers.append(dict(
                ln1=(f32(sd[pre + "ln_1.weight"]), f32(sd[pre + "ln_1.bias"])), ln2=(f32(sd[pre + "ln_2.weight"]), f32(sd[pre + "ln_2.bias"])),
                qkv_w=op(sd[pre + "attn.in_proj_weight"]), qkv_b=f32(sd[pre + "attn.in_proj_bias"]),
                proj_w=op(sd[pre + "attn.out_proj.weight"]), proj_b=f32(sd[pre + "attn.out_proj.bias"]),
                fc1_w=op(sd[pre + "mlp.c_fc.weight"]), fc1_b=f32(sd[pre + "mlp.c_fc.bias"]),
                fc2_w=op(sd[pre + "mlp.c_proj.weight"]), fc2_b=f32(sd[pre + "mlp.c_proj.bias"])))
        self.ln_final = (f32(sd["ln_final.weight"]), f32(sd["ln_final.bias"]))
        self.proj_w = op(sd["text_projection"].float().t())                            # [E, W]
        C, ctx, W, E = self.prompts, g["context"], g["width"], g["embed_dim"]
        z = lambda *s, dtype=adt: torch.zeros(*s, device=dev, dtype=dtype)
        self.ids = z(C, ctx, dtype=torch.int64)
        self.x = z(C * ctx, W, dtype=torch.float32)
        self.ln_buf, self.qkv, self.att, self.fc1 = z(C * ctx, W), z(C * ctx, 3 * W), z(C * ctx, W), z(C * ctx, 4 * W)
        self.eot = z(C, W, dtype=torch.float32)
        self.eot_ln = z(C, W)
        self.feats = z(C, E, dtype=torch.float32)
        self.calls = []
        self._plan()

    def _lin(self, a, w, bias, out, res=None, act=ops.ACT_NONE, name="linear"):
        n, k = w.shape
        self.calls.append(ops.gemm(a, w, n, k, bias=bias, res=res, ld_res=n, out=out, ld_out=n, act=act, impl=self.gemm_impl, name=name))

    def _plan(self):
        g, add = self.geom, self.calls.append
        C, ctx, H = self.prompts, g["context"], g["heads"]
        add(ops.embed_tokens(self.ids, self.table, self.pos, self.x, name="text.embed"))               # clip.py:420-421
        for v in self.layers:                                                                            # clip.py:131-136
            add(ops.layernorm(self.x, v["ln1"][0], v["ln1"][1], self.ln_buf, name="text.ln_1"))
            self._lin(self.ln_buf, v["qkv_w"], v["qkv_b"], self.qkv, name="text.qkv")
            add(ops.attention_causal(self.qkv, self.att, C, ctx, H, name="text.attention"))
            self._lin(self.att, v["proj_w"], v["proj_b"], self.x, res=self.x, name="text.out_proj")
            add(ops.layernorm(self.x, v["ln2"][0], v["ln2"][1], self.ln_buf, name="text.ln_2"))
            self._lin(self.ln_buf, v["fc1_w"], v["fc1_b"], self.fc1, act=ops.ACT_QUICKGELU, name="text.fc1")
            self._lin(self.fc1, v["fc2_w"], v["fc2_b"], self.x, res=self.x, name="text.fc2")
        add(ops.gather_eot(self.x, self.ids, self.eot, name="text.eot"))                                 # clip.py:429
        add(ops.layernorm(self.eot, self.ln_final[0], self.ln_final[1], self.eot_ln, name="text.ln_final"))
        self._lin(self.eot_ln, self.proj_w, None, self.feats, name="text.projection")                    # clip.py:432-433

    def encode(self, ids):
        """ids int64 ``[C, ctx]`` (host or device) -> ``(features [C, E], eot rows [C, W])`` fp32 (views of internal buffers)."""
        g = self.geom
        if ids.dtype != torch.int64 or tuple(ids.shape) != (self.prompts, g["context"]):
            raise ValueError("token ids must be int64 [{}, {}], got {} {}".format(self.prompts, g["context"], ids.dtype, tuple(ids.shape)))
        lo, hi = int(ids.min()), int(ids.max())
        if lo < 0 or hi >= g["vocab"]:
            raise IndexError("token id out of range [0, {}): min {} max {}".format(g["vocab"], lo, hi))   # nn.Embedding raises alike
        self.ids.copy_(ids)
        s = torch.cuda.current_stream(self.device).cuda_stream
        for c in self.calls:
            c.launch(s)
        return self.feats, self.eot
