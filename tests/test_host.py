"""Host-side logic: config schema, registries, builders, weights contract, C ABI surface (no GPU needed)."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from dist_b200 import build as dbuild
from dist_b200.arch import arch_from_cfg, tiny_arch, DistArch
from dist_b200.config import Config
from dist_b200.registry import Registry
from dist_b200.utils import synth

CFG_DIR = os.path.join(ROOT, "configs", "projects", "dist")


def test_registry_semantics():
    reg = Registry("T")

    @reg.register()
    class Foo:
        pass

    assert reg.get("Foo") is Foo and reg.get("Bar") is None and "Foo" in reg.get_all_registered()
    with pytest.raises(AssertionError):
        reg.register()(Foo)


@pytest.mark.parametrize("path,frames,ada,classes,name", [
    ("ssv2/vit-b16-8+16f.yaml", 16, 2, 174, "ViT-B-16"),
    ("ssv2/vit-b16-16+32f.yaml", 32, 2, 174, "ViT-B-16"),
    ("k400/vit-b16-32+64f.yaml", 64, 4, 400, "ViT-B-16"),
    ("k400/vit-l14-32+64f.yaml", 64, 4, 400, "ViT-L-14"),
    ("ssv2/vit-l14-32+64f.yaml", 64, 2, 174, "ViT-L-14"),
])
def test_config_inheritance(path, frames, ada, classes, name):
    cfg = Config.from_file(os.path.join(CFG_DIR, path))
    assert cfg.DATA.NUM_INPUT_FRAMES == frames and cfg.DATA.SPARSE_SAMPLE_ALPHA == 2
    assert cfg.VIDEO.BACKBONE.DIST.ADA_POOLING_LAYERS == ada
    assert cfg.VIDEO.HEAD.NUM_CLASSES == classes and cfg.VIDEO.HEAD.NAME == "ClipVideoTextIdentity"
    assert cfg.VIDEO.BACKBONE.META_ARCH == "ClipVisionTextTransformer" and cfg.VIDEO.BACKBONE.META_ARCH_NAME == name
    assert cfg.VIDEO.BACKBONE.ATTEN_BLOCK == "ResidualAttentionBlockMid"
    assert cfg.OPTIMIZER.BASE_LR == pytest.approx(3.2e-5) and isinstance(cfg.OPTIMIZER.MIN_LR, float)
    assert cfg.DIST_BACKEND == "nccl"                       # from pool/base.yaml
    assert cfg.OPTIMIZER.OPTIM_METHOD == "adamw"            # _BASE_RUN overridden by the project base
    arch = arch_from_cfg(cfg)
    assert arch.tokens == (197 if name == "ViT-B-16" else 257)


def test_config_overrides():
    cfg = Config.from_file(os.path.join(CFG_DIR, "ssv2/vit-b16-8+16f.yaml"),
                           ["DATA.NUM_INPUT_FRAMES", "32", "VIDEO.BACKBONE.DIST.ADA_POOLING_LAYERS", "3", "NUM_GPUS", "0"])
    assert cfg.DATA.NUM_INPUT_FRAMES == 32 and cfg.VIDEO.BACKBONE.DIST.ADA_POOLING_LAYERS == 3 and cfg.NUM_GPUS == 0
    with pytest.raises(AssertionError):
        Config.from_file(os.path.join(CFG_DIR, "ssv2/vit-b16-8+16f.yaml"), ["DATA.NO_SUCH_KEY", "1"])
    with pytest.raises(AssertionError):
        Config.from_file(os.path.join(CFG_DIR, "ssv2/vit-b16-8+16f.yaml"), ["DATA.NUM_INPUT_FRAMES"])


def test_l14_patch_mismatch_is_rejected():
    with pytest.raises(AssertionError):
        DistArch(width=1024, layers=24, patch=14, embed_dim=768, s_patch=16, selected_layers=list(range(24))).validate()


def test_registry_driven_build_and_weights_contract():
    import dist_b200.models.base  # noqa: F401  (registration side effects)
    from dist_b200.models.base.backbone import BACKBONE_REGISTRY
    from dist_b200.models.base.base_blocks import BRANCH_REGISTRY, HEAD_REGISTRY, STEM_REGISTRY
    from dist_b200.models.base.builder import build_model
    from dist_b200.models.base.clip import ATTEN_BLOCK_REGISTRY
    from dist_b200.models.base.models import MODEL_REGISTRY, BaseVideoModel

    assert BACKBONE_REGISTRY.get("ClipVisionTextTransformer") is not None
    assert ATTEN_BLOCK_REGISTRY.get("ResidualAttentionBlockMid") is not None
    assert HEAD_REGISTRY.get("ClipVideoTextIdentity") is not None
    for name in ("DiSTNetwork", "TemporalNet", "IntegrationNetwork", "Integration2TemporalNetwork",
                 "Temporal2IntegrationNetwork", "SpatialTemporalAdaPoolingNetwork"):
        assert BRANCH_REGISTRY.get(name) is not None
    assert STEM_REGISTRY.get("DiSTTemporalStem") is not None
    assert MODEL_REGISTRY.get("clip") is None              # falls back to BaseVideoModel (builder.py:30-32)

    cfg = Config.from_file(os.path.join(CFG_DIR, "ssv2/vit-b16-8+16f.yaml"), ["NUM_GPUS", "0"])
    model, ema = build_model(cfg)
    assert isinstance(model, BaseVideoModel) and ema is None
    enc = model.backbone.base_encoder
    assert model.backbone.get_num_layers() == (12, 0)
    sd = synth.synth_state_dict(arch_from_cfg(cfg), seed=0)
    own = enc.state_dict()
    assert set(own) == set(sd)                              # the reference's key names (checked against it in make_golden.py)
    assert all(own[k].shape == sd[k].shape for k in sd)
    n_dist = sum(v.numel() for k, v in own.items() if k.startswith("dist_net."))
    assert n_dist == 19001184                               # 19.00 M
    # no CPU fallback: running without a GPU must fail loudly, not silently compute something else
    with pytest.raises(RuntimeError):
        model({"video": torch.zeros(1, 3, 16, 224, 224), "texts": torch.zeros(174, 512)})


def test_head_semantics():
    from dist_b200.models.base.base_blocks import ClipVideoTextIdentity
    cfg = Config.from_file(os.path.join(CFG_DIR, "ssv2/vit-b16-8+16f.yaml"))
    head = ClipVideoTextIdentity(cfg).eval()
    x = {"logits_per_image": torch.randn(3, 2, 7)}
    out, passthrough = head(x)
    assert torch.allclose(out, torch.softmax(x["logits_per_image"].mean(1), -1)) and passthrough is x
    head.train()
    out, _ = head(x)
    assert torch.allclose(out, x["logits_per_image"].mean(1))


def test_synth_is_deterministic():
    a = tiny_arch()
    s1, s2 = synth.synth_state_dict(a, 0, "scaled"), synth.synth_state_dict(a, 0, "scaled")
    assert all(torch.equal(s1[k], s2[k]) for k in s1)
    assert not torch.equal(synth.synth_state_dict(a, 1, "scaled")["dist_net.proj"], s1["dist_net.proj"])
    assert torch.equal(synth.synth_clips(2, a, 5), synth.synth_clips(2, a, 5))


def test_flop_model_matches_survey():
    # torch FlopCounterMode on the reference gives 314.3 GF/clip (B/16 8+16f): that count includes 1.85 GF of
    # discarded patch-embed work and omits the attention matmuls (SDPA is not counted on CPU); SURVEY.md 8(d)'s
    # formula, implemented by flops_per_clip, counts attention and only the useful patch embedding.
    a = DistArch()
    f = a.flops_per_clip()
    attn = a.layers * a.sparse_frames * 4 * a.tokens ** 2 * a.width
    assert abs((f["total"] - attn) / 1e9 - (314.3 - 1.85)) < 2.0
    big = DistArch(width=1024, layers=24, patch=14, embed_dim=768, frames=64, s_patch=14, ada_layers=4,
                   selected_layers=list(range(24)))
    fb = big.flops_per_clip()
    attn = big.layers * big.sparse_frames * 4 * big.tokens ** 2 * big.width
    wasted = 2 * big.sparse_frames * big.patches * big.width * 3 * big.patch ** 2
    assert abs((fb["total"] - attn + wasted) / 1e9 - 5458.1) / 5458.1 < 0.01


def test_c_abi_exports_every_declared_symbol():
    """The shared library loads and exports what include/distb200.h declares (no compute without a GPU)."""
    header = open(os.path.join(ROOT, "include", "distb200.h")).read()
    declared = sorted(set(re.findall(r"\b(distb200_[a-z_0-9]+)\s*\(", header)))
    assert "distb200_gemm" in declared and len(declared) >= 11
    lib_path = dbuild.build()
    lib = ctypes.CDLL(lib_path)
    for sym in declared:
        assert hasattr(lib, sym), sym
    lib.distb200_version.restype = ctypes.c_int
    assert lib.distb200_version() == 100
    from dist_b200 import ops
    assert sorted(ops.EXPORTS) == declared
    # the ctypes mirror of the descriptor must have the C layout
    import subprocess, tempfile
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "distb200.h"\nint main(){printf("%zu %zu %zu %zu", sizeof(distb200_gemm_desc), offsetof(distb200_gemm_desc, ldb), offsetof(distb200_gemm_desc, out), offsetof(distb200_gemm_desc, group_dim));return 0;}'
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(td, "t.c"), "-o", os.path.join(td, "t")])
        sizes = [int(x) for x in subprocess.check_output([os.path.join(td, "t")]).split()]
    D = ops.GemmDesc
    assert sizes == [ctypes.sizeof(D), D.ldb.offset, D.out.offset, D.group_dim.offset]


def test_checkpoint_ingest_roundtrip_and_legacy_names(tmp_path):
    """`.pyth` files with the BaseVideoModel prefix, and pre-release `ladder_net.*` names (process_dist_cpkt.py:10-30)."""
    import torch
    from dist_b200.arch import tiny_arch
    from dist_b200.utils import checkpoint, synth
    arch = tiny_arch()
    sd = synth.synth_state_dict(arch, seed=0)
    path = str(tmp_path / "dist.pyth")
    checkpoint.save_checkpoint(path, sd)
    raw = torch.load(path, map_location="cpu", weights_only=False)
    assert all(k.startswith("backbone.base_encoder.") for k in raw["model_state"])
    back = checkpoint.load_state_dict(path)
    assert sorted(back) == sorted(sd) and all(torch.equal(back[k], sd[k]) for k in sd)
    legacy = {}
    inv = [(new, old) for old, new in checkpoint._LEGACY]
    for k, v in sd.items():
        name = k
        if k.startswith("dist_net."):
            for new, old in sorted(inv, key=lambda p: -len(p[0])):
                if name.startswith(new):
                    name = old + name[len(new):]
                    break
        legacy["backbone.base_encoder." + name] = v
    assert any("ladder_net" in k for k in legacy) and not any("dist_net" in k for k in legacy)
    torch.save({"model_state": legacy}, path)
    back = checkpoint.load_state_dict(path)
    assert sorted(back) == sorted(sd) and all(torch.equal(back[k], sd[k]) for k in sd)


def test_graph_branches_dependencies():
    """The two-branch CUDA graph (engine.plan_branches): ViT calls on branch 0, DiST calls on branch 1, and exactly the
    cross-branch edges the tap buffers require (taps alternate between two buffers by ViT block)."""
    from dist_b200.engine import plan_branches
    sel = [0, 2, 3]
    names = ["patchify.dense", "vit.patch_embed", "vit.cls_rows", "vit.ln_pre", "dist.stem"]
    for l in range(4):
        names += ["vit.ln_1" if l == 0 else "vit.ln_1.stats", "vit.qkv", "vit.attention", "vit.out_proj", "vit.ln_2.stats", "vit.fc1", "vit.fc2"]
        if l in sel:
            names += ["dist.tn.ln", "dist.tn.conv_s", "dist.input_linear", "dist.t2i", "dist.int.proj"]
        if l == sel[-1]:
            names.append("dist.cls_mean")
    names += ["ada.init_sp", "tail.proj_spatial_cls", "tail.proj", "head.class_scores"]
    branch, deps = plan_branches(names, sel)
    at = lambda n, j: [k for k, x in enumerate(names) if x == n][j]
    assert all(branch[k] == (0 if n.startswith(("patchify", "vit.")) or n == "dist.cls_mean" else 1) for k, n in enumerate(names))
    assert deps[at("dist.stem", 0)] == [at("patchify.dense", 0)]
    for i, l in enumerate(sel):                                   # DiST layer i reads the tap of block sel[i]
        assert deps[at("dist.input_linear", i)] == [at("vit.fc2", l)]
    assert deps[at("vit.fc2", 2)] == [at("dist.input_linear", 0)]  # block 2 overwrites the buffer of block 0, read by DiST layer 0
    assert at("vit.fc2", 3) not in deps                            # block 1 is not tapped: nothing to wait for
    assert at("vit.fc2", 0) not in deps and at("vit.fc2", 1) not in deps
    assert deps[at("tail.proj_spatial_cls", 0)] == [at("dist.cls_mean", 0)]
    for k, ds in deps.items():                                    # every edge points backwards in plan order and crosses branches
        assert all(d < k and branch[d] != branch[k] for d in ds)


def test_kcat_layout_and_weight_algebra():
    """The K-concatenated DiST operands (engine.kcat_layout / PackedWeights): in fp64 on the CPU, one product with the concatenated
    weights equals input_linear + previous output projection + temporal->integration convolution + cls token
    (dist.py:229,45,80-86,232), and the re-expressed integration->temporal product equals linear_fuse of the PRE-fusion stream
    (dist.py:99-105,231)."""
    from dist_b200.engine import PackedWeights, kcat_layout
    from dist_b200.utils import synth
    arch = tiny_arch(frames=6, alpha=3, resolution=96, selected_layers=[0, 1])
    L = kcat_layout(arch)
    D, Ci, Ct, al, t = arch.width, arch.integration_dim, arch.temporal_dim, arch.alpha, arch.sparse_frames
    Ih, Cm = arch.integration_hidden, arch.integration_temporal_hidden
    assert all(v % 64 == 0 for v in L.values())
    assert L["c_h"] >= D and L["c_x"] >= L["c_h"] + Ih + 2 * Cm and L["c_oh"] >= L["c_x"] + al * Ct and L["c_u"] >= L["c_oh"] + t
    assert L["tw"] >= L["c_u"] + Ci
    big = DistArch()
    LB = kcat_layout(big)
    assert (LB["c_h"], LB["c_x"], LB["c_oh"], LB["c_u"], LB["tw"]) == (768, 1344, 1536, 1600, 1984)

    sd = {k: v.double() for k, v in synth.synth_state_dict(arch, seed=3, init="scaled").items()}
    w = PackedWeights(sd, arch, "cpu", torch.float64)
    g = torch.Generator().manual_seed(0)
    for i in (0, 1):
        d = w.dist[i]
        assert d["cat_w"].shape == (Ci, L["c_u"]) and d["i2t_cat_w"].shape == (Ct, L["c_u"] + Ci - L["c_x"])
        ti = 1
        tap = torch.randn(5, D, generator=g, dtype=torch.float64)
        h_prev = torch.randn(5, Ih + 2 * Cm, generator=g, dtype=torch.float64)       # [ffn hidden | c_fc1 output | temporal hidden]
        xt = torch.randn(5, al, Ct, generator=g, dtype=torch.float64)                # the alpha dense frames of sparse frame ti
        rows = torch.zeros(5, L["c_u"], dtype=torch.float64)
        rows[:, :D] = tap
        rows[:, L["c_h"]:L["c_h"] + Ih + 2 * Cm] = h_prev
        rows[1:, L["c_x"]:L["c_x"] + al * Ct] = xt[1:].reshape(4, -1)                # row 0 plays the class token: no temporal rows,
        rows[0, L["c_oh"] + ti] = 1.0                                                 # a one-hot(ti) instead
        upd = rows @ d["cat_w"].t() + d["cat_b"]
        lin = lambda x, p: x @ sd[p + ".weight"].t() + sd[p + ".bias"]
        want = lin(tap, "dist_net.input_linears.%d" % i)
        if i > 0:
            pit = "dist_net.integration_nets.%d." % (i - 1)
            want = (want + h_prev[:, :Ih] @ sd[pit + "ffn.c_proj.weight"].t() + sd[pit + "ffn.c_proj.bias"]
                    + h_prev[:, Ih + Cm:] @ sd[pit + "temporal_ffn.c_proj.weight"][:, :, 0, 0, 0].t() + sd[pit + "temporal_ffn.c_proj.bias"])
        t2i = "dist_net.temporal2integration_nets.%d." % i
        wt = sd[t2i + "linear_fuse.weight"][:, :, :, 0, 0]                            # [Ci, Ct, alpha]
        v = torch.einsum("rkc,ick->ri", xt, wt) + sd[t2i + "linear_fuse.bias"]
        mid_pre = want.clone()                                                        # input_linear + res: what integration->temporal reads
        want[1:] += v[1:]
        want[0] += sd[t2i + "cls_token"].reshape(t, Ci)[ti]
        assert float((upd - want).abs().max()) < 2e-6                    # the packed operands pass through fp32
        a2 = torch.cat([rows[:, L["c_x"]:], upd], dim=1)
        u = a2 @ d["i2t_cat_w"].t() + d["i2t_cat_b"]
        want_u = lin(mid_pre, "dist_net.integration2temporal_nets.%d.linear_fuse" % i)
        assert float((u[1:] - want_u[1:]).abs().max()) < 2e-6                       # patch rows only (the class row is never read)


def test_torchscript_archive_ingest(tmp_path):
    """OpenAI publishes CLIP as TorchScript archives; the reference reads them with ``torch.jit.load(...).state_dict()``
    (``clip.py:618-623``).  A scripted module with CLIP-shaped parameter names goes through the same path here."""
    import torch.nn as nn
    from dist_b200.utils import checkpoint

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.ln_1 = nn.LayerNorm(8)

        def forward(self, x):
            return self.ln_1(x)

    class Visual(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = nn.Conv2d(3, 8, 4, 4, bias=False)
            self.class_embedding = nn.Parameter(torch.randn(8))
            self.proj = nn.Parameter(torch.randn(8, 6))
            self.resblock = Block()

        def forward(self, x):
            return self.resblock(self.conv1(x).flatten(2).transpose(1, 2)) @ self.proj + self.class_embedding.sum()

    class Clip(nn.Module):
        def __init__(self):
            super().__init__()
            self.visual = Visual()
            self.logit_scale = nn.Parameter(torch.ones([]) * 2.6593)
            self.register_buffer("input_resolution", torch.tensor(8))

        def forward(self, x):
            return self.visual(x) * self.logit_scale.exp()

    torch.manual_seed(0)
    m = Clip().eval()
    path = str(tmp_path / "ViT-tiny.pt")
    torch.jit.script(m).save(path)
    sd = checkpoint.load_state_dict(path)
    want = m.state_dict()
    assert set(sd) == set(want) and all(torch.equal(sd[k], want[k]) for k in want)
    assert {"visual.conv1.weight", "visual.class_embedding", "visual.proj", "visual.resblock.ln_1.weight", "logit_scale"} <= set(sd)
    # a plain state_dict saved under a .pt name (no TorchScript) falls back to torch.load
    plain = str(tmp_path / "plain.pt")
    torch.save(want, plain)
    sd2 = checkpoint.load_state_dict(plain)
    assert set(sd2) == set(want)
