"""CLIP text tower: oracle against the reference fixtures on the CPU; CUDA path against both on the GPU.

Fixtures ``text_tiny`` / ``text_b16`` come from the unmodified reference (``CLIP.encode_text`` and ``CLIP.forward`` with token
ids, ``oracle/make_golden.py::run_text_case``) on seeded synthetic weights and tokenizer-shaped ids.
"""
import os

import pytest
import torch

from conftest import load_golden, rel_l2
from dist_b200.arch import DistArch
from dist_b200.utils import synth
from oracle import dist_oracle

FP32_BAR, BF16_BAR = 1e-4, 1e-2
CFG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs", "projects", "dist", "ssv2", "vit-b16-8+16f.yaml")


def _tiny_cfg(arch):
    from dist_b200.config import Config
    cfg = Config.from_file(CFG, ["NUM_GPUS", "0"])
    cfg.DATA.NUM_INPUT_FRAMES, cfg.DATA.SPARSE_SAMPLE_ALPHA = arch.frames, arch.alpha
    d = cfg.VIDEO.BACKBONE.DIST
    d.INTEGRATION_DIM, d.TEMPORAL_DIM, d.S_PATCH_SIZE, d.ADA_POOLING_LAYERS = arch.integration_dim, arch.temporal_dim, arch.s_patch, arch.ada_layers
    d.SELECTED_LAYERS = list(arch.selected_layers)
    cfg.VIDEO.HEAD.NUM_CLASSES = arch.num_classes
    return cfg


def _inputs(fix):
    tk = fix["text"]
    tsd = synth.synth_text_tower(fix["arch"]["embed_dim"], seed=fix["text_seed"], init=fix["init"], **tk)
    ids = synth.synth_token_ids(fix["prompts"], tk["context"], tk["vocab"], seed=fix["ids_seed"])
    got, want = synth.checksum(tsd), fix["text_checksum"]
    assert all(abs(a - b) <= 1e-6 * max(1.0, abs(b)) for a, b in zip(got, want)), "synthetic text tower differs from the fixture's"
    assert torch.equal(ids, fix["ids"])
    return tsd, ids


@pytest.mark.parametrize("name", ["text_tiny", "text_b16"])
def test_oracle_text_tower_matches_reference(name):
    fix = load_golden(name)
    tsd, ids = _inputs(fix)
    feats, eot = dist_oracle.encode_text({k: v.double() for k, v in tsd.items()}, ids)
    assert rel_l2(feats, fix["feats"]) < 2e-6
    assert rel_l2(eot, fix["eot"]) < 2e-6


def test_oracle_text_tower_is_causal_and_ignores_padding():
    """Tokens after the end-of-text token cannot influence its row (causal mask), so padding is irrelevant."""
    fix = load_golden("text_tiny")
    tsd, ids = _inputs(fix)
    tsd = {k: v.double() for k, v in tsd.items()}
    base, _ = dist_oracle.encode_text(tsd, ids)
    changed = ids.clone()
    for r in range(ids.shape[0]):
        e = int(ids[r].argmax())
        changed[r, e + 1:] = 1                     # any id below the end-of-text token
    other, _ = dist_oracle.encode_text(tsd, changed)
    assert rel_l2(other, base) < 1e-12


def test_text_geometry_and_model_keys():
    """``build_model`` grows the text tower under the reference's key names when the checkpoint carries it (clip.py:586-591)."""
    from dist_b200.models.base import clip
    from dist_b200.text import has_text_tower, text_geometry
    fix = load_golden("text_tiny")
    tsd, _ = _inputs(fix)
    assert has_text_tower(tsd)
    g = text_geometry(tsd)
    assert (g["width"], g["layers"], g["context"], g["vocab"], g["heads"], g["embed_dim"]) == (128, 2, 12, 64, 2, 64)
    arch = DistArch(**fix["arch"]).validate()
    sd = synth.synth_state_dict(arch, seed=0, init="scaled")
    assert not has_text_tower(sd)
    cfg = _tiny_cfg(arch)
    full = dict(sd)
    full.update(tsd)
    model = clip.build_model(cfg, full)
    assert model.has_text_tower and not model.missing_keys
    own = model.state_dict()
    for k, v in tsd.items():
        assert k in own and own[k].shape == v.shape and torch.equal(own[k], v), k
    bare = clip.build_model(cfg, sd)
    assert not bare.has_text_tower
    with pytest.raises(RuntimeError):
        bare.encode_text(torch.zeros(3, 12, dtype=torch.long))


# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name,precision,bar", [("text_tiny", "fp32", FP32_BAR), ("text_tiny", "bf16", BF16_BAR),
                                                ("text_b16", "fp32", FP32_BAR), ("text_b16", "bf16", BF16_BAR)])
def test_text_engine_matches_reference(name, precision, bar):
    from dist_b200.text import TextEngine
    fix = load_golden(name)
    tsd, ids = _inputs(fix)
    eng = TextEngine(tsd, fix["prompts"], device="cuda", precision=precision)
    feats, eot = eng.encode(ids.cuda())
    torch.cuda.synchronize()
    assert torch.isfinite(feats).all()
    assert rel_l2(eot, fix["eot"]) < bar, rel_l2(eot, fix["eot"])
    assert rel_l2(feats, fix["feats"]) < bar, rel_l2(feats, fix["feats"])
    with pytest.raises(IndexError):
        bad = ids.clone()
        bad[0, 1] = fix["text"]["vocab"]
        eng.encode(bad)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("tokens,heads,seqs", [(77, 8, 5), (12, 2, 3), (1, 1, 2), (130, 3, 2)])
def test_attention_causal_kernel(dtype, tokens, heads, seqs):
    """distb200_attention_causal against its fp64 definition (softmax over the keys j <= i only)."""
    from dist_b200 import ops
    g = torch.Generator().manual_seed(3)
    W = heads * 64
    qkv = torch.randn(seqs, tokens, 3 * W, generator=g).to(dtype).cuda()
    out = torch.empty(seqs, tokens, W, dtype=dtype, device="cuda")
    ops.attention_causal(qkv, out, seqs, tokens, heads).launch(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    q, k, v = qkv.double().cpu().split(W, dim=-1)
    sp = lambda z: z.view(seqs, tokens, heads, 64).transpose(1, 2)
    mask = torch.full((tokens, tokens), float("-inf"), dtype=torch.float64).triu_(1)
    want = (torch.softmax(sp(q) @ sp(k).transpose(-1, -2) / 8.0 + mask, dim=-1) @ sp(v)).transpose(1, 2).reshape(seqs, tokens, W)
    assert rel_l2(out, want) < (2e-6 if dtype == torch.float32 else 6e-3)


@pytest.mark.gpu
def test_embed_and_eot_kernels_are_exact():
    from dist_b200 import ops
    g = torch.Generator().manual_seed(4)
    seqs, ctx, width, vocab = 7, 9, 64, 50
    table, pos = torch.randn(vocab, width, generator=g), torch.randn(ctx, width, generator=g)
    ids = torch.randint(0, vocab - 1, (seqs, ctx), generator=g)
    ids[:, 3] = vocab - 1
    ids[2, 5] = vocab - 1                          # a tie resolves to the first position, like torch.argmax
    ids[4] = 0                                     # all equal -> position 0
    x = torch.empty(seqs * ctx, width, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    ops.embed_tokens(ids.cuda(), table.cuda(), pos.cuda(), x).launch(st)
    out = torch.empty(seqs, width, device="cuda")
    ops.gather_eot(x, ids.cuda(), out).launch(st)
    torch.cuda.synchronize()
    want = table[ids] + pos
    assert torch.equal(x.cpu().view(seqs, ctx, width), want)
    assert torch.equal(out.cpu(), want[torch.arange(seqs), ids.argmax(dim=-1)])


@pytest.mark.gpu
@pytest.mark.parametrize("precision,bar", [("fp32", FP32_BAR), ("bf16", BF16_BAR)])
def test_model_forward_with_token_ids_matches_reference(precision, bar):
    """The reference's own entry: ``model({"video", "texts": int64 ids})`` -> text tower once (cache_text) -> class logits."""
    from dist_b200.models.base import clip
    fix = load_golden("text_tiny")
    tsd, ids = _inputs(fix)
    arch = DistArch(**fix["arch"]).validate()
    sd = synth.synth_state_dict(arch, seed=0, init="scaled")
    clips = synth.synth_clips(fix["batch"], arch, seed=fix["clip_seed"], kind="structured")
    cfg = _tiny_cfg(arch)
    full = dict(sd)
    full.update(tsd)
    model = clip.build_model(cfg, full).cuda()
    model.precision = precision
    out = model(clips.cuda(), ids.cuda())
    torch.cuda.synchronize()
    assert rel_l2(out["logits_per_image"], fix["logits"]) < 3 * bar, rel_l2(out["logits_per_image"], fix["logits"])
    feats_first = model._text_cache[0]
    out2 = model(clips.cuda(), ids.cuda())                     # second call reuses the cached label embeddings (clip.py:441-446)
    assert model._text_cache[0] is feats_first
    assert torch.equal(out2["logits_per_image"], out["logits_per_image"])
