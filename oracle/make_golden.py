"""Generate ``tests/golden/*.pt`` by running the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY.

Runs only in the build container (needs ``/root/reference``); the fixtures it writes are committed
and are what travels to the GPU box.  Usage::

    python oracle/make_golden.py            # all cases
    python oracle/make_golden.py tiny_ref   # selected cases

For every case it
  1. builds the synthetic weights / clips / label embeddings with ``dist_b200.utils.synth``
     (seeded, reproducible on any machine with the same torch),
  2. imports the reference with the three module shims of SURVEY.md appendix A (``timm``,
     ``simplejson``, ``oss2`` are not installed here; no reference code is copied or modified),
     builds ``models.base.clip.CLIP`` through the reference's own ``build_model`` and loads the
     synthetic tensors into it with a strict key check on ``visual.*`` / ``dist_net.*``,
  3. runs ``CLIP.forward_without_text`` (``clip.py:466``) and the with-text path through
     ``others['label_embeddings']`` (``clip.py:437-439``) in fp32 on the CPU, capturing
     intermediate activations with forward hooks,
  4. cross-checks ``oracle/dist_oracle.py`` (float64) against those outputs and refuses to write
     the fixture if they disagree by more than 2e-6 rel-L2,
  5. stores outputs, sub-sampled intermediates and the input checksums.
"""

import json
import os
import sys
import time
import types

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, REPO)

from dist_b200.arch import DistArch, tiny_arch  # noqa: E402
from dist_b200.utils import synth  # noqa: E402
from oracle import dist_oracle  # noqa: E402

CASES = {
    # name: (arch, weight init, batch, clip kind)
    "tiny_ref": (tiny_arch(), "reference", 2, "structured"),
    "tiny_scaled": (tiny_arch(), "scaled", 2, "structured"),
    "tiny_a1": (tiny_arch(frames=3, alpha=1, ada_layers=1, selected_layers=[1]), "scaled", 2, "iid"),
    "tiny_a3": (tiny_arch(frames=6, alpha=3, resolution=96, selected_layers=[0, 1]), "scaled", 1, "iid"),
    # taps skipped between DiST layers, first block untapped, three DiST layers (the hidden block of one layer travels in the next one's operand)
    "tiny_skip": (tiny_arch(layers=6, selected_layers=[1, 3, 4], frames=8, alpha=2), "scaled", 3, "structured"),
    "b16_8x16_ref": (DistArch(), "reference", 2, "structured"),
    "b16_8x16_scaled": (DistArch(), "scaled", 2, "structured"),
    "b16_8x16_iid": (DistArch(), "reference", 2, "iid"),
    "b16_32x64_k400": (DistArch(frames=64, ada_layers=4, num_classes=400), "reference", 1, "structured"),
    "l14_32x64_k400": (DistArch(width=1024, layers=24, patch=14, embed_dim=768, frames=64, s_patch=14,
                                ada_layers=4, num_classes=400, selected_layers=list(range(24))),
                       "reference", 1, "structured"),
}


def _install_shims():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    layers = mod(
        "timm.models.layers",
        trunc_normal_=lambda t, mean=0.0, std=1.0, a=-2.0, b=2.0: torch.nn.init.trunc_normal_(t, mean=mean, std=std, a=a, b=b),
        drop_path=lambda x, drop_prob=0.0, training=False: x,
        to_2tuple=lambda x: tuple(x) if isinstance(x, (tuple, list)) else (x, x),
    )
    registry = mod("timm.models.registry", register_model=lambda f: f)
    mod("timm.models", layers=layers, registry=registry)
    mod("timm", models=sys.modules["timm.models"])
    mod("simplejson", dumps=lambda o, **k: json.dumps(o, default=str))
    mod("oss2")


def _reference_cfg(arch):
    from utils.config import Config
    c = Config(load=False, cfg_dict={})
    c.need_initialization = True
    args = types.SimpleNamespace(cfg_file="configs/projects/dist/ssv2/vit-b16-8+16f.yaml", opts=[])
    d = c._merge_cfg_from_base(c._initialize_cfg(), c._load_yaml(args))
    d["DATA"]["NUM_INPUT_FRAMES"] = arch.frames
    d["DATA"]["SPARSE_SAMPLE_ALPHA"] = arch.alpha
    dd = d["VIDEO"]["BACKBONE"]["DIST"]
    dd.update(INTEGRATION_DIM=arch.integration_dim, TEMPORAL_DIM=arch.temporal_dim, S_PATCH_SIZE=arch.s_patch,
              T_PATCH_SIZE=arch.t_patch, TEMPORAL_KERNEL_SIZE=arch.t_kernel, ADA_POOLING_LAYERS=arch.ada_layers,
              SELECTED_LAYERS=list(arch.selected_layers))
    d["VIDEO"]["HEAD"]["NUM_CLASSES"] = arch.num_classes
    d["NUM_GPUS"] = 0
    cfg = Config(load=False, cfg_dict=d)
    cfg.cfg_dict = d
    return cfg


def _text_tower_stub(arch, width=64, layers=1, ctx=8, vocab=32):
    """The keys clip.build_model inspects for the (off-path) text tower (clip.py:586-591)."""
    sd = {
        "text_projection": torch.zeros(width, arch.embed_dim),
        "positional_embedding": torch.zeros(ctx, width),
        "token_embedding.weight": torch.zeros(vocab, width),
        "ln_final.weight": torch.ones(width),
    }
    for i in range(layers):
        sd["transformer.resblocks.%d.ln_1.weight" % i] = torch.ones(width)
    return sd


TEXT_CASES = {
    # name: (visual arch, text tower kwargs, weight init, prompts, clips)
    "text_tiny": (tiny_arch(), dict(width=128, layers=2, context=12, vocab=64), "scaled", 10, 2),
    "text_b16": (tiny_arch(embed_dim=512, num_classes=16), dict(width=512, layers=12, context=77, vocab=49408), "reference", 16, 1),
}


def build_reference(arch, sd, text_sd=None):
    from models.base import clip
    cfg = _reference_cfg(arch)
    full = dict(sd)
    full.update(text_sd if text_sd is not None else _text_tower_stub(arch))
    model = clip.build_model(cfg, dict(full))
    want = {k for k in model.state_dict() if k.startswith(("visual.", "dist_net."))}
    have = {k for k in sd if k.startswith(("visual.", "dist_net."))}
    assert want == have, "weights contract mismatch: missing {} unexpected {}".format(sorted(want - have)[:8], sorted(have - want)[:8])
    msd = model.state_dict()
    for k in have:
        assert msd[k].shape == sd[k].shape, (k, msd[k].shape, sd[k].shape)
        assert torch.equal(msd[k], sd[k]), k
    if text_sd is not None:
        for k, v in text_sd.items():
            assert msd[k].shape == v.shape and torch.equal(msd[k], v), k
    model.prediction_fusion_enable = False          # undefined attribute read at clip.py:519
    return model.eval()


def run_text_case(name):
    """Text tower: ``CLIP.encode_text`` (clip.py:419-434) and the token-id entry of ``CLIP.forward``
    (clip.py:482-533 with FREEZE_TEXT -> cache_text) on seeded synthetic weights and tokenizer-shaped ids."""
    arch, tk, init, prompts, batch = TEXT_CASES[name]
    t0 = time.time()
    sd = synth.synth_state_dict(arch, seed=0, init=init if init == "scaled" else "reference")
    tsd = synth.synth_text_tower(arch.embed_dim, seed=5, init=init, **tk)
    ids = synth.synth_token_ids(prompts, tk["context"], tk["vocab"], seed=11)
    clips = synth.synth_clips(batch, arch, seed=1234, kind="structured")
    model = build_reference(arch, sd, tsd)
    frames = clips.permute(0, 2, 1, 3, 4).reshape(batch * arch.frames, 3, arch.resolution, arch.resolution)
    with torch.no_grad():
        feats, eot, _ = model.encode_text(ids, None)                                  # clip.py:419
        out = model(frames, ids)                                                      # clip.py:460 -> cache_text -> encode_text
    logits = out["logits_per_image"]
    tsd64 = {k: v.double() for k, v in tsd.items()}
    o_feats, o_eot = dist_oracle.encode_text(tsd64, ids)
    o_emb = dist_oracle.forward_arch(sd, clips, arch, dtype=torch.float64)
    o_logits = dist_oracle.class_scores({k: v.double() for k, v in sd.items()}, o_emb, o_feats, softmax=False)
    rel = lambda a, ref: float((a.double() - ref.double()).norm() / ref.double().norm())
    errs = {"feats": rel(o_feats, feats), "eot": rel(o_eot, eot), "logits": rel(o_logits, logits)}
    print("[%s] reference+oracle %.1fs; oracle-vs-reference rel-L2: %s" % (name, time.time() - t0, {k: "%.2e" % v for k, v in errs.items()}))
    assert max(errs.values()) < 2e-6, errs
    fixture = {
        "case": name, "init": init, "arch": dict(arch.__dict__), "text": dict(tk), "prompts": prompts, "batch": batch,
        "weight_seed": 0, "text_seed": 5, "ids_seed": 11, "clip_seed": 1234,
        "text_checksum": synth.checksum(tsd), "ids": ids.clone(),
        "feats": feats.float().clone(), "eot": eot.float().clone(), "logits": logits.float().clone(),
        "oracle_vs_reference": errs, "torch": torch.__version__,
    }
    path = os.path.join(REPO, "tests", "golden", name + ".pt")
    torch.save(fixture, path)
    print("    wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024))


def _sub(x, step_rows, step_cols):
    return x[..., ::step_rows, ::step_cols].contiguous().float()


def run_case(name):
    arch, init, batch, kind = CASES[name]
    t0 = time.time()
    sd = synth.synth_state_dict(arch, seed=0, init=init)
    clips = synth.synth_clips(batch, arch, seed=1234, kind=kind)
    text = synth.synth_text_features(arch.num_classes, arch.embed_dim, seed=77)
    model = build_reference(arch, sd)
    b, T, t, N, g = batch, arch.frames, arch.sparse_frames, arch.tokens, arch.grid

    grabbed = {}

    def hook(key):
        def fn(_m, _i, out):
            grabbed[key] = out[0].detach() if isinstance(out, tuple) else out.detach()
        return fn

    handles = []
    for l in range(arch.layers):
        handles.append(model.visual.transformer.resblocks[l].register_forward_hook(hook("tap.%d" % l)))
    for i in range(len(arch.selected_layers)):
        handles.append(model.dist_net.temporal_nets[i].register_forward_hook(hook("tnet.%d" % i)))
        handles.append(model.dist_net.integration_nets[i].register_forward_hook(hook("res.%d" % i)))
    handles.append(model.dist_net.temporal_stem.register_forward_hook(hook("stem")))
    for j in range(arch.ada_layers):
        handles.append(model.dist_net.adapooling_nets[j].register_forward_hook(hook("ada.%d" % j)))

    frames = clips.permute(0, 2, 1, 3, 4).reshape(b * T, 3, arch.resolution, arch.resolution)  # backbone.py:233
    with torch.no_grad():
        emb = model.forward_without_text(frames)[:, 0]                                         # clip.py:466
        out = model(frames, torch.zeros(arch.num_classes, 8, dtype=torch.long), {"label_embeddings": text})
    logits = out["logits_per_image"]
    # with text the returned vid_logits are the L2-normalised embedding (clip.py:513,532)
    assert torch.allclose(out["vid_logits"][:, 0], emb / emb.norm(dim=1, keepdim=True), atol=1e-6, rtol=1e-5)
    probs = torch.softmax(logits.reshape(b, 1, -1).mean(dim=1), dim=-1)                         # backbone.py:241, base_blocks.py:579-585
    for h in handles:
        h.remove()
    t_ref = time.time() - t0

    # ---- pin the restatement against the reference ----
    t1 = time.time()
    o_emb, parts = dist_oracle.forward_arch(sd, clips, arch, dtype=torch.float64, return_parts=True)
    sd64 = {k: v.double() for k, v in sd.items()}
    o_probs = dist_oracle.class_scores(sd64, o_emb, text.double())
    o_logits = dist_oracle.class_scores(sd64, o_emb, text.double(), softmax=False)
    rel = lambda a, ref: float((a.double() - ref.double()).norm() / ref.double().norm())
    errs = {"emb": rel(o_emb, emb), "logits": rel(o_logits, logits), "probs": rel(o_probs, probs)}
    # reference layouts -> oracle layouts
    ref_parts = {}
    for l in range(arch.layers):
        ref_parts["tap.%d" % l] = grabbed["tap.%d" % l].permute(1, 0, 2)                        # [N,bt,D] -> [bt,N,D]
    ref_parts["stem"] = grabbed["stem"].permute(0, 2, 3, 4, 1)                                  # [b,C,T,g,g] -> [b,T,g,g,C]
    for i in range(len(arch.selected_layers)):
        ref_parts["res.%d" % i] = grabbed["res.%d" % i].permute(1, 0, 2)
    for j in range(arch.ada_layers):
        ref_parts["top.%d" % j] = grabbed["ada.%d" % j].permute(1, 0, 2)                        # [1,b,Ci] -> [b,1,Ci]
    for k, v in ref_parts.items():
        errs[k] = rel(parts[k], v)
    # tnet hook = TemporalNet output = xT before the integration->temporal add; check via a re-run
    for i in range(len(arch.selected_layers)):
        x_in = parts["stem"] if i == 0 else parts["xT.%d" % (i - 1)]
        o = dist_oracle.temporal_net(sd64, "dist_net.temporal_nets.%d" % i, x_in)
        errs["tnet.%d" % i] = rel(o, grabbed["tnet.%d" % i].permute(0, 2, 3, 4, 1))
    worst = max(errs.values())
    print("[%s] reference %.1fs, oracle %.1fs; oracle-vs-reference rel-L2: emb %.2e logits %.2e worst %.2e (%s)" % (
        name, t_ref, time.time() - t1, errs["emb"], errs["logits"], worst, max(errs, key=errs.get)))
    assert worst < 2e-6, errs

    big = arch.width >= 768
    rs, cs = (16, 32) if big else (1, 1)
    fixture = {
        "case": name, "init": init, "clip_kind": kind, "batch": batch, "weight_seed": 0, "clip_seed": 1234, "text_seed": 77,
        "arch": dict(arch.__dict__),
        "weights_checksum": synth.checksum({k: v for k, v in sd.items() if k != "logit_scale"}),
        "clips_checksum": synth.checksum(clips), "text_checksum": synth.checksum(text),
        "emb": emb.float().clone(), "logits": logits.float().clone(), "probs": probs.float().clone(),
        "top1": probs.argmax(dim=-1), "margin": (probs.topk(2, dim=-1).values[:, 0] - probs.topk(2, dim=-1).values[:, 1]),
        "sample_stride": (rs, cs),
        "parts": {k: _sub(v, rs, cs) for k, v in ref_parts.items()},
        "oracle_vs_reference": errs, "torch": torch.__version__,
    }
    path = os.path.join(REPO, "tests", "golden", name + ".pt")
    torch.save(fixture, path)
    print("    wrote %s (%.1f KB); top1 %s margins %s" % (path, os.path.getsize(path) / 1024, fixture["top1"].tolist(),
                                                     [round(float(m), 4) for m in fixture["margin"]]))


def main():
    torch.set_grad_enabled(False)
    _install_shims()
    sys.path.insert(0, REF)
    os.chdir(REF)                      # base.yaml is cwd-relative (utils/config.py:86)
    import models.base                 # noqa: F401  (registration side effects, models/base/__init__.py)
    names = sys.argv[1:] or (list(CASES) + list(TEXT_CASES))
    for n in names:
        (run_text_case if n in TEXT_CASES else run_case)(n)


if __name__ == "__main__":
    main()
