"""Round-2 fixtures from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY (build container; needs ``/root/reference``).

    python oracle/make_golden_r2.py [multi_b16_8x16 multi_b16_32x64 multi_l14_32x64 zeroshot_tiny zeroshot_b16 meter lr_policy
                                     train_b16_16x32 cpu_speed]

  multi_*          eight structured clips per full geometry through ``CLIP.forward`` (``clip.py:460,482-533``): embedding, class
                   logits, top-1 and the top-1/top-2 margin of every clip - the top-1 parity evidence.
                   Only outputs are stored (a few KB).
  zeroshot_*       the zero-shot / prediction-fusion branch (``clip.py:291-298,519-527``): per-frame ``ln_post @ proj`` image
                   embeddings and the 0.5 / 0.5 average of video and mean-frame logits.  The shipped reference reads two
                   attributes it never defines (``prediction_fusion_enable``, ``prediction_fusion_gating_enable``); the script
                   sets them to ``False`` - the un-gated branch, w = 0.5 - and changes nothing else.
  meter            ``utils/meters.py:24-175`` ``TestMeter`` (sum and max ensembles) on seeded scores with missing clips: video score
                   table, labels, clip counts, and the top-1 / top-5 strings of ``finalize_metrics``.
  lr_policy        ``models/utils/lr_policy.py:10-83`` on the DiST fine-tuning schedule and a step schedule; the parameter-group
                   classes of ``models/utils/optimizer.py:138-166`` (the five name lists, recomputed with its own rules because the
                   shipped constructor raises on a slicing typo after filling them).
  train_b16_16x32  BASELINE configs[4] at full size, one clip: loss, logits and the reference autograd's gradients of a dozen
                   ``dist_net`` tensors (sub-sampled) incl. ``temporal_nets.*.c_fc2``, ``integration_nets.*.ffn``, ``adapooling_nets.*.attn``.
  tokenizer        ``dataset/utils/simple_tokenizer.py:64-179`` on prompts with punctuation, contractions, digits, HTML entities, accents,
                   CJK and an over-long prompt (truncate=True): the ``[n, 77]`` id rows and the decoded strings.
  cpu_speed        wall time of the reference's own modules and of the oracle port for one B/16 8+16f clip on this container's
                   cores (how the ``--impl reference`` arm of bench.py, which can only run the port, relates to the real modules).
"""

import json
import os
import sys
import time
import types

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from dist_b200.arch import DistArch, tiny_arch  # noqa: E402
from dist_b200.utils import synth  # noqa: E402
from oracle import dist_oracle  # noqa: E402
from oracle import make_golden as mg  # noqa: E402
from oracle import train_oracle  # noqa: E402

GOLDEN = os.path.join(REPO, "tests", "golden")
L14 = dict(width=1024, layers=24, patch=14, embed_dim=768, frames=64, s_patch=14, ada_layers=4, num_classes=400, selected_layers=list(range(24)))
MULTI = {
    "multi_b16_8x16": (DistArch(), 8),
    "multi_b16_32x64": (DistArch(frames=64, ada_layers=4, num_classes=400), 8),
    "multi_l14_32x64": (DistArch(**L14), 8),
}
ZEROSHOT = {
    "zeroshot_tiny": (tiny_arch(), "scaled", 3),
    "zeroshot_b16": (DistArch(), "reference", 2),
}


def rel(a, ref):
    return float((a.double() - ref.double()).norm() / ref.double().norm().clamp_min(1e-30))


def save(name, fixture):
    path = os.path.join(GOLDEN, name + ".pt")
    torch.save(fixture, path)
    print("    wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024))


def run_multi(name):
    arch, batch = MULTI[name]
    t0 = time.time()
    sd = synth.synth_state_dict(arch, seed=0, init="reference")
    clips = synth.synth_clips(batch, arch, seed=4321, kind="structured")
    text = synth.synth_text_features(arch.num_classes, arch.embed_dim, seed=77)
    model = mg.build_reference(arch, sd)
    embs, logits = [], []
    with torch.no_grad():
        for i in range(batch):                                   # one clip at a time: bounded memory at L/14
            fr = clips[i:i + 1].permute(0, 2, 1, 3, 4).reshape(arch.frames, 3, arch.resolution, arch.resolution)
            out = model(fr, torch.zeros(arch.num_classes, 8, dtype=torch.long), {"label_embeddings": text})
            logits.append(out["logits_per_image"].reshape(1, 1, -1).mean(dim=1))
            embs.append(model.forward_without_text(fr)[:, 0] if i < 2 else None)
    logits = torch.cat(logits)
    top2 = logits.topk(2, dim=-1)
    # the oracle on the first two clips (pins the restatement at this geometry without doubling the run time)
    o = dist_oracle.forward_arch(sd, clips[:2], arch, dtype=torch.float64)
    e = rel(o, torch.cat(embs[:2]))
    assert e < 2e-6, e
    fixture = {"case": name, "arch": dict(arch.__dict__), "batch": batch, "init": "reference", "clip_kind": "structured",
               "weight_seed": 0, "clip_seed": 4321, "text_seed": 77,
               "weights_checksum": synth.checksum({k: v for k, v in sd.items() if k != "logit_scale"}),
               "clips_checksum": synth.checksum(clips), "text_checksum": synth.checksum(text),
               "emb": torch.cat(embs[:2]).float(), "logits": logits.float(), "top1": top2.indices[:, 0].clone(),
               "margin": (top2.values[:, 0] - top2.values[:, 1]).float(), "oracle_vs_reference": {"emb": e}, "torch": torch.__version__}
    print("[%s] %.1fs; top1 %s; logit margins %s; oracle-vs-reference emb %.2e" % (
        name, time.time() - t0, fixture["top1"].tolist(), [round(float(m), 4) for m in fixture["margin"]], e))
    save(name, fixture)


def run_zeroshot(name):
    arch, init, batch = ZEROSHOT[name]
    sd = synth.synth_state_dict(arch, seed=0, init=init)
    clips = synth.synth_clips(batch, arch, seed=1234, kind="structured")
    text = synth.synth_text_features(arch.num_classes, arch.embed_dim, seed=77)
    model = mg.build_reference(arch, sd)
    frames = clips.permute(0, 2, 1, 3, 4).reshape(batch * arch.frames, 3, arch.resolution, arch.resolution)
    ids = torch.zeros(arch.num_classes, 8, dtype=torch.long)
    with torch.no_grad():
        plain = model(frames, ids, {"label_embeddings": text})
        model.zero_shot_test = True                              # TEST.ZEROSHOT.ENABLE (clip.py:327)
        model.prediction_fusion_gating_enable = False            # never defined by the reference: the un-gated branch, w = 0.5 (clip.py:523-526)
        fused = model(frames, ids, {"label_embeddings": text})
    # restatement: img = ln_post(h_L[:, 0]) @ proj on the sparse frames (clip.py:291-298); fused = 0.5 (video + mean_t frame logits)
    sd64 = {k: v.double() for k, v in sd.items()}
    taps = dist_oracle.vit_forward(sd64, clips.double(), arch.alpha)
    img = dist_oracle.image_embeddings(sd64, taps[arch.layers - 1])
    emb = dist_oracle.forward_arch(sd, clips, arch, dtype=torch.float64)
    o_fused = dist_oracle.fused_class_scores(sd64, emb, img, text.double(), arch.sparse_frames)
    errs = {"img_raw": rel(img, plain["img_logits"]), "img_norm": rel(img / img.norm(dim=1, keepdim=True), fused["img_logits"]),
            "fused": rel(o_fused, fused["logits_per_image"])}
    print("[%s] oracle-vs-reference %s" % (name, {k: "%.2e" % v for k, v in errs.items()}))
    assert max(errs.values()) < 2e-6, errs
    save(name, {"case": name, "arch": dict(arch.__dict__), "batch": batch, "init": init, "clip_kind": "structured", "weight_seed": 0,
                "clip_seed": 1234, "text_seed": 77, "weights_checksum": synth.checksum({k: v for k, v in sd.items() if k != "logit_scale"}),
                "clips_checksum": synth.checksum(clips), "text_checksum": synth.checksum(text),
                "img_logits_raw": plain["img_logits"].float(), "img_logits_fused": fused["img_logits"].float(),
                "logits_plain": plain["logits_per_image"].float(), "logits_fused": fused["logits_per_image"].float(),
                "oracle_vs_reference": errs, "torch": torch.__version__})


def run_meter():
    from utils.meters import TestMeter
    import utils.logging as rlog
    rlog.log_json_stats = lambda stats: None                     # printing only
    g = torch.Generator().manual_seed(11)
    V, K, C = 37, 3, 174
    order = torch.randperm(V * K, generator=g)[: V * K - 2]      # two clips never arrive
    batches = list(order.split(16))
    labels_all = torch.randint(0, C, (V,), generator=g)
    # scores that favour the true class for about 60 % of the clips, so that top-1 / top-5 land strictly between 0 and 100 %
    preds = []
    for ids in batches:
        boost = torch.nn.functional.one_hot(labels_all[ids // K], C) * (torch.rand(len(ids), 1, generator=g) < 0.6) * 5.0
        preds.append(torch.softmax(2 * torch.randn(len(ids), C, generator=g) + boost, dim=-1))
    out = {"V": V, "K": K, "C": C, "clip_ids": batches, "labels_all": labels_all, "preds": preds}
    cfg = types.SimpleNamespace(LOG_PERIOD=10)
    for method in ("sum", "max"):
        m = TestMeter(cfg, V, K, C, len(batches), ensemble_method=method)
        for ids, p in zip(batches, preds):
            m.update_stats(p, labels_all[ids // K], ids)
        logged = {}
        rlog.log_json_stats = lambda stats, logged=logged: logged.update(stats)
        import utils.meters as um
        um.logging.log_json_stats = rlog.log_json_stats
        m.finalize_metrics(ks=(1, 5))
        out[method] = {"video_preds": m.video_preds.clone(), "video_labels": m.video_labels.clone(), "clip_count": m.clip_count.clone(),
                       "stats": dict(logged)}
        print("[meter/%s] %s" % (method, logged))
    save("meter", out)


def run_lr_policy():
    from models.utils import lr_policy
    from utils.config import Config
    out = {}
    for tag, path in (("ssv2_16x32", "configs/projects/dist/ssv2/vit-b16-16+32f.yaml"), ("k400_8x16", "configs/projects/dist/k400/vit-b16-8+16f.yaml")):
        c = Config(load=False, cfg_dict={})
        c.need_initialization = True                 # the base file is read on the first merge only (utils/config.py:80-93)
        args = types.SimpleNamespace(cfg_file=path, opts=[])
        d = c._merge_cfg_from_base(c._initialize_cfg(), c._load_yaml(args))
        cfg = Config(load=False, cfg_dict=d)
        o = cfg.OPTIMIZER
        epochs = [0.0, 0.25, 1.0, float(o.WARMUP_EPOCHS) - 0.5, float(o.WARMUP_EPOCHS), float(o.WARMUP_EPOCHS) + 0.37, o.MAX_EPOCH / 2.0,
                  o.MAX_EPOCH - 1.0, o.MAX_EPOCH - 0.01]
        out[tag] = {"optimizer": {k: getattr(o, k) for k in ("BASE_LR", "LR_POLICY", "MAX_EPOCH", "WARMUP_EPOCHS", "WARMUP_START_LR",
                                                             "NEW_NET_LRMULT", "NEW_NET_WEIGHT_DECAY") if hasattr(o, k)},
                    "epochs": epochs, "lr": [lr_policy.get_lr_at_epoch(cfg, e) for e in epochs]}
    steps = types.SimpleNamespace(OPTIMIZER=types.SimpleNamespace(BASE_LR=0.1, LR_POLICY="steps_with_relative_lrs", MAX_EPOCH=30, STEPS=[0, 10, 20],
                                                                  LRS=[1, 0.1, 0.01], WARMUP_EPOCHS=2.0, WARMUP_START_LR=0.01))
    ep = [0.0, 1.0, 2.0, 9.99, 10.0, 19.5, 20.0, 29.0]
    out["steps"] = {"optimizer": dict(vars(steps.OPTIMIZER)), "epochs": ep, "lr": [lr_policy.get_lr_at_epoch(steps, e) for e in ep]}
    # parameter-group classes of construct_DiST_optimizer (optimizer.py:145-166) on the tiny model's named parameters
    arch = tiny_arch()
    model = mg.build_reference(arch, synth.synth_state_dict(arch, seed=0, init="scaled"))
    lists = {k: [] for k in ("no_wd", "ada_normal", "ada_bias", "normal", "bias")}
    for name, p in model.named_parameters():
        if "dist_net" not in name:
            continue
        if name.endswith("cls_token") or name.endswith("positional_embedding"):
            lists["no_wd"].append(name)
        elif "adapooling_nets" in name:
            lists["ada_bias" if ("bias" in name or len(p.shape) == 1) else "ada_normal"].append(name)
        else:
            lists["bias" if ("bias" in name or len(p.shape) == 1) else "normal"].append(name)
    out["groups_tiny"] = lists
    with open(os.path.join(GOLDEN, "lr_policy.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("[lr_policy] wrote tests/golden/lr_policy.json:", {k: v["lr"][:3] for k, v in out.items() if "lr" in v})


SAMPLED = ["temporal_nets.0.temporal_net.c_fc2.weight", "temporal_nets.5.temporal_net.c_fc2.weight", "temporal_nets.11.temporal_net.c_fc1.weight",
           "temporal_nets.3.ln.weight", "integration_nets.0.ffn.c_fc.weight", "integration_nets.11.ffn.c_proj.weight",
           "integration_nets.6.temporal_ffn.c_fc2.weight", "adapooling_nets.0.spatial_transformer.attn.in_proj_weight",
           "adapooling_nets.1.temporal_transformer.attn.out_proj.weight", "input_linears.4.weight", "integration2temporal_nets.2.linear_fuse.weight",
           "temporal2integration_nets.7.linear_fuse.weight", "temporal2integration_nets.7.cls_token", "temporal_stem.weight", "proj",
           "aggregated_cls_token"]


def run_train_full():
    name = "train_b16_16x32"
    arch = DistArch(frames=32)
    t0 = time.time()
    sd = synth.synth_state_dict(arch, seed=0, init="reference")
    clips = synth.synth_clips(1, arch, seed=1234, kind="structured")
    text = synth.synth_text_features(arch.num_classes, arch.embed_dim, seed=77)
    target = synth.synth_soft_targets(1, arch.num_classes, seed=99)
    model = mg.build_reference(arch, sd)
    model.train()
    frames = clips.permute(0, 2, 1, 3, 4).reshape(arch.frames, 3, arch.resolution, arch.resolution)
    with torch.enable_grad():
        out = model(frames, torch.zeros(arch.num_classes, 8, dtype=torch.long), {"label_embeddings": text})
        logits = out["logits_per_image"].reshape(1, 1, -1).mean(dim=1)
        from models.utils.losses import SoftTargetCrossEntropy
        loss = SoftTargetCrossEntropy()(logits, target)
        loss.backward()
    grads = {"dist_net." + k: p.grad.detach().clone() for k, p in model.dist_net.named_parameters() if p.grad is not None}
    norms = {k: float(v.double().norm()) for k, v in grads.items()}
    keep = {}
    for k in SAMPLED:
        gfull = grads["dist_net." + k]
        flat = gfull.reshape(-1)
        step = max(1, flat.numel() // 4096)
        keep["dist_net." + k] = {"stride": step, "values": flat[::step].float().clone(), "norm": norms["dist_net." + k], "shape": tuple(gfull.shape)}
    total = sum(v * v for v in norms.values()) ** 0.5
    print("[%s] reference %.1fs; loss %.6f; %d gradient tensors, global norm %.4e; stored %d sampled tensors" % (
        name, time.time() - t0, float(loss), len(grads), total, len(keep)))
    save(name, {"case": name, "arch": dict(arch.__dict__), "batch": 1, "init": "reference", "clip_kind": "structured", "weight_seed": 0,
                "clip_seed": 1234, "text_seed": 77, "target_seed": 99,
                "weights_checksum": synth.checksum({k: v for k, v in sd.items() if k != "logit_scale"}),
                "clips_checksum": synth.checksum(clips), "text_checksum": synth.checksum(text), "target_checksum": synth.checksum(target),
                "loss": float(loss), "logits": logits.detach().float().clone(), "grad_norms": norms, "grad_global_norm": total, "sampled": keep,
                "torch": torch.__version__})


def run_cpu_speed():
    arch = DistArch()
    sd = synth.synth_state_dict(arch, seed=0, init="reference")
    clips = synth.synth_clips(1, arch, seed=1234, kind="structured")
    model = mg.build_reference(arch, sd)
    fr = clips.permute(0, 2, 1, 3, 4).reshape(arch.frames, 3, arch.resolution, arch.resolution)
    cores = os.cpu_count()
    torch.set_num_threads(cores)

    def med(fn, n=3):
        fn()
        ts = []
        for _ in range(n):
            t0 = time.perf_counter()
            fn()
            ts.append(time.perf_counter() - t0)
        return sorted(ts)[n // 2]

    with torch.no_grad():
        t_ref = med(lambda: model.forward_without_text(fr))
        t_port = med(lambda: dist_oracle.forward_arch(sd, clips, arch, dtype=torch.float32))
    out = {"cores": cores, "config": "DiST ViT-B/16 8+16f, 1 clip, fp32", "reference_s_per_clip": t_ref, "oracle_port_s_per_clip": t_port,
           "port_over_reference": t_port / t_ref, "torch": torch.__version__}
    with open(os.path.join(REPO, "profiles", "r2_cpu_port_vs_reference.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("[cpu_speed]", out)


PROMPTS = [
    "a video of a person pushing something from left to right", "Pretending to put something behind something", "a photo of a dog.",
    "riding a mountain bike", "person's hands can't be seen; it's 3 o'clock!", "tai chi", "playing   ukulele\t(close-up)", "UPPER lower MiXeD",
    "na\u00efve caf\u00e9 \u2014 d\u00e9j\u00e0 vu", "\u4eba\u5728\u8dd1\u6b65", "100% of 42 cats & dogs", "rock &amp; roll &lt;live&gt;", "hello<|endoftext|>world",
    "supercalifragilisticexpialidocious antidisestablishmentarianism", "", "emoji \U0001f600 test", "e=mc^2 ... ok?!", "don't you'll we've I'm he'd she's they're",
]


def run_tokenizer():
    """``dataset/utils/simple_tokenizer.py`` with its own merge list.  ``ftfy`` is not installed in this image; for these well-formed
    prompts ``ftfy.fix_text`` is the identity, which is what the shim supplies."""
    mod = types.ModuleType("ftfy")
    mod.fix_text = lambda t: t
    sys.modules.setdefault("ftfy", mod)
    from dataset.utils import simple_tokenizer as st
    tok = st.tokenize(PROMPTS, context_length=77)
    long_prompt = " ".join(["very long prompt about something"] * 30)
    trunc = st.tokenize([long_prompt], context_length=77, truncate=True)
    out = {"prompts": PROMPTS, "ids": tok.tolist(), "long_prompt": long_prompt, "long_truncated": trunc.tolist(),
           "decoded": [st._tokenizer.decode([t for t in row if t != 0]) for row in tok.tolist()],
           "vocab_size": len(st._tokenizer.encoder), "bpe_file": "dataset/utils/bpe_simple_vocab_16e6.txt.gz"}
    with open(os.path.join(GOLDEN, "tokenizer.json"), "w") as f:
        json.dump(out, f)
    print("[tokenizer] %d prompts, vocabulary %d, first row %s" % (len(PROMPTS), out["vocab_size"], tok[0, :12].tolist()))


def main():
    mg._install_shims()
    sys.path.insert(0, mg.REF)
    os.chdir(mg.REF)
    import models.base  # noqa: F401
    torch.set_grad_enabled(False)
    todo = sys.argv[1:] or (list(MULTI) + list(ZEROSHOT) + ["meter", "lr_policy", "train_b16_16x32", "cpu_speed", "tokenizer"])
    for n in todo:
        if n in MULTI:
            run_multi(n)
        elif n in ZEROSHOT:
            run_zeroshot(n)
        elif n == "meter":
            run_meter()
        elif n == "lr_policy":
            run_lr_policy()
        elif n == "train_b16_16x32":
            run_train_full()
        elif n == "cpu_speed":
            run_cpu_speed()
        elif n == "tokenizer":
            run_tokenizer()
        else:
            raise SystemExit("unknown case " + n)


if __name__ == "__main__":
    main()
