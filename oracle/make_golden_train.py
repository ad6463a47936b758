"""Generate ``tests/golden/train_*.pt``: loss and gradients of one fine-tuning step from the UNMODIFIED reference.
TEST INFRASTRUCTURE ONLY (runs in the build container, where ``/root/reference`` exists).

    python oracle/make_golden_train.py

The reference modules are built exactly as in ``oracle/make_golden.py`` (same shims, same synthetic weights), put in
train mode the way ``runs/train.py`` does, and run through ``CLIP.forward(frames, texts, {"label_embeddings": ...})``
(``models/base/clip.py:460,482-533``; the ViT is under ``no_grad`` at ``clip.py:485-487``), the train-mode head
(``base_blocks.py:579-585``: raw logits) and the reference's own ``SoftTargetCrossEntropy``
(``models/utils/losses.py:20-31``), followed by ``loss.backward()``.  Stored: the loss, the logits, every ``dist_net.*``
gradient, the synthetic soft targets, and the oracle-vs-reference discrepancies (``oracle/train_oracle.py`` in float64
against the reference's fp32 autograd).
"""

import os
import sys
import time

import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)

from dist_b200.arch import tiny_arch  # noqa: E402
from dist_b200.utils import synth  # noqa: E402
from oracle import make_golden as mg  # noqa: E402
from oracle import train_oracle  # noqa: E402

CASES = {
    # name: (arch, weight init, batch, clip kind)
    "train_tiny": (tiny_arch(), "scaled", 2, "structured"),
    "train_tiny_a3": (tiny_arch(frames=6, alpha=3, resolution=96, selected_layers=[0, 1]), "scaled", 1, "iid"),
}


def run_case(name):
    arch, init, batch, kind = CASES[name]
    t0 = time.time()
    sd = synth.synth_state_dict(arch, seed=0, init=init)
    clips = synth.synth_clips(batch, arch, seed=1234, kind=kind)
    text = synth.synth_text_features(arch.num_classes, arch.embed_dim, seed=77)
    target = synth.synth_soft_targets(batch, arch.num_classes, seed=99)
    model = mg.build_reference(arch, sd)
    model.train()                                     # runs/train.py:75; CLIP.forward re-evals the ViT itself (clip.py:485)
    b, T = batch, arch.frames
    frames = clips.permute(0, 2, 1, 3, 4).reshape(b * T, 3, arch.resolution, arch.resolution)          # backbone.py:233
    with torch.enable_grad():
        out = model(frames, torch.zeros(arch.num_classes, 8, dtype=torch.long), {"label_embeddings": text})
        logits = out["logits_per_image"].reshape(b, 1, -1).mean(dim=1)                                  # backbone.py:241, base_blocks.py:579-585
        from models.utils.losses import SoftTargetCrossEntropy
        loss = SoftTargetCrossEntropy()(logits, target)
        loss.backward()
    # the integration->temporal branch of the LAST selected layer feeds a temporal stream nobody reads (dist.py:231-235):
    # its two tensors get no gradient (p.grad is None) and torch.optim.AdamW then leaves them untouched
    unused = sorted("dist_net." + k for k, p in model.dist_net.named_parameters() if p.grad is None)
    grads = {"dist_net." + k: p.grad.detach().clone() for k, p in model.dist_net.named_parameters() if p.grad is not None}
    assert all(p.grad is None for p in model.visual.parameters()), "the ViT must stay frozen"
    t_ref = time.time() - t0

    o_loss, o_logits, o_grads = train_oracle.loss_and_grads(sd, clips, text, target, arch, dtype=torch.float64)
    assert sorted(o_grads) == sorted(grads), (sorted(set(o_grads) ^ set(grads))[:6])
    assert unused == train_oracle.unused_names(arch), unused
    rel = lambda a, ref: float((a.double() - ref.double()).norm() / ref.double().norm().clamp_min(1e-30))
    errs = {"loss": abs(float(o_loss) - float(loss)) / abs(float(loss)), "logits": rel(o_logits, logits)}
    for k in grads:
        errs[k] = rel(o_grads[k], grads[k])
    worst = max(errs, key=errs.get)
    # global measure: all gradients as one vector (individual tiny-norm tensors are dominated by fp32 noise of the reference)
    num = sum(float((o_grads[k].double() - grads[k].double()).pow(2).sum()) for k in grads) ** 0.5
    den = sum(float(grads[k].double().pow(2).sum()) for k in grads) ** 0.5
    errs["all_grads"] = num / den
    print("[%s] reference %.1fs; %d gradient tensors; oracle-vs-reference: loss %.2e all-grads rel-L2 %.2e worst tensor %.2e (%s)" % (
        name, t_ref, len(grads), errs["loss"], errs["all_grads"], errs[worst], worst))
    assert errs["all_grads"] < 5e-6 and errs["loss"] < 1e-6, (errs["all_grads"], errs["loss"])
    assert errs[worst] < 2e-4, (worst, errs[worst])

    fixture = {
        "case": name, "init": init, "clip_kind": kind, "batch": batch, "weight_seed": 0, "clip_seed": 1234, "text_seed": 77,
        "target_seed": 99, "arch": dict(arch.__dict__),
        "weights_checksum": synth.checksum({k: v for k, v in sd.items() if k != "logit_scale"}),
        "clips_checksum": synth.checksum(clips), "text_checksum": synth.checksum(text), "target_checksum": synth.checksum(target),
        "loss": float(loss), "logits": logits.detach().float().clone(),
        "grads": {k: v.float() for k, v in grads.items()}, "unused": unused,
        "oracle_vs_reference": errs, "torch": torch.__version__,
    }
    path = os.path.join(REPO, "tests", "golden", name + ".pt")
    torch.save(fixture, path)
    print("    wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024))


def main():
    mg._install_shims()
    sys.path.insert(0, mg.REF)
    os.chdir(mg.REF)
    import models.base  # noqa: F401
    for n in sys.argv[1:] or list(CASES):
        run_case(n)


if __name__ == "__main__":
    main()
