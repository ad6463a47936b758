"""CPU restatement of one DiST fine-tuning step.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import this file.

What it restates (paths relative to ``/root/reference``):
  * the forward of ``runs/train.py:101`` - the frozen CLIP ViT under ``no_grad`` (``clip.py:470-471``), the DiST
    branches in autograd, raw cosine logits in train mode (``base_blocks.py:579-585`` skips the softmax);
  * ``SoftTargetCrossEntropy`` (``models/utils/losses.py:20-31``): ``mean_b sum_c -target * log_softmax(logits)``;
  * the gradient w.r.t. every ``dist_net.*`` tensor (autograd over the restated forward of ``dist_oracle`` in float64);
  * the parameter groups the reference *intends* (``models/utils/optimizer.py:138-190``; the shipped constructor raises,
    SURVEY.md section 0.3): weight decay ``NEW_NET_WEIGHT_DECAY`` for >= 2-D weights, 0 for biases / 1-D tensors /
    ``cls_token`` / ``positional_embedding``; every group at ``lr * NEW_NET_LRMULT`` (``optimizer.py:207-214``);
  * ``torch.optim.AdamW`` with betas (0.9, 0.999), eps 1e-8 (``optimizer.py:67-73``), written out;
  * the cosine + linear warm-up schedule (``models/utils/lr_policy.py:10-44``).

Parity pin: ``oracle/make_golden_train.py`` runs the unmodified reference modules with autograd on the tiny geometry
and stores loss and gradients under ``tests/golden/train_*.pt``; ``tests/test_oracle_train.py`` checks this file against
them.
"""

import math

import torch

from . import dist_oracle


def trainable_names(sd):
    """``optimizer.py:148``: only tensors whose name contains ``dist_net`` are optimised (``logit_scale`` gets a
    gradient but never enters a parameter group)."""
    return sorted(k for k in sd if k.startswith("dist_net."))


def unused_names(arch):
    """Tensors that never receive a gradient: the integration->temporal branch of the last selected layer writes a
    temporal stream that nothing reads (``dist.py:231-235``), so autograd leaves ``p.grad = None`` and
    ``torch.optim.AdamW`` skips them (no update, no decay)."""
    i = len(arch.selected_layers) - 1
    return sorted("dist_net.integration2temporal_nets.%d.linear_fuse.%s" % (i, s) for s in ("weight", "bias"))


def weight_decay_of(name, shape, wd):
    """Group rule of ``optimizer.py:150-166`` (transformer / non-transformer groups share the same decay)."""
    if name.endswith("cls_token") or name.endswith("positional_embedding"):
        return 0.0
    if "bias" in name or len(shape) == 1:
        return 0.0
    return wd


def soft_target_cross_entropy(logits, target):
    """``losses.py:20-31``."""
    return torch.sum(-target * torch.log_softmax(logits, dim=-1), dim=-1).mean()


def loss_and_grads(sd, video, text_features, target, arch, dtype=torch.float64):
    """One forward + backward.  Returns (loss, logits [b, C], {name: grad}) with grads for every ``dist_net.*`` tensor."""
    sd = {k: v.detach().to(dtype) for k, v in sd.items() if k.startswith(("visual.", "dist_net.", "logit_scale"))}
    names = trainable_names(sd)
    for k in names:
        sd[k] = sd[k].clone().requires_grad_(True)
    video = video.to(dtype)
    with torch.no_grad():                                   # clip.py:470-471: the ViT never sees autograd
        taps = dist_oracle.vit_forward(sd, video, arch.alpha)
    emb = dist_oracle.dist_forward(sd, video, taps, arch.alpha, list(arch.selected_layers), arch.s_patch, arch.ada_layers)
    logits = dist_oracle.class_scores(sd, emb, text_features.to(dtype), softmax=False)
    loss = soft_target_cross_entropy(logits, target.to(dtype))
    grads = torch.autograd.grad(loss, [sd[k] for k in names], allow_unused=True)
    got = {k: g.detach() for k, g in zip(names, grads) if g is not None}
    assert sorted(set(names) - set(got)) == unused_names(arch)
    return loss.detach(), logits.detach(), got


def adamw_update(p, g, m, v, step, lr, weight_decay, beta1=0.9, beta2=0.999, eps=1e-8):
    """torch.optim.AdamW (decoupled decay) for one tensor; ``step`` counts from 1.  Returns (p, m, v)."""
    p = p * (1.0 - lr * weight_decay)
    m = beta1 * m + (1.0 - beta1) * g
    v = beta2 * v + (1.0 - beta2) * g * g
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    p = p - (lr / bc1) * m / denom
    return p, m, v


def lr_at_epoch(cur_epoch, base_lr, max_epoch, warmup_epochs, warmup_start_lr):
    """``lr_policy.py:10-44`` with ``LR_POLICY: cosine``."""
    cosine = lambda e: base_lr * (math.cos(math.pi * e / max_epoch) + 1.0) * 0.5
    if cur_epoch < warmup_epochs:
        alpha = (cosine(warmup_epochs) - warmup_start_lr) / warmup_epochs
        return cur_epoch * alpha + warmup_start_lr
    return cosine(cur_epoch)


def train_step(sd, state, video, text_features, target, arch, lr, weight_decay, dtype=torch.float64):
    """Forward, backward, AdamW on every ``dist_net.*`` tensor.  ``state`` = {"step": int, "m": {...}, "v": {...}} is
    updated in place; returns (loss, new_sd)."""
    loss, _, grads = loss_and_grads(sd, video, text_features, target, arch, dtype=dtype)
    state["step"] = state.get("step", 0) + 1
    new_sd = dict(sd)
    for k, g in grads.items():
        p = sd[k].detach().to(dtype)
        m = state.setdefault("m", {}).get(k, torch.zeros_like(p))
        v = state.setdefault("v", {}).get(k, torch.zeros_like(p))
        p, m, v = adamw_update(p, g, m, v, state["step"], lr, weight_decay_of(k, tuple(p.shape), weight_decay))
        state["m"][k], state["v"][k] = m, v
        new_sd[k] = p
    return loss, new_sd
