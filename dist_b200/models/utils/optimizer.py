"""Parameter groups and learning-rate plumbing of the DiST fine-tuning step (reference: ``models/utils/optimizer.py:138-214``).

Only tensors whose name contains ``dist_net`` are optimised (``optimizer.py:148``); the CLIP towers stay frozen.  The
reference's ``construct_DiST_optimizer`` sorts them into five lists - no-weight-decay tokens, ada-pooling weights / biases,
other weights / biases - and gives every list ``lr_mult = NEW_NET_LRMULT`` and either ``NEW_NET_WEIGHT_DECAY`` or 0.  (As
shipped the constructor raises on a slicing typo; the groups below are the ones it spells out.)  Since all
lists share one multiplier and only two decay values exist, the fused AdamW of ``dist_b200.train`` keeps the tensors in two
contiguous runs of one flat buffer - decayed first - and these groups describe exactly that split.
"""

from . import lr_policy


def decay_class(name, ndim):
    """'no_wd' / 'bias' / 'normal' for a ``dist_net`` tensor (``optimizer.py:150-166``)."""
    if name.endswith("cls_token") or name.endswith("positional_embedding"):
        return "no_wd"
    if "bias" in name or ndim == 1:
        return "bias"
    return "normal"


def weight_decay_of(name, ndim, weight_decay):
    return weight_decay if decay_class(name, ndim) == "normal" else 0.0


def construct_DiST_optimizer(model, cfg):
    """Parameter groups ``[{"names", "params", "weight_decay", "lr_mult"}]`` over the ``dist_net`` tensors of ``model`` (an
    ``nn.Module`` or a ``{name: tensor}`` state dict), in the reference's order: tokens without decay, ada-pooling weights and
    biases, remaining weights and biases.  Empty groups are dropped."""
    items = model.named_parameters() if hasattr(model, "named_parameters") else model.items()
    o = cfg.OPTIMIZER
    wd, mult = float(getattr(o, "NEW_NET_WEIGHT_DECAY", getattr(o, "WEIGHT_DECAY", 0.0))), float(getattr(o, "NEW_NET_LRMULT", 1.0))
    lists = {k: [] for k in ("no_wd", "ada_normal", "ada_bias", "normal", "bias")}
    for name, p in items:
        if "dist_net" not in name or (hasattr(p, "requires_grad") and hasattr(model, "named_parameters") and not p.requires_grad):
            continue
        cls = decay_class(name, p.dim())
        if cls != "no_wd" and "adapooling_nets" in name:
            cls = "ada_" + cls
        lists[cls].append((name, p))
    groups = []
    for key, decay in (("no_wd", 0.0), ("ada_normal", wd), ("ada_bias", 0.0), ("normal", wd), ("bias", 0.0)):
        if lists[key]:
            groups.append({"group": key, "names": [n for n, _ in lists[key]], "params": [p for _, p in lists[key]],
                           "weight_decay": decay, "lr_mult": mult})
    return groups


def get_epoch_lr(cur_epoch, cfg):
    """``optimizer.py:189-198``."""
    return lr_policy.get_lr_at_epoch(cfg, cur_epoch)


def set_lr(optimizer, new_lr):
    """``optimizer.py:201-214``: every group gets ``new_lr`` scaled by its ``lr_mult`` (or ``/ 10`` when flagged ``lr_reduce``).
    ``optimizer`` is anything with ``param_groups`` - a ``torch.optim`` optimiser or ``dist_b200.runs.train.Trainer``."""
    for group in optimizer.param_groups:
        if group.get("lr_reduce"):
            group["lr"] = new_lr / 10
        elif "lr_mult" in group:
            group["lr"] = new_lr * group["lr_mult"]
        else:
            group["lr"] = new_lr
