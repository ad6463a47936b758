"""Fine-tuning loop with the surface of the reference's ``runs/train.py:43-135`` (``train_epoch``).

The reference trains through autograd: ``optimizer.zero_grad(); loss.backward(); optimizer.step()`` on a DDP-wrapped
module.  On this path the forward, the hand-derived backward, the gradient all-reduce and AdamW are one planned CUDA
step (``dist_b200.train.TrainEngine``); the registry-built ``nn.Module`` only owns the parameters.  :class:`Trainer` is the
optimiser-shaped object that ties the two together:

  * it is built from the module's ``state_dict`` and the merged config (precision, weight decay, lr multiplier),
  * it exposes ``param_groups`` (``models/utils/optimizer.py:138-186``) so that ``optimizer.set_lr(trainer, lr)`` works
    unchanged (``optimizer.py:201-214``),
  * ``step(video, soft_targets)`` runs one iteration and returns the loss tensor,
  * ``sync_to_model()`` writes the master weights back into the module (bumping the parameter versions, which makes the
    module's inference engine re-pack) and ``save_checkpoint(path)`` writes ``{"model_state": ...}`` with the reference's
    key prefix (``utils/checkpoint.py:329``).
"""

import math

import torch

from .. import distributed as du
from ..models.utils import optimizer as optim
from ..utils import checkpoint as ckpt


def _core(model):
    """The CLIP module that owns ``visual.*`` / ``dist_net.*`` (``BaseVideoModel.backbone.base_encoder``)."""
    m = model.module if hasattr(model, "module") else model
    return m.backbone.base_encoder if hasattr(m, "backbone") else m


class Trainer:
    def __init__(self, model, cfg, clips_per_step, text_features=None, texts=None, process_group=None):
        from ..train import TrainEngine
        self.model, self.cfg = model, cfg
        core = _core(model)
        dev = core.logit_scale.device
        if dev.type != "cuda":
            raise RuntimeError("dist_b200 trains on a CUDA device only; there is no CPU path (move the model with .cuda())")
        if text_features is None:
            text_features = core.text_features
        if text_features is None and texts is not None:
            text_features = core.cache_text(texts.to(dev))[0]              # FREEZE_TEXT: the label set is encoded once (clip.py:483-486)
        if text_features is None:
            raise ValueError("the training head needs label embeddings: pass text_features [C, E] or token ids `texts` [C, ctx]")
        o = cfg.OPTIMIZER
        self.weight_decay = float(getattr(o, "NEW_NET_WEIGHT_DECAY", getattr(o, "WEIGHT_DECAY", 0.0)))
        sd = {k: v.detach() for k, v in core.state_dict().items()}
        self.engine = TrainEngine(sd, core.arch, clips_per_step, device=dev, precision=core.precision, text_features=text_features,
                                  weight_decay=self.weight_decay, process_group=process_group)
        self.param_groups = optim.construct_DiST_optimizer(sd, cfg)
        for g in self.param_groups:
            g["lr"] = 0.0
        # the fused AdamW keeps two decay classes in one flat buffer: the groups must describe the same split
        flat = self.engine.pt
        for g in self.param_groups:
            for n in g["names"]:
                if n in flat.unused:
                    continue
                assert (flat.offset[n] < flat.n_decay) == (g["group"] in ("normal", "ada_normal")), "parameter group / flat buffer mismatch for " + n
        self.iterations = 0

    # ---- optimiser surface -------------------------------------------------------------------
    @property
    def lr(self):
        lrs = {round(g["lr"], 15) for g in self.param_groups}
        if len(lrs) != 1:
            raise ValueError("the fused AdamW applies one learning rate to all DiST groups (every group carries NEW_NET_LRMULT); got %s" % sorted(lrs))
        return self.param_groups[0]["lr"]

    def zero_grad(self):
        """Nothing to do: the planned backward zeroes its gradient buffer (kept for loops written against torch.optim)."""

    def step(self, video, soft_targets):
        """One iteration: forward, backward, gradient all-reduce, AdamW (``runs/train.py:101-112``).  ``video`` ``[b,3,T,H,W]``
        (device or pinned host), ``soft_targets`` ``[b, C]`` (mixup / label-smoothed, ``losses.py:20-31``).  Returns the loss (device)."""
        loss = self.engine.train_step(video, soft_targets, self.lr)
        self.iterations += 1
        return loss

    def state_dict(self):
        return self.engine.state_dict()

    def sync_to_model(self):
        core = _core(self.model)
        new = self.engine.state_dict()
        own = dict(core.named_parameters())
        with torch.no_grad():
            for k, v in new.items():
                own[k].copy_(v.to(own[k].device, own[k].dtype))

    def save_checkpoint(self, path):
        self.sync_to_model()
        ckpt.save_checkpoint(path, _core(self.model).state_dict())


def soft_targets_of(labels, num_classes, device):
    """``labels["supervised_mixup"]`` when mixup ran (``runs/train.py:91-92``), else one-hot rows of ``labels["supervised"]``."""
    if isinstance(labels, dict):
        if "supervised_mixup" in labels:
            return labels["supervised_mixup"].to(device, torch.float32)
        labels = labels["supervised"]
    labels = labels.to(device)
    if labels.dtype.is_floating_point:
        return labels.float()
    return torch.nn.functional.one_hot(labels.long(), num_classes).float()


def train_epoch(train_loader, model, trainer, cur_epoch, cfg, mixup_fn=None, train_meter=None):
    """``runs/train.py:43-135``: per iteration the fractional-epoch learning rate (``:97-98``), one optimisation step, the NaN
    check of ``utils/misc.py:31``; the loss is averaged over ranks for logging only (``:118-119``).  Returns the mean loss."""
    core = _core(model)
    dev = core.logit_scale.device
    data_size = len(train_loader)
    folds = float(getattr(cfg.TRAIN, "NUM_FOLDS", 1)) if hasattr(cfg, "TRAIN") else 1.0
    total, count = 0.0, 0
    for cur_iter, (inputs, labels, indexes, meta) in enumerate(train_loader):
        video = inputs["video"] if isinstance(inputs, dict) else inputs
        if not video.is_cuda:
            video = (video if video.is_pinned() else video.pin_memory()).to(dev, non_blocking=True)
        if mixup_fn is not None:
            video, mixed = mixup_fn(video, labels["supervised"])
            labels = dict(labels, supervised_mixup=mixed)
        lr = optim.get_epoch_lr(cur_epoch + folds * float(cur_iter) / data_size, cfg)
        optim.set_lr(trainer, lr)
        loss = trainer.step(video, soft_targets_of(labels, trainer.engine.logits.shape[1], dev))
        value = float(loss)
        if math.isnan(value):
            raise RuntimeError("ERROR: Got NaN losses")                                   # utils/misc.py:31
        if du.get_world_size() > 1:
            value = float(du.all_reduce([loss.detach().clone().reshape(1)])[0])
        if train_meter is not None:
            train_meter.update_stats(None, None, value, lr, video.shape[0])
        total, count = total + value, count + 1
    trainer.sync_to_model()
    return total / max(count, 1)
