"""``MODEL_REGISTRY`` and ``BaseVideoModel`` (reference: ``models/base/models.py:10-67``)."""

import torch.nn as nn

from ...registry import Registry
from .backbone import BACKBONE_REGISTRY
from .base_blocks import HEAD_REGISTRY

MODEL_REGISTRY = Registry("Model")


class BaseVideoModel(nn.Module):
    """backbone (``cfg.VIDEO.BACKBONE.META_ARCH``) followed by head (``cfg.VIDEO.HEAD.NAME``)."""

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        backbone_cls = BACKBONE_REGISTRY.get(cfg.VIDEO.BACKBONE.META_ARCH)
        head_cls = HEAD_REGISTRY.get(cfg.VIDEO.HEAD.NAME)
        if backbone_cls is None:
            raise KeyError("backbone {!r} is not registered (have {})".format(cfg.VIDEO.BACKBONE.META_ARCH, sorted(BACKBONE_REGISTRY.get_all_registered())))
        if head_cls is None:
            raise KeyError("head {!r} is not registered (have {})".format(cfg.VIDEO.HEAD.NAME, sorted(HEAD_REGISTRY.get_all_registered())))
        self.backbone = backbone_cls(cfg=cfg)
        self.head = head_cls(cfg=cfg)

    def forward(self, x):
        return self.head(self.backbone(x))

    def train(self, mode=True):
        # models.py:47-67: normalisation layers stay in eval when cfg.BN.FREEZE is set
        self.training = mode
        super().train(mode)
        if getattr(self.cfg.BN, "FREEZE", False):
            for m in self.modules():
                if isinstance(m, (nn.BatchNorm2d, nn.BatchNorm3d, nn.LayerNorm)):
                    m.train(False)
        return self
