"""Registries for stems / branches / heads and the head used by the DiST configs.

``STEM_REGISTRY``, ``BRANCH_REGISTRY`` and ``HEAD_REGISTRY`` are the tables of the reference's
``models/base/base_blocks.py:19-21``; ``ClipVideoTextIdentity`` follows ``base_blocks.py:541-585``.
"""

import torch
import torch.nn as nn

from ...registry import Registry

STEM_REGISTRY = Registry("Stem")
BRANCH_REGISTRY = Registry("Branch")
HEAD_REGISTRY = Registry("Head")


@HEAD_REGISTRY.register()
class ClipVideoTextIdentity(nn.Module):
    """Mean over the view dimension of the cosine logits, softmax (or the configured activation) in eval.

    The backbone already produced class scores against the label embeddings, so the head owns no weights
    (``base_blocks.py:573-585``).  When the backbone ran the fused CUDA head it passes the probabilities along
    and they are returned as is - they are the same numbers.
    """

    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        name = cfg.VIDEO.HEAD.ACTIVATION
        if name == "softmax":
            self.activation = nn.Softmax(dim=-1)
        elif name == "sigmoid":
            self.activation = nn.Sigmoid()
        elif name == "identity":
            self.activation = nn.Identity()
        else:
            raise NotImplementedError("{} is not supported as an activationfunction.".format(name))
        self._softmax = name == "softmax"

    def forward(self, x):
        if isinstance(x, dict):
            fused = x.get("probs_per_image") if (not self.training and self._softmax) else None
            if fused is not None and x["logits_per_image"].shape[1] == 1:
                return fused, x
            out = x["logits_per_image"].mean(dim=1)
        else:
            out = x.mean(dim=1)
        if not self.training:
            out = self.activation(out)
        return out, x
