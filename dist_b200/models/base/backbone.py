"""``BACKBONE_REGISTRY`` and the CLIP + DiST backbone adapter (reference: ``models/base/backbone.py:16,218-256``)."""

import torch.nn as nn

from ...registry import Registry

BACKBONE_REGISTRY = Registry("Backbone")


@BACKBONE_REGISTRY.register()
class ClipVisionTextTransformer(nn.Module):
    """``forward({"video": [B,3,T,H,W], "texts": ...}) -> dict`` with the reference's output keys.

    The reference permutes the clip to ``[B*T,3,H,W]`` before calling CLIP (``backbone.py:232-233``) and the DiST
    stem permutes it straight back (``dist.py:225``); both copies are dropped here - the CUDA path reads the
    clip in its native ``[B,3,T,H,W]`` layout.  ``texts`` is the int64 token matrix ``[C, ctx]`` of the reference
    (encoded once by the CLIP text tower of the checkpoint, ``clip.py:436-452``; or ignored in favour of embeddings cached
    with ``set_text_features``) or directly a float ``[C, E]`` label-embedding matrix.
    """

    def __init__(self, cfg):
        super().__init__()
        from . import clip
        self.base_encoder = clip.load(cfg)

    def forward(self, x):
        video = x["video"]
        b = video.shape[0]
        if "texts" in x:
            out = self.base_encoder(video, x["texts"], x.get("others"))
            lpi = out["logits_per_image"]
            out["logits_per_image"] = lpi.reshape(b, 1, -1)           # backbone.py:241
            return out
        # backbone.py:244-251 (the reference call there has a wrong arity; forward_without_text is the intent)
        return self.base_encoder(video, None)

    def set_text_features(self, feats):
        self.base_encoder.set_text_features(feats)

    def get_num_layers(self):
        return self.base_encoder.visual.transformer.layers, 0
