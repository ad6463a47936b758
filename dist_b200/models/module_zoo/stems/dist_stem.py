"""Tubelet stem of the DiST temporal encoder, registered for config-driven construction.

The reference builds it inline as ``DiSTNetwork.temporal_stem`` (``models/module_zoo/branches/dist.py:178-181``).
"""

import torch.nn as nn

from ...base.base_blocks import STEM_REGISTRY


@STEM_REGISTRY.register()
class DiSTTemporalStem(nn.Conv3d):
    """Conv3d(3 -> Ct, (t_patch, s_patch, s_patch), stride (1, s_patch, s_patch), pad (t_patch // 2, 0, 0)).

    On the CUDA path this is ``t_patch`` row-shifted GEMMs over patchified frames (``engine.py: dist.stem``)."""

    def __init__(self, cfg):
        d = cfg.VIDEO.BACKBONE.DIST
        super().__init__(3, d.TEMPORAL_DIM, kernel_size=(d.T_PATCH_SIZE, d.S_PATCH_SIZE, d.S_PATCH_SIZE),
                         stride=(1, d.S_PATCH_SIZE, d.S_PATCH_SIZE), padding=(d.T_PATCH_SIZE // 2, 0, 0))
