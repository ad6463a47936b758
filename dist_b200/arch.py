"""Shape description of one DiST model instance.

The reference never names these quantities in one place: the ViT geometry is inferred from the
checkpoint (``models/base/clip.py:565-592``) and the DiST geometry is read from the YAML
(``models/module_zoo/branches/dist.py:19-28,51-53,71-75,93-97,117,170-176,190``).  ``DistArch``
collects both so that the engine, the weight packer, the synthetic-weight generator and the tests
agree on them.
"""

from dataclasses import dataclass, field
from typing import List

_VIT_PRESETS = {
    # META_ARCH_NAME -> (width, layers, patch, embed_dim)
    "ViT-B-16": (768, 12, 16, 512),
    "ViT-B/16": (768, 12, 16, 512),
    "ViT-L-14": (1024, 24, 14, 768),
    "ViT-L/14": (1024, 24, 14, 768),
}


@dataclass
class DistArch:
    width: int = 768            # D, ViT channel width
    layers: int = 12            # L, ViT depth
    patch: int = 16             # p, ViT patch size
    resolution: int = 224
    embed_dim: int = 512        # E, output embedding
    frames: int = 16            # T, dense frames (DATA.NUM_INPUT_FRAMES)
    alpha: int = 2              # DATA.SPARSE_SAMPLE_ALPHA
    integration_dim: int = 384  # Ci
    temporal_dim: int = 96      # Ct
    s_patch: int = 16           # DIST.S_PATCH_SIZE
    t_patch: int = 5            # DIST.T_PATCH_SIZE
    t_kernel: int = 3           # DIST.TEMPORAL_KERNEL_SIZE
    temporal_conv_mlp_ratio: float = 1
    integration_mlp_ratio: float = 1
    integration_temporal_mlp_ratio: float = 0.25
    ada_layers: int = 2         # DIST.ADA_POOLING_LAYERS
    selected_layers: List[int] = field(default_factory=lambda: list(range(12)))
    num_classes: int = 174

    # ---- derived -------------------------------------------------------------------------
    @property
    def grid(self):
        return self.resolution // self.patch

    @property
    def patches(self):          # P
        return self.grid * self.grid

    @property
    def tokens(self):           # N
        return self.patches + 1

    @property
    def heads(self):
        return self.width // 64

    @property
    def sparse_frames(self):    # t
        return self.frames // self.alpha

    @property
    def integration_heads(self):
        return self.integration_dim // 64

    @property
    def temporal_hidden(self):  # channels between the two TemporalNet convolutions
        return int(self.temporal_dim * self.temporal_conv_mlp_ratio)

    @property
    def integration_hidden(self):
        return int(self.integration_dim * self.integration_mlp_ratio)

    @property
    def integration_temporal_hidden(self):
        return int(self.integration_dim * self.integration_temporal_mlp_ratio)

    def validate(self):
        assert self.width % 64 == 0 and self.integration_dim % 64 == 0, "head dim is fixed at 64"
        assert self.frames % self.alpha == 0
        assert self.resolution % self.patch == 0
        # the integration->temporal add (dist.py:231) needs equal grids on both streams
        assert self.resolution // self.s_patch == self.grid, (
            "DIST.S_PATCH_SIZE {} does not match the ViT patch {} (the reference's L/14 YAMLs ship 16 and "
            "cannot run; use 14)".format(self.s_patch, self.patch))
        assert self.t_patch % 2 == 1 and self.t_kernel % 2 == 1
        assert len(self.selected_layers) >= 1 and max(self.selected_layers) < self.layers
        return self

    # ---- algorithmic work -------------------------------------------
    def flops_per_clip(self):
        """2 x MACs of one forward, patch-embed counted on the sparse frames only."""
        D, L, N, P, t, T = self.width, self.layers, self.tokens, self.patches, self.sparse_frames, self.frames
        Ci, Ct, Cm = self.integration_dim, self.temporal_dim, self.integration_temporal_hidden
        Ch, Ih = self.temporal_hidden, self.integration_hidden
        vit = L * t * (24 * N * D * D + 4 * N * N * D) + 2 * t * P * D * 3 * self.patch ** 2
        per_layer = (2 * T * P * Ct * Ch * self.t_kernel + 2 * T * P * Ch * Ct * 9      # TemporalNet
                     + 2 * t * N * D * Ci                                               # input linear
                     + 2 * t * P * Ci * Ct                                              # integration -> temporal
                     + 2 * t * P * self.alpha * Ct * Ci                                 # temporal -> integration
                     + 4 * t * N * Ci * Ih + 4 * t * N * Ci * Cm + 2 * self.t_kernel * t * N * Cm * Cm)
        stem = 2 * T * P * Ct * 3 * self.t_patch * self.s_patch ** 2
        ada = self.ada_layers * (4 * t * N * Ci * Ci + 4 * t * N * Ci + 16 * (t + 1) * Ci * Ci
                                 + 4 * (t + 1) * Ci * Ci)
        tail = 2 * D * Ci + 2 * Ci * self.embed_dim
        return {"vit": vit, "dist": len(self.selected_layers) * per_layer + stem + ada + tail,
                "total": vit + len(self.selected_layers) * per_layer + stem + ada + tail}


def arch_from_cfg(cfg, state_dict=None):
    """Build a :class:`DistArch` from a merged config, optionally refined by checkpoint shapes."""
    bb = cfg.VIDEO.BACKBONE
    name = getattr(bb, "META_ARCH_NAME", "ViT-B-16")
    width, layers, patch, embed = _VIT_PRESETS.get(name, _VIT_PRESETS["ViT-B-16"])
    resolution = 224
    if state_dict is not None and "visual.conv1.weight" in state_dict:
        # same inference as models/base/clip.py:568-573,582
        width = state_dict["visual.conv1.weight"].shape[0]
        patch = state_dict["visual.conv1.weight"].shape[-1]
        layers = len([k for k in state_dict if k.startswith("visual.") and k.endswith(".attn.in_proj_weight")])
        grid = round((state_dict["visual.positional_embedding"].shape[0] - 1) ** 0.5)
        resolution = patch * grid
        if "visual.proj" in state_dict:
            embed = state_dict["visual.proj"].shape[1]
    d = bb.DIST
    alpha = int(getattr(cfg.DATA, "SPARSE_SAMPLE_ALPHA", 1))
    arch = DistArch(
        width=width, layers=layers, patch=patch, resolution=resolution, embed_dim=embed,
        frames=int(cfg.DATA.NUM_INPUT_FRAMES), alpha=alpha,
        integration_dim=int(d.INTEGRATION_DIM), temporal_dim=int(d.TEMPORAL_DIM),
        s_patch=int(d.S_PATCH_SIZE), t_patch=int(d.T_PATCH_SIZE), t_kernel=int(d.TEMPORAL_KERNEL_SIZE),
        temporal_conv_mlp_ratio=d.TEMPORAL_CONV_MLP_RATIO, integration_mlp_ratio=d.INTEGRATION_MLP_RATIO,
        integration_temporal_mlp_ratio=d.INTEGRATION_TEMPORAL_MLP_RATIO,
        ada_layers=int(d.ADA_POOLING_LAYERS), selected_layers=list(d.SELECTED_LAYERS),
        num_classes=int(cfg.VIDEO.HEAD.NUM_CLASSES) if cfg.VIDEO.HEAD.NUM_CLASSES is not None else 0,
    )
    return arch.validate()


def tiny_arch(**kw):
    """A geometry small enough for CPU tests that still exercises every code path."""
    base = dict(width=128, layers=2, patch=16, resolution=64, embed_dim=64, frames=4, alpha=2,
                integration_dim=128, temporal_dim=32, s_patch=16, ada_layers=2, selected_layers=[0, 1],
                num_classes=10)
    base.update(kw)
    return DistArch(**base).validate()
