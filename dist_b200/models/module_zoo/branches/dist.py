"""DiST branches: temporal encoder, integration branch, the two fusion networks and the ada-pooling head.

Parameter layout = the reference's (``models/module_zoo/branches/dist.py:16-202``); unlike the reference the
classes are registered in ``BRANCH_REGISTRY`` so a config can name them.  Arithmetic: ``dist_b200.engine``.
"""

from collections import OrderedDict

import torch
import torch.nn as nn

from ...base.base_blocks import BRANCH_REGISTRY, STEM_REGISTRY
from ...base.clip import CrossAttentionBlockGenral, LayerNorm, QuickGELU


def _mlp(d_in, d_hidden, d_out):
    return nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_in, d_hidden)), ("gelu", QuickGELU()), ("c_proj", nn.Linear(d_hidden, d_out))]))


@BRANCH_REGISTRY.register()
class IntegrationNetwork(nn.Module):
    """ffn(LN x) + temporal_ffn(LN_t x) on the integration stream (``dist.py:16-45``)."""

    def __init__(self, cfg, d_model):
        super().__init__()
        d = cfg.VIDEO.BACKBONE.DIST
        ci, kt = d.INTEGRATION_DIM, d.TEMPORAL_KERNEL_SIZE
        hid, cm = int(ci * d.INTEGRATION_MLP_RATIO), int(ci * d.INTEGRATION_TEMPORAL_MLP_RATIO)
        self.ffn = _mlp(ci, hid, ci)
        self.temporal_ffn = nn.Sequential(OrderedDict([
            ("c_fc1", nn.Conv3d(ci, cm, kernel_size=(1, 1, 1))),
            ("c_fc2", nn.Conv3d(cm, cm, kernel_size=(kt, 1, 1), padding=(kt // 2, 0, 0))),
            ("gelu1", QuickGELU()),
            ("c_proj", nn.Conv3d(cm, ci, kernel_size=1)),
        ]))
        self.ln = LayerNorm(ci)
        self.ln_temporal = LayerNorm(ci)
        self.num_frames = cfg.DATA.NUM_INPUT_FRAMES
        self.alpha = int(cfg.DATA.SPARSE_SAMPLE_ALPHA)


@BRANCH_REGISTRY.register()
class TemporalNet(nn.Module):
    """gelu(x + conv(1,3,3)(gelu(conv(kt,1,1)(LN_C x)))) on the dense temporal stream (``dist.py:48-65``)."""

    def __init__(self, cfg, d_model):
        super().__init__()
        d = cfg.VIDEO.BACKBONE.DIST
        ct, kt = d.TEMPORAL_DIM, d.TEMPORAL_KERNEL_SIZE
        hid = int(ct * d.TEMPORAL_CONV_MLP_RATIO)
        self.temporal_net = nn.Sequential(OrderedDict([
            ("c_fc1", nn.Conv3d(ct, hid, kernel_size=(kt, 1, 1), padding=(kt // 2, 0, 0))),
            ("gelu1", QuickGELU()),
            ("c_fc2", nn.Conv3d(hid, ct, kernel_size=(1, 3, 3), padding=(0, 1, 1))),
        ]))
        self.gelu = QuickGELU()
        self.ln = LayerNorm(ct)
        self.num_frames = cfg.DATA.NUM_INPUT_FRAMES


@BRANCH_REGISTRY.register()
class Temporal2IntegrationNetwork(nn.Module):
    """(alpha,1,1)/stride-alpha conv Ct -> Ci plus a learned per-frame class token (``dist.py:68-86``)."""

    def __init__(self, cfg, d_model):
        super().__init__()
        d = cfg.VIDEO.BACKBONE.DIST
        self.alpha = int(cfg.DATA.SPARSE_SAMPLE_ALPHA)
        self.num_frames = cfg.DATA.NUM_INPUT_FRAMES
        self.linear_fuse = nn.Conv3d(d.TEMPORAL_DIM, d.INTEGRATION_DIM, kernel_size=(self.alpha, 1, 1), stride=(self.alpha, 1, 1))
        self.cls_token = nn.Parameter(torch.zeros((1, 1, self.num_frames // self.alpha, d.INTEGRATION_DIM)))


@BRANCH_REGISTRY.register()
class Integration2TemporalNetwork(nn.Module):
    """Linear Ci -> Ct on the patch tokens, nearest upsample x alpha in time (``dist.py:90-105``)."""

    def __init__(self, cfg, d_model):
        super().__init__()
        d = cfg.VIDEO.BACKBONE.DIST
        self.alpha = int(cfg.DATA.SPARSE_SAMPLE_ALPHA)
        self.num_frames = cfg.DATA.NUM_INPUT_FRAMES
        self.linear_fuse = nn.Linear(d.INTEGRATION_DIM, d.TEMPORAL_DIM)


@BRANCH_REGISTRY.register()
class SpatialTemporalAdaPoolingNetwork(nn.Module):
    """Per-frame then per-clip single-query cross attention with token MLPs (``dist.py:108-162``)."""

    def __init__(self, cfg, d_model, layer_id):
        super().__init__()
        ci = cfg.VIDEO.BACKBONE.DIST.INTEGRATION_DIM
        self.num_frames = cfg.DATA.NUM_INPUT_FRAMES
        self.sparse_sample_alpha = getattr(cfg.DATA, "SPARSE_SAMPLE_ALPHA", 1)
        self.integration_dim = ci
        self.temporal_transformer = CrossAttentionBlockGenral(ci, ci // 64)
        self.positional_embedding = nn.Parameter(torch.zeros(1, self.num_frames // self.sparse_sample_alpha, ci))
        self.output_map_cls_token = _mlp(ci, ci * 4, ci)
        self.ln_out_temp_cls_token = LayerNorm(ci)
        self.spatial_transformer = CrossAttentionBlockGenral(ci, ci // 64)
        self.output_map_spatial_cls_token = _mlp(ci, ci * 4, ci)
        self.ln_out_spat_cls_token = LayerNorm(ci)


@BRANCH_REGISTRY.register()
class DiSTNetwork(nn.Module):
    """Container of the whole DiST side (``dist.py:165-202``); ``forward`` is served by the planned engine in ``CLIP``."""

    def __init__(self, cfg, d_model, width, output_dim):
        super().__init__()
        self.cfg = cfg
        d = cfg.VIDEO.BACKBONE.DIST
        self.selected_layers = list(d.SELECTED_LAYERS)
        n = len(self.selected_layers)
        ci = d.INTEGRATION_DIM
        self.alpha = int(cfg.DATA.SPARSE_SAMPLE_ALPHA)
        self.num_frames = cfg.DATA.NUM_INPUT_FRAMES
        self.temporal_stem = STEM_REGISTRY.get("DiSTTemporalStem")(cfg)
        self.input_linears = nn.ModuleList([nn.Linear(d_model, ci) for _ in range(n)])
        self.integration2temporal_nets = nn.ModuleList([Integration2TemporalNetwork(cfg, d_model=d_model) for _ in range(n)])
        self.temporal2integration_nets = nn.ModuleList([Temporal2IntegrationNetwork(cfg, d_model=d_model) for _ in range(n)])
        self.temporal_nets = nn.ModuleList([TemporalNet(cfg, d_model=d_model) for _ in range(n)])
        self.integration_nets = nn.ModuleList([IntegrationNetwork(cfg, d_model=d_model) for _ in range(n)])
        self.adapooling_nets = nn.ModuleList([SpatialTemporalAdaPoolingNetwork(cfg, d_model=d_model, layer_id=0)
                                              for _ in range(d.ADA_POOLING_LAYERS)])
        self.proj_spatial_cls_token = nn.Linear(d_model, ci)
        self.ln_post = LayerNorm(ci)
        self.proj = nn.Parameter((ci ** -0.5) * torch.randn(ci, output_dim))
        self.aggregated_cls_token = nn.Parameter(torch.zeros((1, 1, ci)))
        self.aggregated_spatial_cls_token = nn.Parameter(torch.zeros((1, 1, ci)))
        self._d_model, self._output_dim = d_model, output_dim
        self._engines = {}

    def forward(self, input):
        """``DiSTNetwork.forward`` of the reference (``dist.py:222-247``) as a stand-alone module: reads the taps
        ``input["mid_feat"]["img"][layer_id]`` (``[N, b*t, D]``, frame index ``b*t + ti``) of the selected layers and the frames
        ``input["images"]`` (``[b*T, 3, H, W]``); returns ``(cls_x [b, E], input)``."""
        from ....arch import arch_from_cfg
        from ....engine import DistEngine
        from ....utils import synth
        images = input["images"]
        if not images.is_cuda:
            raise RuntimeError("dist_b200 runs on a CUDA device only; there is no CPU path")
        taps_in = input["mid_feat"]["img"]
        n_tok, bt, c = taps_in[self.selected_layers[0]].shape
        t = self.num_frames // self.alpha
        b = bt // t
        key = (b, str(images.device), sum(p._version for p in self.parameters()))
        if key not in self._engines:
            self._engines.clear()
            own = {"dist_net." + k: v.detach() for k, v in self.state_dict().items()}
            arch = arch_from_cfg(self.cfg)
            arch.width, arch.embed_dim = self._d_model, self._output_dim
            arch.validate()
            sd = synth.synth_state_dict(arch, seed=0)              # the ViT half is never run here; any weights will do
            sd.update(own)
            b200 = getattr(self.cfg, "B200", None)
            self._engines[key] = DistEngine(sd, arch, b, device=images.device, precision=getattr(b200, "PRECISION", "bf16") if b200 else "bf16")
        eng = self._engines[key]
        hh, ww = images.shape[-2:]
        video = images.view(b, self.num_frames, 3, hh, ww).permute(0, 2, 1, 3, 4).float().contiguous()
        taps = {l: taps_in[l].permute(1, 0, 2).contiguous().float() for l in self.selected_layers}
        cls_x = eng.forward_dist(video, taps).clone()
        return cls_x, input
