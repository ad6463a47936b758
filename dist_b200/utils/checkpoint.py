"""Checkpoint ingest for real CLIP / DiST weights (reference: ``models/base/clip.py:614-629``,
``utils/checkpoint.py:277-347,452-529``, ``process_dist_cpkt.py:10-30``).

Formats the reference reads on this path:
  * a released DiST checkpoint ``*.pyth`` = ``{"model_state": state_dict}`` whose keys carry the
    ``backbone.base_encoder.`` prefix of ``BaseVideoModel`` (``utils/checkpoint.py:329``, loaded ``strict=False``);
    checkpoints written before the public release name the branch ``ladder_net.*`` - ``process_dist_cpkt.py`` renames
    them, and so does :func:`rename_legacy_keys`;
  * an OpenAI CLIP TorchScript archive (``ViT-B-16.pt`` / ``ViT-L-14.pt``), read with ``torch.jit.load(...).state_dict()``
    (``clip.py:618-623``) - only the ``visual.*`` tensors and ``logit_scale`` are on this path;
  * a plain ``state_dict`` file.
The result always uses the reference's un-prefixed key names (``visual.*``, ``dist_net.*``, ``logit_scale``), which is
what ``dist_b200.models.base.clip.build_model`` and the engines ingest.
"""

import collections
import os

import torch

_LEGACY = (  # process_dist_cpkt.py:13-24, in the same order (the generic "ladder_net.proj" rule must come last)
    ("ladder_net.temporal_stem", "dist_net.temporal_stem"),
    ("ladder_net.input_map_feat_nets", "dist_net.input_linears"),
    ("ladder_net.s2t_fuse_nets", "dist_net.integration2temporal_nets"),
    ("ladder_net.t2s_fuse_nets", "dist_net.temporal2integration_nets"),
    ("ladder_net.temporal_nets", "dist_net.temporal_nets"),
    ("ladder_net.spatial_nets", "dist_net.integration_nets"),
    ("ladder_net.final_temporal_nets", "dist_net.adapooling_nets"),
    ("ladder_net.proj_spatial_cls_token", "dist_net.proj_spatial_cls_token"),
    ("ladder_net.ln_post", "dist_net.ln_post"),
    ("ladder_net.proj", "dist_net.proj"),
    ("ladder_net.aggregated_cls_token", "dist_net.aggregated_cls_token"),
    ("ladder_net.aggregated_spatial_cls_token", "dist_net.aggregated_spatial_cls_token"),
)
PREFIX = "backbone.base_encoder."


def rename_legacy_keys(model_state):
    out = collections.OrderedDict()
    for name, param in model_state.items():
        if "ladder_net" in name:
            for old, new in _LEGACY:
                name = name.replace(old, new)
            if "ladder_net" in name:
                raise KeyError("unknown legacy DiST key: " + name)
        out[name] = param
    return out


def normalise_state_dict(obj):
    """Any of the accepted containers -> ``{visual.*, dist_net.*, logit_scale, ...}`` with plain tensors."""
    if isinstance(obj, dict) and "model_state" in obj:
        obj = obj["model_state"]
    if hasattr(obj, "state_dict") and not isinstance(obj, dict):           # a TorchScript / nn.Module object
        obj = obj.state_dict()
    sd = rename_legacy_keys(obj)
    out = collections.OrderedDict()
    for k, v in sd.items():
        if k.startswith("module."):
            k = k[len("module."):]
        if k.startswith(PREFIX):
            k = k[len(PREFIX):]
        out[k] = v.detach() if torch.is_tensor(v) else v
    return out


def load_state_dict(path):
    """Read a ``.pyth`` / ``.pth`` checkpoint or an OpenAI CLIP TorchScript archive from ``path``."""
    if not os.path.exists(path):
        raise FileNotFoundError(path)
    if path.endswith((".pyth", ".pth", ".bin")):
        return normalise_state_dict(torch.load(path, map_location="cpu", weights_only=False))
    try:
        return normalise_state_dict(torch.jit.load(path, map_location="cpu"))
    except RuntimeError:
        return normalise_state_dict(torch.load(path, map_location="cpu", weights_only=False))


def save_checkpoint(path, state_dict, prefix=PREFIX):
    """``{"model_state": ...}`` with the ``BaseVideoModel`` prefix, as ``utils/checkpoint.py:329`` writes it."""
    torch.save({"model_state": collections.OrderedDict((prefix + k, v.detach().cpu()) for k, v in state_dict.items())}, path)


def load_test_checkpoint(cfg, model, path=None):
    """``utils/checkpoint.py:452-529`` (``load_test_checkpoint``) for this path: read ``TEST.CHECKPOINT_FILE_PATH`` (or ``path``),
    strip the ``backbone.base_encoder.`` prefix the released DiST checkpoints carry and load the tensors into the model's CLIP
    module.  Fails loudly when the checkpoint does not provide the DiST branches (a silent ``strict=False`` load of a checkpoint
    whose keys all miss - e.g. prefixed keys into the un-prefixed module - would leave ``dist_net`` at its initial values).
    Returns ``(missing_keys, unexpected_keys)``."""
    path = path or getattr(cfg.TEST, "CHECKPOINT_FILE_PATH", "")
    if not path:
        raise ValueError("TEST.CHECKPOINT_FILE_PATH is empty and no path was given")
    sd = load_state_dict(path)
    m = model.module if hasattr(model, "module") else model
    core = m.backbone.base_encoder if hasattr(m, "backbone") else m
    own = core.state_dict()
    bad = [k for k, v in sd.items() if k in own and tuple(own[k].shape) != tuple(v.shape)]
    if bad:
        raise RuntimeError("size mismatch for: %s (DiST weights are frame-count specific: cls_token / positional_embedding)" % bad[:10])
    hit = [k for k in sd if k in own and k.startswith("dist_net.")]
    want = [k for k in own if k.startswith("dist_net.")]
    if len(hit) != len(want):
        raise RuntimeError("the checkpoint provides %d of the %d dist_net tensors (first missing: %s)"
                           % (len(hit), len(want), sorted(set(want) - set(hit))[:5]))
    res = core.load_state_dict({k: v for k, v in sd.items() if k in own}, strict=False)
    if hasattr(core, "refresh_engine"):
        core.refresh_engine()
    return list(res.missing_keys), sorted(k for k in sd if k not in own)
