"""Host-side logic of the fine-tuning surface and of checkpoint ingest (CPU): learning-rate policy and parameter groups against
values produced by the unmodified reference (``tests/golden/lr_policy.json``, ``oracle/make_golden_r2.py``), ``build_model``'s
``load_state_dict(strict=False)`` semantics, and ``load_test_checkpoint``."""
import json
import os
import types

import pytest
import torch

from conftest import GOLDEN, ROOT
from dist_b200.arch import tiny_arch
from dist_b200.config import Config
from dist_b200.models.utils import lr_policy, optimizer as optim
from dist_b200.utils import checkpoint, synth


def _golden():
    return json.load(open(os.path.join(GOLDEN, "lr_policy.json")))


@pytest.mark.parametrize("tag,path", [("ssv2_16x32", "configs/projects/dist/ssv2/vit-b16-16+32f.yaml"),
                                      ("k400_8x16", "configs/projects/dist/k400/vit-b16-8+16f.yaml")])
def test_lr_policy_matches_reference_on_the_dist_schedules(tag, path):
    g = _golden()[tag]
    cfg = Config.from_file(os.path.join(ROOT, path))
    for k, v in g["optimizer"].items():                       # the merged OPTIMIZER section is the reference's
        assert getattr(cfg.OPTIMIZER, k) == pytest.approx(v), k
    for e, want in zip(g["epochs"], g["lr"]):
        assert lr_policy.get_lr_at_epoch(cfg, e) == pytest.approx(want, rel=1e-12, abs=1e-18), e
        assert optim.get_epoch_lr(e, cfg) == pytest.approx(want, rel=1e-12, abs=1e-18)


def test_step_policy_and_unknown_policy():
    g = _golden()["steps"]
    cfg = types.SimpleNamespace(OPTIMIZER=types.SimpleNamespace(**g["optimizer"]))
    for e, want in zip(g["epochs"], g["lr"]):
        assert lr_policy.get_lr_at_epoch(cfg, e) == pytest.approx(want, rel=1e-12)
    cfg.OPTIMIZER.LR_POLICY = "nope"
    with pytest.raises(NotImplementedError):
        lr_policy.get_lr_at_epoch(cfg, 1.0)


def test_parameter_groups_match_reference_lists_and_set_lr():
    g = _golden()["groups_tiny"]
    arch = tiny_arch()
    sd = synth.synth_state_dict(arch, seed=0, init="scaled")
    cfg = Config.from_file(os.path.join(ROOT, "configs/projects/dist/ssv2/vit-b16-16+32f.yaml"))
    groups = optim.construct_DiST_optimizer(sd, cfg)
    by = {x["group"]: x for x in groups}
    assert [x["group"] for x in groups] == [k for k in ("no_wd", "ada_normal", "ada_bias", "normal", "bias") if g[k]]
    for k, names in g.items():
        assert sorted(by[k]["names"]) == sorted(names), k
    assert by["normal"]["weight_decay"] == by["ada_normal"]["weight_decay"] == cfg.OPTIMIZER.NEW_NET_WEIGHT_DECAY
    assert by["no_wd"]["weight_decay"] == by["bias"]["weight_decay"] == by["ada_bias"]["weight_decay"] == 0.0
    assert not any(n.startswith("visual.") for x in groups for n in x["names"])
    holder = types.SimpleNamespace(param_groups=groups + [{"lr_reduce": True}, {}])
    optim.set_lr(holder, 2e-5)
    assert all(x["lr"] == pytest.approx(2e-5 * cfg.OPTIMIZER.NEW_NET_LRMULT) for x in groups)
    assert holder.param_groups[-2]["lr"] == pytest.approx(2e-6) and holder.param_groups[-1]["lr"] == pytest.approx(2e-5)


def _tiny_cfg():
    import dist_b200.models.base  # noqa: F401
    arch = tiny_arch()
    cfg = Config.from_file(os.path.join(ROOT, "configs/projects/dist/ssv2/vit-b16-8+16f.yaml"), ["NUM_GPUS", "0"])
    d = cfg.VIDEO.BACKBONE.DIST
    cfg.DATA.NUM_INPUT_FRAMES, cfg.DATA.SPARSE_SAMPLE_ALPHA = arch.frames, arch.alpha
    d.INTEGRATION_DIM, d.TEMPORAL_DIM, d.S_PATCH_SIZE, d.ADA_POOLING_LAYERS = arch.integration_dim, arch.temporal_dim, arch.s_patch, arch.ada_layers
    d.SELECTED_LAYERS = list(arch.selected_layers)
    cfg.VIDEO.HEAD.NUM_CLASSES = arch.num_classes
    return cfg, arch


def test_build_model_initialises_missing_dist_branches_like_the_reference():
    """A CLIP-only checkpoint (the normal fine-tuning start): dist_net must get trunc_normal(0.02) weights / tokens and zero biases
    (dist.py:77-79,120-121,195-220), not torch's kaiming defaults and zero tokens; the keys are recorded."""
    from dist_b200.models.base.clip import build_model
    cfg, arch = _tiny_cfg()
    full = synth.synth_state_dict(arch, seed=0, init="reference")
    clip_only = {k: v for k, v in full.items() if not k.startswith("dist_net.")}
    model = build_model(cfg, clip_only)
    sd = model.state_dict()
    fresh = [k for k in sd if k.startswith("dist_net.")]
    assert sorted(model.initialised_keys) == sorted(fresh) and sorted(model.missing_keys) == sorted(fresh)
    for k in fresh:
        v = sd[k].float()
        if k.endswith(".bias"):
            assert float(v.abs().max()) == 0.0, k
        elif v.dim() >= 2 and v.numel() >= 256 and not k.endswith("proj") and "in_proj_weight" not in k:      # MHA in-projections keep xavier_uniform
            assert 0.012 < float(v.std()) < 0.028, (k, float(v.std()))      # trunc_normal_(std=.02); its a/b = -2/2 are absolute, i.e. no effective truncation
    for k in ("dist_net.aggregated_cls_token", "dist_net.temporal2integration_nets.0.cls_token", "dist_net.adapooling_nets.0.positional_embedding"):
        assert float(sd[k].abs().max()) > 0, k                        # torch's default would leave the tokens at zero
    assert torch.equal(sd["visual.conv1.weight"], full["visual.conv1.weight"])


def test_build_model_raises_on_shape_mismatch_and_records_unexpected():
    from dist_b200.models.base.clip import build_model
    cfg, arch = _tiny_cfg()
    full = synth.synth_state_dict(arch, seed=0, init="reference")
    bad = dict(full)
    k = "dist_net.temporal2integration_nets.0.cls_token"                  # trained for another frame count
    bad[k] = torch.zeros(1, 1, full[k].shape[2] + 2, full[k].shape[3])
    with pytest.raises(RuntimeError, match="size mismatch"):
        build_model(cfg, bad)
    extra = dict(full, **{"some.other.head.weight": torch.zeros(3)})
    model = build_model(cfg, extra)
    assert model.unexpected_keys == ["some.other.head.weight"] and model.missing_keys == []


def test_load_test_checkpoint_strips_the_prefix_and_fails_loudly(tmp_path):
    import dist_b200.models.base  # noqa: F401
    from dist_b200.models.base.clip import build_model
    cfg, arch = _tiny_cfg()
    a = synth.synth_state_dict(arch, seed=0, init="reference")
    b = synth.synth_state_dict(arch, seed=1, init="reference")
    core = build_model(cfg, a)
    model = types.SimpleNamespace(backbone=types.SimpleNamespace(base_encoder=core))
    path = str(tmp_path / "dist.pyth")
    checkpoint.save_checkpoint(path, b)                                   # {"model_state": {"backbone.base_encoder.<k>": ...}}
    raw = torch.load(path, weights_only=False)["model_state"]
    assert all(k.startswith("backbone.base_encoder.") for k in raw)
    # the recipe the first INTEGRATION.md gave - prefixed keys into the un-prefixed module with strict=False - loads NOTHING, silently
    res = core.load_state_dict(raw, strict=False)
    assert len(res.unexpected_keys) == len(raw) and torch.equal(core.state_dict()["dist_net.proj"], a["dist_net.proj"])
    cfg.TEST.CHECKPOINT_FILE_PATH = path
    missing, unexpected = checkpoint.load_test_checkpoint(cfg, model)
    assert unexpected == [] and all(not k.startswith("dist_net.") for k in missing)
    got = core.state_dict()
    assert all(torch.equal(got[k], b[k]) for k in b if k in got)
    # a checkpoint without the DiST branches must not pass as a test checkpoint
    clip_only = str(tmp_path / "clip_only.pyth")
    checkpoint.save_checkpoint(clip_only, {k: v for k, v in b.items() if not k.startswith("dist_net.")})
    with pytest.raises(RuntimeError, match="dist_net"):
        checkpoint.load_test_checkpoint(cfg, model, clip_only)


def test_module_forward_refuses_autograd_training():
    """The registry-built module runs the planned inference path: train mode + grad must fail loudly, on any device."""
    from dist_b200.models.base.clip import build_model
    cfg, arch = _tiny_cfg()
    core = build_model(cfg, synth.synth_state_dict(arch, seed=0, init="reference"))
    core.train()
    with pytest.raises(RuntimeError, match="Trainer"):
        core(torch.zeros(arch.frames, 3, arch.resolution, arch.resolution), None, {"label_embeddings": torch.zeros(arch.num_classes, arch.embed_dim)})
