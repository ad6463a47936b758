"""The CPU restatement of the fine-tuning step (oracle/train_oracle.py) against the reference's own autograd
(fixtures written by oracle/make_golden_train.py), torch.optim.AdamW and the reference's LR schedule."""
import pytest
import torch

from conftest import load_golden, rel_l2
from helpers import check_inputs, inputs_for
from dist_b200.utils import synth
from oracle import train_oracle


def train_inputs(fix):
    arch, sd, clips, text = inputs_for(fix)
    check_inputs(fix, sd, clips, text)
    target = synth.synth_soft_targets(fix["batch"], arch.num_classes, seed=fix["target_seed"])
    assert abs(synth.checksum(target)[1] - fix["target_checksum"][1]) < 1e-6
    return arch, sd, clips, text, target


@pytest.mark.parametrize("name", ["train_tiny", "train_tiny_a3"])
def test_oracle_gradients_match_reference_autograd(name):
    fix = load_golden(name)
    arch, sd, clips, text, target = train_inputs(fix)
    loss, logits, grads = train_oracle.loss_and_grads(sd, clips, text, target, arch, dtype=torch.float64)
    assert abs(float(loss) - fix["loss"]) < 1e-6 * abs(fix["loss"])
    assert rel_l2(logits, fix["logits"]) < 2e-6
    assert sorted(grads) == sorted(fix["grads"])
    assert fix["unused"] == train_oracle.unused_names(arch)
    num = sum(float((grads[k].double() - fix["grads"][k].double()).pow(2).sum()) for k in grads) ** 0.5
    den = sum(float(fix["grads"][k].double().pow(2).sum()) for k in grads) ** 0.5
    assert num / den < 5e-6
    for k in grads:                       # per tensor: the reference ran in fp32
        assert rel_l2(grads[k], fix["grads"][k]) < 2e-4, k


def test_soft_targets_are_distributions():
    t = synth.synth_soft_targets(16, 174, seed=3)
    assert torch.allclose(t.sum(dim=1), torch.ones(16), atol=1e-6) and float(t.min()) > 0


def test_adamw_restatement_matches_torch():
    torch.manual_seed(0)
    p0 = torch.randn(7, 5, dtype=torch.float64)
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([p], lr=3.2e-4, betas=(0.9, 0.999), weight_decay=1e-4)     # optimizer.py:67-73
    q, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    for step in range(1, 6):
        g = torch.randn(7, 5, dtype=torch.float64)
        p.grad = g.clone()
        opt.step()
        q, m, v = train_oracle.adamw_update(q, g, m, v, step, 3.2e-4, 1e-4)
        assert torch.allclose(q, p.detach(), rtol=1e-12, atol=1e-14)


def test_weight_decay_groups():
    wd = train_oracle.weight_decay_of
    assert wd("dist_net.temporal2integration_nets.0.cls_token", (1, 1, 8, 384), 1e-4) == 0.0
    assert wd("dist_net.adapooling_nets.1.positional_embedding", (1, 8, 384), 1e-4) == 0.0
    assert wd("dist_net.input_linears.3.bias", (384,), 1e-4) == 0.0
    assert wd("dist_net.ln_post.weight", (384,), 1e-4) == 0.0
    assert wd("dist_net.temporal_stem.weight", (96, 3, 5, 16, 16), 1e-4) == 1e-4
    assert wd("dist_net.aggregated_cls_token", (1, 1, 384), 1e-4) == 0.0      # name ends with "cls_token" (optimizer.py:150)
    assert wd("dist_net.proj", (384, 512), 1e-4) == 1e-4


def test_lr_schedule_matches_reference_values():
    # values printed by the reference's models/utils/lr_policy.get_lr_at_epoch for configs/projects/dist/ssv2/vit-b16-16+32f.yaml
    want = {0.0: 8e-08, 0.5: 2.561367205045918e-06, 3.25: 1.6208886832798465e-05, 6.0: 2.9856406460551018e-05,
            10.0: 2.628460175498463e-05, 35.9: 6.09230973259045e-10}
    for e, lr in want.items():
        assert abs(train_oracle.lr_at_epoch(e, 3.2e-5, 36, 6, 8e-8) - lr) <= 1e-12 * max(lr, 1e-9)


def test_train_step_reduces_loss():
    fix = load_golden("train_tiny")
    arch, sd, clips, text, target = train_inputs(fix)
    state = {}
    cur = dict(sd)
    losses = []
    for _ in range(3):
        loss, cur = train_oracle.train_step(cur, state, clips, text, target, arch, lr=2e-4, weight_decay=1e-4)
        losses.append(float(loss))
    assert losses[2] < losses[0]
    for k in train_oracle.unused_names(arch):
        assert torch.equal(cur[k].float(), sd[k].float())
