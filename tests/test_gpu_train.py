"""The fine-tuning step on the GPU (dist_b200/train.py) against the oracle's autograd and the reference's own gradients."""
import pytest
import torch

from conftest import load_golden, rel_l2
from helpers import check_inputs, inputs_for
from dist_b200.utils import synth
from oracle import train_oracle

pytestmark = pytest.mark.gpu


def _engine(fix, precision, **kw):
    from dist_b200.train import TrainEngine
    arch, sd, clips, text = inputs_for(fix)
    check_inputs(fix, sd, clips, text)
    target = synth.synth_soft_targets(fix["batch"], arch.num_classes, seed=fix["target_seed"])
    eng = TrainEngine(sd, arch, fix["batch"], device="cuda", precision=precision, text_features=text, **kw)
    return eng, arch, sd, clips, text, target


def _all_grads_err(got, want):
    num = sum(float((got[k].double() - want[k].double()).pow(2).sum()) for k in want) ** 0.5
    den = sum(float(want[k].double().pow(2).sum()) for k in want) ** 0.5
    return num / den


@pytest.mark.parametrize("name", ["train_tiny", "train_tiny_a3"])
@pytest.mark.parametrize("precision,bar_all,bar_each", [("fp32", 1e-4, 2e-3), ("bf16", 2e-2, 0.2)])
def test_gradients_match_reference(name, precision, bar_all, bar_each):
    fix = load_golden(name)
    eng, arch, sd, clips, text, target = _engine(fix, precision)
    loss = eng.forward_backward(clips.cuda(), target.cuda())
    torch.cuda.synchronize()
    got = eng.gradients()
    want = fix["grads"]
    assert sorted(got) == sorted(want)
    assert abs(float(loss) - fix["loss"]) < (1e-4 if precision == "fp32" else 2e-2) * abs(fix["loss"])
    worst = max(want, key=lambda k: rel_l2(got[k], want[k]))
    err_all = _all_grads_err(got, want)
    print("%s %s: loss %.6f (ref %.6f) all-grads rel-L2 %.2e, worst tensor %s %.2e" % (name, precision, float(loss), fix["loss"], err_all,
                                                                                  worst, rel_l2(got[worst], want[worst])))
    assert err_all < bar_all
    for k in want:
        assert got[k].shape == want[k].shape, k
        assert rel_l2(got[k], want[k]) < bar_each, k


def test_train_steps_follow_the_oracle():
    """Three AdamW steps in fp32 on the GPU track the float64 oracle (parameters and losses)."""
    fix = load_golden("train_tiny")
    eng, arch, sd, clips, text, target = _engine(fix, "fp32")
    state, cur = {}, dict(sd)
    for step in range(3):
        loss = eng.train_step(clips.cuda(), target.cuda(), lr=2e-4)
        o_loss, cur = train_oracle.train_step(cur, state, clips, text, target, arch, lr=2e-4, weight_decay=1e-4)
        assert abs(float(loss) - float(o_loss)) < 2e-4 * abs(float(o_loss)), step
    new = eng.state_dict()
    for k in train_oracle.unused_names(arch):
        assert torch.equal(new[k], sd[k].float()), k
    moved = sum(float((new[k].double() - sd[k].double()).pow(2).sum()) for k in new) ** 0.5
    err = sum(float((new[k].double() - cur[k].double()).pow(2).sum()) for k in new) ** 0.5
    # Adam normalises every element by its own history: elements whose gradient is at fp32 noise level move differently
    assert moved > 0 and err / moved < 1e-2, (err, moved)


def test_graph_replay_equals_eager():
    fix = load_golden("train_tiny")
    eng, arch, sd, clips, text, target = _engine(fix, "bf16")
    l0 = float(eng.forward_backward(clips.cuda(), target.cuda()))
    torch.cuda.synchronize()
    g0 = eng.gradients()
    eng.capture()
    l1 = float(eng.forward_backward(clips.cuda(), target.cuda()))
    torch.cuda.synchronize()
    g1 = eng.gradients()
    assert abs(l0 - l1) < 1e-5 * abs(l0)
    assert _all_grads_err(g1, g0) < 1e-4          # fp32 atomics reorder between runs
