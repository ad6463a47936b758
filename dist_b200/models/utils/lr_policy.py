"""Learning-rate schedules of ``OPTIMIZER.LR_POLICY`` with linear warm-up (reference: ``models/utils/lr_policy.py:10-83``).

``get_lr_at_epoch(cfg, cur_epoch)`` takes a fractional epoch (``runs/train.py:97`` passes
``cur_epoch + NUM_FOLDS * cur_iter / data_size``).  Policies are looked up by name, ``lr_func_<policy>``.
"""

import math


def lr_func_cosine(cfg, cur_epoch):
    """Half-cosine from BASE_LR at epoch 0 to 0 at MAX_EPOCH (``lr_policy.py:31-45``)."""
    o = cfg.OPTIMIZER
    return o.BASE_LR * 0.5 * (1.0 + math.cos(math.pi * cur_epoch / o.MAX_EPOCH))


def get_step_index(cfg, cur_epoch):
    """Index of the last entry of ``OPTIMIZER.STEPS`` that ``cur_epoch`` has reached (``lr_policy.py:59-71``)."""
    o = cfg.OPTIMIZER
    bounds = list(o.STEPS) + [o.MAX_EPOCH]
    ind = 0
    for ind, bound in enumerate(bounds):
        if cur_epoch < bound:
            break
    return ind - 1


def lr_func_steps_with_relative_lrs(cfg, cur_epoch):
    """Piecewise constant: ``LRS[i] * BASE_LR`` inside step ``i`` (``lr_policy.py:48-57``)."""
    return cfg.OPTIMIZER.LRS[get_step_index(cfg, cur_epoch)] * cfg.OPTIMIZER.BASE_LR


_POLICIES = {"cosine": lr_func_cosine, "steps_with_relative_lrs": lr_func_steps_with_relative_lrs}


def get_lr_func(lr_policy):
    if lr_policy not in _POLICIES:
        raise NotImplementedError("Unknown LR policy: {}".format(lr_policy))
    return _POLICIES[lr_policy]


def get_lr_at_epoch(cfg, cur_epoch):
    """The policy's value, replaced during the first ``WARMUP_EPOCHS`` by the straight line from ``WARMUP_START_LR`` to
    the policy's value at the end of the warm-up (``lr_policy.py:10-28``)."""
    o = cfg.OPTIMIZER
    fn = get_lr_func(o.LR_POLICY)
    warm = float(getattr(o, "WARMUP_EPOCHS", 0) or 0)
    if cur_epoch < warm:
        start = o.WARMUP_START_LR
        return start + cur_epoch * (fn(cfg, warm) - start) / warm
    return fn(cfg, cur_epoch)
