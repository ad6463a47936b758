"""Multi-view test loop with the surface of the reference's ``runs/test.py:25-178`` (``perform_test``).

Each video is sampled ``TEST.NUM_ENSEMBLE_VIEWS x TEST.NUM_SPATIAL_CROPS`` times; the softmax scores of the views are
summed per video and compared with the label.  Differences to the reference loop: the scores never leave the device (no
``.cpu()`` per iteration, no per-clip Python loop - see ``dist_b200.meters.TestMeter``), and clips may be decoded uint8
frames ``[B, T, H, W, 3]`` (normalised inside the patch-row kernel).
"""

import torch

from .. import distributed as du


def num_views(cfg):
    """Clips per video (``runs/run.py:38-66`` / ``runs/test.py:232-236``)."""
    t = cfg.TEST
    return int(getattr(t, "NUM_ENSEMBLE_VIEWS", 1)) * int(getattr(t, "NUM_SPATIAL_CROPS", 1))


@torch.no_grad()
def perform_test(test_loader, model, test_meter, cfg, text_tokenizers=None):
    """``for (inputs, labels, video_idx, meta) in loader: preds, _ = model(inputs); meter.update_stats(...)``."""
    model.eval()
    test_meter.iter_tic()
    dev = test_meter.device
    for cur_iter, (inputs, labels, video_idx, meta) in enumerate(test_loader):
        for k, v in inputs.items():
            if torch.is_tensor(v) and not v.is_cuda:
                inputs[k] = (v if v.is_pinned() else v.pin_memory()).to(dev, non_blocking=True)           # runs/test.py:44-53
        if text_tokenizers is not None:
            inputs["texts"] = text_tokenizers                                                             # runs/test.py:71-72
        lab = labels["supervised"] if isinstance(labels, dict) else labels
        lab = lab.to(dev, non_blocking=True)
        video_idx = video_idx.to(dev, non_blocking=True)
        preds, _ = model(inputs)                                                                          # runs/test.py:92
        if du.get_world_size() > 1:
            preds, lab, video_idx = du.all_gather([preds.contiguous(), lab, video_idx])                  # runs/test.py:131-135
        test_meter.iter_toc()
        test_meter.update_stats(preds, lab, video_idx)
        test_meter.log_iter_stats(cur_iter)
        test_meter.iter_tic()
    stats = test_meter.finalize_metrics()
    return stats
