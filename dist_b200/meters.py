"""``TestMeter`` with the surface of the reference's (``utils/meters.py:24-175``), kept on the device.

The reference accumulates the per-clip softmax scores of a multi-view test in a Python loop over CPU tensors
(``update_stats``, ``utils/meters.py:83-115``) after a ``.cpu()`` copy of every batch (``runs/test.py:136-145``) - a host
synchronisation per iteration.  Here the video-level score table, the labels and the clip counts are device tensors and
a batch is folded in by one kernel (``distb200_view_ensemble``); nothing synchronises until ``finalize_metrics``.
"""

import datetime
import time

import torch

from . import ops


class TestMeter:
    __test__ = False          # not a pytest class

    def __init__(self, cfg, num_videos, num_clips, num_cls, overall_iters, ensemble_method="sum", device="cuda"):
        if ensemble_method not in ("sum", "max"):
            raise NotImplementedError("Ensemble Method {} is not supported".format(ensemble_method))      # meters.py:108-112
        ops.lib()
        self.cfg = cfg
        self.num_clips = int(num_clips)
        self.overall_iters = overall_iters
        self.ensemble_method = ensemble_method
        self.device = torch.device(device)
        self.video_preds = torch.zeros((num_videos, num_cls), device=self.device)
        self.video_labels = torch.zeros((num_videos,), dtype=torch.long, device=self.device)
        self.clip_count = torch.zeros((num_videos,), dtype=torch.long, device=self.device)
        self.model_ema_enabled = False
        self._t0 = time.perf_counter()
        self._dt = 0.0
        self.stats = {}
        self.reset()

    def reset(self):
        self.clip_count.zero_()
        self.video_preds.zero_()
        self.video_labels.zero_()

    def update_stats(self, preds, labels, clip_ids):
        """preds [N, C] float, labels [N], clip_ids [N] (device tensors; anything else is moved, which synchronises)."""
        preds = preds.detach().to(self.device, torch.float32).contiguous()
        labels = labels.detach().to(self.device, torch.long).contiguous()
        clip_ids = clip_ids.detach().to(self.device, torch.long).contiguous()
        ops.view_ensemble(preds, labels, clip_ids, self.num_clips, 0 if self.ensemble_method == "sum" else 1, self.video_preds,
                          self.video_labels, self.clip_count, torch.cuda.current_stream(self.device).cuda_stream)

    def iter_tic(self):
        self._t0 = time.perf_counter()

    def iter_toc(self):
        self._dt = time.perf_counter() - self._t0

    def log_iter_stats(self, cur_iter):
        period = getattr(self.cfg, "LOG_PERIOD", 10) if self.cfg is not None else 10
        if (cur_iter + 1) % period != 0:
            return None
        eta = str(datetime.timedelta(seconds=int(self._dt * (self.overall_iters - cur_iter))))
        stats = {"split": "test_iter" if not self.model_ema_enabled else "ema_test_iter", "cur_iter": "{}".format(cur_iter + 1),
                 "eta": eta, "time_diff": self._dt}
        print(stats)
        return stats

    def finalize_metrics(self, ks=(1, 5)):
        """Top-k accuracies of the ensembled scores (``utils/meters.py:135-170``); returns the logged dict."""
        bad = (self.clip_count != self.num_clips).nonzero().flatten()
        if bad.numel():
            print("clip count {} ~= num clips {}".format(
                ", ".join("{}: {}".format(int(i), int(self.clip_count[i])) for i in bad[:32]), self.num_clips))
        ks_dev = torch.tensor(list(ks), dtype=torch.int32, device=self.device)
        correct = torch.zeros(len(ks), dtype=torch.long, device=self.device)
        ops.topk_correct(self.video_preds, self.video_labels, ks_dev, correct, torch.cuda.current_stream(self.device).cuda_stream)
        n = self.video_preds.size(0)
        stats = {"split": "test_final" if not self.model_ema_enabled else "ema_test_final"}
        for k, c in zip(ks, correct.tolist()):
            stats["top{}_acc".format(k)] = "{:.{prec}f}".format(c / n * 100.0, prec=2)
        self.stats = stats
        print(stats)
        return stats

    def set_model_ema_enabled(self, model_ema_enabled):
        self.model_ema_enabled = model_ema_enabled
