"""``build_model(cfg, gpu_id=None) -> (model, model_ema)`` (reference: ``models/base/builder.py:19-74``)."""

import torch

from .models import MODEL_REGISTRY, BaseVideoModel


def build_model(cfg, gpu_id=None):
    model_cls = MODEL_REGISTRY.get(cfg.MODEL.NAME)
    model = BaseVideoModel(cfg) if model_cls is None else model_cls(cfg)   # builder.py:30-36

    if torch.cuda.is_available():
        assert cfg.NUM_GPUS <= torch.cuda.device_count(), "Cannot use more GPU devices than available"
    else:
        assert cfg.NUM_GPUS == 0, "Cuda is not available. Please set `NUM_GPUS: 0 for running on CPUs."

    cur_device = None
    if cfg.NUM_GPUS:
        cur_device = torch.cuda.current_device() if gpu_id is None else gpu_id
        model = model.cuda(device=cur_device)

    model_ema = None
    if cfg.MODEL.EMA.ENABLE:
        raise NotImplementedError("MODEL.EMA is outside the DiST forward path (ModelEmaV2 is undefined in the reference as well, builder.py:13,57)")

    if cfg.NUM_GPUS * cfg.NUM_SHARDS > 1 and torch.distributed.is_available() and torch.distributed.is_initialized():
        # The reference wraps the whole model in DDP with find_unused_parameters=True (builder.py:69-74), which
        # all-reduces the frozen CLIP parameters as zero buckets.  Here only dist_net parameters are trainable
        # (the rest are frozen by flag), so DDP reduces exactly the 19-40 M trainable values.
        for name, p in model.named_parameters():
            if ".dist_net." not in name:
                p.requires_grad_(False)
        if any(p.requires_grad for p in model.parameters()):
            model = torch.nn.parallel.DistributedDataParallel(module=model, device_ids=[cur_device], output_device=cur_device)
    return model, model_ema
