"""Process-group helpers with the surface of the reference's ``utils/distributed.py`` (``:19-304``).

One process per GPU; NCCL over NVLink/NVSwitch on the GPU box, Gloo for the CPU tests.  The DiST forward path
shards clips across ranks with replicated weights, so inference needs exactly one collective - the gather of the
per-clip class scores (``runs/test.py:133``) - and fine-tuning one all-reduce of the ``dist_net`` gradients, which
:func:`all_reduce_flat` performs on a single flat buffer instead of the reference's DDP buckets over every
parameter of the model (``models/base/builder.py:69-74``).
"""

import os

import torch
import torch.distributed as dist


def is_initialized():
    return dist.is_available() and dist.is_initialized()


def get_world_size():
    return dist.get_world_size() if is_initialized() else 1


def get_rank():
    return dist.get_rank() if is_initialized() else 0


def get_local_rank():
    return int(os.environ.get("LOCAL_RANK", "0"))


def get_local_size():
    return int(os.environ.get("LOCAL_WORLD_SIZE", str(get_world_size())))


def is_master_proc(num_gpus=8):
    return get_rank() % num_gpus == 0 if is_initialized() else True


def init_process_group(local_rank=None, backend=None, init_method=None, world_size=None, rank=None):
    """Env-based (torchrun) or explicit rendezvous; ``nccl`` when CUDA is present, else ``gloo``."""
    if is_initialized():
        return
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    kw = {}
    if init_method is not None:
        kw.update(init_method=init_method, world_size=world_size, rank=rank)
    if backend == "nccl":
        local_rank = get_local_rank() if local_rank is None else local_rank
        torch.cuda.set_device(local_rank)
        kw["device_id"] = torch.device("cuda", local_rank)
    dist.init_process_group(backend=backend, **kw)


def init_distributed_training(cfg):
    """Kept for drivers that call it (``utils/distributed.py:262-277``); per-machine groups are not needed here."""
    return None


def synchronize():
    if is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def all_gather(tensors):
    """Concatenate every tensor of the list over ranks along dim 0 (``utils/distributed.py:19-38``)."""
    ws = get_world_size()
    if ws == 1:
        return list(tensors)
    out = []
    for t in tensors:
        t = t.contiguous()
        buf = torch.empty((ws * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, t)
        out.append(buf)
    return out


def all_reduce(tensors, average=True):
    """In-place sum (or mean) over ranks (``utils/distributed.py:41-57``)."""
    ws = get_world_size()
    if ws == 1:
        return tensors
    for t in tensors:
        dist.all_reduce(t)
        if average:
            t.mul_(1.0 / ws)
    return tensors


def all_reduce_flat(tensors, average=True):
    """One collective for many small tensors: flatten, all-reduce once, scatter back in place."""
    ws = get_world_size()
    tensors = [t for t in tensors if t is not None]
    if ws == 1 or not tensors:
        return tensors
    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat)
    if average:
        flat.mul_(1.0 / ws)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t))
        off += n
    return tensors


def shard_clips(num_clips, rank=None, world_size=None):
    """Contiguous clip range of this rank (clip i lives on rank i // ceil(num_clips / world_size))."""
    rank = get_rank() if rank is None else rank
    ws = get_world_size() if world_size is None else world_size
    per = (num_clips + ws - 1) // ws
    lo = min(rank * per, num_clips)
    return lo, min(lo + per, num_clips)
