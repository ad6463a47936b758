"""world_size-2 Gloo tests of the multi-process host logic (clip sharding, score gather, flat gradient all-reduce)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from dist_b200 import distributed as du
    du.init_process_group(backend="gloo")
    assert du.get_world_size() == world and du.get_rank() == rank
    # clip sharding + the logits gather of runs/test.py:133
    total = 7
    lo, hi = du.shard_clips(total)
    scores = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 3) if hi > lo else torch.zeros(0, 3)
    pad = torch.zeros(4 - scores.shape[0], 3) - 1
    gathered, = du.all_gather([torch.cat([scores, pad])])
    valid = gathered[gathered[:, 0] >= 0]
    assert torch.equal(valid[:, 0], torch.arange(total, dtype=torch.float32))
    # mean all-reduce, list form and flat form agree
    g = [torch.full((3,), float(rank + 1)), torch.full((2, 2), float(10 * (rank + 1)))]
    a = du.all_reduce([t.clone() for t in g], average=True)
    b = du.all_reduce_flat([t.clone() for t in g], average=True)
    assert all(torch.allclose(x, y) for x, y in zip(a, b))
    assert torch.allclose(a[0], torch.full((3,), 1.5)) and torch.allclose(a[1], torch.full((2, 2), 15.0))
    du.synchronize()
    results[rank] = True
    dist.destroy_process_group()


def test_gloo_world_size_2():
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, port, results), nprocs=2, join=True)
    assert results.get(0) and results.get(1)


def test_shard_clips_covers_everything():
    from dist_b200 import distributed as du
    for n in (0, 1, 7, 32, 33):
        for ws in (1, 2, 4, 8):
            seen = []
            for r in range(ws):
                lo, hi = du.shard_clips(n, r, ws)
                seen += list(range(lo, hi))
            assert seen == list(range(n))
