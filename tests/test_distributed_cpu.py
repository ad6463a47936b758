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


def _train_worker(rank, world, port, results):
    """Data-parallel fine-tuning, host side: every rank holds the same flat parameter table, computes the gradient of
    ITS clips (here with the oracle, in place of the CUDA backward), the gradients are summed with one all-reduce of
    the flat buffer and the update applies 1 / world - the result must equal single-process training on all clips."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from dist_b200 import distributed as du
    from dist_b200.arch import tiny_arch
    from dist_b200.train import ParamTable, reduce_gradients
    from dist_b200.utils import synth
    from oracle import train_oracle
    du.init_process_group(backend="gloo")
    arch = tiny_arch()
    sd = synth.synth_state_dict(arch, seed=0, init="scaled")
    clips = synth.synth_clips(2, arch, seed=1234, kind="structured")
    text = synth.synth_text_features(arch.num_classes, arch.embed_dim, seed=77)
    target = synth.synth_soft_targets(2, arch.num_classes, seed=99)
    pt = ParamTable(sd, arch, "cpu", torch.float32, 1e-4)
    assert 0 < pt.n_decay < pt.n_used < pt.total
    lo, hi = du.shard_clips(2)
    _, _, grads = train_oracle.loss_and_grads(sd, clips[lo:hi], text, target[lo:hi], arch)
    # write this rank's gradients into the flat buffer (packed layouts)
    for k, g in grads.items():
        packed = pt.mat(k).g if g.dim() == 5 else None
        view = pt._view(pt.g, k)
        if g.dim() == 5:
            from dist_b200.train import _pack_conv
            view.copy_(_pack_conv(k, g.float())[0])
        else:
            view.copy_(g.float().reshape(view.shape))
    scale = reduce_gradients(pt)
    assert scale == 0.5
    got = {k: v * scale for k, v in pt.export(pt.g).items() if k not in pt.unused}
    _, _, want = train_oracle.loss_and_grads(sd, clips, text, target, arch)       # mean loss over both clips
    num = sum(float((got[k].double() - want[k]).pow(2).sum()) for k in want) ** 0.5
    den = sum(float(want[k].pow(2).sum()) for k in want) ** 0.5
    assert num / den < 1e-6, num / den
    for k in pt.unused:                                                           # never reduced, never updated
        assert float(pt._view(pt.g, k).abs().max()) == 0.0
    results[rank] = True
    dist.destroy_process_group()


def test_gloo_data_parallel_gradients():
    port = _free_port()
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_train_worker, args=(2, port, results), nprocs=2, join=True)
    assert results.get(0) and results.get(1)


def test_param_table_layout_and_roundtrip():
    from dist_b200.arch import tiny_arch
    from dist_b200.train import ParamTable
    from dist_b200.utils import synth
    from oracle import train_oracle
    arch = tiny_arch()
    sd = synth.synth_state_dict(arch, seed=0, init="scaled")
    pt = ParamTable(sd, arch, "cpu", torch.float32, 1e-4)
    back = pt.export()
    for k in back:
        assert torch.equal(back[k], sd[k].float()), k                  # pack -> unpack is the identity
    assert sorted(pt.unused) == train_oracle.unused_names(arch)
    for k in pt.names:
        decayed = pt.offset[k] < pt.n_decay
        if k in pt.unused:
            assert pt.offset[k] >= pt.n_used
        else:
            assert decayed == (train_oracle.weight_decay_of(k, tuple(sd[k].shape), 1e-4) > 0), k
        assert pt.offset[k] % 4 == 0
