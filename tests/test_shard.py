"""CPU tests of the N>1 host logic: contiguous frame/clip sharding and the end-of-step gather of the
fixed-size detection buffers, run with two gloo ranks (the GPU box uses NCCL for the same call)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tdrn_b200.utils import shard


def test_shard_range_partitions_every_unit_once():
    for n in (0, 1, 7, 16, 32, 33):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = shard.shard_range(n, r, world)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            sizes = [shard.shard_range(n, r, world)[1] - shard.shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def test_shard_clips_keeps_clips_whole():
    lengths = [16, 16, 5, 16, 9]
    frames = []
    for r in range(2):
        clips, ranges = shard.shard_clips(lengths, r, 2)
        for c, (lo, hi) in zip(clips, ranges):
            assert hi - lo == lengths[c]
            frames += list(range(lo, hi))
    assert frames == list(range(sum(lengths)))


def test_gather_without_process_group_is_identity():
    x = torch.arange(24.).view(2, 3, 2, 2)
    assert torch.equal(shard.gather_detections(x), x)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_units, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        full = torch.arange(n_units * 3 * 4 * 5, dtype=torch.float32).view(n_units, 3, 4, 5)   # [B, C, top_k, 5]
        local = shard.shard_batch(full, rank, world).clone()
        got = shard.gather_detections(local, n_units)
        ok = torch.equal(got, full)
        # timing contract of bench.py: max over ranks
        t = torch.tensor([float(rank + 1)])
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        q.put((rank, bool(ok), float(t.item())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n_units', [8, 7])
def test_two_rank_gather_reassembles_the_batch(n_units):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_units, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res == [(0, True, 2.0), (1, True, 2.0)]


def _async_worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        n_slots, steps = 3, 7
        ag = shard.AsyncGather(n_slots, (2, 3, 4, 5))
        bufs = [torch.zeros(2, 3, 4, 5) for _ in range(n_slots)]
        ok = True
        for i in range(steps):
            s = i % n_slots
            prev = ag.wait(s)                              # the gather that read bufs[s] three steps ago is complete
            if i >= n_slots:
                exp = torch.cat([torch.full((2, 3, 4, 5), float(100 * r + i - n_slots)) for r in range(world)], 0)
                ok = ok and torch.equal(prev, exp)
            bufs[s].fill_(float(100 * rank + i))           # "the step": rewrite the slot's buffer
            ag.issue(s, bufs[s])
        ag.wait_all()
        last = steps - 1
        exp = torch.cat([torch.full((2, 3, 4, 5), float(100 * r + last)) for r in range(world)], 0)
        ok = ok and torch.equal(ag.out[last % n_slots], exp)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_two_rank_async_gather_slots():
    """AsyncGather (what bench.py uses at N > 1): per-slot asynchronous all_gather, waited for only when the slot is reused."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_async_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]
