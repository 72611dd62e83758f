"""N>1 host logic on CPU: world_size-2 gloo process groups (no GPU): replica aggregation, replay sharding and the
data-parallel exchange step (all-reduce mean of the gradient slab between the two phases of a step)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deepqnetwork_b200 import distributed as D


def _free_port():
    s = socket.socket(); s.bind(('127.0.0.1', 0)); p = s.getsockname()[1]; s.close(); return p


class FakeLearner:
    """CPU stand-in with the two-phase interface: phase 0 writes a rank-dependent 'gradient', phase 1 applies SGD."""

    def __init__(self, rank, n=1003):
        self.rank = rank
        self.w = torch.ones(n, dtype=torch.float32)
        self.g = torch.zeros(n, dtype=torch.float32)
        self.calls = []

    def train_step_phase(self, phase, scale):
        self.calls.append(phase)
        if phase == 0:
            # the local gradient arrives scaled by grad_scale = 1/world: the SUM over ranks is the global-batch mean
            self.g.copy_(torch.arange(self.g.numel(), dtype=torch.float32) * (self.rank + 1) * scale)
        else:
            self.w.sub_(0.5 * self.g)


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # replicas: sum of steps / max time
        v, ms, steps = D.aggregate_throughput(1000, 100.0 + 50.0 * rank)
        # data parallel: mean of the rank gradients, identical weights afterwards
        L = FakeLearner(rank)
        dp = D.DataParallelLearner(L, grads=L.g, stream_sync=lambda: None)
        dp.train_step()
        wl = [torch.zeros_like(L.w) for _ in range(world)]
        dist.all_gather(wl, L.w)
        out[rank] = dict(v=v, ms=ms, steps=steps, g=L.g.clone().numpy(), w_equal=bool(all(torch.equal(wl[0], x) for x in wl)),
                         w=L.w.clone().numpy(), calls=list(L.calls))
    finally:
        dist.destroy_process_group()


def test_world2_replica_aggregation_and_dp_exchange():
    world, port = 2, _free_port()
    mgr = mp.Manager(); out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    base = np.arange(1003, dtype=np.float32)
    for r in range(world):
        o = out[r]
        assert o['steps'] == 2000 and o['ms'] == 150.0 and abs(o['v'] - 2000 / 0.150) < 1e-6
        assert np.allclose(o['g'], base * 1.5)           # mean of 1x and 2x
        assert o['w_equal'] and np.allclose(o['w'], 1.0 - 0.5 * 1.5 * base)
        assert o['calls'] == [0, 1]                      # exchange sits between the two phases


def test_single_process_is_identity():
    v, ms, steps = D.aggregate_throughput(500, 250.0)
    assert (v, ms, steps) == (2000.0, 250.0, 500)


@pytest.mark.parametrize('total,world', [(1000000, 8), (10, 3), (7, 8)])
def test_replay_shards_are_disjoint_and_cover(total, world):
    seen = []
    for r in range(world):
        s, n = D.shard_capacity(total, r, world)
        seen += list(range(s, s + n)) if total < 100 else [(s, n)]
    if total < 100:
        assert seen == list(range(total))
    else:
        assert sum(n for _, n in seen) == total and all(seen[i][0] + seen[i][1] == seen[i + 1][0] for i in range(world - 1))


def test_gradient_buckets_cover_the_slab_once():
    """The exchange buckets of the data-parallel step: the two dense hidden kernels as one contiguous range each (sent
    first, while the conv backward still runs), everything else coalesced -- together every slab element exactly once."""
    tensors, off = [], 0
    for name, cnt in (('conv2d/kernel', 8192), ('conv2d/bias', 32), ('conv2d_1/kernel', 32768), ('conv2d_1/bias', 64),
                      ('conv2d_2/kernel', 65536), ('conv2d_2/bias', 64), ('dense/kernel', 7680 * 512), ('dense/bias', 512),
                      ('dense_1/kernel', 2048), ('dense_1/bias', 4), ('dense_2/kernel', 7680 * 512), ('dense_2/bias', 512),
                      ('dense_3/kernel', 512), ('dense_3/bias', 1)):
        tensors.append((name, cnt, off)); off += (cnt + 3) // 4 * 4
    big, small = D.gradient_buckets(tensors)
    assert [c for _, c in big] == [7680 * 512] * 2
    cover = np.zeros(off, np.int32)
    for o, c in big + small:
        cover[o:o + c] += 1
    assert (cover == 1).all() and len(small) == 3
    assert sum(c for _, c in small) * 4 < 500000  # the late bucket is < 0.5 MB
