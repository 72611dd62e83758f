"""-m gpu: the large-batch data-parallel step (BASELINE.json configs[4]; a capability the north star
adds, the reference is single-device).  World size > 1 is emulated on ONE GPU, as the profiling guide prescribes for
boxes with fewer GPUs than ranks: every "rank" is its own learner handle, the all-reduce is a sum of the handles'
gradient slabs.  Checked: (a) the big-batch kernels (rows 128 .. 1024 go through the generic tcgen05 implicit-GEMM tiles,
not the batch-32 image-resident ones) against the fp64 oracle; (b) n ranks x B/n == one rank x B: the summed, 1/n-scaled
shard gradients equal the full-batch gradient and the weights after the optimizer phase agree; (c) every rank ends with
identical weights.  The NCCL path itself (bucketed, overlapped with the backward pass) is exercised by
`bench.py --config c5` under torchrun and by tools/dp_check.py on a multi-GPU box."""
import numpy as np
import pytest
import torch

from oracle import nn_ref
from tests.gpu_util import need_gpu, rel_err
from tests.parity_util import check_step_against_oracle, oracle_state_from_gpu, oracle_step_with_device_gates, snapshot
from tests.test_gpu_nn import CONFIGS, make, rand_batch

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(autouse=True)
def _gpu():
    need_gpu()


def load_batch(L, b, isw=None):
    """Put rows into the learner's replay and make them the device batch (row order preserved)."""
    n = len(b['actions'])
    L.replay_clear()
    L.replay_append(b['states'], b['actions'], b['rewards'], b['next_states'], b['continues'],
                    np.full(n, 0.5, np.float32) if L.use_per else None)
    L.uniform_sample(np.arange(n))


@pytest.mark.parametrize('B', [128, 256])
def test_big_batch_step_matches_oracle(B):
    """One split step (phase 4 + phase 1) at a data-parallel per-rank batch size against the fp64 oracle."""
    cfg = CONFIGS['c1_vanilla']  # uniform sampler: the device batch is loaded explicitly
    L, spec, shapes, on, tg = make(cfg, batch=B, capacity=B, math_mode='tc3x', keep_grads=1, max_rows=B)
    b = rand_batch(cfg, B, 3, rewards=(0, 1, 4, 7))
    load_batch(L, b)
    o, t, opt = oracle_state_from_gpu(L, shapes, spec)
    before = snapshot(o, opt)
    L.train_step_phase(4, 1.0)
    L.train_step_phase(1, 1.0)
    L.sync()
    assert L.get_step() == 1
    r = oracle_step_with_device_gates(L, spec, o, t, opt, b, None, B, 'dp-big-batch[%d]' % B, 0)
    losses = L.debug_read('losses', (B,))
    assert rel_err(losses, r[1]) < TOL
    assert rel_err(L.debug_read('q_online', (B, cfg['a'])), r[3]['q']) < TOL
    check_step_against_oracle(L, spec, shapes, before, r[3], o, opt, 'dp-big-batch[%d]' % B, 0)
    L.close()


@pytest.mark.parametrize('world,B', [(2, 256), (4, 256), (2, 1024)])
def test_shards_reproduce_the_full_batch_step(world, B):
    """n ranks x B/n rows with grad_scale = 1/n and a SUM exchange == one rank x B rows."""
    # double + dueling towers with the uniform sampler: the device batch can be loaded row for row (under PER the rows come
    # from each shard's own tree; the IS-weighted loss itself is covered by the batch-32 parity tests)
    cfg = dict(h=89, w=80, a=4, dueling=True, double=True, per=False)
    b = rand_batch(cfg, B, 11, rewards=(0, 1, 4, 7))

    def run(rows, scale, n_total):
        L, spec, shapes, on, tg = make(cfg, batch=len(rows), capacity=max(len(rows), 64), math_mode='tc3x', keep_grads=1, max_rows=len(rows))
        load_batch(L, {k: v[rows] for k, v in b.items()})
        L.train_step_phase(4, scale)
        L.sync()
        return L, shapes

    full, shapes = run(np.arange(B), 1.0, B)
    g_full = {k: full.get_param(k, slot=3).astype(np.float64) for k in shapes}
    shards = [run(np.arange(r * B // world, (r + 1) * B // world), 1.0 / world, B) for r in range(world)]
    g_sum = {k: sum(s.get_param(k, slot=3).astype(np.float64) for s, _ in shards) for k in shapes}
    for k in shapes:
        scale = max(float(np.max(np.abs(g_full[k]))), 1e-30)
        assert float(np.max(np.abs(g_sum[k] - g_full[k]))) <= TOL * scale, (k, float(np.max(np.abs(g_sum[k] - g_full[k]))) / scale)
    # the exchange: every rank receives the sum, then applies the same optimizer step
    for s, _ in shards:
        for k in shapes:
            s.set_param(k, g_sum[k].astype(np.float32), slot=3)
        s.train_step_phase(1, 1.0 / world); s.sync()
    full.train_step_phase(1, 1.0); full.sync()
    lr = 0.00025
    for k in shapes:
        w0 = shards[0][0].get_param(k)
        for s, _ in shards[1:]:
            assert np.array_equal(s.get_param(k), w0), ('ranks diverged', k)  # identical inputs -> identical weights
        wf = full.get_param(k)
        # same gradient to 1e-4 (measured 1e-6 .. 3e-5: different summation order of the tensor-core partial sums): Adam moves an element by <= lr; elements whose gradient is within that noise of 0 may differ by 2 lr
        d = np.abs(w0.astype(np.float64) - wf)
        assert float(d.max()) <= 2.1 * lr and float(np.mean(d > 0.05 * lr)) < 2e-2, (k, float(d.max()) / lr, float(np.mean(d > 0.05 * lr)))
    for s, _ in shards:
        assert s.get_step() == 1
        s.close()
    full.close()
