"""-m gpu parity: Q-network forward / TD target / loss / backward / optimizer through the C ABI vs the torch-CPU
fp64 restatement of rl/deep_q_model.py (oracle/nn_ref.py).  Tolerance: 1e-4 relative per tensor
(||a-b||inf / ||b||inf) for Q-values, losses, loss, gradients, updated weights and optimizer
slots in the fp32 mode; argmax actions bit-exact."""
import random

import numpy as np
import pytest
import torch

from oracle import nn_ref
from tests.gpu_util import need_gpu, rel_err
from tests.parity_util import (DENSE_KERNELS, check_step_against_oracle, oracle_state_from_gpu,
                               oracle_step_with_device_gates, snapshot)

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(autouse=True)
def _gpu():
    need_gpu()


CONFIGS = {
    'c1_vanilla': dict(h=89, w=80, a=4, dueling=False, double=False, per=False),
    'c2_breakout_ddp': dict(h=89, w=80, a=4, dueling=True, double=True, per=True),
    'c3_mspacman_ddp': dict(h=86, w=80, a=9, dueling=True, double=True, per=True),
    'double_only_momentum': dict(h=89, w=80, a=6, dueling=False, double=True, per=False, momentum=True),
    'dueling_only': dict(h=86, w=80, a=9, dueling=True, double=False, per=False),
    'c2_momentum': dict(h=89, w=80, a=4, dueling=True, double=True, per=True, momentum=True),
    'nature84': dict(h=84, w=84, a=4, dueling=False, double=False, per=False, variant='nature'),
    'nature84_ddp': dict(h=84, w=84, a=4, dueling=True, double=True, per=True, variant='nature'),
}


def make(cfg, batch=32, capacity=256, seed=0, math_mode='fp32', keep_grads=1, **learner_kw):
    from deepqnetwork_b200.learner import Learner
    from deepqnetwork_b200.deep_q_model import init_tower_params, tower_param_shapes
    variant = cfg.get('variant', 'reference')
    L = Learner(cfg['h'], cfg['w'], 4, num_actions=cfg['a'], use_dueling=cfg['dueling'], use_double=cfg['double'],
                use_per=cfg['per'], use_momentum=cfg.get('momentum', False), batch_size=batch, replay_capacity=capacity,
                conv_variant=variant, math_mode=math_mode, **learner_kw)
    spec = nn_ref.NetSpec(cfg['h'], cfg['w'], 4, cfg['a'], use_dueling=cfg['dueling'], use_double=cfg['double'],
                          use_per=cfg['per'], use_momentum=cfg.get('momentum', False), variant=variant)
    shapes = tower_param_shapes(cfg['h'], cfg['w'], 4, cfg['a'], 512, cfg['dueling'], variant)
    assert {k: tuple(v) for k, v in shapes.items()} == {k: tuple(v) for k, v in spec.param_shapes().items()}
    assert [t[0] for t in L.tensors()] == list(shapes)
    assert L.num_params == spec.num_params()
    on = init_tower_params(shapes, seed + 1); tg = init_tower_params(shapes, seed + 2)
    # non-zero biases so bias handling is exercised
    rng = np.random.default_rng(seed)
    for p in (on, tg):
        for k in p:
            if k.endswith('/bias'):
                p[k] = (rng.normal(0, 0.05, p[k].shape)).astype(np.float32)
    L.set_tower(on, 0); L.set_tower(tg, 1)
    # keep_grads=1: dqn_train_batch leaves every gradient readable in slot 3 (split dense wgrad + streaming optimizer);
    # keep_grads=0 is what ships: the fused dense wgrad+optimizer kernel, whose gradient the tests recover from the m slot
    L.set_option('keep_grads', keep_grads)
    return L, spec, shapes, on, tg


def slab_grad_names(L, shapes, mode, keep_grads):
    """Tensors whose gradient is in the gradient slab after a step."""
    if mode == 'fp32' or keep_grads:
        return None
    return [k for k in shapes if k not in DENSE_KERNELS]


def rand_batch(cfg, n, seed, rewards=(0, 1, 4, 7, 200)):
    rng = np.random.default_rng(seed)
    shp = (n, cfg['h'], cfg['w'], 4)
    # image-like: large constant regions plus noise, so ReLU gates are a mix of on/off
    st = (rng.integers(0, 256, shp) * (rng.random(shp) < 0.5)).astype(np.uint8)
    ns = (rng.integers(0, 256, shp) * (rng.random(shp) < 0.5)).astype(np.uint8)
    return dict(states=st, next_states=ns, actions=rng.integers(0, cfg['a'], n).astype(np.uint8),
                rewards=rng.choice(list(rewards), n).astype(np.uint8), continues=rng.random(n) < 0.8)


MODES = ['fp32', 'tc3x']


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('name', list(CONFIGS))
def test_predict_matches_oracle(name, mode):
    cfg = CONFIGS[name]
    L, spec, shapes, on, tg = make(cfg, math_mode=mode)
    for n in (1, 32, 77):
        x = rand_batch(cfg, n, n)['states']
        for tower, params in ((False, on), (True, tg)):
            q = L.predict(x, use_target=tower)
            ref = nn_ref.forward(nn_ref.to_torch(params, torch.float64), x, spec, torch.float64).numpy()
            assert q.shape == ref.shape and rel_err(q, ref) < TOL, (name, n, rel_err(q, ref))
            assert np.array_equal(np.argmax(q, 1), np.argmax(ref, 1))  # argmax actions bit-exact
    L.close()


PATHS = [('fp32', 1), ('tc3x', 1), ('tc3x', 0)]  # (math mode, keep_grads); ('tc3x', 0) is the shipped default


@pytest.mark.parametrize('mode,keep', PATHS, ids=['fp32', 'tc3x-keep_grads', 'tc3x-shipped'])
@pytest.mark.parametrize('name', list(CONFIGS))
def test_train_batch_matches_oracle(name, mode, keep):
    """DeepQModel.train (two runs + host target) for consecutive steps on fresh batches, each step starting the fp64
    oracle from the device's own state.  FFMA fp32 and tcgen05 TF32 x3 are held to the same 1e-4; the shipped path
    (fused dense wgrad+optimizer kernel, no gradient round trip) runs 5 steps and has its m / v slots and delta-w
    compared with the oracle like the others (tests/parity_util.py)."""
    import copy
    cfg = CONFIGS[name]
    L, spec, shapes, on, tg = make(cfg, math_mode=mode, keep_grads=keep)
    tag = 'train_batch[%s-%s-%d]' % (name, mode, keep)
    for it in range(3 if keep else 5):
        b = rand_batch(cfg, 32, 100 + it)
        w = np.random.default_rng(it).random(32).astype(np.float32) + 0.1 if cfg['per'] else None
        o, t, opt = oracle_state_from_gpu(L, shapes, spec)
        before = snapshot(o, opt)
        step, losses, loss = L.train_batch(b['states'], b['actions'], b['rewards'], b['continues'], b['next_states'], w)
        rstep, rlosses, rloss, aux = oracle_step_with_device_gates(L, spec, o, t, opt, b, w, 32, tag, it)
        assert step == rstep == it + 1
        A = cfg['a']
        assert rel_err(L.debug_read('q_online', (32, A)), aux['q']) < TOL
        assert rel_err(L.debug_read('max_q', (32,)), aux['max_q']) < TOL
        assert rel_err(L.debug_read('y', (32,)), aux['y']) < TOL
        assert rel_err(losses, rlosses) < TOL and abs(float(loss) - rloss) <= TOL * abs(rloss)
        check_step_against_oracle(L, spec, shapes, before, aux, o, opt, tag, it, slab_grad_names(L, shapes, mode, keep))
        for k in shapes:  # the target tower only changes through copy_network
            assert np.array_equal(L.get_param(k, tower=1), tg[k].reshape(-1))
    L.copy_network()
    for k in shapes:
        assert np.array_equal(L.get_param(k, tower=1), L.get_param(k, tower=0))
    L.close()


def test_double_dqn_masked_max_clamps_negative_targets():
    """D6: reduce_max(target_q * one_hot(argmax online_q)) clamps a negative selected Q to 0 (deep_q_model.py:247-252)."""
    cfg = CONFIGS['c2_breakout_ddp']
    L, spec, shapes, on, tg = make(cfg)
    tg = dict(tg); tg['dense_3/bias'] = np.array([-50.0], np.float32)  # value stream bias: all target Q << 0
    L.set_tower(tg, 1)
    b = rand_batch(cfg, 32, 5)
    w = np.ones(32, np.float32)
    L.train_batch(b['states'], b['actions'], b['rewards'], b['continues'], b['next_states'], w)
    mq = L.debug_read('max_q', (32,)); y = L.debug_read('y', (32,))
    assert np.all(mq == 0.0) and np.array_equal(y, b['rewards'].astype(np.float32))
    ref = nn_ref.max_q_values(nn_ref.to_torch(on, torch.float64), nn_ref.to_torch(tg, torch.float64), b['next_states'], spec, torch.float64)
    assert np.all(ref.numpy() == 0.0)
    L.close()


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('name', ['c1_vanilla', 'c2_breakout_ddp'])
def test_get_losses_matches_oracle(name, mode):
    """DeepQModel.get_losses on the 128-row PER insert batch (deep_q_agent.py:481-485)."""
    cfg = CONFIGS[name]
    L, spec, shapes, on, tg = make(cfg, math_mode=mode)
    b = rand_batch(cfg, 128, 9)
    got = L.get_losses(b['states'], b['actions'], b['rewards'], b['continues'], b['next_states'])
    ref = nn_ref.losses_only(nn_ref.to_torch(on, torch.float64), nn_ref.to_torch(tg, torch.float64), b, spec, torch.float64)
    assert rel_err(got, ref) < TOL
    assert L.get_step() == 0  # no optimizer step
    L.close()


def fused_loop_vs_oracle(name, mode, steps, draws, tag, N=2000, per_b_start=0.4, per_b_end=1.0, **learner_kw):
    """DeepQAgent._train_run_train + _train_save_model's target sync, `steps` times, fully on device (dqn_train_step fed
    the draws), each step against the oracle restatement of rl/deep_q_agent.py:326-360,500-534: sampled rows and gathered
    bytes bit-exact, IS weights (with the per_calculate_steps refresh cadence and the beta anneal) 1e-6, losses 1e-4,
    gradients / slots / delta-w per tests/parity_util.py, priority write-back and f16 loss column bit-exact given the
    device priorities, target tower equal to the online one exactly when step % copy_network_steps == 0.
    Returns the learner (caller closes)."""
    from oracle.sumtree_c import SumTreeC
    cfg = CONFIGS[name]
    B = 32
    L, spec, shapes, on, tg = make(cfg, capacity=N, math_mode=mode, keep_grads=0, per_b_start=per_b_start,
                                   per_b_end=per_b_end, **learner_kw)
    L.replay_fill_synthetic(N, seed=77)
    ref_tree = SumTreeC(N)
    anneal = learner_kw.get('use_per_annealing', False)
    calc = learner_kw.get('per_calculate_steps', 5000)
    copy_steps = learner_kw.get('copy_network_steps', 10000)
    max_w = None
    for it in range(steps):
        o, t, opt = oracle_state_from_gpu(L, shapes, spec)
        before = snapshot(o, opt)
        tg_before = {k: v.numpy().reshape(-1).astype(np.float32) for k, v in t.items()}
        assert opt.step == it
        if cfg['per']:
            tree, data, size, write = L.tree_read()
            ref_tree.load(tree, data, size, write)
            u = draws(it, B, N)
            step, losses, loss = L.train_step(uniforms=u)
            ti, di, p, mi = ref_tree.sample(u)
            # deep_q_agent.py:445-450: per_b follows the step counter; :505-510: normaliser every per_calculate_steps
            per_b = nn_ref.per_beta(it, per_b_start, per_b_end, learner_kw.get('per_anneal_steps', 2000000)) if anneal else per_b_start
            if max_w is None or it % calc == 0:
                leaves = tree[N - 1:N - 1 + size]
                max_w = nn_ref.max_weight(leaves.min(), tree[0], B, per_b)
            isw = nn_ref.is_weights(p, tree[0], B, per_b, max_w)
            assert rel_err(L.debug_read('is_weights', (B,)), isw) < 1e-6, (tag, it, 'is_weights')
            st = L.get_stats()
            assert abs(st['per_b'] - per_b) < 1e-6 and abs(st['last_max_weight'] - max_w) <= 1e-6 * max_w, (tag, it, st, per_b, max_w)
        else:
            mi = draws(it, B, N)
            step, losses, loss = L.train_step(rows=mi)
            isw = None
        got = L.batch_read(); rows = L.replay_read(mi)
        for k in ('states', 'next_states', 'actions', 'rewards', 'continues'):
            assert np.array_equal(got[k], rows[k]), k
        rstep, rlosses, rloss, aux = oracle_step_with_device_gates(L, spec, o, t, opt, rows, isw, B, tag, it)
        assert step == rstep == it + 1
        assert rel_err(losses, rlosses) < TOL and abs(float(loss) - rloss) <= TOL * abs(rloss)
        check_step_against_oracle(L, spec, shapes, before, aux, o, opt, tag, it, slab_grad_names(L, shapes, mode, 0))
        if cfg['per']:
            newp = L.debug_read('new_priorities', (B,))
            assert rel_err(newp, nn_ref.make_priority(losses)) < 1e-6  # powf vs np.power, float tolerance
            ref_tree.update(ti, newp)
            assert np.array_equal(L.tree_read()[0], ref_tree.tree)  # order-preserving scatter, bit-exact
            last = {}
            for r, v in zip(mi.tolist(), newp.astype(np.float16)):
                last[r] = v  # later duplicates win, as the sequential loop does
            assert all(L.replay_read([r], ('losses',))['losses'][0] == v for r, v in last.items())
        # deep_q_agent.py:358-360: target <- online when the post-increment step is a multiple of copy_network_steps
        for k in shapes:
            tgt = L.get_param(k, tower=1)
            if copy_steps > 0 and step % copy_steps == 0:
                assert np.array_equal(tgt, L.get_param(k, tower=0)), (tag, it, 'target sync missing', k)
            else:
                assert np.array_equal(tgt, tg_before[k]), (tag, it, 'target tower changed without a sync', k)
    return L, shapes


def host_random_draws(per):
    from oracle.sumtree_ref import uniform_rows
    rnd = random.Random(0)
    if per:
        return lambda it, B, N: np.array([rnd.random() for _ in range(B)])
    return lambda it, B, N: uniform_rows(N, B, rnd)


@pytest.mark.parametrize('mode', MODES)
@pytest.mark.parametrize('name', ['c2_breakout_ddp', 'c1_vanilla', 'c3_mspacman_ddp'])
def test_fused_train_steps_match_reference_loop(name, mode):
    """4 consecutive fused steps fed CPython random.random() / randint draws, as the reference's samplers make them."""
    L, _ = fused_loop_vs_oracle(name, mode, 4, host_random_draws(CONFIGS[name]['per']), 'fused[%s-%s]' % (name, mode))
    L.close()


CADENCE = dict(copy_network_steps=3, per_calculate_steps=2, use_per_annealing=True, per_anneal_steps=10)


@pytest.mark.parametrize('name', ['c2_breakout_ddp', 'c1_vanilla'])
def test_cadences_inside_train_steps_match_oracle_loop(name):
    """Target sync every copy_network_steps, PER normaliser refresh every per_calculate_steps and the beta anneal, all
    crossed several times inside ONE dqn_train_steps(9) call (CUDA-graph replay, look-ahead sampling, device Philox).
    Leg 1: the same nine steps issued one by one with the device's own draws restated on the host
    (oracle/philox_ref.py), every step checked against the oracle loop.  Leg 2: dqn_train_steps(9) must end bit-identical
    to leg 1 -- online and target towers, both optimizer slots, sum tree, step, beta powers, PER statistics."""
    from oracle import philox_ref
    cfg = CONFIGS[name]
    seed = 7
    if cfg['per']:
        draws = lambda it, B, N: philox_ref.per_uniforms(B, it, seed)
    else:
        draws = lambda it, B, N: philox_ref.uniform_rows(B, N, it, seed)
    steps = 9
    L1, shapes = fused_loop_vs_oracle(name, 'tc3x', steps, draws, 'cadence[%s]' % name, **CADENCE)
    L2, *_ = make(cfg, capacity=2000, math_mode='tc3x', keep_grads=0, **CADENCE)
    L2.replay_fill_synthetic(2000, seed=77)
    L2.train_steps(steps); L2.sync()
    assert L1.get_step() == L2.get_step() == steps
    assert L1.get_beta_powers() == L2.get_beta_powers()
    for tower, slot in ((0, 0), (1, 0), (0, 1), (0, 2)):
        for k in shapes:
            assert np.array_equal(L1.get_param(k, tower=tower, slot=slot), L2.get_param(k, tower=tower, slot=slot)), (tower, slot, k)
    if cfg['per']:
        assert np.array_equal(L1.tree_read()[0], L2.tree_read()[0])
        s1, s2 = L1.get_stats(), L2.get_stats()
        for key in ('min', 'max', 'avg', 'total', 'last_max_weight'):
            assert s1[key] == s2[key], (key, s1, s2)
        # per_b is "beta of the last sample drawn": leg 1 last sampled for step index 8; the pipelined replay has already
        # sampled the batch of step index 9 (look-ahead), which is the value the reference's agent holds after 9 steps
        # (_update_priority_sampling runs after the train call, deep_q_agent.py:341-343)
        assert abs(s1['per_b'] - nn_ref.per_beta(steps - 1, 0.4, 1.0, 10)) < 1e-6 and abs(s2['per_b'] - nn_ref.per_beta(steps, 0.4, 1.0, 10)) < 1e-6
    b1, b2 = L1.batch_read(), L2.batch_read()  # the last trained batch
    for k in b1:
        assert np.array_equal(b1[k], b2[k]), k
    L1.close(); L2.close()


def test_tree_stats_getter_leaves_the_per_normaliser_alone():
    """SumTree.get_min/get_max/get_average/total are read-only in the reference (sum_tree.py:139-165); the [train] report
    line reads `total` every train_report_interval steps (deep_q_agent.py:391).  The normaliser last_max_weight must keep
    the per_calculate_steps cadence however often the getters run (dqn_tree_stats used to overwrite it)."""
    cfg = CONFIGS['c2_breakout_ddp']
    L, spec, shapes, on, tg = make(cfg, capacity=2000, math_mode='tc3x', keep_grads=0, per_calculate_steps=1000)
    L.replay_fill_synthetic(2000, seed=77)
    L.train_step(fetch=False)
    st0 = L.get_stats()
    for _ in range(3):
        L.tree_stats()
        L.train_step(fetch=False)
    mn, mx, avg, total = L.tree_stats()
    st1 = L.get_stats()
    assert st1['last_max_weight'] == st0['last_max_weight'] and st1['min'] == st0['min'] and st1['total'] == st0['total']
    tree = L.tree_read()[0]
    assert mn == tree[1999:].min() and mx == tree[1999:].max() and total == tree[0] and total != st0['total']
    L.close()


def test_per_stats_getters_match_numpy():
    cfg = CONFIGS['c2_breakout_ddp']
    L, *_ = make(cfg, capacity=2000, math_mode='tc3x', keep_grads=0)
    L.replay_fill_synthetic(1500, seed=3)
    tree = L.tree_read()[0]
    leaves = tree[1999:1999 + 1500]
    mn, mx, avg, total = L.tree_stats()
    assert mn == leaves.min() and mx == leaves.max() and total == tree[0]
    assert abs(avg - np.average(leaves)) <= 2e-6 * abs(np.average(leaves))
    L.close()


@pytest.mark.parametrize('name', ['c2_breakout_ddp', 'c1_vanilla'])
def test_async_steps_equal_synchronous_steps(name):
    """dqn_train_step_async / dqn_train_step_result (two steps in flight, results read one step late) give bit-identical step
    numbers, losses and weights to the same host draws handed to the synchronous dqn_train_step."""
    cfg = CONFIGS[name]
    rng = np.random.RandomState(11)
    draws = [rng.random_sample(32) if cfg.get('per') else None for _ in range(7)]
    rows = [None if cfg.get('per') else rng.randint(0, 4096, 32) for _ in range(7)]
    out = []
    for mode in ('sync', 'async'):
        L, spec, shapes, on, tg = make(cfg, capacity=4096, math_mode='tc3x')
        L.replay_fill_synthetic(4096, seed=5)
        res = []
        if mode == 'sync':
            for u, r in zip(draws, rows):
                res.append(L.train_step(uniforms=u, rows=r))
        else:
            for i, (u, r) in enumerate(zip(draws, rows)):
                L.train_step_async(uniforms=u, rows=r)
                if i >= 1:
                    res.append(L.train_step_result())
            res.append(L.train_step_result())
            with pytest.raises(RuntimeError):
                L.train_step_result()  # nothing in flight any more
        L.sync()
        out.append((res, L.get_tower(0), L.tree_read()[0]))
        L.close()
    for a, b in zip(out[0][0], out[1][0]):
        assert a[0] == b[0] and np.array_equal(a[1], b[1]) and a[2] == b[2]
    for k in out[0][1]:
        assert np.array_equal(out[0][1][k], out[1][1][k]), k
    assert np.array_equal(out[0][2], out[1][2])


@pytest.mark.parametrize('math_mode', MODES)
def test_graph_replay_equals_eager_steps(math_mode):
    """dqn_train_steps (CUDA-graph replay, device Philox) == the same steps launched one by one."""
    cfg = CONFIGS['c2_breakout_ddp']
    res = []
    for mode in ('eager', 'graph'):
        L, spec, shapes, on, tg = make(cfg, capacity=4096, math_mode=math_mode)
        L.replay_fill_synthetic(4096, seed=5)
        if mode == 'eager':
            for _ in range(6):
                L.train_step(fetch=False)
        else:
            L.train_steps(6)
        L.sync()
        assert L.get_step() == 6
        res.append((L.get_tower(0), L.tree_read()[0], L.launch_count()))
        L.close()
    for k in res[0][0]:
        assert np.array_equal(res[0][0][k], res[1][0][k]), k
    assert np.array_equal(res[0][1], res[1][1])
    # the pipelined replay samples one batch ahead: one extra sample + gather launch pair over the whole run
    assert res[0][2] > 6 * 20 and res[1][2] - res[0][2] in (0, 2)


def test_save_restore_roundtrip(tmp_path):
    cfg = CONFIGS['c2_breakout_ddp']
    L, spec, shapes, on, tg = make(cfg, capacity=512)
    L.replay_fill_synthetic(512, seed=5)
    L.train_steps(3); L.set_game_count(17); L.sync()
    prefix = str(tmp_path / 'rl.game_environment.BreakoutEnvironment')
    L.save(prefix)
    before = (L.get_tower(0), L.get_tower(1), L.get_tower(0, slot=1), L.get_tower(0, slot=2), L.get_beta_powers())
    L.close()
    L2, *_ = make(cfg, capacity=512, seed=9)
    assert L2.restore(str(tmp_path / 'missing')) is False
    assert L2.restore(prefix) is True
    assert L2.get_step() == 3 and L2.get_game_count() == 17 and L2.get_beta_powers() == before[4]
    after = (L2.get_tower(0), L2.get_tower(1), L2.get_tower(0, slot=1), L2.get_tower(0, slot=2))
    for a, b in zip(before[:4], after):
        for k in a:
            assert np.array_equal(a[k], b[k])
    L2.close()


def test_tc1x_single_pass_tf32_is_close_but_not_fp32_grade():
    """DQN_MATH_TC1X (single-pass TF32) is the fast mode reported separately: Q-values within 5e-3, not 1e-4."""
    cfg = CONFIGS['c2_breakout_ddp']
    L, spec, shapes, on, tg = make(cfg, math_mode='tc1x')
    x = rand_batch(cfg, 32, 3)['states']
    q = L.predict(x)
    ref = nn_ref.forward(nn_ref.to_torch(on, torch.float64), x, spec, torch.float64).numpy()
    assert 1e-6 < rel_err(q, ref) < 5e-3
    L.close()


@pytest.mark.parametrize('math_mode', MODES)
def test_fused_optimizer_and_overlap_do_not_change_results(math_mode):
    """The dense-wgrad+Adam epilogue and the two-stream overlap are pure scheduling: weights, slots and the sum tree
    after 5 steps are bit-identical with both switched off."""
    cfg = CONFIGS['c2_breakout_ddp']
    res = []
    for fuse, overlap in ((1, 1), (0, 0), (1, 0), (0, 1)):
        L, spec, shapes, on, tg = make(cfg, capacity=4096, math_mode=math_mode)
        L.set_option('fuse_fc_adam', fuse); L.set_option('overlap', overlap)
        L.set_option('fc_rows_ctas', 0)  # compare the IEEE-division paths (the default fused row kernel uses rcp/sqrt.approx)
        L.replay_fill_synthetic(4096, seed=5)
        L.train_steps(5); L.sync()
        res.append((L.get_tower(0), L.get_tower(0, slot=1), L.get_tower(0, slot=2), L.tree_read()[0]))
        L.close()
    for r in res[1:]:
        for a, b in zip(res[0][:3], r[:3]):
            for k in a:
                assert np.array_equal(a[k], b[k]), k
        assert np.array_equal(res[0][3], r[3])


def test_ffma_dense_wgrad_adam_option_is_fp32_grade():
    """Option fc_ffma_adam: dense wgrad + optimizer as one fp32 FFMA pass.  Not bit-identical to the tensor-core
    wgrad (different summation), so it is held to the end-to-end weight bound: 1e-4 * |w|_inf + 2.5 * lr."""
    cfg = CONFIGS['c2_breakout_ddp']
    res = []
    for opt in (0, 1):
        L, spec, shapes, on, tg = make(cfg, capacity=4096, math_mode='tc3x')
        L.set_option('fc_ffma_adam', opt)
        L.replay_fill_synthetic(4096, seed=5)
        L.train_steps(3); L.sync()
        res.append((L.get_tower(0), L.tree_read()[0]))
        L.close()
    lr = 0.00025 if not hasattr(spec, 'learning_rate') else spec.learning_rate
    for k in res[0][0]:
        a, b = res[0][0][k].astype(np.float64), res[1][0][k].astype(np.float64)
        assert np.max(np.abs(a - b)) <= TOL * max(np.max(np.abs(a)), 1e-30) + 2.5 * lr, k


def _run_steps(cfg_name, opts, steps=5):
    L, spec, shapes, on, tg = make(CONFIGS[cfg_name], capacity=4096, math_mode='tc3x')
    for k, v in opts.items():
        L.set_option(k, v)
    L.replay_fill_synthetic(4096, seed=5)
    L.train_steps(steps)
    L.sync()
    assert L.get_step() == steps
    out = (L.get_tower(0), L.get_tower(0, slot=1), L.get_tower(0, slot=2), L.tree_read()[0])
    L.close()
    return out


@pytest.mark.parametrize('opts', [dict(prefetch=0), dict(priorities=0), dict(fc_split_streams=0), dict(defer_adam=1), dict(merge_towers=0), dict(tower_cluster=4), dict(tower_cluster=2), dict(pack_in_tail=0), dict(fuse_td=0),
                                  dict(fc_tma_adam=1), dict(fc_tma_adam=1, fc_tma_ctas=148), dict(fc_tma_adam=1, fc_tma_ctas=61)],
                         ids=lambda o: '+'.join('%s=%d' % kv for kv in o.items()))
def test_scheduling_options_are_bit_identical(opts):
    """Pipelined sampling, launch priorities, per-stream dense wgrad/optimizer launches, the deferred optimizer pass and
    the TMA-staged fused dense wgrad+optimizer kernels only reschedule or re-tile the same arithmetic: weights,
    optimizer slots and the sum tree after 5 fused steps are bit-identical to the default configuration."""
    base = dict(fc_rows_ctas=0)  # the split wgrad + streaming-Adam path these options act on (IEEE division / square root)
    ref = _run_steps('c2_breakout_ddp', base)
    got = _run_steps('c2_breakout_ddp', dict(base, **opts))
    for slot in range(3):
        for k in ref[slot]:
            assert np.array_equal(ref[slot][k], got[slot][k]), (slot, k)
    assert np.array_equal(ref[3], got[3])


@pytest.mark.parametrize('opts', [dict(tail_join=0), dict(fc_after_dgrad=0), dict(tail_join=0, fc_after_dgrad=0)],
                         ids=lambda o: '+'.join('%s=%d' % kv for kv in o.items()))
def test_end_of_step_join_is_bit_identical(opts):
    """Default path (fused dense wgrad+optimizer kernel): letting the optimizer tail run beside that kernel, with the end-of-step
    bookkeeping done by whichever finishes second, and forking the kernel's branch behind the dense dgrad GEMM instead of behind
    the sum+gate kernel, only reschedule: weights, slots and tree after 5 fused steps are bit-identical to the serial arrangement."""
    ref = _run_steps('c2_breakout_ddp', dict(opts))
    got = _run_steps('c2_breakout_ddp', {})
    for slot in range(3):
        for k in ref[slot]:
            assert np.array_equal(ref[slot][k], got[slot][k]), (slot, k)
    assert np.array_equal(ref[3], got[3])


def test_fused_row_kernel_momentum_matches_split_path():
    """Same comparison for the Nesterov-momentum optimizer (MomentumOptimizer, deep_q_model.py:385): no division, so the two
    paths differ only by the summation order of the gradient."""
    ref = _run_steps('c2_momentum', dict(fc_rows_ctas=0))
    got = _run_steps('c2_momentum', dict(fc_rows_ctas=148))
    for slot in (0, 1):
        for k in ref[slot]:
            scale = float(np.abs(ref[slot][k]).max()) + 1e-20
            assert float(np.abs(ref[slot][k] - got[slot][k]).max()) <= 1e-5 * scale, (slot, k)


@pytest.mark.parametrize('ctas', [112, 148, 61])
def test_fused_row_kernel_matches_split_path(ctas):
    """Default dense wgrad + optimizer kernel (fcwg_adam_rows_kernel: transposed product, approximate rcp / sqrt in the
    update term) against tcgen05 wgrad + streaming Adam: weights within 2e-5 of the tensor scale after 5 steps (the largest of 3.9 M
    Adam-amplified differences; measured 1.1e-5), first / second moments within 1e-4 (5x / 1x tighter than the parity bar); independent of the CTA count (bit-identical between counts)."""
    ref = _run_steps('c2_breakout_ddp', dict(fc_rows_ctas=0))
    got = _run_steps('c2_breakout_ddp', dict(fc_rows_ctas=ctas))
    got2 = _run_steps('c2_breakout_ddp', dict(fc_rows_ctas=100))
    for slot, tol in ((0, 2e-5), (1, 1e-4), (2, 1e-4)):
        for k in ref[slot]:
            scale = float(np.abs(ref[slot][k]).max()) + 1e-20
            assert float(np.abs(ref[slot][k] - got[slot][k]).max()) <= tol * scale, (slot, k)
            assert np.array_equal(got[slot][k], got2[slot][k]), (slot, k)


@pytest.mark.parametrize('opts', [dict(img_conv=0), dict(img_dgrad=0), dict(fc_tma_fwd=1), dict(tower_fused=0), dict(tower_fused=0, merge_towers=0)],
                         ids=lambda o: '+'.join('%s=%d' % kv for kv in o.items()))
@pytest.mark.parametrize('name', ['c2_breakout_ddp', 'c3_mspacman_ddp', 'nature84_ddp'])
def test_kernel_variants_agree(name, opts):
    """Image-resident conv kernels (forward / dgrad / wgrad), the fused conv tower (layer 0 on the frame bytes x bf16 pieces of
    w / 255, two accumulators per band) and the TMA-streamed dense forward against the per-layer / generic implicit-GEMM
    kernels: a different tiling and summation order of fp32-grade products -- weights after 3 fused
    steps agree to 1e-5 of the tensor's scale (+ 2 % of the Adam steps taken) for >= 95 % of the elements, and to that plus the
    Adam steps taken (2.5 lr each) for the elements whose gradient is within rounding of 0."""
    ref = _run_steps(name, {}, steps=3)
    got = _run_steps(name, opts, steps=3)
    # sum tree (priorities from the TD errors): the paths differ by ~1e-6 |q|, which is 1e-3 of a TD error of 1e-3
    # (a priority is (|delta| + eps) ** alpha: near delta = 0 its slope is unbounded, hence the absolute term)
    assert np.allclose(ref[3], got[3], rtol=2e-3, atol=2e-3)
    for k in ref[0]:
        scale = float(np.abs(ref[0][k]).max()) + 1e-12
        err = np.abs(ref[0][k] - got[0][k])
        assert float(np.mean(err > 1e-5 * scale + 2.5 * 0.00025 * 3)) == 0.0, k
        # Adam normalises every element by the root of its own second moment: an element whose gradient is 1e-3 of the tensor's
        # largest and differs by 1e-6 of it between the two kernels moves 1e-3 of an Adam step differently; a gradient within
        # rounding of 0 becomes +-lr.  >= 95 % of the elements within 1e-5 of the scale + 2 % of the steps taken
        # (the fused tower's layer 0 multiplies the frame BYTES with w / 255 -- exact products -- where the per-layer kernels
        # multiply round(x / 255) with w: activations agree to 1.3e-6 after a step, scratch/act_cmp.py, but steps 2 and 3 divide
        # by second moments of two steps, and a quarter of the conv kernels' smallest gradients then moves by > 2 % of a step)
        assert float(np.mean(err > 1e-5 * scale + 2e-2 * 0.00025 * 3)) < (0.25 if 'tower_fused' in opts else 5e-2), k


def test_train_batch_default_path_uses_fused_dense_kernel_and_agrees():
    """dqn_train_batch without keep_grads (the default, what the e2e seam path runs) applies the dense-kernel update with
    the fused row kernel; weights agree with the gradient-keeping path at the parity bar (1e-4 of the tensor scale; the two paths sum the 3xTF32
    products in a different order and Adam amplifies gradients within rounding of 0), optimizer slots to 1e-3."""
    cfg = CONFIGS['c2_breakout_ddp']
    out = []
    for keep in (1, 0):
        L, spec, shapes, on, tg = make(cfg, math_mode='tc3x')
        L.set_option('keep_grads', keep)
        for it in range(3):
            b = rand_batch(cfg, 32, 100 + it)
            w = np.random.default_rng(it).uniform(0.3, 1.0, 32).astype(np.float32)
            L.train_batch(b['states'], b['actions'], b['rewards'], b['continues'], b['next_states'], w)
        out.append((L.get_tower(0), L.get_tower(0, slot=1), L.get_tower(0, slot=2)))
        L.close()
    # weights: the LARGEST difference over 3.9 M elements is an Adam-amplified one (an element whose gradient is within 1e-6 of the
    # tensor's largest of 0 moves by a different fraction of lr): 1e-4 of the scale + 1 % of the three steps taken
    for slot, tol, extra in ((0, 1e-4, 1e-2 * 3 * 0.00025), (1, 1e-3, 0.0), (2, 1e-3, 0.0)):
        for k in out[0][slot]:
            scale = float(np.abs(out[0][slot][k]).max()) + 1e-20
            assert float(np.abs(out[0][slot][k] - out[1][slot][k]).max()) <= tol * scale + extra, (slot, k)
