"""-m gpu: the reference's on-disk formats end to end through the plugin surface -- a training run leaves
`<prefix>.index / .data-00000-of-00001 / checkpoint / _replay_memory.hdf5` behind (rl/model/model.py:60-67,
rl/deep_q_agent.py:410-432) and the next run resumes from them (rl/model/model.py:70-83, rl/deep_q_agent.py:216-221):
device state bit-identical, sum tree rebuilt from the float16 losses, training continues under the oracle's eye."""
import os
import random

import numpy as np
import pytest

from tests.gpu_util import need_gpu
from tests.stub_env import StubEnvironment, drive

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _gpu():
    need_gpu()


def _state(L):
    return (L.get_tower(0), L.get_tower(1), L.get_tower(0, slot=1), L.get_tower(0, slot=2))


def test_learner_checkpoint_round_trip_in_the_tensorflow_format(tmp_path):
    from deepqnetwork_b200.formats import tf_bundle
    from tests.test_gpu_nn import CONFIGS, make
    cfg = CONFIGS['c2_breakout_ddp']
    L, spec, shapes, on, tg = make(cfg, capacity=512)
    L.replay_fill_synthetic(512, seed=5)
    L.train_steps(3); L.set_game_count(17); L.sync()
    prefix = str(tmp_path / 'rl.game_environment.BreakoutEnvironment')
    L.save(prefix, fmt='reference', shapes=shapes)
    before, betas = _state(L), L.get_beta_powers()
    L.close()
    r = tf_bundle.BundleReader(prefix)
    assert len(r.names()) == 60 and int(r.tensor('train/step')) == 3 and int(r.tensor('game_count')) == 17
    b1p = np.float32(0.9)
    for _ in range(3):
        b1p = np.float32(b1p * np.float32(0.9))
    assert np.float32(r.tensor('train/beta1_power')) == b1p == np.float32(betas[0])   # TF: starts at beta1, multiplied per step
    assert np.array_equal(r.tensor('q_networks/online/dense_2/kernel').reshape(-1), before[0]['dense_2/kernel'])
    assert r.tensor('train/q_networks/online/conv2d/kernel/Adam_1').shape == (8, 8, 4, 32)
    L2, *_ = make(cfg, capacity=512, seed=9)
    assert L2.restore(str(tmp_path / 'missing'), fmt='reference', shapes=shapes) is False
    assert L2.restore(prefix, fmt='reference', shapes=shapes) is True
    assert L2.get_step() == 3 and L2.get_game_count() == 17 and L2.get_beta_powers() == betas
    for a, b in zip(before, _state(L2)):
        for k in a:
            assert np.array_equal(a[k], b[k]), k
    L2.close()


def test_agent_resumes_from_a_reference_format_save_dir(tmp_path):
    from deepqnetwork_b200.data.replay_memory_disk import ReplayMemoryDisk
    from deepqnetwork_b200.formats import tf_bundle
    from oracle.sumtree_c import SumTreeC
    from tests.test_gpu_agent import make_agent
    random.seed(3); np.random.seed(3)
    agent, holder, conf = make_agent(tmp_path, True)
    env = StubEnvironment(seed=5)
    with agent:
        L = holder['learner']
        drive(agent, env, 460)
        n = L.replay_info()[0]
        rows = L.replay_read(np.arange(n))
        state, betas, step, games = _state(L), L.get_beta_powers(), L.get_step(), L.get_game_count()
        assert step == conf['max_num_training_steps'] and n >= 256
    prefix = os.path.join(conf['save_dir'], conf['environment'])
    # what the run left behind parses on its own and holds the final state
    r = tf_bundle.BundleReader(prefix)
    assert int(r.tensor('train/step')) == step and int(r.tensor('game_count')) == games
    d = ReplayMemoryDisk(prefix + '_replay_memory.hdf5')
    assert (len(d), d.cur_idx, d.is_full, d.max_size) == (n, n, False, conf['replay_max_memory_length'])
    assert (d.input_height, d.input_width, d.input_channels, d.state_type) == (89, 80, 4, 'uint8')
    for k in ('states', 'next_states', 'actions', 'rewards', 'continues', 'losses'):
        assert np.array_equal(np.asarray(getattr(d, k)[:n]), rows[k]), k
    assert d.continues.dtype == np.bool_ and d.losses.dtype == np.float16
    d.close()
    # the next run: same save_dir, four more training steps
    random.seed(4); np.random.seed(4)
    agent2, holder, conf2 = make_agent(tmp_path, True, max_num_training_steps=step + 4)
    with agent2:
        L2 = holder['learner']
        assert L2 is not L and agent2.step == step == L2.get_step() and L2.get_game_count() == games and L2.get_beta_powers() == betas
        for a, b in zip(state, _state(L2)):
            for k in a:
                assert np.array_equal(a[k], b[k]), k
        assert L2.replay_info() == (n, n % conf['replay_max_memory_length'], False) and len(agent2.replay_sampler) == n
        back = L2.replay_read(np.arange(n))
        for k in rows:
            assert np.array_equal(back[k], rows[k]), k
        # ReplaySamplerPriority.add_losses (rl/data/replay_sampler_priority.py:57-59): add(losses[i], i) with the float16 values
        ref = SumTreeC(L2.capacity)
        ref.add(rows['losses'].astype(np.float32), np.arange(n).astype(np.uint32))
        tree, data, size, write = L2.tree_read()
        assert np.array_equal(tree, ref.tree) and np.array_equal(data, ref.data) and (size, write) == (n, n % L2.capacity)
        drive(agent2, StubEnvironment(seed=6), 200)   # every step checked against the oracle by the CheckingLearner
        assert agent2.step == step + 4 and agent2.is_training_done()
        n2 = L2.replay_info()[0]
    # deep_q_agent.py:416-429 as written: the exit appends the whole ring to the existing file at the file's own cur_idx
    d = ReplayMemoryDisk(prefix + '_replay_memory.hdf5')
    cap = conf['replay_max_memory_length']
    assert (len(d), d.cur_idx, d.is_full) == (min(n + n2, cap), (n + n2) % cap, n + n2 >= cap)
    d.close()
    assert int(tf_bundle.BundleReader(prefix).tensor('train/step')) == step + 4


def test_native_container_is_still_available(tmp_path):
    from tests.test_gpu_agent import make_agent
    random.seed(3); np.random.seed(3)
    agent, holder, conf = make_agent(tmp_path, False, persist_format='native', max_num_training_steps=3)
    with agent:
        drive(agent, StubEnvironment(seed=5), 300)
        n = holder['learner'].replay_info()[0]
    prefix = os.path.join(conf['save_dir'], conf['environment'])
    assert os.path.exists(prefix + '.dqnb200') and os.path.exists(prefix + '_replay_memory.dqnb200')
    assert not os.path.exists(prefix + '.index') and not os.path.exists(prefix + '_replay_memory.hdf5')
    agent2, holder, _ = make_agent(tmp_path, False, persist_format='native', max_num_training_steps=3)
    with agent2:
        assert agent2.step == 3 and holder['learner'].replay_info()[0] == n
