"""-m gpu: the plugin surface itself.  DeepQAgent.train(game_state) / get_action are driven the way GameRunner drives
them (tests/stub_env.py) for a few hundred synthetic transitions; every device call the mirror classes make is
intercepted and checked against the oracle on the spot, and the sequence of calls is compared with an independent
restatement of the reference's control flow (rl/deep_q_agent.py:271-306 warm-up gate and every-Nth-step rule, :467-497
the 128-row PER insert, :326-365 sample/train/write-back/target sync, :368-407 report line)."""
import logging
import os
import random
import re

import numpy as np
import pytest
import torch

from oracle import nn_ref
from tests.gpu_util import need_gpu, rel_err
from tests.parity_util import (DENSE_KERNELS, check_step_against_oracle, oracle_state_from_gpu,
                               oracle_step_with_device_gates, snapshot)
from tests.stub_env import StubEnvironment, drive

pytestmark = pytest.mark.gpu
TOL = 1e-4
G = os.path.join(os.path.dirname(__file__), 'golden')


@pytest.fixture(autouse=True)
def _gpu():
    need_gpu()


class CheckingLearner:
    """Delegates to the real Learner; checks each compute call against the oracle and logs the call sequence."""

    def __init__(self, real, spec, shapes, per, calc_steps, anneal_steps):
        from oracle.sumtree_c import SumTreeC
        self._r, self.spec, self.shapes, self.per = real, spec, shapes, per
        self.events = []
        self.ring = []          # host copy of every appended row, in ring order (capacity never wraps in these tests)
        self.ref_tree = SumTreeC(real.capacity) if per else None
        self.max_w = None
        self.calc_steps, self.anneal_steps = calc_steps, anneal_steps
        self.predict_checked = 0

    def __getattr__(self, name):
        return getattr(self._r, name)

    # -- replay insert: ReplayMemory.append (+ SumTree.add under PER, replay_sampler_priority.py:22-24) ---------------
    def replay_append(self, states, actions, rewards, next_states, continues, losses=None):
        n = len(np.atleast_1d(actions))
        self.events.append(('append', n))
        size0 = self._r.replay_info()[0]
        self._r.replay_append(states, actions, rewards, next_states, continues, losses)
        st = np.asarray(states).reshape((n,) + self._r.state_shape); ns = np.asarray(next_states).reshape((n,) + self._r.state_shape)
        for i in range(n):
            self.ring.append(dict(state=st[i].copy(), next_state=ns[i].copy(), action=np.atleast_1d(actions)[i],
                                  reward=np.atleast_1d(rewards)[i], cont=bool(np.atleast_1d(continues)[i]),
                                  loss=None if losses is None else np.float32(np.atleast_1d(losses)[i])))
        assert self._r.replay_info()[0] == size0 + n
        if self.per:
            self.ref_tree.add(np.asarray(losses, np.float32), np.arange(size0, size0 + n).astype(np.uint32))
            tree, data, size, write = self._r.tree_read()
            assert np.array_equal(tree, self.ref_tree.tree) and np.array_equal(data, self.ref_tree.data)
            assert (size, write) == (self.ref_tree.size, self.ref_tree.write)

    # -- DeepQModel.get_losses on the insert batch (deep_q_agent.py:481-487) ------------------------------------------
    def get_losses(self, states, actions, rewards, continues, next_states):
        n = len(actions)
        self.events.append(('get_losses', n))
        got = self._r.get_losses(states, actions, rewards, continues, next_states)
        o, t, _ = oracle_state_from_gpu(self._r, self.shapes, self.spec)
        b = dict(states=np.asarray(states), actions=np.asarray(actions), rewards=np.asarray(rewards),
                 continues=np.asarray(continues), next_states=np.asarray(next_states))
        ref = nn_ref.losses_only(o, t, b, self.spec, torch.float64)
        assert rel_err(got, ref) < TOL
        return got

    # -- DeepQModel.predict for get_action (deep_q_agent.py:559-592): every 16th call is checked ---------------------
    def predict(self, states, use_target=False):
        q = self._r.predict(states, use_target=use_target)
        self.events.append(('predict', len(q)))
        if sum(1 for e in self.events if e[0] == 'predict') % 16 == 1:
            params = {k: self._r.get_param(k, tower=1 if use_target else 0).reshape(self.shapes[k]) for k in self.shapes}
            ref = nn_ref.forward(nn_ref.to_torch(params, torch.float64), np.asarray(states), self.spec, torch.float64).numpy()
            assert rel_err(q, ref) < TOL and np.array_equal(np.argmax(q, 1), np.argmax(ref, 1))
            self.predict_checked += 1
        return q

    # -- DeepQAgent._train_run_train as one device call ------------------------------------------------------------------
    def train_step(self, uniforms=None, rows=None, fetch=True):
        L, spec, shapes = self._r, self.spec, self.shapes
        B = L.batch_size
        it = L.get_step()
        self.events.append(('train_step', it))
        o, t, opt = oracle_state_from_gpu(L, shapes, spec)
        before = snapshot(o, opt)
        if self.per:
            assert uniforms is not None and rows is None
            tree, data, size, write = L.tree_read()
            assert np.array_equal(tree, self.ref_tree.tree)
        out = L.train_step(uniforms=uniforms, rows=rows, fetch=True)
        step, losses, loss = out
        if self.per:
            ti, di, p, mi = self.ref_tree.sample(uniforms)
            per_b = nn_ref.per_beta(it, L.cfg.per_b_start, L.cfg.per_b_end, self.anneal_steps) if L.cfg.use_per_annealing else L.cfg.per_b_start
            if self.max_w is None or it % self.calc_steps == 0:  # deep_q_agent.py:505-510
                leaves = tree[L.capacity - 1:L.capacity - 1 + size]
                self.max_w = nn_ref.max_weight(leaves.min(), tree[0], B, per_b)
            isw = nn_ref.is_weights(p, tree[0], B, per_b, self.max_w)
            assert rel_err(L.debug_read('is_weights', (B,)), isw) < 1e-6, ('is_weights', it)
        else:
            assert rows is not None and uniforms is None
            mi, isw = np.asarray(rows), None
        batch = {k: np.stack([self.ring[r][src] for r in mi]) for k, src in
                 (('states', 'state'), ('next_states', 'next_state'), ('actions', 'action'), ('rewards', 'reward'), ('continues', 'cont'))}
        batch['actions'] = batch['actions'].astype(np.uint8); batch['rewards'] = batch['rewards'].astype(np.uint8)
        got = L.batch_read()
        for k in batch:
            assert np.array_equal(got[k], batch[k]), ('gathered batch', k)
        rstep, rlosses, rloss, aux = oracle_step_with_device_gates(L, spec, o, t, opt, batch, isw, B, 'agent', it)
        assert step == rstep == it + 1
        assert rel_err(losses, rlosses) < TOL and abs(float(loss) - rloss) <= TOL * abs(rloss)
        check_step_against_oracle(L, spec, shapes, before, aux, o, opt, 'agent', it, [k for k in shapes if k not in DENSE_KERNELS])
        if self.per:
            newp = L.debug_read('new_priorities', (B,))
            assert rel_err(newp, nn_ref.make_priority(losses)) < 1e-6
            self.ref_tree.update(ti, newp)
            assert np.array_equal(L.tree_read()[0], self.ref_tree.tree)
        return out if fetch else None

    def copy_network(self):
        self.events.append(('copy_network', self._r.get_step()))
        self._r.copy_network()
        for k in self.shapes:
            assert np.array_equal(self._r.get_param(k, tower=1), self._r.get_param(k, tower=0))

    def save(self, prefix, **kw):
        self.events.append(('save', self._r.get_step()))
        self._r.save(prefix, **kw)


def expected_events(calls, conf, per):
    """The reference's control flow restated from rl/deep_q_agent.py:271-306,316-323,352-365,453-497 on the recorded
    game states: which device calls must happen, in which order.  predict events are added by the caller."""
    ev = []
    appended = 0       # len(replay_sampler)
    pending = 0        # len(per_memory_batch)
    game_steps = 0
    step = 0
    games = 0
    for c in calls:
        if step >= conf['max_num_training_steps']:  # is_training_done -> train() returns False, nothing else happens
            break
        if c['game_done']:
            games += 1
        if c['old_state'] is not None:
            if per:
                pending += 1
                if pending >= 128:
                    ev += [('get_losses', 128), ('append', 128)]
                    appended += 128; pending = 0
            else:
                ev.append(('append', 1)); appended += 1
        if appended >= conf['num_game_frames_before_training']:
            game_steps += 1
            if game_steps % conf['num_game_steps_per_train'] == 0:
                ev.append(('train_step', step)); step += 1
                if step % conf['copy_network_steps'] == 0:
                    ev.append(('copy_network', step))
                if step % conf['save_model_steps'] == 0:
                    ev.append(('save', step))
        yield_predict = True  # GameRunner asks for an action after every in-episode train() call (cont == 1)
        if c['cont'] == 1:
            ev.append(('predict', 1))
    return ev, step, games, appended


HOLDER = {}


def _checked_model_class():
    from deepqnetwork_b200.deep_q_model import BreakoutModel, tower_param_shapes

    class CheckedBreakoutModel(BreakoutModel):
        def open(self):
            super().open()
            c = self.conf
            spec = nn_ref.NetSpec(89, 80, 4, self.num_outputs, use_dueling=bool(c['use_dueling']), use_double=bool(c['use_double']),
                                  use_per=bool(c['use_per']))
            shapes = tower_param_shapes(89, 80, 4, self.num_outputs, 512, bool(c['use_dueling']))
            self.learner = CheckingLearner(self.learner, spec, shapes, bool(c['use_per']), c['per_calculate_steps'], c['per_anneal_steps'])
            HOLDER['learner'] = self.learner
            return self

    return CheckedBreakoutModel


def __getattr__(name):  # resolved by import_class('tests.test_gpu_agent.CheckedBreakoutModel'), like any --model path
    if name == 'CheckedBreakoutModel':
        cls = _checked_model_class()
        globals()[name] = cls
        return cls
    raise AttributeError(name)


def make_agent(tmp_path, per, **over):
    from deepqnetwork_b200.deep_q_agent import DeepQAgent
    conf = dict(is_training=True, save_dir=str(tmp_path / 'save'), environment='tests.stub_env.StubEnvironment',
                model='tests.test_gpu_agent.CheckedBreakoutModel', action_space=4, use_double=per, use_dueling=per, use_per=per,
                use_per_annealing=per, per_anneal_steps=40, per_calculate_steps=7, copy_network_steps=4,
                save_model_steps=9, replay_max_memory_length=640, num_game_frames_before_training=256,
                num_game_steps_per_train=4, batch_size=32, max_num_training_steps=14, train_report_interval=5,
                use_memory=True)
    conf.update(over)
    return DeepQAgent(conf), HOLDER, conf


@pytest.mark.parametrize('per', [True, False], ids=['per+double+dueling', 'uniform-vanilla'])
def test_agent_seam_drives_the_device_like_the_reference(tmp_path, per, caplog):
    from deepqnetwork_b200.utils import logging as rl_logging
    random.seed(3); np.random.seed(3)
    agent, holder, conf = make_agent(tmp_path, per)
    env = StubEnvironment(seed=5)
    lines = []
    orig_log = rl_logging.log
    import deepqnetwork_b200.deep_q_agent as agent_mod
    agent_mod.log = lambda *a: lines.append(' '.join(str(x) for x in a))
    try:
        with agent:
            L = holder['learner']
            assert L.cfg.math_mode == 1  # the default arithmetic is the benchmarked one (tcgen05 TF32x3), not FFMA
            calls = drive(agent, env, 460)
            assert agent.is_training_done() and agent.train(dict(calls[-1])) is False  # agent seam: False == done
            exp, steps, games, appended = expected_events(calls, agent.conf, per)
            got = list(L.events)
            assert [e for e in got if e[0] != 'predict'] == [e for e in exp if e[0] != 'predict']
            assert sum(1 for e in got if e[0] == 'predict') >= sum(1 for e in exp if e[0] == 'predict') - 1
            assert steps == conf['max_num_training_steps'] == agent.step == L.get_step()
            assert L.get_game_count() == games == agent.model.get_game_count()
            assert len(agent.replay_sampler) == appended == L.replay_info()[0]
            assert L.predict_checked >= 10
            # ring contents: the transitions the agent was handed, in order (uint8 wrap of the raw reward included)
            trans = [c for c in calls if c['old_state'] is not None][:appended]
            rows = L.replay_read(np.arange(appended))
            assert np.array_equal(rows['states'], np.stack([c['old_state'] for c in trans]))
            assert np.array_equal(rows['next_states'], np.stack([c['state'] for c in trans]))
            assert np.array_equal(rows['actions'], np.array([c['action'] for c in trans], np.uint8))
            assert np.array_equal(rows['rewards'], (np.array([c['reward'] for c in trans]).astype(np.int64) & 255).astype(np.uint8))
            assert np.array_equal(rows['continues'], np.array([bool(c['cont']) for c in trans]))
            assert any(c['reward'] == 400 for c in trans)  # the wrap case occurred
            # the replay-seam mirrors answer like the reference objects
            mem = agent.memories
            assert len(mem) == appended and mem.cur_idx == appended % mem.max_size and not mem.is_full
            r7 = mem.get(7)
            assert np.array_equal(r7['state'], trans[7]['old_state']) and r7['action'] == trans[7]['action']
            if per:
                st = agent.replay_sampler.sum_tree
                tree = L.tree_read()[0]
                assert len(st) == appended and st.write == appended % st.capacity and st.total == tree[0]
                leaves = tree[st.capacity - 1:st.capacity - 1 + appended]
                assert agent.replay_sampler.get_min() == leaves.min() and agent.replay_sampler.get_max() == leaves.max()
                from oracle.sumtree_ref import SumTreeRef
                ref = SumTreeRef(st.capacity)
                ref.tree[:], ref.data[:], ref.size = tree, L.tree_read()[1], appended
                for frac in (0.0, 0.37, 0.999, 1.0, 1.0001):  # incl. the "float rounding" clamp past the last filled leaf
                    probe = np.float32(float(tree[0]) * frac)
                    t, d, sc, dat = st.get_with_info(probe)
                    assert (t, d, sc, dat) == tuple(ref.get_with_info(probe)), frac
                    assert st.get_data_idx(t) == d and st.get_data(t) == dat and st.get_score(t) == sc == tree[t]
    finally:
        agent_mod.log = orig_log
    # report line (deep_q_agent.py:400-407), every train_report_interval steps; PER fields only under PER
    reports = [l for l in lines if l.startswith('[train] step')]
    assert [int(re.match(r'\[train\] step (\d+) ', l).group(1)) for l in reports] == [5, 10]
    for l in reports:
        assert re.search(r'avg loss: \d+\.\d{5} ', l) and re.search(r'mem: \d+ ', l) and re.search(r'fr: \d+\.\d eps: \d\.\d\d', l)
        assert bool(re.search(r'per_ab: 0\.60/0\.\d\d avg/min/max loss: .* max_weight: .* sum total: ', l)) == per
    assert any('copying online to target dqn' in l for l in lines)
    # save-dir layout (deep_q_agent.py:126-133): <save_dir>/<environment>.conf + the checkpoint at the same prefix
    prefix = os.path.join(conf['save_dir'], conf['environment'])
    # (persist_format 'reference', the default: what tf.train.Saver and ReplayMemoryDisk leave behind, model.py:60-67 and
    # deep_q_agent.py:416-432,537-541)
    for suffix in ('.conf', '.index', '.data-00000-of-00001', '_replay_memory.hdf5'):
        assert os.path.exists(prefix + suffix), suffix
    assert os.path.exists(os.path.join(conf['save_dir'], 'checkpoint')) and not os.path.exists(prefix + '.dqnb200')


def test_uniform_sampler_draw_rows_matches_reference_execution():
    """ReplaySampler.draw_rows (the mirror's host half of rl/data/replay_sampler.py:27-30) against rows the reference's
    own ReplaySampler selected (tests/golden/replay_uniform_cases.npz, produced by oracle/make_golden.py)."""
    from deepqnetwork_b200.data.replay_memory import ReplayMemory
    from deepqnetwork_b200.data.replay_sampler import ReplaySampler
    g = np.load(os.path.join(G, 'replay_uniform_cases.npz'))
    for k in range(int(g['count'])):
        length, cap, bsz, seed = g['u%d_meta' % k].tolist()
        mem = ReplayMemory(8, 8, 4, max_size=cap)
        v = np.zeros((length, 8, 8, 4), np.uint8)
        v[:, 0, 0, 0] = np.arange(length) & 255; v[:, 0, 0, 1] = np.arange(length) >> 8
        mem.append_rows(v, np.zeros(length, np.uint8), np.zeros(length, np.uint8), v, np.ones(length, bool))
        smp = ReplaySampler(mem)
        random.seed(seed)
        rows = smp.draw_rows(bsz)
        assert rows.tolist() == g['u%d_rows' % k].tolist(), k
        if bsz == mem.learner.batch_size:  # the device gather of those rows, through the sampler seam
            random.seed(seed)
            assert smp.sample_memories(None, bsz).tolist() == rows.tolist()
            got = mem.learner.batch_read()['states']
            assert np.array_equal(got[:, 0, 0, 0].astype(np.int64) + 256 * got[:, 0, 0, 1].astype(np.int64), rows)
        mem.close()
