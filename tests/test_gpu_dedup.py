"""-m gpu: the frame-de-duplicated replay layout (dqn_config.frame_dedup; BASELINE.json configs[2] "10M-step-scale replay
resident in HBM").  The reference stores the full s and s' stacks of every transition (rl/data/replay_memory.py:26-39);
the new layout stores each frame once.  Everything observable must be byte-identical to the stacked layout: rows read
back, sampled batches, and therefore every training step."""
import random

import numpy as np
import pytest

from tests.gpu_util import need_gpu
from tests.stub_env import StubEnvironment, drive

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _gpu():
    need_gpu()


class _Recorder:
    """Agent-seam stand-in that only records what GameRunner hands over."""

    def __init__(self, actions):
        self.rng = np.random.default_rng(1); self.actions = actions

    def train(self, gs):
        return True

    def get_action(self, state):
        return int(self.rng.integers(0, self.actions))

    def is_training_done(self):
        return False


def transitions_from_runner(n_calls, h=86, w=80, seed=3, episode_len=23, lives=2):
    calls = drive(_Recorder(9), StubEnvironment(h, w, 9, seed=seed, episode_len=episode_len, lives=lives), n_calls)
    tr = [c for c in calls if c['old_state'] is not None]
    assert any(np.array_equal(c['old_state'], c['state']) for c in tr)  # the s == s' transition after a life loss (Q5)
    games = sum(1 for c in calls if c['old_state'] is None)             # first call of a game: no transition yet
    return tr, games


def make_pair(cap, h=86, w=80, a=9, per=True, **kw):
    from deepqnetwork_b200.learner import Learner
    from deepqnetwork_b200.deep_q_model import init_tower_params, tower_param_shapes
    shapes = tower_param_shapes(h, w, 4, a, 512, True)
    out = []
    for dedup in (False, True):
        L = Learner(h, w, 4, num_actions=a, use_dueling=True, use_double=True, use_per=per, batch_size=32, replay_capacity=cap,
                    math_mode='tc3x', frame_dedup=dedup, max_rows=256, **kw)
        L.set_tower(init_tower_params(shapes, 1), 0); L.set_tower(init_tower_params(shapes, 2), 1)
        out.append(L)
    return out


@pytest.mark.parametrize('cap,n_calls', [(600, 420), (200, 520)], ids=['partial-ring', 'wrapped-ring'])
def test_dedup_ring_is_byte_identical_to_the_stacked_ring(cap, n_calls):
    from deepqnetwork_b200.data.replay_memory import ReplayMemory
    tr, games = transitions_from_runner(n_calls)
    Ls, Ld = make_pair(cap, frame_capacity=cap + 96)
    mems = [ReplayMemory(86, 80, 4, max_size=cap, learner=L) for L in (Ls, Ld)]
    rng = np.random.default_rng(0)
    for i0 in range(0, len(tr), 128):  # the 128-row PER insert batches of deep_q_agent.py:467-497
        chunk = tr[i0:i0 + 128]
        pri = (rng.random(len(chunk)) * 0.9 + 0.05).astype(np.float32)
        for m in mems:
            m.append_rows(np.stack([c['old_state'] for c in chunk]), [c['action'] for c in chunk], [c['reward'] for c in chunk],
                          np.stack([c['state'] for c in chunk]), [c['cont'] for c in chunk], pri)
    n = min(len(tr), cap)
    assert len(mems[0]) == len(mems[1]) == n and mems[0].cur_idx == mems[1].cur_idx
    # one new frame per environment step (+3 when a game starts with an empty deque): 8x fewer frames than 2 x 4 per row
    assert mems[1].frames_uploaded <= len(tr) + 4 * (games + 1)
    rows = np.arange(n)
    a, b = Ls.replay_read(rows), Ld.replay_read(rows)
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(Ls.tree_read()[0], Ld.tree_read()[0])
    # sampled batches and three fused training steps: identical draws -> identical bytes -> bit-identical learners
    rnd = random.Random(5)
    for it in range(3):
        u = np.array([rnd.random() for _ in range(32)])
        ra, rb = Ls.train_step(uniforms=u), Ld.train_step(uniforms=u)
        ba, bb = Ls.batch_read(), Ld.batch_read()
        for k in ba:
            assert np.array_equal(ba[k], bb[k]), (it, k)
        assert ra[0] == rb[0] and np.array_equal(ra[1], rb[1]) and ra[2] == rb[2]
    for name, _, _ in Ls.tensors():
        assert np.array_equal(Ls.get_param(name), Ld.get_param(name)), name
    assert np.array_equal(Ls.tree_read()[0], Ld.tree_read()[0])
    for m in mems:
        m.close()
    Ls.close(); Ld.close()


def test_frame_ring_overrun_is_an_error_not_corruption():
    from deepqnetwork_b200.data.replay_memory import ReplayMemory
    tr, _ = transitions_from_runner(300)
    Ls, Ld = make_pair(200, frame_capacity=64)  # far fewer frame slots than live transitions need
    mem = ReplayMemory(86, 80, 4, max_size=200, learner=Ld)
    with pytest.raises(RuntimeError, match='frame ring too small'):
        for c in tr:
            mem.append(c['old_state'], c['action'], c['reward'], c['state'], c['cont'], 0.5)
    Ls.close(); Ld.close()


def test_synthetic_fill_and_gather_in_the_dedup_layout():
    """The device-side synthetic fill of the dedup layout (bench.py --config c3) against its host restatement, through
    both gathers (dqn_uniform_sample and the read-back path), and a graph-replayed run on it."""
    from deepqnetwork_b200.learner import Learner
    from oracle import philox_ref
    H, W, N, seed = 86, 80, 50000, 1234
    L = Learner(H, W, 4, num_actions=9, use_dueling=True, use_double=True, use_per=True, batch_size=32, replay_capacity=N,
                math_mode='tc3x', frame_dedup=True)
    L.replay_fill_synthetic(N, seed=seed)
    rows = np.array(sorted(np.random.default_rng(0).choice(N, 32, replace=False)), np.int64)
    rows[0], rows[-1] = 0, N - 1
    fb = H * W
    def stack(t0):
        f = philox_ref.synthetic_state_rows(np.arange(t0, t0 + 4), fb, seed).reshape(4, H, W)
        return np.stack(list(f), axis=2)
    exp_s = np.stack([stack(t) for t in rows]); exp_n = np.stack([stack(t + 1) for t in rows])
    L.uniform_sample(rows)
    got = L.batch_read()
    assert np.array_equal(got['states'], exp_s) and np.array_equal(got['next_states'], exp_n)
    rb = L.replay_read(rows)
    assert np.array_equal(rb['states'], exp_s) and np.array_equal(rb['next_states'], exp_n)
    for k in ('actions', 'rewards', 'continues', 'losses'):
        assert np.array_equal(got[k], rb[k]), k
    L.train_steps(5); L.sync()
    assert L.get_step() == 5
    L.close()


@pytest.mark.parametrize('dedup', [False, True], ids=['stacked', 'frames'])
def test_replay_save_restore_follows_the_reference_reload(tmp_path, dedup):
    """DeepQAgent._train_finish / _train_init_memory (rl/deep_q_agent.py:410-432, :211-222): the replay is written at exit,
    rows 0 .. len-1 in index order; at the next start the rows are appended from index 0 and the sum tree is rebuilt from the
    float16 losses column (ReplaySamplerPriority.add_losses).  Rows byte-identical, tree bit-identical to the oracle's
    rebuild, and training continues identically on both layouts."""
    from deepqnetwork_b200.data.replay_memory import ReplayMemory
    from oracle.sumtree_c import SumTreeC
    tr, _ = transitions_from_runner(330)
    cap = 400
    L = make_pair(cap, frame_capacity=cap + 96)[1 if dedup else 0]
    mem = ReplayMemory(86, 80, 4, max_size=cap, learner=L)
    rng = np.random.default_rng(0)
    pri = (rng.random(len(tr)) * 0.9 + 0.05).astype(np.float32)
    mem.append_rows(np.stack([c['old_state'] for c in tr]), [c['action'] for c in tr], [c['reward'] for c in tr],
                    np.stack([c['state'] for c in tr]), [c['cont'] for c in tr], pri)
    rnd = random.Random(1)
    for _ in range(3):
        L.train_step(uniforms=np.array([rnd.random() for _ in range(32)]), fetch=False)  # priority write-back changes losses + tree
    n = len(tr)
    before = L.replay_read(np.arange(n))
    prefix = str(tmp_path / 'rl.game_environment.MsPacmanEnvironment')
    L.save_replay(prefix)
    L2 = make_pair(cap, frame_capacity=cap + 96)[1 if dedup else 0]
    assert L2.restore_replay(str(tmp_path / 'missing')) is None
    assert L2.restore_replay(prefix) == n
    assert L2.replay_info() == (n, n % cap, False)
    after = L2.replay_read(np.arange(n))
    for k in before:
        assert np.array_equal(before[k], after[k]), k
    ref = SumTreeC(cap)
    ref.add(before['losses'].astype(np.float32), np.arange(n).astype(np.uint32))  # add(losses[i], i), float16 values
    tree, data, size, write = L2.tree_read()
    assert np.array_equal(tree, ref.tree) and np.array_equal(data, ref.data) and (size, write) == (n, n % cap)
    u = np.array([rnd.random() for _ in range(32)])
    t, d, p, m = ref.sample(u)
    L2.train_step(uniforms=u, fetch=False)
    got = L2.batch_read()
    assert np.array_equal(got['states'], before['states'][m]) and np.array_equal(got['next_states'], before['next_states'][m])
    L.close(); L2.close()
