"""Generate tests/golden/*.npz by EXECUTING the reference's own code.

TEST INFRASTRUCTURE.  Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

The replay/PER modules (rl/data/sum_tree.py, replay_memory.py,
replay_sampler.py, replay_sampler_priority.py) import and run unmodified.
rl/deep_q_agent.py and rl/deep_q_model.py import tensorflow/gym/h5py/PIL at
module top; those imports are satisfied with empty stub modules so that the
reference's *host-side NumPy arithmetic* (``_sample_memories``,
``_make_priority``, ``_update_priority_sampling``, ``_get_target_q_values``)
can be executed verbatim on hand-built instances.  No TensorFlow op is ever
called, so the NN arithmetic stays unpinned (see oracle/__init__.py).

Nothing on the GPU box reads /root/reference: the fixtures written here are
committed and are the only thing tests consume.
"""
import os
import random
import sys
import types

import numpy as np

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests', 'golden')


def _import_reference():
    sys.path.insert(0, REF)
    for name in ('tensorflow', 'gym', 'h5py', 'PIL', 'PIL.Image', 'PIL.ImageDraw',
                 'matplotlib', 'matplotlib.pyplot', 'matplotlib.animation'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['PIL'].Image = sys.modules['PIL.Image']
    sys.modules['PIL'].ImageDraw = sys.modules['PIL.ImageDraw']
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.modules['matplotlib'].animation = sys.modules['matplotlib.animation']
    from rl.data.sum_tree import SumTree
    from rl.data.replay_memory import ReplayMemory
    from rl.data.replay_sampler import ReplaySampler
    from rl.data.replay_sampler_priority import ReplaySamplerPriority
    return SumTree, ReplayMemory, ReplaySampler, ReplaySamplerPriority


def draws(seed, n):
    random.seed(seed)
    return np.array([random.random() for _ in range(n)], dtype=np.float64)


def fill_rows(mem, n, h, w, c, rng):
    for i in range(n):
        mem.append(state=rng.integers(0, 256, (h, w, c), dtype=np.uint8),
                   action=int(rng.integers(0, 9)), reward=int(rng.integers(0, 256)),
                   next_state=rng.integers(0, 256, (h, w, c), dtype=np.uint8),
                   cont=bool(rng.integers(0, 2)))


def g1_uniform(ReplayMemory, ReplaySampler):
    mem = ReplayMemory(2, 2, 1, max_size=100)
    for i in range(100):
        mem.append(state=np.full((2, 2, 1), i, np.uint8), action=i % 4, reward=i % 3,
                   next_state=np.full((2, 2, 1), (i + 1) % 256, np.uint8), cont=(i % 7 != 0))
    batch = ReplayMemory(2, 2, 1, max_size=4)
    random.seed(0)
    ReplaySampler(mem).sample_memories(batch, batch_size=4)
    return dict(rows=batch.states[:, 0, 0, 0].astype(np.int64), actions=batch.actions,
                rewards=batch.rewards, continues=batch.continues, next_states=batch.next_states)


def uniform_cases(ReplayMemory, ReplaySampler):
    """Index selection of ReplaySampler for several (length, B, seed)."""
    out = {}
    k = 0
    for length, cap in ((100, 100), (37, 64), (1000, 1000), (33, 40), (32, 32), (4097, 4097)):
        for bsz in (4, 32):
            if length < bsz:
                continue
            mem = ReplayMemory(1, 1, 4, max_size=cap)
            for i in range(length):
                v = np.array([i & 255, (i >> 8) & 255, 0, 0], np.uint8).reshape(1, 1, 4)
                mem.append(state=v, action=0, reward=0, next_state=v, cont=True)
            batch = ReplayMemory(1, 1, 4, max_size=bsz)
            seed = 100 + k
            random.seed(seed)
            ReplaySampler(mem).sample_memories(batch, batch_size=bsz)
            rows = batch.states[:, 0, 0, 0].astype(np.int64) + 256 * batch.states[:, 0, 0, 1].astype(np.int64)
            out['u%d_meta' % k] = np.array([length, cap, bsz, seed], np.int64)
            out['u%d_rows' % k] = rows
            k += 1
    out['count'] = np.array(k)
    return out


def per_trace(SumTree, capacity, n_add, bsz, rounds, seed, zero_frac=0.0, dup_heavy=False):
    """Run add / sample / update rounds on the reference SumTree and record everything."""
    rng = np.random.default_rng(seed)
    tree = SumTree(capacity)
    add_scores = np.power(np.minimum(np.abs(rng.normal(0, 0.3, n_add)).astype(np.float32)
                                     + np.float32(0.01), np.float32(1.0)), np.float32(0.6)).astype(np.float32)
    if zero_frac > 0:
        add_scores[rng.random(n_add) < zero_frac] = 0.0
    if dup_heavy:
        # one leaf dominates so that several strata land on the same leaf
        add_scores[:] = np.float32(1e-3)
        add_scores[n_add // 2] = np.float32(1.0)
    for i in range(n_add):
        tree.add(add_scores[i], i % capacity)
    rec = dict(meta=np.array([capacity, n_add, bsz, rounds, seed], np.int64),
               add_scores=add_scores, tree_after_add=tree.tree.copy(), data_after_add=tree.data.copy(),
               size_write=np.array([tree.size, tree.write, int(tree.max_reached)], np.int64),
               stats_after_add=np.array([tree.get_min(), tree.get_max(), tree.get_average(), tree.total],
                                        np.float32))
    U, T, D, P, M, S, ROOT = [], [], [], [], [], [], []
    for r in range(rounds):
        u = draws(seed * 1000 + r, bsz)
        random.seed(seed * 1000 + r)
        size = tree.total / bsz
        t = np.zeros(bsz, np.int64); d = np.zeros(bsz, np.int64)
        p = np.zeros(bsz, np.float64); m = np.zeros(bsz, np.int64)
        for i in range(bsz):
            # same expression as replay_sampler_priority.py:33
            s = random.random() * size + i * size
            t[i], d[i], p[i], m[i] = tree.get_with_info(s)
        new = np.power(np.minimum(np.abs(rng.normal(0, 0.5, bsz)).astype(np.float32)
                                  + np.float32(0.01), np.float32(1.0)), np.float32(0.6)).astype(np.float32)
        for ti, pi in zip(t, new):
            tree.update_score(ti, pi)
        U.append(u); T.append(t); D.append(d); P.append(p); M.append(m); S.append(new)
        ROOT.append(tree.total)
    rec.update(uniforms=np.array(U), tree_idx=np.array(T), data_idx=np.array(D), prio=np.array(P),
               mem_idx=np.array(M), new_scores=np.array(S), roots=np.array(ROOT, np.float32),
               tree_final=tree.tree.copy())
    return rec


def sampler_priority_g2(ReplayMemory, ReplaySamplerPriority):
    """capacity 10 (two leaf depths) through the sampler class itself."""
    mem = ReplayMemory(1, 1, 4, max_size=10)
    smp = ReplaySamplerPriority(mem)
    for i in range(7):
        v = np.full((1, 1, 4), i, np.uint8)
        smp.append(v, i % 4, i, v, True, np.float32(0.1 * (i + 1)))
    tree0 = smp.sum_tree.tree.copy()
    batch = ReplayMemory(1, 1, 4, max_size=4)
    t = np.zeros(4, dtype=int); p = np.zeros(4, dtype=float)
    random.seed(0)
    smp.sample_memories(batch, batch_size=4, tree_idxes=t, priorities=p)
    rows = batch.states[:, 0, 0, 0].astype(np.int64)
    smp.update_sum_tree(t, np.array([.5, .25, .125, 1.0], np.float32))
    return dict(tree0=tree0, tree_idx=t.astype(np.int64), prio=p, rows=rows,
                uniforms=draws(0, 4), tree1=smp.sum_tree.tree.copy(), losses_f16=mem.losses.copy())


def small_cases(SumTree):
    out = {}
    # G3 fix-up clamp
    t = SumTree(8)
    for i in range(3):
        t.add(np.float32(1.0), i)
    a = t.get_with_info(t.total)
    b = t.get_with_info(t.total * np.float32(1.0001))
    out['g3'] = np.array([a[0], a[1], b[0], b[1]], np.int64)
    # G4 wrap
    t = SumTree(4)
    for i in range(1, 7):
        t.add(np.float32(i), i)
    out['g4_tree'] = t.tree.copy(); out['g4_data'] = t.data.copy()
    out['g4_sw'] = np.array([t.size, t.write], np.int64)
    # G5 drift
    rng = np.random.default_rng(3)
    t = SumTree(1000)
    sc = rng.random(1000).astype(np.float32)
    for i in range(1000):
        t.add(sc[i], i)
    idx = rng.integers(999, 1999, 20000)
    val = rng.random(20000).astype(np.float32)
    for i, v in zip(idx, val):
        t.update_score(int(i), v)
    out['g5_scores'] = sc; out['g5_idx'] = idx.astype(np.int64); out['g5_val'] = val
    out['g5_tree'] = t.tree.copy()
    return out


def agent_math():
    """Execute the reference agent/model host arithmetic verbatim (TF stubbed, never called)."""
    from rl.deep_q_agent import DeepQAgent
    from rl.deep_q_model import DeepQModel
    from rl.data.replay_memory import ReplayMemory
    from rl.data.replay_sampler_priority import ReplaySamplerPriority
    out = {}
    rng = np.random.default_rng(11)

    # _make_priority  (deep_q_agent.py:527-534)
    ag = object.__new__(DeepQAgent)
    ag.per_a = 0.6
    losses = np.concatenate([np.abs(rng.normal(0, 1.0, 500)), [0.0, 0.99, 1.0, 5.0, 1e-8]]).astype(np.float32)
    out['prio_in'] = losses
    out['prio_out'] = ag._make_priority(losses)
    assert out['prio_out'].dtype == np.float32

    # _sample_memories  (deep_q_agent.py:500-524): IS weights
    h, w, c, cap, n, bsz = 3, 2, 4, 300, 257, 32
    mem = ReplayMemory(h, w, c, max_size=cap)
    smp = ReplaySamplerPriority(mem)
    pri = ag._make_priority(np.abs(rng.normal(0, 0.3, n)).astype(np.float32))
    st = rng.integers(0, 256, (n, h, w, c), dtype=np.uint8)
    ns = rng.integers(0, 256, (n, h, w, c), dtype=np.uint8)
    ac = rng.integers(0, 9, n).astype(np.uint8); rw = rng.integers(0, 256, n).astype(np.uint8)
    ct = rng.integers(0, 2, n).astype(bool)
    for i in range(n):
        smp.append(st[i], ac[i], rw[i], ns[i], ct[i], pri[i])
    ag.use_per = True; ag.step = 0; ag.per_calculate_steps = 5000; ag.last_max_weight = None
    ag.replay_sampler = smp; ag.batch_size = bsz; ag.per_b = 0.4
    ag.train_batch = ReplayMemory(h, w, c, max_size=bsz)
    ag.tree_idxes = np.zeros(bsz, dtype=int); ag.priorities = np.zeros(bsz, dtype=float)
    random.seed(5)
    ag._sample_memories()
    out.update(is_states=st, is_next_states=ns, is_actions=ac, is_rewards=rw, is_continues=ct,
               is_priorities_in=pri, is_uniforms=draws(5, bsz), is_meta=np.array([h, w, c, cap, n, bsz], np.int64),
               is_tree_idx=ag.tree_idxes.astype(np.int64), is_prio=ag.priorities.copy(),
               is_weights=np.asarray(ag.is_weights, np.float64),
               is_stats=np.array([ag.last_avg_loss, ag.last_min_loss, ag.last_max_loss, ag.last_max_weight,
                                  smp.total], np.float64),
               is_batch_states=ag.train_batch.states.copy(), is_batch_next=ag.train_batch.next_states.copy(),
               is_batch_actions=ag.train_batch.actions.copy(), is_batch_rewards=ag.train_batch.rewards.copy(),
               is_batch_continues=ag.train_batch.continues.copy(), is_batch_losses=ag.train_batch.losses.copy())

    # _update_priority_sampling  (deep_q_agent.py:445-450)
    ag.conf = {'use_per_annealing': True}
    ag.per_b_start = 0.4; ag.per_b_end = 1; ag.per_anneal_steps = 2000000
    steps = np.array([0, 1, 12345, 1999999, 2000000, 5000000], np.int64)
    betas = []
    for s in steps:
        ag.step = int(s); ag.per_b = ag.per_b_start
        ag._update_priority_sampling()
        betas.append(ag.per_b)
    out['beta_steps'] = steps; out['beta_vals'] = np.array(betas, np.float64)

    # _get_target_q_values host arithmetic (deep_q_model.py:106-114), TF run() stubbed
    md = object.__new__(DeepQModel)
    md.discount_rate = 0.99; md.X_state = 'x'; md.max_q_values = 'q'
    maxq = rng.normal(0, 2, 64).astype(np.float32)
    md.run = lambda fetches, feed_dict=None: [maxq]
    rew = rng.integers(0, 256, 64).astype(np.uint8); con = rng.integers(0, 2, 64).astype(bool)
    y = md._get_target_q_values(rew, con, None)
    out.update(y_maxq=maxq, y_rewards=rew, y_continues=con, y_out=np.asarray(y), y_dtype=np.array(str(y.dtype)))
    return out


def preprocess_cases():
    """Frames pushed through the reference environments' own preprocess_observation (rl/game_environment.py:110-160;
    instances made with __new__: the constructor only wraps gym), plus the CRC of the full RGB cube through Breakout's."""
    import zlib
    from rl.game_environment import BreakoutEnvironment, SpaceInvadersEnvironment, MsPacmanEnvironment
    rng = np.random.default_rng(7)
    out = {}
    for name, cls in (('Breakout', BreakoutEnvironment), ('SpaceInvaders', SpaceInvadersEnvironment), ('MsPacman', MsPacmanEnvironment)):
        env = cls.__new__(cls)
        obs = rng.integers(0, 256, (2, 210, 160, 3), dtype=np.uint8)
        obs[1] = 255                      # saturated frame: 262.65 -> wraps to 6
        obs[1, ::7, ::5] = (0, 0, 255); obs[1, 1::7, ::3] = (255, 255, 0); obs[1, 2::9] = np.arange(160)[:, None]
        out[name + '_obs'] = obs
        out[name + '_out'] = np.stack([env.preprocess_observation(o) for o in obs])
    # every (r, g, b) triple through Breakout's method, 7120 triples per call, r-major order
    env = BreakoutEnvironment.__new__(BreakoutEnvironment)
    r, g, b = np.meshgrid(np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), np.arange(256, dtype=np.uint8), indexing='ij')
    cube = np.stack([r, g, b], axis=-1).reshape(-1, 3)
    per = 89 * 80
    crc = 0
    for i0 in range(0, len(cube), per):
        chunk = cube[i0:i0 + per]
        pad = np.zeros((per, 3), np.uint8); pad[:len(chunk)] = chunk
        obs = np.zeros((210, 160, 3), np.uint8)
        obs[16:-16:2, ::2] = pad.reshape(89, 80, 3)
        res = env.preprocess_observation(obs).reshape(-1)[:len(chunk)]
        crc = zlib.crc32(res.tobytes(), crc)
    out['cube_crc'] = np.array(crc, np.uint64)
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    SumTree, ReplayMemory, ReplaySampler, ReplaySamplerPriority = _import_reference()
    np.savez_compressed(os.path.join(OUT, 'replay_g1_uniform.npz'), **g1_uniform(ReplayMemory, ReplaySampler))
    np.savez_compressed(os.path.join(OUT, 'replay_uniform_cases.npz'), **uniform_cases(ReplayMemory, ReplaySampler))
    np.savez_compressed(os.path.join(OUT, 'per_g2_sampler.npz'), **sampler_priority_g2(ReplayMemory, ReplaySamplerPriority))
    np.savez_compressed(os.path.join(OUT, 'per_small_cases.npz'), **small_cases(SumTree))
    traces = [
        # capacity, n_add, B, rounds, seed, zero_frac, dup_heavy
        (2, 2, 4, 3, 1, 0.0, False),
        (3, 2, 4, 3, 2, 0.0, False),
        (7, 5, 8, 4, 3, 0.0, False),
        (10, 7, 4, 4, 4, 0.0, False),
        (37, 37, 32, 6, 5, 0.0, False),
        (64, 40, 32, 6, 6, 0.2, False),
        (64, 150, 32, 6, 7, 0.0, False),      # wrapped ring
        (1000, 1000, 32, 8, 8, 0.0, False),
        (1024, 700, 32, 8, 9, 0.3, False),
        (4097, 4097, 32, 8, 10, 0.0, False),
        (1000, 1000, 32, 4, 11, 0.0, True),   # many duplicate leaves per batch
        (5000, 5000, 128, 3, 12, 0.0, False),
        (300, 300, 33, 3, 13, 0.0, False),    # B not a multiple of 32
    ]
    for k, (cap, n_add, bsz, rounds, seed, zf, dup) in enumerate(traces):
        np.savez_compressed(os.path.join(OUT, 'per_trace_%02d.npz' % k),
                            **per_trace(SumTree, cap, n_add, bsz, rounds, seed, zf, dup))
    np.savez_compressed(os.path.join(OUT, 'agent_math.npz'), **agent_math())
    np.savez_compressed(os.path.join(OUT, 'preprocess_cases.npz'), **preprocess_cases())
    print('wrote', sorted(os.listdir(OUT)))


if __name__ == '__main__':
    main()
