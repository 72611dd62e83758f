"""Oracle NN/agent-maths restatement vs fixtures produced by executing the reference's
host arithmetic (oracle/make_golden.py: agent_math) and vs SURVEY.md's geometry facts.  CPU only."""
import os

import numpy as np
import torch

from oracle import nn_ref
from oracle.sumtree_ref import ReplayRingRef, SumTreeRef, per_sample

G = os.path.join(os.path.dirname(__file__), 'golden')


def test_param_counts_and_geometry():
    S = nn_ref.NetSpec
    assert S(89, 80, 4, 4).num_params() == 4041380
    assert S(89, 80, 4, 4, use_dueling=True).num_params() == 7974565
    assert S(86, 80, 4, 9, use_dueling=True).num_params() == 7321770
    assert S(84, 84, 4, 4, variant='nature').num_params() == 1686180
    l = S(89, 80, 4, 4).conv_layers()
    assert [(x['oh'], x['ow']) for x in l] == [(23, 20), (12, 10), (12, 10)]
    assert [(x['pt'], x['pb'], x['pl'], x['pr']) for x in l] == [(3, 4, 2, 2), (1, 2, 1, 1), (1, 2, 1, 2)]
    l = S(86, 80, 4, 9).conv_layers()
    assert [(x['oh'], x['ow']) for x in l] == [(22, 20), (11, 10), (11, 10)]
    assert (l[0]['pt'], l[0]['pb']) == (3, 3) and (l[1]['pt'], l[1]['pb']) == (1, 1)
    assert list(S(89, 80, 4, 4, use_dueling=True).param_shapes())[6:] == [
        'dense/kernel', 'dense/bias', 'dense_1/kernel', 'dense_1/bias',
        'dense_2/kernel', 'dense_2/bias', 'dense_3/kernel', 'dense_3/bias']


def test_make_priority_matches_reference_execution():
    g = np.load(os.path.join(G, 'agent_math.npz'))
    out = nn_ref.make_priority(g['prio_in'])
    assert out.dtype == np.float32 and np.array_equal(out, g['prio_out'])


def test_is_weights_match_reference_execution():
    g = np.load(os.path.join(G, 'agent_math.npz'))
    h, w, c, cap, n, bsz = g['is_meta'].tolist()
    tree = SumTreeRef(cap); ring = ReplayRingRef(h, w, c, cap)
    for i in range(n):
        tree.add(g['is_priorities_in'][i], ring.cur_idx)
        ring.append(g['is_states'][i], g['is_actions'][i], g['is_rewards'][i], g['is_next_states'][i],
                    g['is_continues'][i], g['is_priorities_in'][i])
    avg, mn, mx, maxw, total = g['is_stats']
    assert tree.total == np.float32(total) and np.float32(tree.get_average()) == np.float32(avg)
    mw = nn_ref.max_weight(tree.get_min(), tree.total, bsz, 0.4)
    assert np.float64(mw) == maxw
    t, d, p, m = per_sample(tree, bsz, g['is_uniforms'])
    assert np.array_equal(t, g['is_tree_idx']) and np.array_equal(p, g['is_prio'])
    wts = nn_ref.is_weights(p, tree.total, bsz, 0.4, mw)
    assert np.array_equal(np.asarray(wts, np.float64), g['is_weights'])
    b = ring.gather(m)
    for k, gk in (('states', 'is_batch_states'), ('next_states', 'is_batch_next'), ('actions', 'is_batch_actions'),
                  ('rewards', 'is_batch_rewards'), ('continues', 'is_batch_continues'), ('losses', 'is_batch_losses')):
        assert np.array_equal(b[k], g[gk]), k


def test_beta_anneal_and_host_target():
    g = np.load(os.path.join(G, 'agent_math.npz'))
    assert [nn_ref.per_beta(int(s)) for s in g['beta_steps']] == g['beta_vals'].tolist()
    y = nn_ref.host_target(g['y_rewards'], g['y_continues'], g['y_maxq'], 0.99)
    assert y.dtype == np.float64 == np.dtype(str(g['y_dtype'])) and np.array_equal(y, g['y_out'])


def test_train_step_fp32_tracks_fp64():
    """The fp32 restatement stays within 1e-4 of the fp64 truth for one step (noise yardstick)."""
    from deepqnetwork_b200.deep_q_model import init_tower_params
    spec = nn_ref.NetSpec(89, 80, 4, 4, use_dueling=True, use_double=True, use_per=True)
    rng = np.random.default_rng(0)
    on = init_tower_params(spec.param_shapes(), seed=1); tg = init_tower_params(spec.param_shapes(), seed=2)
    B = 8
    batch = dict(states=rng.integers(0, 256, (B, 89, 80, 4), dtype=np.uint8),
                 next_states=rng.integers(0, 256, (B, 89, 80, 4), dtype=np.uint8),
                 actions=rng.integers(0, 4, B).astype(np.uint8), rewards=rng.integers(0, 3, B).astype(np.uint8),
                 continues=rng.integers(0, 2, B).astype(bool))
    w = rng.random(B).astype(np.float32)
    res = {}
    for dt in (torch.float32, torch.float64):
        o = nn_ref.to_torch(on, dt); t = nn_ref.to_torch(tg, dt)
        opt = nn_ref.OptState(o, spec, dt)
        step, losses, loss, aux = nn_ref.train_step(o, t, opt, batch, w, spec, dt)
        res[dt] = (losses, loss, aux['q'], o)
        assert step == 1
    l32, l64 = res[torch.float32], res[torch.float64]
    rel = lambda a, b: np.max(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64))) / max(np.max(np.abs(np.asarray(b))), 1e-30)
    assert rel(l32[0], l64[0]) < 1e-4 and rel(l32[2], l64[2]) < 1e-4 and abs(l32[1] - l64[1]) < 1e-4 * abs(l64[1])
    for k in l64[3]:
        assert rel(l32[3][k].numpy(), l64[3][k].numpy()) < 1e-4, k
