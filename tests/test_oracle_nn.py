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


# ---- second, independent restatement (no torch, explicit index loops) on tiny shapes ---------------------------------
# oracle/nn_ref.py is "parity unpinned" (TensorFlow cannot be run here).  What can be done without TF is to make sure the
# restatement does not rest on a misreading shared with the device code: the functions below are written directly from
# the published TF formulas -- conv2d 'SAME': out[b,i,j,k] = sum_{di,dj,q} in[b, s*i+di-pad_top, s*j+dj-pad_left, q] *
# filter[di,dj,q,k] with pad_total = max((ceil(in/s)-1)*s + k - in, 0), pad_top = pad_total // 2 (the extra row/column goes
# to the bottom/right); tf.layers.flatten = C-order reshape of NHWC; dense = x @ kernel + bias; ApplyAdam as documented in
# tf.train.AdamOptimizer -- with a hand-written backward pass (no autograd), and compared with nn_ref element by element.
def _np_conv_same(x, w, b, s):
    n, ih, iw, cin = x.shape
    k, _, _, cout = w.shape
    oh, ow = -(-ih // s), -(-iw // s)
    pt = max((oh - 1) * s + k - ih, 0) // 2
    pl = max((ow - 1) * s + k - iw, 0) // 2
    out = np.zeros((n, oh, ow, cout))
    for i in range(oh):
        for j in range(ow):
            for di in range(k):
                for dj in range(k):
                    y, xx = s * i + di - pt, s * j + dj - pl
                    if 0 <= y < ih and 0 <= xx < iw:
                        out[:, i, j, :] += x[:, y, xx, :] @ w[di, dj]
    return out + b, (pt, pl)


def _np_conv_same_backward(x, w, s, pads, dy):
    n, ih, iw, cin = x.shape
    k = w.shape[0]
    pt, pl = pads
    dx, dw = np.zeros_like(x), np.zeros_like(w)
    for i in range(dy.shape[1]):
        for j in range(dy.shape[2]):
            for di in range(k):
                for dj in range(k):
                    y, xx = s * i + di - pt, s * j + dj - pl
                    if 0 <= y < ih and 0 <= xx < iw:
                        dw[di, dj] += x[:, y, xx, :].T @ dy[:, i, j, :]
                        dx[:, y, xx, :] += dy[:, i, j, :] @ w[di, dj].T
    return dx, dw, dy.sum((0, 1, 2))


def _np_tower(p, x_u8, dueling, keep=None):
    a = x_u8.astype(np.float64) / 255
    tape = []
    for i, s in enumerate((4, 2, 1)):
        nm = 'conv2d' if i == 0 else 'conv2d_%d' % i
        pre, pads = _np_conv_same(a, p[nm + '/kernel'], p[nm + '/bias'], s)
        tape.append((nm, a, s, pads, pre))
        a = np.maximum(pre, 0)
    flat = a.reshape(a.shape[0], -1)
    if dueling:
        pa, pv = flat @ p['dense/kernel'] + p['dense/bias'], flat @ p['dense_2/kernel'] + p['dense_2/bias']
        ha, hv = np.maximum(pa, 0), np.maximum(pv, 0)
        adv = ha @ p['dense_1/kernel'] + p['dense_1/bias']
        val = hv @ p['dense_3/kernel'] + p['dense_3/bias']
        q = val + (adv - adv.mean(1, keepdims=True))
        head = (flat, pa, pv, ha, hv)
    else:
        ph = flat @ p['dense/kernel'] + p['dense/bias']
        hh = np.maximum(ph, 0)
        q = hh @ p['dense_1/kernel'] + p['dense_1/bias']
        head = (flat, ph, hh)
    return q, tape, head, a.shape


def _np_backward(p, dq, tape, head, act_shape, dueling):
    g = {}
    if dueling:
        flat, pa, pv, ha, hv = head
        A = dq.shape[1]
        dval = dq.sum(1, keepdims=True)
        dadv = dq - dq.sum(1, keepdims=True) / A
        g['dense_1/kernel'], g['dense_1/bias'] = ha.T @ dadv, dadv.sum(0)
        g['dense_3/kernel'], g['dense_3/bias'] = hv.T @ dval, dval.sum(0)
        dpa = (dadv @ p['dense_1/kernel'].T) * (pa > 0)
        dpv = (dval @ p['dense_3/kernel'].T) * (pv > 0)
        g['dense/kernel'], g['dense/bias'] = flat.T @ dpa, dpa.sum(0)
        g['dense_2/kernel'], g['dense_2/bias'] = flat.T @ dpv, dpv.sum(0)
        dflat = dpa @ p['dense/kernel'].T + dpv @ p['dense_2/kernel'].T
    else:
        flat, ph, hh = head
        g['dense_1/kernel'], g['dense_1/bias'] = hh.T @ dq, dq.sum(0)
        dph = (dq @ p['dense_1/kernel'].T) * (ph > 0)
        g['dense/kernel'], g['dense/bias'] = flat.T @ dph, dph.sum(0)
        dflat = dph @ p['dense/kernel'].T
    da = dflat.reshape(act_shape)
    for nm, x, s, pads, pre in reversed(tape):
        dpre = da * (pre > 0)
        da, g[nm + '/kernel'], g[nm + '/bias'] = _np_conv_same_backward(x, p[nm + '/kernel'], s, pads, dpre)
    return g


def _tiny_case(dueling, double, per, seed):
    from deepqnetwork_b200.deep_q_model import init_tower_params
    H, W, A, hid, B = 13, 10, 3, 16, 3
    spec = nn_ref.NetSpec(H, W, 4, A, num_hidden=hid, use_dueling=dueling, use_double=double, use_per=per)
    rng = np.random.default_rng(seed)
    on = init_tower_params(spec.param_shapes(), seed + 1); tg = init_tower_params(spec.param_shapes(), seed + 2)
    for pz in (on, tg):
        for k in pz:
            if k.endswith('/bias'):
                pz[k] = rng.normal(0, 0.1, pz[k].shape).astype(np.float32)
    batch = dict(states=rng.integers(0, 256, (B, H, W, 4), dtype=np.uint8),
                 next_states=rng.integers(0, 256, (B, H, W, 4), dtype=np.uint8),
                 actions=rng.integers(0, A, B).astype(np.uint8), rewards=np.array([0, 1, 200], np.uint8),
                 continues=np.array([True, False, True]))
    return spec, on, tg, batch, rng.random(B).astype(np.float32) + 0.2


def _close(a, b, tol=1e-9):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.max(np.abs(a - b)) <= tol * max(np.max(np.abs(b)), 1e-30)


import pytest


@pytest.mark.parametrize('dueling,double,per', [(False, False, False), (True, True, True), (True, False, False), (False, True, True)])
def test_nn_oracle_agrees_with_an_independent_loop_restatement(dueling, double, per):
    spec, on, tg, batch, isw = _tiny_case(dueling, double, per, 11)
    l = spec.conv_layers()
    assert [(x['oh'], x['ow']) for x in l] == [(4, 3), (2, 2), (2, 2)] and spec.flat() == 256
    P = {k: v.astype(np.float64) for k, v in on.items()}; T = {k: v.astype(np.float64) for k, v in tg.items()}
    A, B = spec.num_actions, 3
    # target (deep_q_model.py:247-255, :114)
    qt = _np_tower(T, batch['next_states'], dueling)[0]
    if double:
        qo = _np_tower(P, batch['next_states'], dueling)[0]
        sel = np.argmax(qo, 1)
        max_q = np.array([max(qt[b, sel[b]], 0.0) for b in range(B)])  # max over a row that is zero except at sel
    else:
        max_q = qt.max(1)
    y = batch['rewards'].astype(np.float64) + batch['continues'].astype(np.float64) * 0.99 * max_q
    q, tape, head, ashape = _np_tower(P, batch['states'], dueling)
    qa = q[np.arange(B), batch['actions']]
    err = qa - y
    if per:  # deep_q_model.py:372-376
        loss = np.mean(isw.astype(np.float64) * err ** 2); losses = np.abs(err)
        dqa = 2 * isw.astype(np.float64) * err / B
    else:    # tf.losses.huber_loss(delta=1), mean
        ab = np.abs(err); quad = np.minimum(ab, 1.0)
        losses = 0.5 * quad ** 2 + (ab - quad); loss = losses.mean()
        dqa = np.clip(err, -1.0, 1.0) / B
    dq = np.zeros_like(q); dq[np.arange(B), batch['actions']] = dqa
    grads = _np_backward(P, dq, tape, head, ashape, dueling)
    # ApplyAdam, first step (tf.train.AdamOptimizer docs): lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m,v from zero
    f32 = lambda v: float(np.float32(v))
    lr, b1, b2, eps = f32(0.00025), f32(0.9), f32(0.999), f32(1e-8)
    new = {}
    for k in P:
        m = (1 - b1) * grads[k]; v = (1 - b2) * grads[k] ** 2
        new[k] = P[k] - lr * np.sqrt(1 - b2) / (1 - b1) * m / (np.sqrt(v) + eps)
    # the torch restatement
    o = nn_ref.to_torch(on, torch.float64); t = nn_ref.to_torch(tg, torch.float64)
    opt = nn_ref.OptState(o, spec, torch.float64)
    step, rlosses, rloss, aux = nn_ref.train_step(o, t, opt, batch, isw if per else None, spec, torch.float64)
    assert step == 1
    assert _close(q, aux['q']) and _close(max_q, aux['max_q']) and _close(y, aux['y'])
    assert _close(losses, rlosses) and abs(loss - rloss) <= 1e-9 * abs(rloss)
    for k in P:
        assert _close(grads[k], aux['grads'][k]), ('grad', k)
        assert _close(new[k], o[k].numpy(), 1e-9), ('weight', k)
    # two more Adam steps on a scalar, against the closed form of the documented recurrence
    w, m, v, b1p, b2p = 0.5, 0.0, 0.0, b1, b2
    pw = {'w': torch.tensor([0.5], dtype=torch.float64)}
    so = nn_ref.OptState(pw, spec, torch.float64)
    for g in (0.3, -0.7, 0.05):
        m = b1 * m + (1 - b1) * g; v = b2 * v + (1 - b2) * g * g
        w -= lr * np.sqrt(1 - b2p) / (1 - b1p) * m / (np.sqrt(v) + eps)
        b1p *= b1; b2p *= b2
        nn_ref.apply_optimizer(pw, {'w': torch.tensor([g], dtype=torch.float64)}, so, spec, torch.float64)
        assert abs(float(pw['w']) - w) < 1e-15


def test_nn_oracle_nesterov_momentum_matches_documented_recurrence():
    spec = nn_ref.NetSpec(13, 10, 4, 3, num_hidden=16, use_momentum=True)
    f32 = lambda v: float(np.float32(v))
    lr, mom = f32(0.00025), f32(0.95)
    pw = {'w': torch.tensor([0.5], dtype=torch.float64)}
    so = nn_ref.OptState(pw, spec, torch.float64)
    w, acc = 0.5, 0.0
    for g in (0.3, -0.7, 0.05):
        acc = acc * mom + g            # ApplyMomentum: accum = accum * momentum + grad
        w -= g * lr + acc * mom * lr   # use_nesterov: var -= grad * lr + accum * momentum * lr
        nn_ref.apply_optimizer(pw, {'w': torch.tensor([g], dtype=torch.float64)}, so, spec, torch.float64)
        assert abs(float(pw['w']) - w) < 1e-15


def test_gated_oracle_equals_free_oracle_when_gates_agree():
    """nn_ref.forward(gates=...) with the oracle's own on/off pattern is the plain ReLU network (values and gradients)."""
    spec, on, tg, batch, isw = _tiny_case(True, True, True, 5)
    res = []
    for use_gates in (False, True):
        o = nn_ref.to_torch(on, torch.float64); t = nn_ref.to_torch(tg, torch.float64)
        opt = nn_ref.OptState(o, spec, torch.float64)
        gates = None
        if use_gates:
            acts = []
            with torch.no_grad():
                nn_ref.forward(o, batch['states'], spec, torch.float64, acts=acts)
            gates = [a > 0 for a in acts]
        res.append(nn_ref.train_step(o, t, opt, batch, isw, spec, torch.float64, gates=gates))
    assert np.array_equal(res[0][1], res[1][1]) and res[0][2] == res[1][2]
    for k in res[0][3]['grads']:
        assert np.array_equal(res[0][3]['grads'][k], res[1][3]['grads'][k]), k
