"""Pin the oracle's replay/PER restatement to executions of the reference code
(fixtures written by oracle/make_golden.py).  CPU only."""
import glob
import os
import random

import numpy as np
import pytest

from oracle.sumtree_ref import (ReplayRingRef, SumTreeRef, per_sample, per_update,
                                stratified_probe, uniform_rows)

G = os.path.join(os.path.dirname(__file__), 'golden')


def load(name):
    return np.load(os.path.join(G, name))


def test_g1_uniform_rows():
    g = load('replay_g1_uniform.npz')
    rows = uniform_rows(100, 4, random.Random(0))
    assert rows.tolist() == g['rows'].tolist() == [12, 49, 63, 76]


def test_uniform_cases():
    g = load('replay_uniform_cases.npz')
    for k in range(int(g['count'])):
        length, cap, bsz, seed = g['u%d_meta' % k].tolist()
        rows = uniform_rows(length, bsz, random.Random(seed))
        assert rows.tolist() == g['u%d_rows' % k].tolist()


def test_g2_sampler_two_leaf_depths():
    g = load('per_g2_sampler.npz')
    t = SumTreeRef(10)
    for i in range(7):
        t.add(np.float32(0.1 * (i + 1)), i)
    assert np.array_equal(t.tree, g['tree0'])
    ti, di, p, mi = per_sample(t, 4, g['uniforms'])
    assert ti.tolist() == g['tree_idx'].tolist() == [15, 11, 12, 14]
    assert np.array_equal(p, g['prio'])
    assert mi.tolist() == g['rows'].tolist()
    losses = np.zeros(10, np.float16)
    for i in range(7):
        losses[i] = np.float32(0.1 * (i + 1))
    per_update(t, ti, np.array([.5, .25, .125, 1.0], np.float32), losses)
    assert np.array_equal(t.tree, g['tree1'])
    assert np.array_equal(losses, g['losses_f16'])


def test_small_cases_g3_g4_g5():
    g = load('per_small_cases.npz')
    t = SumTreeRef(8)
    for i in range(3):
        t.add(1.0, i)
    a = t.get_with_info(t.total)
    b = t.get_with_info(t.total * np.float32(1.0001))
    assert [a[0], a[1], b[0], b[1]] == g['g3'].tolist() == [9, 2, 9, 2]
    t = SumTreeRef(4)
    for i in range(1, 7):
        t.add(np.float32(i), i)
    assert np.array_equal(t.tree, g['g4_tree']) and np.array_equal(t.data, g['g4_data'])
    assert [t.size, t.write] == g['g4_sw'].tolist()
    t = SumTreeRef(1000)
    for i, s in enumerate(g['g5_scores']):
        t.add(s, i)
    for i, v in zip(g['g5_idx'], g['g5_val']):
        t.update_score(int(i), v)
    assert np.array_equal(t.tree, g['g5_tree'])
    # f32 delta-propagation drift is part of the contract: root != exact leaf sum
    assert float(t.total) != float(np.sum(t.tree[999:].astype(np.float64)))


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(G, 'per_trace_*.npz'))))
def test_per_traces(path):
    g = np.load(path)
    cap, n_add, bsz, rounds, seed = g['meta'].tolist()
    t = SumTreeRef(cap)
    for i, s in enumerate(g['add_scores']):
        t.add(s, i % cap)
    assert np.array_equal(t.tree, g['tree_after_add'])
    assert np.array_equal(t.data, g['data_after_add'])
    assert [t.size, t.write, int(t.max_reached)] == g['size_write'].tolist()
    st = g['stats_after_add']
    assert t.get_min() == st[0] and t.get_max() == st[1] and t.total == st[3]
    assert np.float32(t.get_average()) == st[2]
    for r in range(rounds):
        ti, di, p, mi = per_sample(t, bsz, g['uniforms'][r])
        assert np.array_equal(ti, g['tree_idx'][r])
        assert np.array_equal(di, g['data_idx'][r])
        assert np.array_equal(p, g['prio'][r])
        assert np.array_equal(mi, g['mem_idx'][r])
        per_update(t, ti, g['new_scores'][r])
        assert t.total == g['roots'][r]
    assert np.array_equal(t.tree, g['tree_final'])


def test_legacy_promotion_differs_only_in_rounding():
    # legacy_f64 descends in float64; it must agree with the f32 path except
    # for draws that sit within rounding distance of a node boundary.
    g = load('per_trace_07.npz')
    cap = int(g['meta'][0])
    a = SumTreeRef(cap, 'numpy2_f32'); b = SumTreeRef(cap, 'legacy_f64')
    for i, s in enumerate(g['add_scores']):
        a.add(s, i); b.add(s, i)
    assert np.array_equal(a.tree, b.tree)
    rng = random.Random(1)
    diff = 0
    for _ in range(200):
        u = [rng.random() for _ in range(32)]
        ta = per_sample(a, 32, u)[0]; tb = per_sample(b, 32, u)[0]
        diff += int((ta != tb).sum())
        assert np.all(np.abs(ta - tb) <= 1)
    assert diff < 64
    s32 = stratified_probe(a.total, 32, 3, 0.123456789, 'numpy2_f32')
    s64 = stratified_probe(a.total, 32, 3, 0.123456789, 'legacy_f64')
    assert isinstance(s32, np.float32) and isinstance(s64, np.float64)


def test_ring_gather_and_wrap():
    r = ReplayRingRef(2, 2, 1, 4)
    for i in range(6):
        r.append(np.full((2, 2, 1), i, np.uint8), i, i * 40, np.full((2, 2, 1), i + 1, np.uint8), i % 2 == 0, 0.5)
    assert len(r) == 4 and r.cur_idx == 2 and r.is_full
    b = r.gather([0, 1, 2, 3])
    assert b['states'][:, 0, 0, 0].tolist() == [4, 5, 2, 3]
    assert b['rewards'].tolist() == [160, 200, 80, 120]


@pytest.mark.parametrize('path', sorted(glob.glob(os.path.join(G, 'per_trace_*.npz'))))
def test_c_restatement_matches_reference_traces(path):
    """oracle/sumtree_ref.c (used for full-size parity runs) against the same reference executions."""
    from oracle.sumtree_c import SumTreeC
    g = np.load(path)
    cap, n_add, bsz, rounds, seed = g['meta'].tolist()
    t = SumTreeC(cap)
    t.add(g['add_scores'], (np.arange(n_add) % cap).astype(np.uint32))
    assert np.array_equal(t.tree, g['tree_after_add']) and np.array_equal(t.data, g['data_after_add'])
    assert [t.size, t.write] == g['size_write'][:2].tolist()
    for r in range(rounds):
        ti, di, p, mi = t.sample(g['uniforms'][r])
        assert np.array_equal(ti, g['tree_idx'][r]) and np.array_equal(di, g['data_idx'][r])
        assert np.array_equal(p, g['prio'][r]) and np.array_equal(mi, g['mem_idx'][r])
        t.update(ti, g['new_scores'][r])
        assert t.total == g['roots'][r]
    assert np.array_equal(t.tree, g['tree_final'])
