"""Python handle over the C ABI: one Learner = one dqn_learner (params + replay ring + sum tree in HBM).

The reference keeps the model (rl/deep_q_model.py), the replay ring (rl/data/replay_memory.py) and the sum
tree (rl/data/sum_tree.py) as three host objects; here they are three views of one device-resident handle so
that a training step never leaves the GPU.  The mirror classes in deep_q_model.py / data/*.py share a Learner.
"""
import ctypes as C

import os

import numpy as np

from . import _capi
from ._capi import check, ptr, as_u8


class Learner:
    def __init__(self, height, width, channels=4, num_actions=4, num_hidden=512, use_dueling=False,
                 use_double=False, use_per=False, use_momentum=False, learning_rate=0.00025, momentum=0.95,
                 discount_rate=0.99, batch_size=32, replay_capacity=10000, per_a=0.6, per_b_start=0.4,
                 per_b_end=1.0, per_anneal_steps=2000000, use_per_annealing=False, per_calculate_steps=5000,
                 copy_network_steps=10000, max_rows=0, conv_variant='reference', math_mode='tc3x', seed=7,
                 device=0, frame_dedup=False, frame_capacity=0):
        self._lib = _capi.lib()
        cfg = _capi.DqnConfig()
        cfg.struct_size = C.sizeof(_capi.DqnConfig)
        cfg.device = device
        cfg.input_height, cfg.input_width, cfg.input_channels = height, width, channels
        cfg.num_actions, cfg.num_hidden = num_actions, num_hidden
        cfg.conv_variant = {'reference': _capi.DQN_CONV_REFERENCE, 'nature': _capi.DQN_CONV_NATURE}[conv_variant]
        cfg.use_dueling, cfg.use_double = int(bool(use_dueling)), int(bool(use_double))
        cfg.use_per, cfg.use_momentum = int(bool(use_per)), int(bool(use_momentum))
        cfg.learning_rate, cfg.momentum, cfg.discount_rate = learning_rate, momentum, discount_rate
        cfg.batch_size, cfg.replay_capacity = batch_size, replay_capacity
        cfg.per_a, cfg.per_b_start, cfg.per_b_end = per_a, per_b_start, per_b_end
        cfg.per_anneal_steps, cfg.use_per_annealing = per_anneal_steps, int(bool(use_per_annealing))
        cfg.per_calculate_steps, cfg.copy_network_steps = per_calculate_steps, copy_network_steps
        cfg.max_rows = max_rows
        cfg.math_mode = {'fp32': _capi.DQN_MATH_FP32, 'tc3x': _capi.DQN_MATH_TC3X, 'tc1x': _capi.DQN_MATH_TC1X}[math_mode]
        cfg.seed = seed
        cfg.frame_dedup, cfg.frame_capacity = int(bool(frame_dedup)), int(frame_capacity)
        self.frame_dedup = bool(frame_dedup)
        self.cfg = cfg
        self.height, self.width, self.channels = height, width, channels
        self.num_actions, self.batch_size, self.capacity = num_actions, batch_size, replay_capacity
        self.use_per = bool(use_per)
        self.state_shape = (height, width, channels)
        self.state_bytes = height * width * channels
        h = C.c_void_p()
        rc = self._lib.dqn_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise _capi.DqnError((self._lib.dqn_last_error(None) or b'').decode())
        self._h = h
        self._tensors = None

    # -- lifecycle -------------------------------------------------------------
    def close(self):
        if getattr(self, '_h', None):
            self._lib.dqn_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _c(self, rc):
        check(self._h, rc)

    def sync(self):
        self._c(self._lib.dqn_sync(self._h))

    def set_stream(self, cuda_stream_ptr):
        self._c(self._lib.dqn_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0)))

    # -- parameters ------------------------------------------------------------
    def tensors(self):
        """[(name, count, offset)] in TF creation order."""
        if self._tensors is None:
            n = C.c_int32(); tot = C.c_int64()
            self._c(self._lib.dqn_num_params(self._h, C.byref(n), C.byref(tot)))
            out = []
            for i in range(n.value):
                name = C.c_char_p(); cnt = C.c_int64(); off = C.c_int64()
                self._c(self._lib.dqn_param_info(self._h, i, C.byref(name), C.byref(cnt), C.byref(off)))
                out.append((name.value.decode(), cnt.value, off.value))
            self._tensors = out
            self.num_params = tot.value
        return self._tensors

    def get_param(self, name, tower=0, slot=0):
        cnt = dict((n, c) for n, c, _ in self.tensors())[name]
        out = np.empty(cnt, np.float32)
        self._c(self._lib.dqn_get_param(self._h, tower, slot, name.encode(), ptr(out), cnt))
        return out

    def set_param(self, name, value, tower=0, slot=0):
        v = np.ascontiguousarray(value, dtype=np.float32).reshape(-1)
        self._c(self._lib.dqn_set_param(self._h, tower, slot, name.encode(), ptr(v), v.size))

    def set_tower(self, params, tower=0):
        for name, _, _ in self.tensors():
            self.set_param(name, params[name], tower=tower)

    def get_tower(self, tower=0, slot=0, shapes=None):
        out = {}
        for name, _, _ in self.tensors():
            v = self.get_param(name, tower=tower, slot=slot)
            out[name] = v.reshape(shapes[name]) if shapes else v
        return out

    def device_ptr(self, what):
        p = C.c_void_p(); b = C.c_int64()
        self._c(self._lib.dqn_device_ptr(self._h, what.encode(), C.byref(p), C.byref(b)))
        return p.value, b.value

    def get_step(self):
        s = C.c_int64()
        self._c(self._lib.dqn_get_step(self._h, C.byref(s)))
        return s.value

    def set_step(self, step):
        self._c(self._lib.dqn_set_step(self._h, int(step)))

    def get_game_count(self):
        s = C.c_int64()
        self._c(self._lib.dqn_get_game_count(self._h, C.byref(s)))
        return s.value

    def set_game_count(self, n):
        self._c(self._lib.dqn_set_game_count(self._h, int(n)))

    def get_beta_powers(self):
        a = C.c_float(); b = C.c_float()
        self._c(self._lib.dqn_get_beta_powers(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_beta_powers(self, b1p, b2p):
        """train/beta1_power, train/beta2_power of tf.train.AdamOptimizer (checkpoint restore)."""
        self._c(self._lib.dqn_set_beta_powers(self._h, float(b1p), float(b2p)))

    def copy_network(self):
        self._c(self._lib.dqn_copy_network(self._h))

    def get_stats(self):
        out = np.zeros(8, np.float32)
        self._c(self._lib.dqn_get_stats(self._h, ptr(out)))
        return dict(zip(('min', 'max', 'avg', 'total', 'last_max_weight', 'per_b', 'loss', 'loss_sum'), out.tolist()))

    # -- replay ring -------------------------------------------------------------
    def replay_append(self, states, actions, rewards, next_states, continues, losses=None):
        st = as_u8(states).reshape((-1,) + self.state_shape)
        n = st.shape[0]
        ns = as_u8(next_states).reshape((-1,) + self.state_shape)
        ac = as_u8(np.atleast_1d(actions)); rw = as_u8(np.atleast_1d(rewards)); ct = as_u8(np.atleast_1d(continues))
        assert ns.shape[0] == n and ac.size == n and rw.size == n and ct.size == n
        ls = None if losses is None else np.ascontiguousarray(np.atleast_1d(losses), dtype=np.float32)
        self._c(self._lib.dqn_replay_append(self._h, n, ptr(st), ptr(ac), ptr(rw), ptr(ns), ptr(ct), ptr(ls)))

    def frames_append(self, frames):
        """frame_dedup layout: store n new [H,W] uint8 frames; returns their slots in the frame ring."""
        f = as_u8(np.asarray(frames)).reshape((-1, self.height, self.width))
        first = C.c_int64()
        self._c(self._lib.dqn_frames_append(self._h, f.shape[0], ptr(f), C.byref(first)))
        cap, _ = self.frames_info()
        return (first.value + np.arange(f.shape[0], dtype=np.int64)) % cap

    def frames_info(self):
        c = C.c_int64(); w = C.c_int64()
        self._c(self._lib.dqn_frames_info(self._h, C.byref(c), C.byref(w)))
        return c.value, w.value

    def replay_append_indexed(self, frame_slots, actions, rewards, continues, losses=None):
        """frame_dedup layout: n transitions whose s / s' stacks are the given frame slots ([n, 8] = s0..s3, s'0..s'3)."""
        sl = np.ascontiguousarray(frame_slots, dtype=np.uint32).reshape(-1, 8)
        n = sl.shape[0]
        ac = as_u8(np.atleast_1d(actions)); rw = as_u8(np.atleast_1d(rewards)); ct = as_u8(np.atleast_1d(continues))
        assert ac.size == n and rw.size == n and ct.size == n
        ls = None if losses is None else np.ascontiguousarray(np.atleast_1d(losses), dtype=np.float32)
        self._c(self._lib.dqn_replay_append_indexed(self._h, n, ptr(sl), ptr(ac), ptr(rw), ptr(ct), ptr(ls)))

    PREPROCESS = {'Breakout': (16, 16), 'SpaceInvaders': (20, 12), 'MsPacman': (0, 38)}  # rows cropped top / bottom

    def preprocess(self, obs, game=None, top=None, bottom=None, stride=2, to_frame_ring=False):
        """GameEnvironment.preprocess_observation for a batch of raw RGB observations [n, raw_h, raw_w, 3] -> uint8 frames
        [n, oh, ow] (and, with to_frame_ring on a frame_dedup learner, their slots in the frame ring)."""
        o = as_u8(np.asarray(obs))
        o = o.reshape((-1,) + o.shape[-3:])
        if game is not None:
            top, bottom = self.PREPROCESS[game]
        n, rh, rw, _ = o.shape
        oh, ow = -(-(rh - bottom - top) // stride), -(-rw // stride)
        out = np.empty((n, oh, ow), np.uint8)
        first = C.c_int64(-1)
        self._c(self._lib.dqn_preprocess(self._h, n, ptr(o), rh, rw, int(top), int(bottom), int(stride), ptr(out), int(bool(to_frame_ring)), C.byref(first)))
        if to_frame_ring:
            cap, _ = self.frames_info()
            return out, (first.value + np.arange(n, dtype=np.int64)) % cap
        return out

    def replay_read(self, rows, fields=('states', 'actions', 'rewards', 'next_states', 'continues', 'losses')):
        rows = np.ascontiguousarray(np.atleast_1d(rows), dtype=np.int64)
        n = rows.size
        out = {}
        if 'states' in fields: out['states'] = np.empty((n,) + self.state_shape, np.uint8)
        if 'next_states' in fields: out['next_states'] = np.empty((n,) + self.state_shape, np.uint8)
        if 'actions' in fields: out['actions'] = np.empty(n, np.uint8)
        if 'rewards' in fields: out['rewards'] = np.empty(n, np.uint8)
        if 'continues' in fields: out['continues'] = np.empty(n, np.uint8)
        if 'losses' in fields: out['losses'] = np.empty(n, np.float16)
        self._c(self._lib.dqn_replay_read(self._h, n, ptr(rows), ptr(out.get('states')), ptr(out.get('actions')),
                                          ptr(out.get('rewards')), ptr(out.get('next_states')),
                                          ptr(out.get('continues')), ptr(out.get('losses'))))
        if 'continues' in out:
            out['continues'] = out['continues'].view(np.bool_)
        return out

    def replay_info(self):
        s = C.c_int64(); c = C.c_int64(); f = C.c_int32()
        self._c(self._lib.dqn_replay_info(self._h, C.byref(s), C.byref(c), C.byref(f)))
        return s.value, c.value, bool(f.value)

    def replay_clear(self):
        self._c(self._lib.dqn_replay_clear(self._h))

    def replay_fill_synthetic(self, n, seed=1234):
        self._c(self._lib.dqn_replay_fill_synthetic(self._h, int(n), int(seed)))

    # -- sum tree ------------------------------------------------------------------
    def tree_add(self, scores, data=None):
        s = np.ascontiguousarray(np.atleast_1d(scores), dtype=np.float32)
        d = None if data is None else np.ascontiguousarray(np.atleast_1d(data), dtype=np.uint32)
        self._c(self._lib.dqn_tree_add(self._h, s.size, ptr(s), ptr(d)))

    def tree_update(self, tree_idx, scores, store_losses=False):
        t = np.ascontiguousarray(np.atleast_1d(tree_idx), dtype=np.int64)
        s = np.ascontiguousarray(np.atleast_1d(scores), dtype=np.float32)
        assert t.size == s.size
        self._c(self._lib.dqn_tree_update(self._h, t.size, ptr(t), ptr(s), int(store_losses)))

    def tree_sample(self, uniforms):
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        b = u.size
        t = np.empty(b, np.int64); d = np.empty(b, np.int64); p = np.empty(b, np.float64); m = np.empty(b, np.int64)
        self._c(self._lib.dqn_tree_sample(self._h, b, ptr(u), ptr(t), ptr(d), ptr(p), ptr(m)))
        return t, d, p, m

    def tree_retrieve(self, scores):
        """SumTree.get_with_info for caller-computed probes -> (tree_idx, data_idx, priority, mem_idx) arrays."""
        sc = np.ascontiguousarray(np.atleast_1d(scores), dtype=np.float64)
        b = sc.size
        t = np.empty(b, np.int64); d = np.empty(b, np.int64); p = np.empty(b, np.float64); m = np.empty(b, np.int64)
        self._c(self._lib.dqn_tree_retrieve(self._h, b, ptr(sc), ptr(t), ptr(d), ptr(p), ptr(m)))
        return t, d, p, m

    def tree_stats(self):
        v = [C.c_float() for _ in range(4)]
        self._c(self._lib.dqn_tree_stats(self._h, *[C.byref(x) for x in v]))
        return tuple(np.float32(x.value) for x in v)  # min, max, avg, total

    def tree_read(self):
        tree = np.empty(2 * self.capacity - 1, np.float32); data = np.empty(self.capacity, np.uint32)
        s = C.c_int64(); w = C.c_int64()
        self._c(self._lib.dqn_tree_read(self._h, ptr(tree), ptr(data), C.byref(s), C.byref(w)))
        return tree, data, s.value, w.value

    def tree_info(self, with_total=False):
        """(size, write[, total]) -- host bookkeeping plus, on request, the 4-byte root; the tree is not moved."""
        s = C.c_int64(); w = C.c_int64(); t = C.c_float()
        self._c(self._lib.dqn_tree_info(self._h, C.byref(s), C.byref(w), C.byref(t) if with_total else None))
        return (s.value, w.value, np.float32(t.value)) if with_total else (s.value, w.value)

    def tree_get(self, tree_idx):
        """(scores f32[n], data u32[n]) of individual tree nodes."""
        t = np.ascontiguousarray(np.atleast_1d(tree_idx), dtype=np.int64)
        sc = np.empty(t.size, np.float32); d = np.empty(t.size, np.uint32)
        self._c(self._lib.dqn_tree_get(self._h, t.size, ptr(t), ptr(sc), ptr(d)))
        return sc, d

    # -- samplers ---------------------------------------------------------------------
    def per_sample(self, uniforms, refresh_stats, per_b):
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        assert u.size == self.batch_size
        t = np.empty(u.size, np.int64); p = np.empty(u.size, np.float64); w = np.empty(u.size, np.float64)
        self._c(self._lib.dqn_per_sample(self._h, ptr(u), int(refresh_stats), float(per_b), ptr(t), ptr(p), ptr(w)))
        return t, p, w

    def uniform_sample(self, rows):
        r = np.ascontiguousarray(rows, dtype=np.int64)
        assert r.size == self.batch_size
        self._c(self._lib.dqn_uniform_sample(self._h, ptr(r)))

    def batch_read(self):
        b = self.batch_size
        out = dict(states=np.empty((b,) + self.state_shape, np.uint8), next_states=np.empty((b,) + self.state_shape, np.uint8),
                   actions=np.empty(b, np.uint8), rewards=np.empty(b, np.uint8), continues=np.empty(b, np.uint8),
                   losses=np.empty(b, np.float16))
        self._c(self._lib.dqn_batch_read(self._h, ptr(out['states']), ptr(out['actions']), ptr(out['rewards']),
                                         ptr(out['next_states']), ptr(out['continues']), ptr(out['losses'])))
        out['continues'] = out['continues'].view(np.bool_)
        return out

    # -- model ---------------------------------------------------------------------------
    def train_batch(self, states, actions, rewards, continues, next_states, is_weights=None):
        st = as_u8(states).reshape((-1,) + self.state_shape); n = st.shape[0]
        ns = as_u8(next_states).reshape((-1,) + self.state_shape)
        ac, rw, ct = as_u8(actions), as_u8(rewards), as_u8(continues)
        w = None if is_weights is None else np.ascontiguousarray(is_weights, dtype=np.float32)
        step = C.c_int64(); loss = C.c_float(); losses = np.empty(n, np.float32)
        self._c(self._lib.dqn_train_batch(self._h, n, ptr(st), ptr(ac), ptr(rw), ptr(ct), ptr(ns), ptr(w),
                                          C.byref(step), ptr(losses), C.byref(loss)))
        return step.value, losses, np.float32(loss.value)

    def predict(self, states, use_target=False):
        st = as_u8(np.asarray(states)).reshape((-1,) + self.state_shape); n = st.shape[0]
        q = np.empty((n, self.num_actions), np.float32)
        self._c(self._lib.dqn_predict(self._h, n, ptr(st), int(bool(use_target)), ptr(q)))
        return q

    def get_losses(self, states, actions, rewards, continues, next_states):
        st = as_u8(states).reshape((-1,) + self.state_shape); n = st.shape[0]
        ns = as_u8(next_states).reshape((-1,) + self.state_shape)
        losses = np.empty(n, np.float32)
        self._c(self._lib.dqn_get_losses(self._h, n, ptr(st), ptr(as_u8(actions)), ptr(as_u8(rewards)),
                                         ptr(as_u8(continues)), ptr(ns), ptr(losses)))
        return losses

    def debug_read(self, what, shape):
        out = np.empty(shape, np.float32)
        self._c(self._lib.dqn_debug_read(self._h, what.encode(), ptr(out), out.size))
        return out

    # -- fused step -------------------------------------------------------------------------
    def train_step(self, uniforms=None, rows=None, fetch=True):
        """One DeepQAgent._train_run_train on the device.  uniforms/rows: the reference's host draws (parity)."""
        u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64)
        r = None if rows is None else np.ascontiguousarray(rows, dtype=np.int64)
        if not fetch:
            self._c(self._lib.dqn_train_step(self._h, ptr(u), ptr(r), None, None, None))
            return None
        step = C.c_int64(); loss = C.c_float(); losses = np.empty(self.batch_size, np.float32)
        self._c(self._lib.dqn_train_step(self._h, ptr(u), ptr(r), C.byref(step), ptr(losses), C.byref(loss)))
        return step.value, losses, np.float32(loss.value)

    def train_step_async(self, uniforms=None, rows=None):
        """Queue one step with host draws (copied before the call returns); at most two in flight."""
        u = None if uniforms is None else np.ascontiguousarray(uniforms, dtype=np.float64)
        r = None if rows is None else np.ascontiguousarray(rows, dtype=np.int64)
        self._c(self._lib.dqn_train_step_async(self._h, ptr(u), ptr(r)))

    def train_step_result(self):
        """(step, losses[B], loss) of the oldest queued step that has not been read yet; blocks until it has finished."""
        step = C.c_int64(); loss = C.c_float(); losses = np.empty(self.batch_size, np.float32)
        self._c(self._lib.dqn_train_step_result(self._h, C.byref(step), ptr(losses), C.byref(loss)))
        return step.value, losses, np.float32(loss.value)

    def train_steps(self, n):
        self._c(self._lib.dqn_train_steps(self._h, int(n)))

    def train_step_phase(self, phase, grad_scale=1.0):
        self._c(self._lib.dqn_train_step_phase(self._h, int(phase), float(grad_scale)))

    def phase_wait(self, what, cuda_stream_ptr):
        """Make a CUDA stream wait for 'fc_grads' / 'grads' of the phase 0 just issued."""
        self._c(self._lib.dqn_phase_wait(self._h, what.encode(), C.c_void_p(cuda_stream_ptr)))

    def set_per_normaliser(self, last_max_weight):
        self._c(self._lib.dqn_set_per_normaliser(self._h, float(last_max_weight)))

    def set_option(self, name, value):
        self._c(self._lib.dqn_set_option(self._h, name.encode(), int(value)))

    def launch_count(self):
        n = C.c_int64()
        self._c(self._lib.dqn_launch_count(self._h, C.byref(n)))
        return n.value

    def set_profile(self, enabled):
        self._c(self._lib.dqn_set_profile(self._h, int(bool(enabled))))

    def get_profile(self):
        cap = 32
        names = (C.c_char_p * cap)(); ms = (C.c_float * cap)(); n = C.c_int32(); steps = C.c_int64()
        self._c(self._lib.dqn_get_profile(self._h, C.byref(n), names, ms, cap, C.byref(steps)))
        return {names[i].decode(): ms[i] for i in range(n.value)}, steps.value

    def selftest_gemm(self, mode, A, B, a_kcontig=True, b_kcontig=True):
        """C = A @ B.T-style product through the chosen GEMM engine; A:[M,K], B:[N,K] logical (host float32)."""
        M, K = A.shape; N = B.shape[0]
        a = np.ascontiguousarray(A if a_kcontig else A.T, np.float32)
        b = np.ascontiguousarray(B if b_kcontig else B.T, np.float32)
        c = np.empty((M, N), np.float32)
        m = {'fp32': 0, 'tc3x': 1, 'tc1x': 2}[mode]
        self._c(self._lib.dqn_selftest_gemm(self._h, m, M, N, K, int(a_kcontig), int(b_kcontig), ptr(a), ptr(b), ptr(c)))
        return c

    def bench_gemm(self, mode, M, N, K, a_kcontig=True, b_kcontig=True, iters=20):
        ms = C.c_float()
        m = {'fp32': 0, 'tc3x': 1, 'tc1x': 2}[mode]
        self._c(self._lib.dqn_bench_gemm(self._h, m, M, N, K, int(a_kcontig), int(b_kcontig), iters, C.byref(ms)))
        return ms.value

    def debug_timeline(self, mode, M, N, K, a_kcontig=True, b_kcontig=True):
        m = mode if isinstance(mode, int) else {'fp32': 0, 'tc3x': 1, 'tc1x': 2}[mode]
        out = np.zeros(1024 if m >= 20 else 512, np.uint64)  # modes 20-23: fused conv tower / conv weight gradients, 128 rows of marks
        self._c(self._lib.dqn_debug_timeline(self._h, m, M, N, K, int(a_kcontig), int(b_kcontig), ptr(out)))
        return out.reshape(-1, 8)

    # -- persistence ---------------------------------------------------------------------------
    def save(self, prefix, fmt='native', shapes=None, use_momentum=False):
        """fmt 'native': <prefix>.dqnb200 (dqn_save).  fmt 'reference': the files tf.train.Saver().save leaves behind
        (<prefix>.index, <prefix>.data-00000-of-00001, `checkpoint`; formats/tf_bundle.py), which needs the tower's
        `shapes` ({variable name: shape}, deep_q_model.tower_param_shapes)."""
        if fmt == 'reference':
            from .formats import tf_bundle
            tf_bundle.save_model(self, str(prefix), shapes, use_momentum)
            return
        self._c(self._lib.dqn_save(self._h, str(prefix).encode()))

    def restore(self, prefix, fmt='native', shapes=None, use_momentum=False):
        if fmt == 'reference':
            from .formats import tf_bundle
            if not os.path.exists(str(prefix) + '.index'):
                return False
            return tf_bundle.restore_model(self, str(prefix), shapes, use_momentum)
        rc = self._lib.dqn_restore(self._h, str(prefix).encode())
        if rc == 3:
            return False
        self._c(rc)
        return True

    def save_replay(self, prefix):
        self._c(self._lib.dqn_save_replay(self._h, str(prefix).encode()))

    def restore_replay(self, prefix):
        """-> number of rows loaded, or None when <prefix>_replay_memory.dqnb200 does not exist."""
        n = C.c_int64()
        rc = self._lib.dqn_restore_replay(self._h, str(prefix).encode(), C.byref(n))
        if rc == 3:
            return None
        self._c(rc)
        return n.value
