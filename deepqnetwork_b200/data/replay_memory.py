"""Replay seam: the reference's ``ReplayMemory`` interface (rl/data/replay_memory.py) over the HBM ring.

``ReplayMemory(h, w, c, max_size, state_type, learner=...)`` is a view of the uint8 SoA ring that lives inside a
device learner (states, actions, rewards, next_states, continues, losses-f16: replay_memory.py:26-43).  Rows are
appended with one host->device copy per field; they are read back only for parity/debug -- the training step
gathers rows on the device (csrc/replay.cu).  ``host=True`` gives the small NumPy staging container the agent
uses for its 128-row insert buffer (deep_q_agent.py:243-247); it is a buffer, not a compute path.
"""
import numpy as np


class _RingField:
    """Read-only array-like over one ring column (device -> host on access)."""

    def __init__(self, mem, name):
        self._mem, self._name = mem, name

    def __getitem__(self, idx):
        mem = self._mem
        if isinstance(idx, slice):
            rows = np.arange(*idx.indices(mem.max_size))
            return mem.learner.replay_read(rows, (self._name,))[self._name]
        scalar = np.ndim(idx) == 0
        rows = np.atleast_1d(np.asarray(idx, dtype=np.int64))
        out = mem.learner.replay_read(rows, (self._name,))[self._name]
        return out[0] if scalar else out

    def __len__(self):
        return self._mem.max_size

    def __array__(self, dtype=None, copy=None):
        a = self[:]
        return a.astype(dtype) if dtype is not None else a


class _FrameDedup:
    """Host side of the frame-de-duplicated ring (dqn_config.frame_dedup): turns the stacked states the agent seam hands
    over (H x W x 4, built by GameRunner from its 4-frame deque, rl/game_runner.py:139-148,236-244) back into single
    frames.  Frames are content-addressed through a small cache of the most recent ones, so the 3 frames a stack shares
    with its predecessor -- and the s == s' transition the runner emits after a life loss -- are stored once;
    only frames not seen among the last CACHE appended ones are uploaded.  Nothing is assumed about HOW stacks overlap:
    a stack whose frames are all new simply costs four uploads, so episode and game boundaries need no special case.

    The frame ring reuses a slot after frame_capacity appends.  A monotonic queue over the live transitions tracks the
    oldest frame any of them references; overwriting a referenced slot raises instead of corrupting the replay."""
    CACHE = 16

    def __init__(self, learner, max_size):
        from collections import OrderedDict, deque
        self.learner = learner
        self.fcap = learner.frames_info()[0]
        self.cache = OrderedDict()      # frame bytes -> append count of the frame
        self.count = learner.frames_info()[1]
        self.max_size = max_size
        self.rows = learner.replay_info()[0]  # transitions in the ring (non-zero after a replay restore)
        self.window = deque()           # (row number, oldest frame count of that row), increasing in both: sliding minimum
        if self.rows:
            # restored rows: their frame references are not re-read; assume the density of a runner stream (one new frame per
            # transition plus a few per game) so that the overrun check stays meaningful for them
            self.window.append((self.rows - 1, max(0, self.count - (self.rows + self.rows // 16 + 32))))

    def append_rows(self, states, actions, rewards, next_states, continues, losses):
        n = len(actions)
        H, W = self.learner.height, self.learner.width
        new_frames, slots = [], np.empty((n, 8), np.uint32)
        count0 = self.count
        oldest = []
        for i in range(n):
            row_oldest = None
            for c in range(8):
                src = states[i] if c < 4 else next_states[i]
                key = np.ascontiguousarray(src[..., c & 3]).tobytes()
                k = self.cache.get(key)
                if k is not None and self.count + len(new_frames) - k >= self.fcap // 2:
                    k = None  # too old to be trusted to outlive this transition: store it again
                if k is None:
                    k = count0 + len(new_frames)
                    new_frames.append(np.frombuffer(key, np.uint8).reshape(H, W))
                self.cache[key] = k
                self.cache.move_to_end(key)
                while len(self.cache) > self.CACHE:
                    self.cache.popitem(last=False)
                slots[i, c] = k % self.fcap
                row_oldest = k if row_oldest is None else min(row_oldest, k)
            oldest.append(row_oldest)
        # slots about to be reused must not be referenced by a transition that is still in the ring after this append
        end = count0 + len(new_frames)
        first_live = self.rows + n - self.max_size
        while self.window and self.window[0][0] < first_live:
            self.window.popleft()
        live = [o for r, o in zip(range(self.rows, self.rows + n), oldest) if r >= first_live]
        if self.window:
            live.append(self.window[0][1])  # front of the monotonic queue = oldest frame among the earlier live rows
        live_min = min(live) if live else end
        if end - live_min > self.fcap:
            raise RuntimeError('frame ring too small: a live transition references a frame %d appends old but frame_capacity '
                               'is %d; create the learner with a larger frame_capacity' % (end - live_min, self.fcap))
        if new_frames:
            got = self.learner.frames_append(np.stack(new_frames))
            assert int(got[0]) == count0 % self.fcap
        self.count = end
        self.learner.replay_append_indexed(slots, _as_u8_store(actions), _as_u8_store(rewards), np.asarray(continues, dtype=bool), losses)
        for r, o in zip(range(self.rows, self.rows + n), oldest):
            while self.window and self.window[-1][1] >= o:
                self.window.pop()
            self.window.append((r, o))
        self.rows += n
        return len(new_frames)


class ReplayMemory:
    MAX_SIZE = 10000
    STATE_TYPE = 'uint8'

    def __init__(self, input_height, input_width, input_channels, max_size=MAX_SIZE, state_type=STATE_TYPE,
                 learner=None, host=False):
        if state_type != 'uint8':
            raise ValueError('the HBM replay ring stores uint8 frames (STATE_TYPE)')
        self.max_size = int(max_size)
        self.shape = (input_height, input_width, input_channels)
        self.host = bool(host)
        self._owns = False
        if self.host:
            self.learner = None
            self._cur, self._full = 0, False
            self.states = np.zeros((self.max_size,) + self.shape, np.uint8)
            self.next_states = np.zeros((self.max_size,) + self.shape, np.uint8)
            self.actions = np.zeros(self.max_size, np.uint8)
            self.rewards = np.zeros(self.max_size, np.uint8)
            self.continues = np.zeros(self.max_size, np.bool_)
            self.losses = np.zeros(self.max_size, np.float16)
            return
        if learner is None:
            from ..learner import Learner
            learner = Learner(input_height, input_width, input_channels, num_hidden=8, replay_capacity=self.max_size,
                              batch_size=min(32, self.max_size), math_mode='fp32')  # ring only
            self._owns = True
        if learner.capacity != self.max_size or learner.state_shape != self.shape:
            raise ValueError('learner ring geometry does not match this ReplayMemory')
        self.learner = learner
        self._dedup = _FrameDedup(learner, self.max_size) if getattr(learner, 'frame_dedup', False) else None
        self.frames_uploaded = 0
        for name in ('states', 'next_states', 'actions', 'rewards', 'continues', 'losses'):
            setattr(self, name, _RingField(self, name))

    # -- bookkeeping ----------------------------------------------------------------
    @property
    def cur_idx(self):
        return self._cur if self.host else self.learner.replay_info()[1]

    @property
    def is_full(self):
        return self._full if self.host else self.learner.replay_info()[2]

    def __len__(self):
        if self.host:
            return self.max_size if self._full else self._cur
        return self.learner.replay_info()[0]

    def clear(self):
        if self.host:
            self._cur, self._full = 0, False
        else:
            self.learner.replay_clear()

    def increment_idx(self):
        if not self.host:
            raise NotImplementedError('the device ring advances inside append()')
        if self._cur + 1 >= self.max_size:
            self._full = True
        self._cur = (self._cur + 1) % self.max_size

    # -- writes ---------------------------------------------------------------------------
    def append(self, state=None, action=None, reward=None, next_state=None, cont=None, loss=None):
        if self.host:
            self.set(self._cur, state=state, action=action, reward=reward, next_state=next_state, cont=cont, loss=loss)
            self.increment_idx()
            return
        self.append_rows(np.asarray(state)[None], [action], [reward], np.asarray(next_state)[None], [cont],
                         None if loss is None else [loss])

    def append_rows(self, states, actions, rewards, next_states, continues, losses=None):
        """n rows with one call (the batched form of n append()s; same ring order)."""
        if self.host:
            for i in range(len(actions)):
                self.append(states[i], actions[i], rewards[i], next_states[i], continues[i],
                            None if losses is None else losses[i])
            return
        if self._dedup is not None:
            self.frames_uploaded += self._dedup.append_rows(np.asarray(states), actions, rewards, np.asarray(next_states), continues, losses)
            return
        self.learner.replay_append(states, _as_u8_store(actions), _as_u8_store(rewards), next_states,
                                   np.asarray(continues, dtype=bool), losses)

    def extend(self, target):
        n = len(target)
        if n == 0:
            return
        rows = np.arange(n)
        self.append_rows(np.asarray(target.states[rows]), np.asarray(target.actions[rows]),
                         np.asarray(target.rewards[rows]), np.asarray(target.next_states[rows]),
                         np.asarray(target.continues[rows]), np.asarray(target.losses[rows], dtype=np.float32))

    def set(self, idx, state=None, action=None, reward=None, next_state=None, cont=None, loss=None):
        if not self.host:
            raise NotImplementedError('in-place row writes on the device ring go through append()/update_sum_tree()')
        if state is not None:
            self.states[idx] = state
        if action is not None:
            self.actions[idx] = _as_u8_store(action)
        if reward is not None:
            self.rewards[idx] = _as_u8_store(reward)
        if next_state is not None:
            self.next_states[idx] = next_state
        if cont is not None:
            self.continues[idx] = cont
        if loss is not None:
            self.losses[idx] = loss

    # -- reads -------------------------------------------------------------------------------
    def get(self, idx):
        if self.host:
            return {'state': self.states[idx], 'action': self.actions[idx], 'reward': self.rewards[idx],
                    'next_state': self.next_states[idx], 'cont': self.continues[idx], 'loss': self.losses[idx]}
        r = self.learner.replay_read([idx])
        return {'state': r['states'][0], 'action': r['actions'][0], 'reward': r['rewards'][0],
                'next_state': r['next_states'][0], 'cont': r['continues'][0], 'loss': r['losses'][0]}

    def copy(self, idx, target, target_idx):
        """Row idx -> target[target_idx] (replay_memory.py:113-122).  Host-side convenience; the training step's
        gather runs on the device."""
        row = self.get(idx)
        target.set(target_idx, state=row['state'], action=row['action'], reward=row['reward'],
                   next_state=row['next_state'], cont=row['cont'], loss=row['loss'])

    def __getitem__(self, idx):
        if isinstance(idx, (int, np.integer)):
            return self.get(int(idx))
        return [self.get(i) for i in range(*idx.indices(self.max_size))]

    def __iter__(self):
        for i in range(len(self)):
            yield self.get(i)

    def close(self):
        if self._owns and self.learner is not None:
            self.learner.close()
            self.learner = None


def _as_u8_store(v):
    """NumPy's C-cast store into the reference's uint8 columns: float rewards wrap (400.0 -> 144)."""
    a = np.asarray(v)
    if a.dtype.kind == 'f':
        return (a.astype(np.int64) & 255).astype(np.uint8)
    if a.dtype.kind in 'iu' and a.dtype != np.uint8:
        return (a.astype(np.int64) & 255).astype(np.uint8)
    return a.astype(np.uint8)
