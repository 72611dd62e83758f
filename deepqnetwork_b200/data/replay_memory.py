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
