"""``ReplayMemoryDisk`` (rl/data/replay_memory_disk.py): the replay ring as an HDF5 file, through the hand-written
reader / writer in ``formats/hdf5.py`` instead of h5py.

Same file schema (:71-102): root attributes ``cur_idx, is_full, max_size, input_height, input_width, input_channels,
state_type`` and the datasets ``states, actions, rewards, next_states, continues, losses`` of ``max_size`` rows each; same
open semantics (:33-55): a missing file is created from the constructor's geometry, an existing one dictates its own; same
ring arithmetic (``append/extend/set/increment_idx``, :129-170,259-267), with ``cur_idx`` / ``is_full`` written through to
the file's attributes on every change (:336-355).

Differences, all on the host side of the seam: the datasets are ``numpy.memmap`` views of the file's contiguous storage, so
``extend`` copies blocks of rows instead of one row per call; the reference's row cache (``cache_size``) is accepted and
ignored -- the operating system's page cache already serves repeated reads of a mapped file.  This is the persistence format
of the agent (``_train_finish`` / ``_train_init``), not a training path: sampling runs on the HBM ring.
"""
import os

import numpy as np

from ..formats import hdf5

FIELDS = ('states', 'actions', 'rewards', 'next_states', 'continues', 'losses')


class ReplayMemoryDisk:
    MAX_SIZE = 2000000
    CACHE_SIZE = 350000
    BLOCK = 1024  # rows per copy in extend()

    def __init__(self, fn, input_height=1, input_width=1, input_channels=1, state_type='uint8', max_size=MAX_SIZE,
                 cache_size=CACHE_SIZE):
        self.filename = fn
        if not os.path.exists(fn):
            shape = (int(max_size), int(input_height), int(input_width), int(input_channels))
            hdf5.create_file(fn, {'cur_idx': 0, 'is_full': False, 'max_size': int(max_size), 'input_height': int(input_height),
                                  'input_width': int(input_width), 'input_channels': int(input_channels),
                                  'state_type': str(state_type)},
                             {'states': (shape, state_type), 'actions': (shape[:1], 'uint8'), 'rewards': (shape[:1], 'uint8'),
                              'next_states': (shape, state_type), 'continues': (shape[:1], 'bool'),
                              'losses': (shape[:1], 'float16')})
        self.data_file = hdf5.File(fn, 'r+')
        a = self.data_file.attrs
        missing = [k for k in ('cur_idx', 'is_full', 'max_size') if k not in a] + [k for k in FIELDS if k not in self.data_file.datasets]
        if missing:
            self.data_file.close()
            raise ValueError('%s is not a replay memory file: no %s' % (fn, ', '.join(missing)))
        self._cur_idx = int(a['cur_idx'])
        self._is_full = bool(a['is_full'])
        self.max_size = int(a['max_size'])
        self.input_height = int(a.get('input_height', input_height))
        self.input_width = int(a.get('input_width', input_width))
        self.input_channels = int(a.get('input_channels', input_channels))
        self.state_type = a.get('state_type', state_type)
        if isinstance(self.state_type, bytes):
            self.state_type = self.state_type.decode()
        self.cache = None
        self._arrays = {}

    # -- datasets ------------------------------------------------------------------------------------------------------
    def _array(self, name, write=False):
        key = (name, write)
        if key not in self._arrays:
            self._arrays[key] = self.data_file.array(name, 'r+' if write else 'r')
        return self._arrays[key]

    states = property(lambda self: self._array('states'))
    actions = property(lambda self: self._array('actions'))
    rewards = property(lambda self: self._array('rewards'))
    next_states = property(lambda self: self._array('next_states'))
    continues = property(lambda self: self._array('continues'))
    losses = property(lambda self: self._array('losses'))

    # -- ring bookkeeping (written through to the file) --------------------------------------------------------------------
    @property
    def cur_idx(self):
        return self._cur_idx

    @cur_idx.setter
    def cur_idx(self, idx):
        self._cur_idx = int(idx)
        self.data_file.set_attr('cur_idx', self._cur_idx)

    @property
    def is_full(self):
        return self._is_full

    @is_full.setter
    def is_full(self, val):
        self._is_full = bool(val)
        self.data_file.set_attr('is_full', self._is_full)

    @property
    def cache_size(self):
        return 0

    def __len__(self):
        return self.max_size if self._is_full else self._cur_idx

    def clear(self):
        self.cur_idx = 0
        self.is_full = False

    def increment_idx(self):
        if self.cur_idx + 1 >= self.max_size:
            self.is_full = True
        self.cur_idx = (self.cur_idx + 1) % self.max_size

    # -- rows ----------------------------------------------------------------------------------------------------------
    def set(self, idx, state=None, action=None, reward=None, next_state=None, cont=None, loss=None):
        for name, v in (('states', state), ('actions', action), ('rewards', reward), ('next_states', next_state),
                        ('continues', cont), ('losses', loss)):
            if v is not None:
                self._array(name, True)[idx] = v

    def append(self, state=None, action=None, reward=None, next_state=None, cont=None, loss=None):
        self.set(self.cur_idx, state=state, action=action, reward=reward, next_state=next_state, cont=cont, loss=loss)
        self.increment_idx()

    def extend(self, target):
        """Rows 0 .. len(target)-1 of `target` appended at cur_idx, wrapping (replay_memory_disk.py:144-151), a block of
        rows at a time."""
        n = len(target)
        for s0 in range(0, n, self.BLOCK):
            rows = np.arange(s0, min(n, s0 + self.BLOCK))
            block = {name: np.asarray(getattr(target, name)[rows]) for name in FIELDS}
            m = len(rows)
            done = 0
            while done < m:
                k = min(m - done, self.max_size - self._cur_idx)
                for name in FIELDS:
                    self._array(name, True)[self._cur_idx:self._cur_idx + k] = block[name][done:done + k]
                done += k
                if self._cur_idx + k >= self.max_size:
                    self._is_full = True
                self._cur_idx = (self._cur_idx + k) % self.max_size
        self.flush()
        self.is_full = self._is_full
        self.cur_idx = self._cur_idx

    def get_row(self, idx):
        return {'state': self.states[idx], 'action': self.actions[idx], 'reward': self.rewards[idx],
                'next_state': self.next_states[idx], 'cont': self.continues[idx], 'loss': self.losses[idx]}

    get = get_row

    def copy(self, idx, target, target_idx):
        for name in FIELDS:
            getattr(target, name)[target_idx] = getattr(self, name)[idx]

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            stop = self.max_size - 1 if idx.stop is None else idx.stop  # the reference's slice quirk (:280-283)
            return [self.get(i) for i in range(idx.start or 0, stop, idx.step or 1)]
        return self.get(idx)

    def __iter__(self):
        for i in range(len(self)):
            yield self.get(i)

    # -- lifecycle -------------------------------------------------------------------------------------------------------
    def flush(self):
        for a in self._arrays.values():
            if isinstance(a, np.memmap):
                a.flush()

    def close(self):
        if self.data_file is not None:
            self.flush()
            self._arrays = {}
            self.data_file.close()
            self.data_file = None

    def __enter__(self):
        return self

    def __exit__(self, ty, value, tb):
        self.close()
