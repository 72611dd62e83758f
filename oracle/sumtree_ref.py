"""NumPy restatement of the reference sum tree and samplers.  TEST INFRASTRUCTURE.

Follows (file:line into /root/reference):
  rl/data/sum_tree.py:14-21     constructor (2N-1 f32 heap, u32 data, size/write)
  rl/data/sum_tree.py:23-42     add
  rl/data/sum_tree.py:45-52     update_score
  rl/data/sum_tree.py:101-110   _propagate  (recursive there, a loop here)
  rl/data/sum_tree.py:113-129   _retrieve   (recursive there, a loop here)
  rl/data/sum_tree.py:76-91     get_with_info / get_data_idx
  rl/data/sum_tree.py:139-165   total / get_min / get_max / get_average
  rl/data/replay_sampler_priority.py:27-43   stratified proportional sampling
  rl/data/replay_sampler_priority.py:46-50   update_sum_tree
  rl/data/replay_sampler.py:21-32            stratified uniform sampling

Pinned by tests/test_oracle_replay.py against tests/golden/replay_*.npz, which
oracle/make_golden.py produced by running the reference's own modules.

Promotion switch
----------------
The in-container reference run (NumPy 2, NEP 50) evaluates the probe
``s = random.random()*size + i*size`` and the descent compare/subtract in
float32; the lock-file NumPy 1.16 (value-based casting) evaluates them in
float64.  ``promotion='numpy2_f32'`` is the pinned behaviour;
``'legacy_f64'`` is the documented alternative.
"""
import numpy as np

F32 = np.float32


class SumTreeRef:
    def __init__(self, capacity, promotion='numpy2_f32'):
        assert promotion in ('numpy2_f32', 'legacy_f64')
        self.capacity = int(capacity)
        self.promotion = promotion
        self.tree = np.zeros(2 * self.capacity - 1, dtype=np.float32)
        self.data = np.zeros(self.capacity, dtype=np.uint32)
        self.size = 0
        self.max_reached = False
        self.write = 0

    # -- writes -----------------------------------------------------------
    def add(self, score, data):
        idx = self.write + self.capacity - 1
        last = self.write
        self.data[self.write] = data
        self.update_score(idx, score)
        if not self.max_reached:
            self.size += 1
        self.write += 1
        if self.write >= self.capacity:
            self.write = 0
            self.max_reached = True
        return last

    def update_score(self, tree_idx, score):
        # scores reaching here are f32 (or f16) array elements in the
        # reference's call sites, so the subtraction is an f32 subtraction.
        score = F32(score)
        tree = self.tree
        idx = int(tree_idx)
        change = F32(score - tree[idx])
        tree[idx] = score
        while True:
            idx = (idx - 1) // 2
            tree[idx] = F32(tree[idx] + change)
            if idx == 0:
                break

    # -- reads ------------------------------------------------------------
    def get_data_idx(self, tree_idx):
        return tree_idx - self.capacity + 1

    def retrieve(self, score):
        tree = self.tree
        n = tree.shape[0]
        f64 = self.promotion == 'legacy_f64'
        score = np.float64(score) if f64 else F32(score)
        idx = 0
        while True:
            left = 2 * idx + 1
            if left >= n:
                over = self.get_data_idx(idx) - self.size + 1
                if over > 0:
                    idx -= over
                return idx
            lv = tree[left]
            if score <= lv:
                idx = left
            else:
                score = (np.float64(score) - np.float64(lv)) if f64 else F32(score - lv)
                idx = left + 1

    def get_with_info(self, s):
        t = self.retrieve(s)
        d = self.get_data_idx(t)
        return t, d, self.tree[t], self.data[d]

    @property
    def total(self):
        return self.tree[0]

    def _leaves(self):
        s = self.capacity - 1
        return self.tree[s:s + self.size]

    def get_min(self):
        return np.min(self._leaves())

    def get_max(self):
        return np.max(self._leaves())

    def get_average(self):
        return np.average(self._leaves())

    def __len__(self):
        return self.size


def stratified_probe(total_f32, batch_size, i, u, promotion='numpy2_f32'):
    """``s`` for stratum i given uniform draw u (replay_sampler_priority.py:30-33)."""
    if promotion == 'legacy_f64':
        size = np.float64(total_f32) / batch_size
        return np.float64(u) * size + i * size
    size = F32(F32(total_f32) / F32(batch_size))
    return F32(F32(F32(u) * size) + F32(F32(i) * size))


def per_sample(tree, batch_size, uniforms):
    """Restates ReplaySamplerPriority.sample_memories' index selection.

    Returns (tree_idx i64[B], data_idx i64[B], priority f64[B], memory_idx i64[B]).
    The row copy into the batch is ReplayRingRef.gather.
    """
    total = tree.total
    t_idx = np.zeros(batch_size, dtype=np.int64)
    d_idx = np.zeros(batch_size, dtype=np.int64)
    prio = np.zeros(batch_size, dtype=np.float64)
    m_idx = np.zeros(batch_size, dtype=np.int64)
    for i in range(batch_size):
        s = stratified_probe(total, batch_size, i, uniforms[i], tree.promotion)
        t, d, score, mem = tree.get_with_info(s)
        t_idx[i], d_idx[i], prio[i], m_idx[i] = t, d, score, mem
    return t_idx, d_idx, prio, m_idx


def per_update(tree, tree_idxes, scores_f32, losses_f16=None):
    """Restates ReplaySamplerPriority.update_sum_tree (in order, duplicates allowed)."""
    for t, p in zip(tree_idxes, scores_f32):
        tree.update_score(int(t), p)
        if losses_f16 is not None:
            losses_f16[tree.data[tree.get_data_idx(int(t))]] = p


def uniform_rows(length, batch_size, rng):
    """Restates ReplaySampler.sample_memories' index selection
    (replay_sampler.py:27-30): one inclusive randint per equal segment."""
    size = length / batch_size
    return np.array([rng.randint(int(i * size), int((i + 1) * size - 1))
                     for i in range(batch_size)], dtype=np.int64)


class ReplayRingRef:
    """Restates rl/data/replay_memory.py:17-43 (SoA ring), :54-67 (append),
    :94-110 (set), :113-122 (copy == gather of one row), :125-132 (ring index)."""

    def __init__(self, h, w, c, max_size):
        self.max_size = int(max_size)
        self.cur_idx = 0
        self.is_full = False
        self.states = np.zeros((max_size, h, w, c), np.uint8)
        self.next_states = np.zeros((max_size, h, w, c), np.uint8)
        self.actions = np.zeros(max_size, np.uint8)
        self.rewards = np.zeros(max_size, np.uint8)
        self.continues = np.zeros(max_size, np.bool_)
        self.losses = np.zeros(max_size, np.float16)

    def append(self, state, action, reward, next_state, cont, loss=None):
        i = self.cur_idx
        self.states[i] = state
        self.actions[i] = action
        self.rewards[i] = reward
        self.next_states[i] = next_state
        self.continues[i] = cont
        if loss is not None:
            self.losses[i] = loss
        if self.cur_idx + 1 >= self.max_size:
            self.is_full = True
        self.cur_idx = (self.cur_idx + 1) % self.max_size

    def __len__(self):
        return self.max_size if self.is_full else self.cur_idx

    def gather(self, rows):
        rows = np.asarray(rows, dtype=np.int64)
        return dict(states=self.states[rows], actions=self.actions[rows],
                    rewards=self.rewards[rows], next_states=self.next_states[rows],
                    continues=self.continues[rows], losses=self.losses[rows])
