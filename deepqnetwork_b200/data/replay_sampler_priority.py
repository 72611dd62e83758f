"""Prioritised sampler (rl/data/replay_sampler_priority.py): stratified proportional sampling over the GPU sum
tree, rows gathered on the device, priorities scattered back in order.  The B uniform draws come from the host
``random`` module exactly as in the reference (:33), so a seeded run selects the same leaves."""
import random

import numpy as np

from .replay_sampler import _fill_host_batch
from .sum_tree import SumTree


class ReplaySamplerPriority:
    MAX_DUPLICATE_RETRIES = 100

    def __init__(self, replay_memory):
        self.replay_memory = replay_memory
        learner = replay_memory.learner
        if not learner.use_per:
            raise ValueError('the learner behind this replay memory was created without use_per')
        self.sum_tree = SumTree(replay_memory.max_size, dtype='uint32', learner=learner)
        self.add_losses()

    def append(self, state, action, reward, next_state, cont, loss):
        # tree.add(loss, cur_idx) + ring append happen together inside dqn_replay_append
        self.replay_memory.append(state, action, reward, next_state, cont, loss)

    def append_rows(self, states, actions, rewards, next_states, continues, losses):
        self.replay_memory.append_rows(states, actions, rewards, next_states, continues, losses)

    def sample_memories(self, target, batch_size=32, priorities=None, tree_idxes=None, skip_duplicates=True,
                        per_b=None, refresh_stats=False):
        learner = self.replay_memory.learner
        if batch_size != learner.batch_size:
            raise ValueError('batch_size differs from the learner batch')
        u = np.array([random.random() for _ in range(batch_size)], np.float64)
        t, p, w = learner.per_sample(u, refresh_stats, learner.cfg.per_b_start if per_b is None else per_b)
        if priorities is not None:
            priorities[:] = p
        if tree_idxes is not None:
            tree_idxes[:] = t
        if target is not None and getattr(target, 'host', False):
            _fill_host_batch(target, learner.batch_read())
        self.last_is_weights = w
        return u

    def update_sum_tree(self, tree_idxes, losses):
        self.replay_memory.learner.tree_update(tree_idxes, np.asarray(losses, np.float32), store_losses=True)

    def add_losses(self):
        """Rebuild priorities from the stored f16 losses (restart path, :57-59)."""
        n = len(self.replay_memory)
        if n and len(self.sum_tree) == 0:
            self.sum_tree.add_many(np.asarray(self.replay_memory.losses[np.arange(n)], np.float32),
                                   np.arange(n, dtype=np.uint32))

    def get_min(self):
        return self.sum_tree.get_min()

    def get_max(self):
        return self.sum_tree.get_max()

    def get_average(self):
        return self.sum_tree.get_average()

    def close(self):
        self.replay_memory.close()

    def __getitem__(self, idx):
        return self.replay_memory[idx]

    def __len__(self):
        return len(self.replay_memory)

    @property
    def total(self):
        return self.sum_tree.total
