"""Stratified uniform sampler (rl/data/replay_sampler.py:21-32): one ``random.randint`` per equal index range,
drawn on the host from the same ``random`` stream as the reference, rows gathered on the device."""
import random

import numpy as np


class ReplaySampler:
    def __init__(self, replay_memory):
        self.replay_memory = replay_memory

    def append(self, state, action, reward, next_state, cont):
        self.replay_memory.append(state, action, reward, next_state, cont)

    def draw_rows(self, batch_size):
        size = len(self) / batch_size
        return np.array([random.randint(int(i * size), int((i + 1) * size - 1)) for i in range(batch_size)], np.int64)

    def sample_memories(self, target, batch_size=32):
        """Fills the learner's device batch; `target` (a host ReplayMemory) is also filled when given."""
        rows = self.draw_rows(batch_size)
        learner = self.replay_memory.learner
        learner.uniform_sample(rows)
        if target is not None and getattr(target, 'host', False):
            _fill_host_batch(target, learner.batch_read())
        return rows

    def close(self):
        self.replay_memory.close()

    def __getitem__(self, idx):
        return self.replay_memory[idx]

    def __len__(self):
        return len(self.replay_memory)


def _fill_host_batch(target, batch):
    n = batch['actions'].shape[0]
    target.states[:n] = batch['states']; target.next_states[:n] = batch['next_states']
    target.actions[:n] = batch['actions']; target.rewards[:n] = batch['rewards']
    target.continues[:n] = batch['continues']; target.losses[:n] = batch['losses']
