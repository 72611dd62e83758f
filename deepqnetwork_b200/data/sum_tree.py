"""The reference's ``SumTree`` interface (rl/data/sum_tree.py) over the GPU heap array (csrc/sumtree.cu).

Same layout and arithmetic as the NumPy original -- 2N-1 float32 nodes, leaves at N-1.., delta propagation in
update order -- so ``tree`` read back from the device is bit-identical to the reference's array after the same
sequence of add/update calls.
"""
import numpy as np


class SumTree:
    def __init__(self, capacity, dtype='uint32', learner=None):
        if np.dtype(dtype) != np.uint32:
            raise ValueError('data slots are uint32 ring indices')
        self.capacity = int(capacity)
        self._owns = learner is None
        if learner is None:
            from ..learner import Learner
            learner = Learner(8, 8, 4, num_hidden=8, replay_capacity=self.capacity, batch_size=32, use_per=True,
                              max_rows=4096, math_mode='fp32')  # ring/tree only: the tiny network is never run
        if learner.capacity != self.capacity:
            raise ValueError('learner capacity does not match this SumTree')
        self.learner = learner

    # -- writes -------------------------------------------------------------------------
    def add(self, score, data):
        last = self.write
        self.learner.tree_add([score], [data])
        return last

    def add_many(self, scores, data=None):
        self.learner.tree_add(scores, data)

    def update_score(self, tree_idx, score):
        self.learner.tree_update([tree_idx], [score])

    def update_scores(self, tree_idxes, scores):
        """In-order batched update_score (what ReplaySamplerPriority.update_sum_tree loops over)."""
        self.learner.tree_update(tree_idxes, scores)

    # -- reads ---------------------------------------------------------------------------
    def get_data_idx(self, tree_idx):
        return tree_idx - self.capacity + 1

    def retrieve_batch(self, uniforms):
        """The B stratified probes of one sample_memories call -> (tree_idx, data_idx, score, data)."""
        return self.learner.tree_sample(uniforms)

    @property
    def tree(self):
        """The whole heap array (device -> host copy; parity / debug)."""
        return self.learner.tree_read()[0]

    @property
    def data(self):
        return self.learner.tree_read()[1]

    # size / write are host-side bookkeeping and total is the 4-byte root: none of them moves the tree
    @property
    def size(self):
        return self.learner.tree_info()[0]

    @property
    def write(self):
        return self.learner.tree_info()[1]

    @property
    def total(self):
        return self.learner.tree_info(with_total=True)[2]

    def get_min(self):
        return self.learner.tree_stats()[0]

    def get_max(self):
        return self.learner.tree_stats()[1]

    def get_average(self):
        return self.learner.tree_stats()[2]

    def get_score(self, tree_idx):
        return self.learner.tree_get([tree_idx])[0][0]

    def get_data(self, tree_idx):
        return self.learner.tree_get([tree_idx])[1][0]

    def get_with_info(self, score):
        """(tree_idx, data_idx, score, data) of the leaf a probe lands on (sum_tree.py:76-83)."""
        t, d, p, m = self.learner.tree_retrieve([score])
        return int(t[0]), int(d[0]), np.float32(p[0]), np.uint32(m[0])

    def get(self, score):
        return self.get_with_info(score)[3]

    def __len__(self):
        return self.size

    def __getitem__(self, item):
        return self.data[item]

    def close(self):
        if self._owns and self.learner is not None:
            self.learner.close()
            self.learner = None
