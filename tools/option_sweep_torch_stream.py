"""tail_join on/off with the learner on a torch pool stream (torch.cuda.Stream(priority=-1) + dqn_set_stream), the conditions
bench.py used until the end of round 2.  Source of the last block of profiles/r02b_tail_join_sweeps.txt."""
import sys, time; sys.path.insert(0, '.')
import torch
from deepqnetwork_b200.learner import Learner
from deepqnetwork_b200.deep_q_model import init_tower_params, tower_param_shapes
H, W, A, B = 89, 80, 4, 32
L = Learner(H, W, 4, num_actions=A, use_dueling=True, use_double=True, use_per=True, batch_size=B, replay_capacity=200000,
            math_mode='tc3x', seed=7, max_rows=256)
stream = torch.cuda.Stream(priority=-1); torch.cuda.set_stream(stream); L.set_stream(stream.cuda_stream)
shapes = tower_param_shapes(H, W, 4, A, 512, True, 'reference')
L.set_tower(init_tower_params(shapes, 42), 0); L.set_tower(init_tower_params(shapes, 43), 1)
L.replay_fill_synthetic(200000, seed=1234); L.sync()
for tj in (0, 1, 0, 1):
    L.set_option('tail_join', tj)
    L.train_steps(150); L.sync(); t = time.perf_counter(); L.train_steps(1500); L.sync()
    print('torch-stream tail_join=%d steps/s %.0f' % (tj, 1500 / (time.perf_counter() - t)), flush=True)
L.close()
