"""A/B of dqn_set_option values on one C2 learner (own stream, no torch): `python tools/option_sweep.py x k=v+k2=v2 k=v ...`;
each spec sets its options (the others keep their last value), then 200 warm-up + 3000 timed fused steps (wall clock around
dqn_train_steps + dqn_sync).  Source of profiles/r02b_tail_join_sweeps.txt."""
import sys, time; sys.path.insert(0,'.')
from deepqnetwork_b200.learner import Learner
from deepqnetwork_b200.deep_q_model import init_tower_params, tower_param_shapes
H,W,A,B=89,80,4,32
L = Learner(H,W,4,num_actions=A,use_dueling=True,use_double=True,use_per=True,batch_size=B,replay_capacity=1000000,math_mode='tc3x')
shapes = tower_param_shapes(H,W,4,A,512,True,'reference')
L.set_tower(init_tower_params(shapes,42),0); L.set_tower(init_tower_params(shapes,43),1)
L.replay_fill_synthetic(1000000, seed=1)
name = sys.argv[1]
for spec in sys.argv[2:]:
    for kv in spec.split('+'):
        k, v = kv.split('='); L.set_option(k, int(v))
    L.train_steps(200); L.sync(); t=time.perf_counter(); L.train_steps(3000); L.sync(); print(spec,'steps/s %.0f'%(3000/(time.perf_counter()-t)), flush=True)
L.close()
