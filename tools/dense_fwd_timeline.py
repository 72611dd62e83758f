import sys; sys.path.insert(0,'.')  # run from the repo root
import numpy as np
from deepqnetwork_b200.learner import Learner
from deepqnetwork_b200.deep_q_model import init_tower_params, tower_param_shapes
H,W,A,B=89,80,4,32
L = Learner(H,W,4,num_actions=A,use_dueling=True,use_double=True,use_per=True,batch_size=B,replay_capacity=20000,math_mode='tc3x')
L.set_option('fc_tma_fwd',1)
shapes = tower_param_shapes(H,W,4,A,512,True,'reference')
L.set_tower(init_tower_params(shapes,42),0); L.set_tower(init_tower_params(shapes,43),1)
L.replay_fill_synthetic(20000, seed=1); L.train_steps(3); L.sync()
t = L.debug_timeline(16+3, 64, 0, 0).astype(np.int64)
t0 = t[63,0]
r = np.where(t>0, t-t0, -1)
print('special: start prologue pdl accum epi_done', r[63,:5])
print('j: TMA:issue | CV:wfull CV:done | B:start B:arrive | MMA:afull MMA:bfull MMA:issued')
for st in range(0, 30):
    if (t[st]>0).any(): print('%2d '%st + ' '.join('%7d'%x for x in r[st,:8]))
L.close()
