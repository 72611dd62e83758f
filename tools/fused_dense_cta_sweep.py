import sys, time; sys.path.insert(0,'.')  # run from the repo root
from deepqnetwork_b200.learner import Learner
from deepqnetwork_b200.deep_q_model import init_tower_params, tower_param_shapes
H,W,A,B=89,80,4,32
L = Learner(H,W,4,num_actions=A,use_dueling=True,use_double=True,use_per=True,batch_size=B,replay_capacity=1000000,math_mode='tc3x')
shapes = tower_param_shapes(H,W,4,A,512,True,'reference')
L.set_tower(init_tower_params(shapes,42),0); L.set_tower(init_tower_params(shapes,43),1)
L.replay_fill_synthetic(1000000, seed=1)
for c in [int(x) for x in sys.argv[1].split(',')]:
    L.set_option('fc_rows_ctas', c)
    L.train_steps(200); L.sync(); t=time.perf_counter(); L.train_steps(2000); L.sync(); print('ctas',c,'steps/s %.0f'%(2000/(time.perf_counter()-t)))
L.close()
