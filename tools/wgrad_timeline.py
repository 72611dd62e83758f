import sys; sys.path.insert(0,'.')
import numpy as np
from deepqnetwork_b200.learner import Learner
from deepqnetwork_b200.deep_q_model import init_tower_params, tower_param_shapes
H,W,A,B=89,80,4,32
L = Learner(H,W,4,num_actions=A,use_dueling=True,use_double=True,use_per=True,batch_size=B,replay_capacity=20000,math_mode='tc3x')
shapes = tower_param_shapes(H,W,4,A,512,True,'reference')
L.set_tower(init_tower_params(shapes,42),0); L.set_tower(init_tower_params(shapes,43),1)
L.replay_fill_synthetic(20000, seed=1); L.train_steps(3); L.sync()
for layer in (int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else '0,1,2').split(',')):
    t = L.debug_timeline(21 + layer, 32, 0, 0).astype(np.int64)
    t0 = t[127,0]
    r = np.where(t>0, t-t0, -1)
    print('conv%d wgrad: start prologue pdl staged synced epi_end end' % (layer+1), r[127,:7])
    print(' epilogue tile starts', [int(r[100+i,0]) for i in range(8) if t[100+i,0] > 0])
    print(' step: CV:start CV:done | MMA:ready MMA:issued')
    for st in range(0, 40):
        if (t[st]>0).any(): print(' %2d '%st + ' '.join('%7d'%x for x in r[st,:5]))
L.close()
