import sys, time; sys.path.insert(0,'.')
import numpy as np, torch
from deepqnetwork_b200.learner import Learner
from deepqnetwork_b200.deep_q_model import init_tower_params, tower_param_shapes
from deepqnetwork_b200.deep_q_agent import make_priority
H,W,A,B=89,80,4,32
L = Learner(H,W,4,num_actions=A,use_dueling=True,use_double=True,use_per=True,batch_size=B,replay_capacity=200000,math_mode='tc3x')
shapes = tower_param_shapes(H,W,4,A,512,True,'reference')
L.set_tower(init_tower_params(shapes,42),0); L.set_tower(init_tower_params(shapes,43),1)
L.replay_fill_synthetic(200000, seed=1); L.sync()
hs=[torch.randint(0,256,(B,H,W,4),dtype=torch.uint8).pin_memory() for _ in range(4)]
ha=np.zeros(B,np.uint8); hr=np.zeros(B,np.uint8); hc=np.ones(B,np.uint8)
rs=np.random.default_rng(0)
T=[0,0,0,0]
def it(i, rec):
    st,ns=hs[(2*i)%4].numpy(),hs[(2*i+1)%4].numpy()
    t0=time.perf_counter(); u=rs.random(B)
    t,p,w=L.per_sample(u, False, 0.4); t1=time.perf_counter()
    step,losses,loss=L.train_batch(st,ha,hr,hc,ns,w); t2=time.perf_counter()
    pr=make_priority(losses); t3=time.perf_counter()
    L.tree_update(t,pr,store_losses=True); t4=time.perf_counter()
    if rec:
        T[0]+=t1-t0; T[1]+=t2-t1; T[2]+=t3-t2; T[3]+=t4-t3
for i in range(20): it(i,False)
N=300
t0=time.perf_counter()
for i in range(N): it(i,True)
L.sync(); dt=time.perf_counter()-t0
print('per step us: total %.1f per_sample %.1f train_batch %.1f make_priority %.1f tree_update %.1f'%(dt/N*1e6,*[x/N*1e6 for x in T]))
# fused device step for comparison
L.train_steps(50); L.sync(); t0=time.perf_counter(); L.train_steps(300); L.sync(); print('fused us/step %.1f'%((time.perf_counter()-t0)/300*1e6))
L.close()
