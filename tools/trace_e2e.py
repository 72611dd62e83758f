import sys, ctypes, time; sys.path.insert(0,'.')  # run from the repo root
import numpy as np, torch
T = ctypes.CDLL('tools/libcupti_trace.so')
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
def it(i):
    st,ns=hs[(2*i)%4].numpy(),hs[(2*i+1)%4].numpy()
    t,p,w=L.per_sample(rs.random(B), False, 0.4)
    step,losses,loss=L.train_batch(st,ha,hr,hc,ns,w)
    L.tree_update(t,make_priority(losses),store_losses=True)
for i in range(20): it(i)
L.sync()
assert T.trace_begin()==0
for i in range(6): it(i)
L.sync()
T.trace_end(b'gpurun_out/trace_e2e.csv')
import csv
rows=list(csv.DictReader(open('gpurun_out/trace_e2e.csv')))
rows.sort(key=lambda r:int(r['start_ns']))
idx=[i for i,r in enumerate(rows) if 'tree_sample' in r['name']]
a,b=idx[2],idx[3]
t0=int(rows[a]['start_ns'])
for r in rows[a:b+1]:
    n=r['name']; n=n[:n.find('(')] if '(' in n else n
    print('%8.1f %8.1f %6.1f  s%-3s g%-7s %s'%((int(r['start_ns'])-t0)/1000,(int(r['end_ns'])-t0)/1000,(int(r['end_ns'])-int(r['start_ns']))/1000,r['stream'],r['grid'],n[:60]))
L.close()
