import sys, ctypes, collections; sys.path.insert(0,'.')  # run from the repo root
import numpy as np, torch
T = ctypes.CDLL('tools/libcupti_trace.so')
from deepqnetwork_b200.learner import Learner
from deepqnetwork_b200.deep_q_model import init_tower_params, tower_param_shapes
H,W,A,B=89,80,4,32
cap = int(sys.argv[1]) if len(sys.argv)>1 else 200000
L = Learner(H,W,4,num_actions=A,use_dueling=True,use_double=True,use_per=True,batch_size=B,replay_capacity=cap,math_mode='tc3x')
for a in sys.argv[2:]:
    k,v=a.split('='); L.set_option(k,int(v))
shapes = tower_param_shapes(H,W,4,A,512,True,'reference')
L.set_tower(init_tower_params(shapes,42),0); L.set_tower(init_tower_params(shapes,43),1)
L.replay_fill_synthetic(cap, seed=1); L.train_steps(50); L.sync()
assert T.trace_begin()==0
L.train_steps(12); L.sync()
T.trace_end(b'gpurun_out/trace.csv')
import csv
rows=list(csv.DictReader(open('gpurun_out/trace.csv')))
rows.sort(key=lambda r:int(r['start_ns']))
# pick a middle step: find tree_sample occurrences; in pipelined mode step boundary = optimizer_tail end
ends=[i for i,r in enumerate(rows) if 'optimizer_tail' in r['name']]
a=ends[5]; b=ends[6]
# step = kernels starting after pack kernels of previous... print from first kernel after ends[5]'s pack kernels
t0=int(rows[a]['end_ns'])
print('step wall us:', (int(rows[b]['end_ns'])-int(rows[a]['end_ns']))/1000)
for r in rows[a+1:b+4]:
    n=r['name']; n=n[:n.find('(')] if '(' in n else n
    n=n.replace('void ','').replace('ts::','').replace('tc::','').replace('cv::','')[:70]
    print('%8.1f %8.1f %6.1f  s%-3s g%-5s %s'%((int(r['start_ns'])-t0)/1000,(int(r['end_ns'])-t0)/1000,(int(r['end_ns'])-int(r['start_ns']))/1000,r['stream'],r['grid'],n))
L.close()
