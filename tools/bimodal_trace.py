import sys, time, ctypes, csv; sys.path.insert(0,'.')  # run from the repo root
T = ctypes.CDLL('tools/libcupti_trace.so')
from deepqnetwork_b200.learner import Learner
from deepqnetwork_b200.deep_q_model import init_tower_params, tower_param_shapes
H,W,A,B=89,80,4,32
cap=200000
shapes = tower_param_shapes(H,W,4,A,512,True,'reference')
seen=set()
for rep in range(8):
    L = Learner(H,W,4,num_actions=A,use_dueling=True,use_double=True,use_per=True,batch_size=B,replay_capacity=cap,math_mode='tc3x')
    L.set_tower(init_tower_params(shapes,42),0); L.set_tower(init_tower_params(shapes,43),1)
    L.replay_fill_synthetic(cap, seed=1)
    L.train_steps(200); L.sync(); t=time.perf_counter(); L.train_steps(1500); L.sync(); v=round(1500/(time.perf_counter()-t))
    mode='slow' if v<5800 else 'fast'
    print('learner',rep,v,mode,flush=True)
    if mode not in seen:
        seen.add(mode)
        assert T.trace_begin()==0
        L.train_steps(12); L.sync()
        T.trace_end(('gpurun_out/trace_%s.csv'%mode).encode())
        rows=list(csv.DictReader(open('gpurun_out/trace_%s.csv'%mode)))
        rows.sort(key=lambda r:int(r['start_ns']))
        ends=[i for i,r in enumerate(rows) if 'optimizer_tail' in r['name']]
        a=ends[5]; b=ends[6]; t0=int(rows[a]['end_ns'])
        print(mode,'step wall us',(int(rows[b]['end_ns'])-int(rows[a]['end_ns']))/1000)
        for r in rows[a+1:b+1]:
            n=r['name']; n=n[:n.find('(')] if '(' in n else n
            print('  %7.1f %7.1f %6.1f s%-3s g%-5s %s'%((int(r['start_ns'])-t0)/1000,(int(r['end_ns'])-t0)/1000,(int(r['end_ns'])-int(r['start_ns']))/1000,r['stream'],r['grid'],n[:44]))
    L.close()
    if len(seen)==2: break
