"""CUPTI kernel trace of one large-batch (data-parallel per-rank) split step on one GPU: python tools/trace_dp.py [rows]"""
import sys, ctypes, collections; sys.path.insert(0, '.')
import numpy as np
T = ctypes.CDLL('tools/libcupti_trace.so')
from deepqnetwork_b200.learner import Learner
from deepqnetwork_b200.deep_q_model import init_tower_params, tower_param_shapes
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = Learner(89, 80, 4, num_actions=4, use_dueling=True, use_double=True, use_per=True, batch_size=B, replay_capacity=100000, math_mode='tc3x', max_rows=B)
shapes = tower_param_shapes(89, 80, 4, 4, 512, True)
L.set_tower(init_tower_params(shapes, 42), 0); L.set_tower(init_tower_params(shapes, 43), 1)
L.replay_fill_synthetic(100000, seed=1)
for _ in range(3):
    L.train_step_phase(0, 1.0); L.train_step_phase(1, 1.0)
L.sync()
assert T.trace_begin() == 0
L.train_step_phase(0, 1.0); L.train_step_phase(1, 1.0); L.sync()
T.trace_end(b'gpurun_out/trace_dp.csv')
import csv
rows = list(csv.DictReader(open('gpurun_out/trace_dp.csv')))
rows.sort(key=lambda r: int(r['start_ns']))
t0 = int(rows[0]['start_ns'])
print('step wall us:', (max(int(r['end_ns']) for r in rows) - t0) / 1000)
for r in rows:
    n = r['name']; n = n[:n.find('(')] if '(' in n else n
    print('%8.1f %8.1f %7.1f  s%-3s g%-6s %s' % ((int(r['start_ns']) - t0) / 1000, (int(r['end_ns']) - t0) / 1000, (int(r['end_ns']) - int(r['start_ns'])) / 1000, r['stream'], r['grid'], n[:90]))
L.close()
