"""Actor-side device operations, timed on one B200 through the Python mirrors / C ABI with host
buffers, wall clock: batch-1 get_action (predict), vectorised predict, the 128-row PER insert (get_losses + append; stacked
and frame-de-duplicated layouts) and frame preprocessing.  Prints one JSON line.   python tools/actor_latency.py"""
import json, statistics, sys, time
sys.path.insert(0, '.')
import numpy as np
from deepqnetwork_b200.learner import Learner
from deepqnetwork_b200.deep_q_model import init_tower_params, tower_param_shapes
from deepqnetwork_b200.data.replay_memory import ReplayMemory
from deepqnetwork_b200.deep_q_agent import make_priority

H, W, A = 89, 80, 4
shapes = tower_param_shapes(H, W, 4, A, 512, True)
rng = np.random.default_rng(0)


def timed(fn, reps, warm=10):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e6)
    return dict(median_us=round(statistics.median(ts), 1), p90_us=round(sorted(ts)[int(0.9 * len(ts))], 1), reps=reps)


out = {}
for layout in ('stacked', 'frames'):
    L = Learner(H, W, 4, num_actions=A, use_dueling=True, use_double=True, use_per=True, batch_size=32, replay_capacity=100000,
                math_mode='tc3x', frame_dedup=layout == 'frames')
    L.set_tower(init_tower_params(shapes, 1), 0); L.set_tower(init_tower_params(shapes, 2), 1)
    mem = ReplayMemory(H, W, 4, max_size=100000, learner=L)
    if layout == 'stacked':
        for n in (1, 16, 64, 256):
            x = rng.integers(0, 256, (n, H, W, 4), dtype=np.uint8)
            out['predict_n%d' % n] = timed(lambda: L.predict(x), 200)
        x128 = rng.integers(0, 256, (128, 210, 160, 3), dtype=np.uint8)
        out['preprocess_128_obs'] = timed(lambda: L.preprocess(x128, 'Breakout'), 50)
    # the 128-row PER insert of deep_q_agent.py:467-497 on a sliding 4-frame stream (what GameRunner produces)
    frames = rng.integers(0, 256, (128 + 5, H, W), dtype=np.uint8)
    st = np.stack([np.stack(list(frames[t:t + 4]), axis=2) for t in range(128)])
    ns = np.stack([np.stack(list(frames[t + 1:t + 5]), axis=2) for t in range(128)])
    ac = rng.integers(0, A, 128).astype(np.uint8); rw = np.zeros(128, np.uint8); ct = np.ones(128, bool)

    def insert():
        losses = L.get_losses(st, ac, rw, ct, ns)
        mem.append_rows(st, ac, rw, ns, ct, make_priority(losses))

    out['per_insert_128_rows_' + layout] = timed(insert, 30, warm=3)
    out['get_losses_128_rows'] = timed(lambda: L.get_losses(st, ac, rw, ct, ns), 50)
    mem.close() if False else None
    L.close()
print(json.dumps(out))
