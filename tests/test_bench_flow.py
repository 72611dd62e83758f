"""bench.py's B200 arm walked on the CPU with the device taken away: torch.cuda and the Learner are replaced by recorders, so
that the control flow around the measurement (stream choice, warm-up / timed step counts, every e2e leg, teardown order, the
JSON line and its keys) is executed by the `-m "not gpu"` suite.  Nothing is measured here and no number from this test means
anything: it exists because a typo in bench.py would otherwise first show on the GPU box."""
import json
import os
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeLearner:
    instances = []

    def __init__(self, h, w, c, **kw):
        self.kw, self.h, self.w = kw, h, w
        self.B = kw['batch_size']
        self.step = self.launches = 0
        self.calls = []
        self.stream_set = None
        self.closed = False
        self.pending = 0
        FakeLearner.instances.append(self)

    def _advance(self, n=1):
        assert not self.closed
        self.step += n; self.launches += 22 * n

    def _result(self):
        return self.step, np.full(self.B, 0.5, np.float32), 0.25

    def device_ptr(self, what):
        self.calls.append(('device_ptr', what))
        return (0xABC0, 0) if what == 'stream' else (0x1000, 64)

    def set_stream(self, ptr): self.stream_set = ptr
    def set_option(self, k, v): self.calls.append(('set_option', k, v))
    def set_tower(self, params, tower=0): self.calls.append(('set_tower', tower))
    def replay_fill_synthetic(self, n, seed=0): self.calls.append(('fill', n))
    def sync(self): pass
    def tensors(self): return [('dense/kernel', 8, 0)]
    def train_steps(self, n): self.calls.append(('train_steps', n)); self._advance(n)
    def get_step(self): return self.step
    def launch_count(self): return self.launches
    def set_profile(self, on): self.calls.append(('profile', bool(on)))
    def get_profile(self): return {'fc_wgrad_adam': 0.047 * 20, 'conv_tower_fwd': 0.040 * 20, 'optimizer': 0.007 * 20}, 20
    def get_stats(self): return {}

    def train_step(self, uniforms=None, rows=None, fetch=True):
        assert (uniforms is None) != (rows is None)
        self._advance(); return self._result()

    def train_step_async(self, uniforms=None, rows=None):
        assert self.pending <= 1
        self.pending += 1; self._advance()

    def train_step_result(self):
        assert self.pending >= 1
        self.pending -= 1; return self._result()

    def per_sample(self, u, refresh, per_b):
        cap = self.kw['replay_capacity']
        return np.arange(self.B, dtype=np.int64) + cap - 1, np.ones(self.B), np.ones(self.B, np.float32)

    def train_batch(self, s, a, r, c, s2, w):
        assert s.shape == (self.B, self.h, self.w, 4) and s2.shape == s.shape and len(a) == self.B
        self._advance(); return self._result()

    def tree_update(self, t, p, store_losses=False): assert len(t) == len(p) == self.B
    def close(self): self.calls.append(('close', FakeCuda.current)); self.closed = True


class FakeStream:
    def __init__(self, priority=0, ptr=None):
        self.priority, self.cuda_stream, self.external = priority, ptr if ptr is not None else 0x7700, ptr is not None


class FakeEvent:
    clock = 0.0

    def __init__(self, enable_timing=False): self.t = None
    def record(self, stream=None):
        assert stream is FakeCuda.current, 'events must be recorded on the stream the learner works on'
        FakeEvent.clock += 40.0; self.t = FakeEvent.clock
    def elapsed_time(self, other): return other.t - self.t


class FakeCuda:
    current = 'default'
    log = []


@pytest.fixture()
def bench_mod(monkeypatch):
    sys.path.insert(0, ROOT)
    import bench
    import deepqnetwork_b200.build as lib_build
    import deepqnetwork_b200.learner as learner_mod
    FakeLearner.instances.clear(); FakeCuda.current = 'default'; FakeCuda.log.clear()
    monkeypatch.setattr(learner_mod, 'Learner', FakeLearner)
    monkeypatch.setattr(lib_build, 'build', lambda *a, **k: None)
    c = torch.cuda
    monkeypatch.setattr(c, 'is_available', lambda: True)
    monkeypatch.setattr(c, 'set_device', lambda d: None)
    monkeypatch.setattr(c, 'synchronize', lambda *a: None)
    monkeypatch.setattr(c, 'Stream', lambda priority=0: FakeStream(priority))
    monkeypatch.setattr(c, 'ExternalStream', lambda ptr, device=None: FakeStream(ptr=ptr))
    monkeypatch.setattr(c, 'Event', FakeEvent)
    monkeypatch.setattr(c, 'default_stream', lambda *a: 'default')
    monkeypatch.setattr(c, 'set_stream', lambda s: (FakeCuda.log.append(s), setattr(FakeCuda, 'current', s)))
    monkeypatch.setattr(torch.Tensor, 'pin_memory', lambda self, *a, **k: self)
    lines = []
    monkeypatch.setattr(bench, 'emit', lambda line: lines.append(json.loads(json.dumps(line))))
    monkeypatch.setattr(bench.os, 'dup', lambda fd: 99)
    monkeypatch.setattr(bench.os, 'dup2', lambda a, b: None)
    for k in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE'):
        monkeypatch.delenv(k, raising=False)
    return bench, lines


def _run(bench, monkeypatch, argv):
    monkeypatch.setattr(sys, 'argv', ['bench.py'] + argv)
    bench.main()


@pytest.mark.parametrize('config', ['c2', 'c1'])
def test_b200_arm_control_flow(bench_mod, monkeypatch, config):
    bench, lines = bench_mod
    _run(bench, monkeypatch, ['--steps', '60', '--warmup', '5', '--capacity', '4096', '--config', config, '--no-cpu-baseline'])
    L, = FakeLearner.instances
    assert len(lines) == 1
    line = lines[0]
    # timed on the learner's own stream: no set_stream, torch's current stream wraps the handle's, restored before close
    assert ('device_ptr', 'stream') in L.calls and L.stream_set is None
    first = FakeCuda.log[0]
    assert isinstance(first, FakeStream) and first.external and first.cuda_stream == 0xABC0
    assert FakeCuda.log[-1] == 'default' and L.calls[-1] == ('close', 'default')
    assert 'own stream' in line['config']['stream']
    # W warm-up steps, exactly K timed steps, value = K / elapsed
    runs = [c[1] for c in L.calls if c[0] == 'train_steps']
    assert runs[:2] == [5, 60]
    assert line['steps'] == 60 and line['warmup'] == 5 and line['n_gpus'] == 1
    assert line['ms_per_step'] == pytest.approx(40.0 / 60) and line['value'] == pytest.approx(60 / 0.040)
    assert line['gpu_launches'] == 22 * 60
    for key in ('metric', 'unit', 'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'clocks', 'e2e', 'roofline', 'cpu_baseline'):
        assert key in line, key
    assert line['scaling'] == 'weak' and line['vs_baseline'] is None and line['cpu_baseline'] is None
    r = line['roofline']
    assert r['kernel'] == 'fc_wgrad_adam' and r['bound'] == 'hbm' and r['us_per_launch'] == pytest.approx(47.0) and 0 < r['frac'] < 2
    e = line['e2e']
    assert e['h2d_bytes_per_step'] == 32 * 8 and e['value'] > 0 and e['synchronous']['value'] > 0 and e['host_batch']['value'] > 0
    assert L.pending == 0 and L.closed


def test_b200_arm_falls_back_to_a_torch_stream(bench_mod, monkeypatch):
    bench, lines = bench_mod

    def no_stream_query(self, what):
        raise RuntimeError('unknown device buffer: stream')
    monkeypatch.setattr(FakeLearner, 'device_ptr', no_stream_query)
    _run(bench, monkeypatch, ['--steps', '30', '--warmup', '3', '--capacity', '4096', '--no-cpu-baseline', '--no-e2e'])
    L, = FakeLearner.instances
    assert L.stream_set == 0x7700 and not FakeCuda.log[0].external and FakeCuda.log[0].priority == -1
    assert 'dqn_set_stream' in lines[0]['config']['stream'] and lines[0]['e2e'] is None
