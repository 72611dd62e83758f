"""The -m gpu parity checks (tests/parity_util.py) exercised on the CPU: the torch fp32 restatement plays the device.
They must accept an honest fp32 step and reject a skipped update, a doubled update, a stale second-moment slot and a
wrong gradient -- the faults the round-1 tolerances (1e-4*|w| + 2.5*lr, relaxed after any gate flip) could not see."""
import copy
import types

import numpy as np
import pytest
import torch

from oracle import nn_ref
from tests import parity_util as pu


class Fp32Device:
    """Just enough of the Learner read-back surface, backed by an fp32 run of the oracle."""

    def __init__(self, spec, on, tg, momentum=False):
        self.spec = spec
        self.cfg = types.SimpleNamespace(use_double=int(spec.use_double))
        self.o = nn_ref.to_torch(on, torch.float32); self.t = nn_ref.to_torch(tg, torch.float32)
        self.opt = nn_ref.OptState(self.o, spec, torch.float32)
        self.grads, self.acts = None, None

    def train_batch(self, b, w):
        step, losses, loss, aux = nn_ref.train_step(self.o, self.t, self.opt, b, w, self.spec, torch.float32)
        self.grads, self.acts, self.dq = aux['grads'], aux['acts'], aux['dq']
        return step, losses, loss

    def get_param(self, k, tower=0, slot=0):
        src = {(0, 0): self.o, (1, 0): self.t, (0, 1): self.opt.m, (0, 2): self.opt.v}.get((tower, slot))
        if slot == 3:
            return np.asarray(self.grads[k], np.float32).reshape(-1)
        return src[k].numpy().astype(np.float32).reshape(-1)

    def debug_read(self, name, shape):
        if name == 'dq':
            return np.asarray(self.dq, np.float32).reshape(shape)
        a = self.acts[('act1_online', 'act2_online', 'act3_online', 'hid_online').index(name)]
        out = np.zeros(shape, np.float32)
        out[:a.shape[0]] = a
        return out

    def get_beta_powers(self):
        return float(self.opt.beta1_power), float(self.opt.beta2_power)

    def get_step(self):
        return self.opt.step


def _setup(momentum=False):
    from deepqnetwork_b200.deep_q_model import init_tower_params
    spec = nn_ref.NetSpec(29, 24, 4, 4, num_hidden=32, use_dueling=True, use_double=True, use_per=True, use_momentum=momentum)
    on = init_tower_params(spec.param_shapes(), 1); tg = init_tower_params(spec.param_shapes(), 2)
    rng = np.random.default_rng(0)
    def batch(seed):
        r = np.random.default_rng(seed)
        shp = (8, 29, 24, 4)
        return dict(states=(r.integers(0, 256, shp) * (r.random(shp) < 0.5)).astype(np.uint8),
                    next_states=(r.integers(0, 256, shp) * (r.random(shp) < 0.5)).astype(np.uint8),
                    actions=r.integers(0, 4, 8).astype(np.uint8), rewards=r.choice([0, 1, 4], 8).astype(np.uint8),
                    continues=r.random(8) < 0.8), r.random(8).astype(np.float32) + 0.1
    return spec, on, tg, batch


def _one_step(dev, spec, shapes, b, w, it, slab=None, tamper=None):
    o, t, opt = pu.oracle_state_from_gpu(dev, shapes, spec)
    before = pu.snapshot(o, opt)
    dev.train_batch(b, w)
    if tamper:
        tamper(dev, before)
    r = pu.oracle_step_with_device_gates(dev, spec, o, t, opt, b, w, 8, 'cpu-check', it)
    return pu.check_step_against_oracle(dev, spec, shapes, before, r[3], o, opt, 'cpu-check', it, slab)


@pytest.mark.parametrize('momentum', [False, True])
@pytest.mark.parametrize('recover', [False, True], ids=['slab-grads', 'grads-recovered-from-m'])
def test_checker_accepts_an_honest_fp32_run(momentum, recover):
    spec, on, tg, batch = _setup(momentum)
    shapes = spec.param_shapes()
    dev = Fp32Device(spec, on, tg)
    slab = [k for k in shapes if k not in pu.DENSE_KERNELS] if recover else None
    for it in range(4):
        b, w = batch(it)
        _one_step(dev, spec, shapes, b, w, it, slab)


def _skip_update(dev, before):
    dev.o['dense/kernel'] = torch.as_tensor(before[0]['dense/kernel']).to(torch.float32)


def _double_update(dev, before):
    w0 = torch.as_tensor(before[0]['conv2d_1/kernel']).to(torch.float32)
    dev.o['conv2d_1/kernel'] = w0 + 2 * (dev.o['conv2d_1/kernel'] - w0)


def _one_element_off_by_lr(dev, before):
    w = dev.o['dense_2/kernel'].clone()
    g = np.abs(np.asarray(dev.grads['dense_2/kernel']))
    i = np.unravel_index(np.argmax(g), g.shape)  # an element with a well-determined gradient
    w[i] += 0.00025
    dev.o['dense_2/kernel'] = w


def _stale_v(dev, before):
    dev.opt.v['dense/kernel'] = torch.as_tensor(before[2]['dense/kernel']).to(torch.float32)


def _wrong_gradient(dev, before):
    dev.grads['conv2d/kernel'] = np.asarray(dev.grads['conv2d/kernel']) * 1.001


@pytest.mark.parametrize('tamper', [_skip_update, _double_update, _one_element_off_by_lr, _stale_v, _wrong_gradient])
def test_checker_rejects(tamper):
    spec, on, tg, batch = _setup()
    shapes = spec.param_shapes()
    dev = Fp32Device(spec, on, tg)
    b, w = batch(0)
    _one_step(dev, spec, shapes, b, w, 0)  # second step: m, v non-zero
    b, w = batch(1)
    with pytest.raises(AssertionError):
        _one_step(dev, spec, shapes, b, w, 1, tamper=tamper)
