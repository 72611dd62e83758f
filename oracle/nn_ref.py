"""torch-CPU restatement of the reference's DeepQModel graph.  TEST INFRASTRUCTURE.

PARITY UNPINNED: TensorFlow (Pipfile:13 tensorflow==1.12.3; Pipfile.lock:256
1.9.0) is an un-vendored dependency that cannot be installed here and the
reference holds no tests or golden vectors for this arithmetic.  This file
restates the published TF semantics at the reference's call sites:

  rl/deep_q_model.py:201-234  _make_inputs      u8 -> f32, divide by 255
  rl/deep_q_model.py:266-354  _make_q_network   conv(32,8,s4) conv(64,4,s2) conv(64,4,s1), all
                                                padding='same' + relu; NHWC flatten; dense 512 relu;
                                                dueling: V + (A - mean_a A)
  rl/deep_q_model.py:237-263  _make_network     max_q: reduce_max(target * one_hot(argmax(online)))
                                                (double) or reduce_max(target)
  rl/deep_q_model.py:106-114  _get_target_q_values   y = r(u8) + c(bool)*gamma*maxq(f32) in float64
  rl/deep_q_model.py:358-392  _make_train       q_a, abs/huber/IS-weighted squared loss, Adam/Momentum
  rl/deep_q_model.py:75-103   train             two runs; returns (step, losses, loss)
  rl/deep_q_model.py:144-157  get_losses
TF semantics restated ([TF-ext] in SURVEY.md): SAME padding out=ceil(in/s),
pad_total=max((out-1)*s+k-in,0), before=pad_total//2; conv = cross-correlation with HWIO
kernels; tf.layers.flatten on NHWC; tf.argmax first maximum; tf.losses.huber_loss(delta=1);
ApplyAdam: lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m += (g-m)(1-b1); v += (g*g-v)(1-b2);
var -= lr_t*m/(sqrt(v)+eps); ApplyMomentum(nesterov): acc = acc*mom + g;
var -= g*lr + acc*mom*lr; global_step post-increment.

`dtype=torch.float64` is the truth the CUDA path is graded against (1e-4 relative);
`dtype=torch.float32` is the noise yardstick and the timed CPU baseline.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

CONV_REFERENCE = dict(maps=(32, 64, 64), kernels=(8, 4, 4), strides=(4, 2, 1), padding='same')
# canonical Nature DQN geometry: a labelled extra, NOT what the reference builds (deep_q_model.py:281 is commented out)
CONV_NATURE = dict(maps=(32, 64, 64), kernels=(8, 4, 3), strides=(4, 2, 1), padding='valid')


class NetSpec:
    def __init__(self, height=89, width=80, channels=4, num_actions=4, num_hidden=512,
                 use_dueling=False, use_double=False, use_per=False, use_momentum=False,
                 learning_rate=0.00025, momentum=0.95, discount_rate=0.99, variant='reference'):
        self.height, self.width, self.channels = height, width, channels
        self.num_actions, self.num_hidden = num_actions, num_hidden
        self.use_dueling, self.use_double, self.use_per = use_dueling, use_double, use_per
        self.use_momentum = use_momentum
        self.learning_rate, self.momentum, self.discount_rate = learning_rate, momentum, discount_rate
        self.variant = variant

    def conv_layers(self):
        cfg = CONV_REFERENCE if self.variant == 'reference' else CONV_NATURE
        h, w, cin = self.height, self.width, self.channels
        out = []
        for cout, k, s in zip(cfg['maps'], cfg['kernels'], cfg['strides']):
            if cfg['padding'] == 'same':
                oh, ow = -(-h // s), -(-w // s)
                ph = max((oh - 1) * s + k - h, 0); pw = max((ow - 1) * s + k - w, 0)
                pt, pl = ph // 2, pw // 2
                pb, pr = ph - pt, pw - pl
            else:
                oh, ow = (h - k) // s + 1, (w - k) // s + 1
                pt = pb = pl = pr = 0
            out.append(dict(cin=cin, cout=cout, k=k, s=s, ih=h, iw=w, oh=oh, ow=ow, pt=pt, pb=pb, pl=pl, pr=pr))
            h, w, cin = oh, ow, cout
        return out

    def flat(self):
        last = self.conv_layers()[-1]
        return last['oh'] * last['ow'] * last['cout']

    def param_shapes(self):
        """TF variable names relative to the tower scope, in creation order (deep_q_model.py:292-343)."""
        shapes = OrderedDict()
        for i, l in enumerate(self.conv_layers()):
            name = 'conv2d' if i == 0 else 'conv2d_%d' % i
            shapes[name + '/kernel'] = (l['k'], l['k'], l['cin'], l['cout'])
            shapes[name + '/bias'] = (l['cout'],)
        flat, hid, a = self.flat(), self.num_hidden, self.num_actions
        if self.use_dueling:
            dims = [(flat, hid), (hid, a), (flat, hid), (hid, 1)]
        else:
            dims = [(flat, hid), (hid, a)]
        for i, (fi, fo) in enumerate(dims):
            name = 'dense' if i == 0 else 'dense_%d' % i
            shapes[name + '/kernel'] = (fi, fo)
            shapes[name + '/bias'] = (fo,)
        return shapes

    def num_params(self):
        return sum(int(np.prod(s)) for s in self.param_shapes().values())


def forward(params, x_u8, spec, dtype=torch.float32, acts=None, gates=None):
    """One tower.  params: name -> torch tensor (HWIO / [in,out]).  x_u8: [n,H,W,C] uint8 (numpy or torch).

    gates: optional [conv1, conv2, conv3 (NHWC bool), hidden ([n,NH] bool)] on/off pattern that replaces each ReLU's own
    `pre > 0` decision (y = pre * gate).  A pre-activation within fp32 rounding of zero can land on either side in an
    fp32 implementation; feeding that implementation's pattern keeps the fp64 gradient on the same branch, so the
    gradient comparison stays at the plain tolerance instead of being relaxed when a gate differs.  The forward value of
    such a unit changes by less than fp32 resolution."""
    def relu(v, i, nhwc):
        if gates is None:
            return F.relu(v)
        g = torch.as_tensor(np.asarray(gates[i]))
        if nhwc:
            g = g.permute(0, 3, 1, 2)
        return v * g.to(dtype)
    x = torch.as_tensor(np.asarray(x_u8)) if not torch.is_tensor(x_u8) else x_u8
    x = x.to(dtype) / 255
    x = x.permute(0, 3, 1, 2)  # NHWC -> NCHW for torch
    for i, l in enumerate(spec.conv_layers()):
        name = 'conv2d' if i == 0 else 'conv2d_%d' % i
        w = params[name + '/kernel'].to(dtype).permute(3, 2, 0, 1)  # HWIO -> OIHW
        x = F.pad(x, (l['pl'], l['pr'], l['pt'], l['pb']))
        x = relu(F.conv2d(x, w, params[name + '/bias'].to(dtype), stride=l['s']), i, True)
        if acts is not None:
            acts.append(x.permute(0, 2, 3, 1).detach().numpy())  # NHWC, post-ReLU
    flat = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)  # tf.layers.flatten on NHWC
    nh = spec.num_hidden
    def dense(i, inp, use_relu):
        name = 'dense' if i == 0 else 'dense_%d' % i
        y = inp @ params[name + '/kernel'].to(dtype) + params[name + '/bias'].to(dtype)
        if not use_relu:
            return y
        if gates is None:
            return F.relu(y)
        g = torch.as_tensor(np.asarray(gates[3]))
        g = g[:, nh:] if i == 2 else g[:, :nh]  # hidden gates: [advantage (or only) stream | value stream]
        return y * g.to(dtype)
    if spec.use_dueling:
        ha, hv = dense(0, flat, True), dense(2, flat, True)
        adv = dense(1, ha, False)
        val = dense(3, hv, False)
        q = val + (adv - adv.mean(dim=1, keepdim=True))
        if acts is not None:
            acts.append(torch.cat([ha, hv], dim=1).detach().numpy())
    else:
        hh = dense(0, flat, True)
        q = dense(1, hh, False)
        if acts is not None:
            acts.append(hh.detach().numpy())
    return q


def max_q_values(online, target, next_states, spec, dtype=torch.float32):
    """deep_q_model.py:247-255, including the one-hot-masked reduce_max of the double variant."""
    with torch.no_grad():
        qt = forward(target, next_states, spec, dtype)
        if spec.use_double:
            qo = forward(online, next_states, spec, dtype)
            onehot = F.one_hot(torch.argmax(qo, dim=1), spec.num_actions).to(dtype)
            return (qt * onehot).max(dim=1).values
        return qt.max(dim=1).values


def host_target(rewards_u8, continues_bool, max_q, discount_rate):
    """deep_q_model.py:114 -- NumPy float64 arithmetic on (u8, bool, python float, f32)."""
    return np.asarray(rewards_u8) + np.asarray(continues_bool) * discount_rate * np.asarray(max_q)


def huber(y, q):
    err = q - y
    a = err.abs()
    quad = torch.clamp(a, max=1.0)
    return 0.5 * quad * quad + (a - quad)


def losses_only(online, target, batch, spec, dtype=torch.float32):
    """DeepQModel.get_losses (deep_q_model.py:144-157)."""
    mq = max_q_values(online, target, batch['next_states'], spec, dtype)
    y = host_target(batch['rewards'], batch['continues'], mq.numpy(), spec.discount_rate)
    y = torch.as_tensor(y).to(torch.float32 if dtype == torch.float32 else dtype)
    with torch.no_grad():
        q = forward(online, batch['states'], spec, dtype)
        qa = (q * F.one_hot(torch.as_tensor(np.asarray(batch['actions'])).long(), spec.num_actions).to(dtype)).sum(1)
        return ((y.to(dtype) - qa).abs() if spec.use_per else huber(y.to(dtype), qa)).numpy()


class OptState:
    """Adam slots m,v + beta powers, or Momentum accumulators; plus train/step."""

    def __init__(self, params, spec, dtype=torch.float32):
        self.m = {k: torch.zeros_like(v, dtype=dtype) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v, dtype=dtype) for k, v in params.items()}
        # TF converts the python-float hyper-parameters to tensors of the variable dtype (float32) before ApplyAdam
        # uses them, so the true constants are the f32 roundings (1 - f32(0.999) differs from 0.001 by 1.3e-5)
        f32 = lambda x: float(np.float32(x))
        self.beta1, self.beta2, self.eps = f32(0.9), f32(0.999), f32(1e-8)
        self.beta1_power = torch.tensor(self.beta1, dtype=dtype)
        self.beta2_power = torch.tensor(self.beta2, dtype=dtype)
        self.step = 0


def apply_optimizer(params, grads, opt, spec, dtype):
    lr = torch.tensor(float(np.float32(spec.learning_rate)), dtype=dtype)
    if spec.use_momentum:
        mom = torch.tensor(float(np.float32(spec.momentum)), dtype=dtype)
        for k in params:
            g = grads[k]
            opt.m[k] = opt.m[k] * mom + g
            params[k] = params[k] - (g * lr + opt.m[k] * mom * lr)
    else:
        b1 = torch.tensor(opt.beta1, dtype=dtype); b2 = torch.tensor(opt.beta2, dtype=dtype)
        eps = torch.tensor(opt.eps, dtype=dtype)
        lr_t = lr * torch.sqrt(1 - opt.beta2_power) / (1 - opt.beta1_power)
        for k in params:
            g = grads[k]
            opt.m[k] = opt.m[k] + (g - opt.m[k]) * (1 - b1)
            opt.v[k] = opt.v[k] + (g * g - opt.v[k]) * (1 - b2)
            params[k] = params[k] - (opt.m[k] * lr_t) / (torch.sqrt(opt.v[k]) + eps)
        opt.beta1_power = opt.beta1_power * b1
        opt.beta2_power = opt.beta2_power * b2
    opt.step += 1


def train_step(online, target, opt, batch, is_weights, spec, dtype=torch.float32, gates=None, dq_override=None):
    """DeepQModel.train (deep_q_model.py:75-103): returns (step, losses, loss); updates `online`, `opt` in place.

    online/target: dict name -> torch tensor of `dtype`.  batch: dict of numpy arrays
    (states u8 [B,H,W,C], actions u8, rewards u8, continues bool, next_states u8).
    """
    mq = max_q_values(online, target, batch['next_states'], spec, dtype)
    y64 = host_target(batch['rewards'], batch['continues'], mq.numpy(), spec.discount_rate)
    # fed through a tf.float32 placeholder; the fp64 truth keeps full precision
    y = torch.as_tensor(y64.astype(np.float32)).to(dtype) if dtype == torch.float32 else torch.as_tensor(y64)
    leaf = OrderedDict((k, v.detach().clone().requires_grad_(True)) for k, v in online.items())
    acts = []
    q = forward(leaf, batch['states'], spec, dtype, acts=acts, gates=gates)
    onehot = F.one_hot(torch.as_tensor(np.asarray(batch['actions'])).long(), spec.num_actions).to(dtype)
    qa = (q * onehot).sum(1)
    abs_l = (y - qa).abs()
    hub = huber(y, qa)
    if spec.use_per:
        w = torch.as_tensor(np.asarray(is_weights, dtype=np.float32)).to(dtype)
        loss = (w * (y - qa) ** 2).mean()
        losses = abs_l
    else:
        loss = hub.mean()
        losses = hub
    # dL/dq_a per sample, the vector the backward pass starts from (aux['dq']).  dq_override: evaluate the backward pass
    # from a caller-supplied vector instead (the device's own TD-error gradient), so that a test can hold the backward
    # arithmetic to its tolerance separately from the forward error that |y - q_a| << |q_a| amplifies.
    dq_ref = torch.autograd.grad(loss, qa, retain_graph=True)[0]
    if dq_override is None:
        grads = torch.autograd.grad(loss, list(leaf.values()))
    else:
        up = torch.as_tensor(np.asarray(dq_override, np.float64)).to(dtype)
        grads = torch.autograd.grad(qa, list(leaf.values()), grad_outputs=up)
    grads = dict(zip(leaf.keys(), grads))
    with torch.no_grad():
        apply_optimizer(online, grads, opt, spec, dtype)
    return opt.step, losses.detach().numpy(), float(loss.detach()), dict(q=q.detach().numpy(), y=y.numpy(),
                                                                        grads={k: g.numpy() for k, g in grads.items()},
                                                                        max_q=mq.numpy(), acts=acts, dq=dq_ref.detach().numpy())


def to_torch(params_np, dtype=torch.float32):
    return OrderedDict((k, torch.as_tensor(np.asarray(v)).to(dtype).clone()) for k, v in params_np.items())


def make_priority(losses_f32, per_a=0.6, lo=0.01, hi=1.0):
    """deep_q_agent.py:527-534, all float32 (losses is an f32 array)."""
    l = np.asarray(losses_f32, dtype=np.float32)
    return np.power(np.minimum(l + np.float32(lo), np.float32(hi)), np.float32(per_a)).astype(np.float32)


def is_weights(priorities_f64, total_f32, batch_size, per_b, last_max_weight):
    """deep_q_agent.py:518-519."""
    probs = np.asarray(priorities_f64, np.float64) / total_f32
    return np.power(batch_size * probs, -per_b) / last_max_weight


def max_weight(min_p, total_f32, batch_size, per_b, lo=0.01):
    """deep_q_agent.py:506-510 (python max/min/pow on np.float32 scalars)."""
    last_min = max(min_p, lo)
    return pow(batch_size * (last_min / total_f32), -per_b)


def per_beta(step, b_start=0.4, b_end=1, anneal_steps=2000000):
    """deep_q_agent.py:445-450."""
    return min(b_end, b_start + step * ((b_end - b_start) / anneal_steps))
