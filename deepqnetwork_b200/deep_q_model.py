"""Model seam: the reference's ``DeepQModel`` interface over the B200 learner.

Mirrors rl/deep_q_model.py (class attrs :42-59, ctor :62-72, train :75-103, predict :118-131, get_action
:134-141, get_losses :144-157, copy_network :160-165, get_training_step :168-173, set/get_game_count :176-188)
and the session/saver runtime of rl/model/model.py (open :86-99, close :138-149, save :60-68, restore :71-83,
context manager :39-50).  There is no TensorFlow graph: ``open()`` creates the device learner
(libdqn_b200.so), initialises both towers independently like ``global_variables_initializer`` does
(online != target until the first copy, rl/deep_q_model.py:243-244) and restores a checkpoint when one exists.
"""
import os

import numpy as np

from .learner import Learner
from .utils.logging import log


def conv_geometry(height, width, channels, variant='reference'):
    """Layer geometry of rl/deep_q_model.py:278-283 under TF 'same' padding (or the labelled Nature variant)."""
    maps, strides = (32, 64, 64), (4, 2, 1)
    kernels = (8, 4, 4) if variant == 'reference' else (8, 4, 3)
    h, w, cin, out = height, width, channels, []
    for cout, k, s in zip(maps, kernels, strides):
        if variant == 'reference':
            oh, ow = -(-h // s), -(-w // s)
        else:
            oh, ow = (h - k) // s + 1, (w - k) // s + 1
        out.append(dict(k=k, s=s, cin=cin, cout=cout, oh=oh, ow=ow))
        h, w, cin = oh, ow, cout
    return out


def tower_param_shapes(height, width, channels, num_actions, num_hidden=512, use_dueling=False, variant='reference'):
    """TF variable names (relative to the tower scope) -> shapes, creation order of rl/deep_q_model.py:292-343."""
    shapes = {}
    geo = conv_geometry(height, width, channels, variant)
    for i, g in enumerate(geo):
        base = 'conv2d' if i == 0 else 'conv2d_%d' % i
        shapes[base + '/kernel'] = (g['k'], g['k'], g['cin'], g['cout'])
        shapes[base + '/bias'] = (g['cout'],)
    flat = geo[-1]['oh'] * geo[-1]['ow'] * geo[-1]['cout']
    dense = [(flat, num_hidden), (num_hidden, num_actions)]
    if use_dueling:
        dense += [(flat, num_hidden), (num_hidden, 1)]
    for i, (fi, fo) in enumerate(dense):
        base = 'dense' if i == 0 else 'dense_%d' % i
        shapes[base + '/kernel'] = (fi, fo)
        shapes[base + '/bias'] = (fo,)
    return shapes


def init_tower_params(shapes, seed):
    """Initial values with the distributions TF 1.x gives the reference's layers: conv kernels glorot-uniform
    (tf.layers.conv2d default), dense kernels tf.contrib.layers.variance_scaling_initializer() (factor 2, fan-in,
    truncated normal; rl/deep_q_model.py:273), biases zero.  Streams differ from TF's RNG by construction."""
    rng = np.random.default_rng(seed)
    out = {}
    for name, shp in shapes.items():
        if name.endswith('/bias'):
            out[name] = np.zeros(shp, np.float32)
        elif len(shp) == 4:
            rf = shp[0] * shp[1]
            limit = np.sqrt(6.0 / (rf * shp[2] + rf * shp[3]))
            out[name] = rng.uniform(-limit, limit, shp).astype(np.float32)
        else:
            std = np.sqrt(1.3 * 2.0 / shp[0])
            v = rng.normal(0.0, std, shp)
            bad = np.abs(v) > 2 * std
            while bad.any():  # truncated normal: redraw beyond two standard deviations
                v[bad] = rng.normal(0.0, std, int(bad.sum()))
                bad = np.abs(v) > 2 * std
            out[name] = v.astype(np.float32)
    return out


class DeepQModel:
    DEFAULT_OPTIONS = {
        'use_dueling': False,
        'learning_rate': 0.00025,
        'use_momentum': False,
        'momentum': 0.95,
        'use_per': False,
        'discount_rate': 0.99,
    }

    INPUT_HEIGHT = 110
    INPUT_WIDTH = 80
    INPUT_CHANNELS = 4
    USE_CONV = True
    USE_ENCODER = False
    STATE_TYPE = 'uint8'
    NUM_HIDDEN = 512
    NUM_FRAMES_PER_STATE = 4
    CONV_VARIANT = 'reference'

    def __init__(self, conf):
        # Reference quirk (rl/deep_q_model.py:65-69): the agent conf is merged with `if key not in self.conf`,
        # so the six keys of DEFAULT_OPTIONS (use_dueling, use_per, use_momentum, learning_rate, momentum,
        # discount_rate) can never be overridden -- `--dueling`/`--per` do not reach the reference's graph.
        # honor_model_flags=True (default) makes the flags effective, which is what README.md:44-50 and the
        # benchmark configs describe; honor_model_flags=False reproduces the shadowing exactly.
        self.conf = dict(self.DEFAULT_OPTIONS)
        honor = conf.get('honor_model_flags', True)
        for key, value in conf.items():
            if key not in self.conf or (honor and value is not None):
                self.conf[key] = value
        self.num_outputs = self.conf['action_space']
        self.discount_rate = self.conf['discount_rate']
        self.learner = None
        self.load_model = True
        self.save_model = False
        if not self.USE_CONV:
            raise NotImplementedError('only the convolutional image path (USE_CONV) is built for B200')

    # -- lifecycle (rl/model/model.py) ------------------------------------------------
    def param_shapes(self):
        return tower_param_shapes(self.INPUT_HEIGHT, self.INPUT_WIDTH, self.INPUT_CHANNELS, self.num_outputs,
                                  self.NUM_HIDDEN, bool(self.conf.get('use_dueling')), self.CONV_VARIANT)

    def open(self):
        c = self.conf
        self.learner = Learner(
            self.INPUT_HEIGHT, self.INPUT_WIDTH, self.INPUT_CHANNELS, num_actions=self.num_outputs,
            num_hidden=self.NUM_HIDDEN, use_dueling=c.get('use_dueling', False), use_double=c.get('use_double', False),
            use_per=c.get('use_per', False), use_momentum=c.get('use_momentum', False),
            learning_rate=c.get('learning_rate', 0.00025), momentum=c.get('momentum', 0.95),
            discount_rate=c.get('discount_rate', 0.99), batch_size=c.get('batch_size', 32) or 32,
            replay_capacity=c.get('replay_max_memory_length', 10000) or 10000,
            per_a=c.get('per_a', 0.6), per_b_start=c.get('per_b_start', 0.4), per_b_end=c.get('per_b_end', 1),
            per_anneal_steps=c.get('per_anneal_steps', 2000000), use_per_annealing=c.get('use_per_annealing', False),
            per_calculate_steps=c.get('per_calculate_steps', 5000), copy_network_steps=0,
            conv_variant=self.CONV_VARIANT, math_mode=c.get('math_mode', 'tc3x'), seed=c.get('device_seed', 7),
            device=c.get('device', 0), frame_dedup=c.get('frame_dedup', False), frame_capacity=c.get('frame_capacity', 0) or 0)
        loaded = False
        prefix = self.conf.get('save_path_prefix')
        if self.load_model and prefix:
            loaded = self.restore(prefix)
        if not loaded:
            shapes = self.param_shapes()
            seed = self.conf.get('init_seed', 42)
            self.learner.set_tower(init_tower_params(shapes, seed), tower=0)
            self.learner.set_tower(init_tower_params(shapes, seed + 1), tower=1)
        return self

    def close(self, ty=None, value=None, tb=None):
        if self.learner is not None:
            if self.save_model and self.conf.get('save_path_prefix'):
                self.save(self.conf['save_path_prefix'])
            self.learner.close()
            self.learner = None

    def __enter__(self):
        return self.open()

    def __exit__(self, ty, value, tb):
        self.close(ty, value, tb)

    def save(self, save_path_prefix):
        """rl/model/model.py:60-67.  conf['persist_format'] = 'reference' (default) writes what `tf.train.Saver().save` leaves
        behind -- <prefix>.index, <prefix>.data-00000-of-00001 and the directory's `checkpoint` file, with the reference
        graph's variable names (formats/tf_bundle.py) -- so a save-dir moves between the reference and this build in both
        directions; 'native' writes the <prefix>.dqnb200 container (adds the device RNG counter of device_rng runs)."""
        log('saving model: ', save_path_prefix)
        if self.conf.get('persist_format', 'reference') == 'native':
            self.learner.save(save_path_prefix)
        else:
            self.learner.save(save_path_prefix, fmt='reference', shapes=self.param_shapes(),
                              use_momentum=bool(self.conf.get('use_momentum')))
        log('saved model')

    def restore(self, save_path_prefix):
        """rl/model/model.py:70-83: False when there is no checkpoint.  A TensorFlow checkpoint (<prefix>.index) is read
        whatever wrote it; a <prefix>.dqnb200 container is used when only that exists."""
        if os.path.exists(save_path_prefix + '.index'):
            log('  restoring model: ', save_path_prefix)
            ok = self.learner.restore(save_path_prefix, fmt='reference', shapes=self.param_shapes(),
                                      use_momentum=bool(self.conf.get('use_momentum')))
        elif os.path.exists(save_path_prefix + '.dqnb200'):
            log('  restoring model: ', save_path_prefix + '.dqnb200')
            ok = self.learner.restore(save_path_prefix)
        else:
            log('  model does not exist:', save_path_prefix)
            return False
        log('  restored model')
        return ok

    # -- compute (rl/deep_q_model.py) ------------------------------------------------------
    def train(self, X_states, X_actions, rewards, continues, next_states, is_weights=None):
        """-> (step, losses f32[B], loss).  Host-buffer form of the step (seam 2)."""
        if self.conf.get('use_per') and is_weights is None:
            raise ValueError('use_per: is_weights must be fed')
        return self.learner.train_batch(X_states, X_actions, rewards, continues, next_states,
                                        is_weights if self.conf.get('use_per') else None)

    def predict(self, X_states, use_target=False):
        return self.learner.predict(np.asarray(X_states), use_target=use_target)

    def get_action(self, X_states, use_target=False):
        return np.argmax(self.predict(X_states, use_target=use_target), axis=1)

    def get_losses(self, X_states, actions, rewards, continues, next_states):
        return self.learner.get_losses(X_states, actions, rewards, continues, next_states)

    def copy_network(self):
        self.learner.copy_network()

    def get_training_step(self):
        return self.learner.get_step()

    def set_game_count(self, count):
        self.learner.set_game_count(count)

    def get_game_count(self):
        return self.learner.get_game_count()


class BreakoutModel(DeepQModel):
    INPUT_HEIGHT = 89
    INPUT_WIDTH = 80
    INPUT_CHANNELS = 4


class SpaceInvadersModel(DeepQModel):
    INPUT_HEIGHT = 89
    INPUT_WIDTH = 80
    INPUT_CHANNELS = 4


class MsPacmanModel(DeepQModel):
    INPUT_HEIGHT = 86
    INPUT_WIDTH = 80
    INPUT_CHANNELS = 4


class NatureModel(DeepQModel):
    """Labelled extra, NOT a reference class: canonical Nature-DQN geometry (84x84x4, conv 8/4 4/2 3/1 'valid')
    that BASELINE.json's config strings name; the reference itself builds 8/4/4 'same' (deep_q_model.py:281-283)."""
    INPUT_HEIGHT = 84
    INPUT_WIDTH = 84
    INPUT_CHANNELS = 4
    CONV_VARIANT = 'nature'
