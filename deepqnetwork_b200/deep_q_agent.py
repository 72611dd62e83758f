"""Agent seam: the reference's ``DeepQAgent`` (rl/deep_q_agent.py) over the B200 learner.

Same conf handling (DEFAULT_OPTIONS :33-67, three-layer merge :91-117, one ``<environment>.conf`` per save dir
:120-133), same training cadence (warm-up gate and every-Nth-game-step rule :271-306, target sync / save :352-365,
report line :368-407) and the same PER maths (:445-450, :500-534).  What changes is where the step runs: replay
ring, sum tree, batch and both networks live in HBM, and ``_train_run_train`` is ONE fused device call
(``dqn_train_step``) fed with the same ``random`` draws the reference would make.
"""
import json
import os
import random
import time

import numpy as np

from .agent.game_agent import GameAgent
from .data.replay_memory import ReplayMemory
from .data.replay_memory_disk import ReplayMemoryDisk
from .data.replay_sampler import ReplaySampler
from .data.replay_sampler_priority import ReplaySamplerPriority
from .utils.config import get_conf
from .utils.import_class import import_class
from .utils.logging import log

MIN_ERROR_PRIORITY = 0.01
MAX_ERROR_PRIORITY = 1.0


def make_priority(losses, per_a=0.6):
    """min(|delta| + 0.01, 1.0) ** per_a in float32 (rl/deep_q_agent.py:527-534)."""
    l = np.asarray(losses, dtype=np.float32)
    return np.power(np.minimum(l + np.float32(MIN_ERROR_PRIORITY), np.float32(MAX_ERROR_PRIORITY)), np.float32(per_a))


def importance_weights(priorities, total, batch_size, per_b, last_max_weight):
    """(B * p / total) ** -beta / last_max_weight (rl/deep_q_agent.py:518-519)."""
    probs = np.asarray(priorities, np.float64) / total
    return np.power(batch_size * probs, -per_b) / last_max_weight


def max_weight(min_priority, total, batch_size, per_b):
    """pow(B * (max(min_p, 0.01) / total), -beta) (rl/deep_q_agent.py:508-510)."""
    return pow(batch_size * (max(min_priority, MIN_ERROR_PRIORITY) / total), -per_b)


class DeepQAgent(GameAgent):
    DEFAULT_OPTIONS = {
        'save_dir': './data', 'game_id': None,
        'eps_min': 0.1, 'eps_max': 1.0, 'eps_decay_steps': 2000000,
        'discount_rate': 0.99,
        'save_model_steps': 10000, 'copy_network_steps': 10000,
        'batch_size': 32, 'model_save_prefix': None,
        'replay_max_memory_length': 300000, 'replay_cache_size': 300000,
        'max_num_training_steps': 2000000, 'num_game_frames_before_training': 10000,
        'num_game_steps_per_train': 4, 'num_train_steps_save_video': None,
        'train_report_interval': 100,
        'use_episodes': True, 'use_dueling': False, 'use_double': False, 'use_per': False,
        'use_per_annealing': False, 'use_momentum': False, 'use_memory': True, 'use_log': True,
        'frame_skip': 1, 'max_game_length': 50000, 'tf_log_level': 3,
        'per_a': 0.6, 'per_b_start': 0.4, 'per_b_end': 1, 'per_anneal_steps': 2000000, 'per_calculate_steps': 5000,
    }
    MAX_MEMORY_BATCH_SIZE = 128
    MIN_ERROR_PRIORITY = MIN_ERROR_PRIORITY
    MAX_ERROR_PRIORITY = MAX_ERROR_PRIORITY

    def __init__(self, conf):
        self._init_conf(conf)
        log('config values:')
        for key, value in self.conf.items():
            log('  {}: {}'.format(key, value))
        self.model = import_class(self.conf['model'])(conf=self.conf)
        self.is_training = self.conf['is_training']

    # -- configuration ----------------------------------------------------------------------
    def _init_conf(self, conf):
        self.save_dir = conf['save_dir']
        if not os.path.exists(self.save_dir):
            os.makedirs(self.save_dir, exist_ok=True)
            self.conf = dict(self.DEFAULT_OPTIONS)
        else:
            self.conf = get_conf(self.save_dir)
            if self.conf is None:  # directory exists but holds no .conf yet
                self.conf = dict(self.DEFAULT_OPTIONS)
        for key, value in conf.items():  # every non-None caller value wins (store_true flags are never None)
            if value is not None:
                self.conf[key] = value
        self.save_path_prefix = os.path.join(self.save_dir, self.conf['environment'])
        self.conf['save_path_prefix'] = self.save_path_prefix
        conf_path = self.save_path_prefix + '.conf'
        if not os.path.exists(conf_path):
            with open(conf_path, 'w+') as fo:
                json.dump(self.conf, fo, sort_keys=True, indent=4)

    # -- lifecycle --------------------------------------------------------------------------------
    def open(self):
        self.model.open()
        if self.is_training:
            self._train_init()
        else:
            self.use_epsilon = self.conf.get('use_epsilon', False)
            self.step = self.model.get_training_step()

    def close(self, ty=None, value=None, tb=None):
        if self.is_training:
            self._train_finish()
        self.model.close()

    def __enter__(self):
        return self.open()

    def __exit__(self, ty, value, tb):
        self.close(ty, value, tb)

    def _train_init(self):
        c = self.conf
        self.max_num_training_steps = c['max_num_training_steps']
        self.replay_max_memory_length = c['replay_max_memory_length']
        self.num_game_frames_before_training = c['num_game_frames_before_training']
        self.batch_size = c['batch_size']
        self.save_steps = c['save_model_steps']
        self.copy_steps = c['copy_network_steps']
        self.train_report_interval = c['train_report_interval']
        self.num_game_steps_per_train = c['num_game_steps_per_train']
        self.use_per = bool(c['use_per'])
        self.eps_min, self.eps_max, self.eps_decay_steps = c['eps_min'], c['eps_max'], c['eps_decay_steps']
        self.train_start_time = time.time()
        self.game_step_count = 0
        self.step = self.model.get_training_step()
        self.train_report_start_time = time.time()
        self.train_report_last_step = self.step
        self.total_losses = []
        self.device_rng = bool(c.get('device_rng', False))

        m = self.model
        learner = m.learner
        # the replay ring lives in the model's device learner (HBM); the reference's RAM/HDF5 choice (--memory /
        # --disk) only decides where the *reference* keeps it, so both map to the HBM ring here
        native_replay = '{}_replay_memory.dqnb200'.format(self.save_path_prefix)
        old_memories = None
        if c.get('use_memory', True) and os.path.exists(self._get_replay_memory_path(reference=True)):
            # deep_q_agent.py:216-221: the HDF5 replay written at the last exit (by the reference or by this build)
            log('loading old memories from', self._get_replay_memory_path(reference=True))
            old_memories = ReplayMemoryDisk(self._get_replay_memory_path(reference=True))
            log('found', len(old_memories))
        elif c.get('use_memory', True) and os.path.exists(native_replay):
            # own container (persist_format 'native'); before the sampler is built: under PER its tree is rebuilt from the
            # stored float16 losses, as ReplaySamplerPriority.add_losses does
            log('loading old memories from', native_replay)
            log('found', learner.restore_replay(self.save_path_prefix))
        self.memories = ReplayMemory(m.INPUT_HEIGHT, m.INPUT_WIDTH, m.INPUT_CHANNELS, max_size=self.replay_max_memory_length,
                                     state_type=m.STATE_TYPE, learner=learner)
        if old_memories is not None:
            # self.memories.extend(old_memories) (:221), a block of rows at a time; under PER every row also enters the sum
            # tree with its stored float16 loss as priority, which is what ReplaySamplerPriority.add_losses rebuilds (:57-59)
            n = len(old_memories)
            if n > self.replay_max_memory_length:
                raise ValueError('replay file holds %d rows, replay_max_memory_length is %d' % (n, self.replay_max_memory_length))
            for s0 in range(0, n, 1024):
                r = slice(s0, min(n, s0 + 1024))
                self.memories.append_rows(np.asarray(old_memories.states[r]), np.asarray(old_memories.actions[r]),
                                          np.asarray(old_memories.rewards[r]), np.asarray(old_memories.next_states[r]),
                                          np.asarray(old_memories.continues[r]),
                                          np.asarray(old_memories.losses[r], dtype=np.float32) if self.use_per else None)
            old_memories.close()
        self.train_batch = None  # the B-row batch is device resident (learner.batch_read() for inspection)
        self._train_inflight = False  # pipeline_train: a queued step whose result has not been collected yet
        if self.use_per:
            self.per_memory_batch = ReplayMemory(m.INPUT_HEIGHT, m.INPUT_WIDTH, m.INPUT_CHANNELS,
                                                 max_size=self.MAX_MEMORY_BATCH_SIZE, state_type=m.STATE_TYPE, host=True)
            self.replay_sampler = ReplaySamplerPriority(self.memories)
            self.per_a = c['per_a']
            self.per_b_start, self.per_b_end = c['per_b_start'], c['per_b_end']
            self.per_b = self.per_b_start
            self.per_anneal_steps = c['per_anneal_steps']
            self.per_calculate_steps = c['per_calculate_steps']
            self.last_min_loss = self.last_max_loss = self.last_max_weight = self.last_avg_loss = None
            self.tree_idxes = np.zeros(self.batch_size, dtype=int)
            self.priorities = np.zeros(self.batch_size, dtype=float)
        else:
            self.replay_sampler = ReplaySampler(self.memories)
        self.is_weights = None

    # -- the agent seam ---------------------------------------------------------------------------------
    def train(self, game_state):
        if not self.is_training:
            return True
        if self.is_training_done():
            return False
        if game_state['game_done']:
            self.model.set_game_count(self.model.get_game_count() + 1)
        if game_state['old_state'] is not None:
            self._add_memories(state=game_state['old_state'], action=game_state['action'], reward=game_state['reward'],
                               cont=game_state['cont'], next_state=game_state['state'])
        if len(self.replay_sampler) < self.num_game_frames_before_training:
            return True
        if self.game_step_count == 0:
            log('[train] train start')
        self.game_step_count += 1
        if self.game_step_count % self.num_game_steps_per_train != 0:
            return True
        self._train_step()
        return True

    def is_training_done(self):
        return self.step >= self.max_num_training_steps

    def get_action(self, state):
        q_values = self.model.predict([state])
        if self.is_training or self.use_epsilon:
            return self._epsilon_greedy(q_values, self.step)
        return np.argmax(q_values)

    def get_actions(self, states):
        """Vectorised actors: the states of n environments -> n actions with ONE batched forward pass
        (DeepQModel.predict on [n, H, W, 4]) and the reference's per-call epsilon-greedy rule (deep_q_agent.py:559-592)
        applied per environment."""
        q_values = self.model.predict(np.asarray(states))
        greedy = np.argmax(q_values, axis=1)
        if not (self.is_training or self.use_epsilon):
            return greedy
        eps = self._get_epsilon(self.step)
        explore = np.random.rand(len(greedy)) < eps
        return np.where(explore, np.random.randint(self.model.num_outputs, size=len(greedy)), greedy)

    # -- one training step -------------------------------------------------------------------------------------
    def _train_step(self):
        self._train_run_train()
        self._train_save_model()
        self._train_report_stats()

    def _train_run_train(self):
        """sample -> train -> priority write-back as one device call (rl/deep_q_agent.py:326-349)."""
        learner = self.model.learner
        if self.device_rng:
            u = rows = None
        elif self.use_per:
            u, rows = np.array([random.random() for _ in range(self.batch_size)]), None  # replay_sampler_priority.py:33
        else:
            u, rows = None, self.replay_sampler.draw_rows(self.batch_size)                # replay_sampler.py:30
        if self.conf.get('pipeline_train'):
            # queue this step and collect the previous one's loss (it finished while the environment was stepping): the host never
            # waits for the step it has just launched.  One optimizer step per call, so the step counter is known without a read-back
            learner.train_step_async(uniforms=u, rows=rows)
            if self._train_inflight:
                _, losses, loss = learner.train_step_result()
                self.total_losses.append(loss)
            self._train_inflight = True
            self.step += 1
        else:
            self.step, losses, loss = learner.train_step(uniforms=u, rows=rows)
            self.total_losses.append(loss)

    def _refresh_per_stats(self):
        """The PER statistics of the report line (deep_q_agent.py:505-510 keeps them on the host; here they live on the device and
        are read back when a line is printed, not every step)."""
        st = self.model.learner.get_stats()
        self.per_b = st['per_b'] if not self.conf['use_per_annealing'] else self._annealed_beta(self.step)
        self.last_avg_loss, self.last_max_weight = st['avg'], st['last_max_weight']
        self.last_min_loss = max(st['min'], self.MIN_ERROR_PRIORITY)
        self.last_max_loss = min(st['max'], self.MAX_ERROR_PRIORITY)

    def _annealed_beta(self, step):
        return min(self.per_b_end, self.per_b_start + step * ((self.per_b_end - self.per_b_start) / self.per_anneal_steps))

    def _train_save_model(self):
        if self.step % self.copy_steps == 0:
            log('copying online to target dqn')
            self.model.copy_network()
        if self.step % self.save_steps == 0:
            self.model.save(self.save_path_prefix)

    def _train_report_stats(self):
        if self.step % self.train_report_interval != 0:
            return
        elapsed = time.time() - self.train_report_start_time
        frame_rate = (self.step - self.train_report_last_step) / elapsed if elapsed > 0 else 0.0
        self.train_report_last_step = self.step
        self.train_report_start_time = time.time()
        avg_loss = sum(self.total_losses) / len(self.total_losses) if self.total_losses else 0
        self.total_losses.clear()
        per_str = ''
        if self.use_per:
            self._refresh_per_stats()
            per_str = ('per_ab: {:0.2f}/{:0.2f}'.format(self.per_a, self.per_b) +
                       ' avg/min/max loss: {:0.3f}/{:0.3f}/{:0.3f}'.format(self.last_avg_loss, self.last_min_loss, self.last_max_loss) +
                       ' max_weight: {:0.3f} sum total: {:0.1f} '.format(self.last_max_weight, self.replay_sampler.total))
        log('[train] step {} game {}/{}'.format(self.step, self.game_step_count, self.model.get_game_count()) +
            ' avg loss: {:0.5f} {}mem: {:d} '.format(avg_loss, per_str, len(self.replay_sampler)) +
            ' fr: {:0.1f} eps: {:0.2f} '.format(frame_rate, self._get_epsilon(self.step)))

    def _train_finish(self):
        log('train finish')
        if self._train_inflight:  # pipeline_train: the last queued step
            self.total_losses.append(self.model.learner.train_step_result()[2])
            self._train_inflight = False
        if self.conf.get('use_memory', True):
            log('saving', len(self.memories), 'memories to disk')
            if self.conf.get('persist_format', 'reference') == 'native':
                self.model.learner.save_replay(self.save_path_prefix)
            else:
                # deep_q_agent.py:416-429 as written: the file is opened (created with this run's geometry when missing) and
                # the ring is appended at the file's own cur_idx
                m = self.model
                old_memories = ReplayMemoryDisk(self._get_replay_memory_path(reference=True), m.INPUT_HEIGHT, m.INPUT_WIDTH,
                                                m.INPUT_CHANNELS, state_type=m.STATE_TYPE,
                                                max_size=self.replay_max_memory_length, cache_size=0)
                old_memories.extend(self.memories)
                old_memories.close()
            log('saved memory to disk')
        self.model.save(self.save_path_prefix)
        elapsed = time.time() - self.train_start_time
        log('train finished in {:0.1f} seconds / {:0.1f} mins / {:0.1f} hours'.format(elapsed, elapsed / 60, elapsed / 3600))

    # -- replay insert (actor side) ----------------------------------------------------------------
    def _add_memories(self, state, action, reward, cont, next_state):
        if not self.use_per:
            self.replay_sampler.append(state=state, action=action, reward=reward, next_state=next_state, cont=cont)
            return
        pb = self.per_memory_batch
        pb.append(state=state, action=action, reward=reward, cont=cont, next_state=next_state)
        if len(pb) >= self.MAX_MEMORY_BATCH_SIZE:
            n = len(pb)
            losses = self.model.get_losses(pb.states[:n], pb.actions[:n], pb.rewards[:n], pb.continues[:n], pb.next_states[:n])
            self.replay_sampler.append_rows(pb.states[:n], pb.actions[:n], pb.rewards[:n], pb.next_states[:n],
                                            pb.continues[:n], make_priority(losses, self.per_a))
            pb.clear()

    def _make_priority(self, losses):
        return make_priority(losses, self.per_a)

    def _get_replay_memory_path(self, reference=None):
        """deep_q_agent.py:537-541: <prefix>_replay_memory.hdf5; the own container of persist_format 'native' sits beside it."""
        if reference is None:
            reference = self.conf.get('persist_format', 'reference') != 'native'
        return '{}_replay_memory.{}'.format(self.save_path_prefix, 'hdf5' if reference else 'dqnb200')

    # -- epsilon-greedy -----------------------------------------------------------------------------------------------
    def _get_epsilon(self, step):
        return max(self.eps_min, self.eps_max - (self.eps_max - self.eps_min) * step / self.eps_decay_steps)

    def _epsilon_greedy(self, q_values, step):
        if np.random.rand() < self._get_epsilon(step):
            return np.random.randint(self.model.num_outputs)
        return np.argmax(q_values)
