"""A deterministic stand-in for the emulator side of the agent seam (tests only; no gym / ALE / ROMs here).

StubEnvironment follows the env seam of rl/game_environment.py:20-98 (get_action_space / reset / step ->
(obs u8[H,W,1], reward, done, info with 'ale.lives')).  drive() restates the caller side of the agent seam -- what
rl/game_runner.py:115-244 does between env.step() and agent.train()/get_action(): the 4-frame deque, old_state/state
hand-over, cont flag, the extra end-of-episode train() call with cont = 0 -- and records every game_state it
hands to the agent, so that a test knows exactly which transitions the agent was given."""
from collections import deque

import numpy as np


class StubEnvironment:
    GAME_ID = 'Stub-v0'

    def __init__(self, height=89, width=80, actions=4, seed=0, episode_len=37, lives=3):
        self.h, self.w, self.actions = height, width, actions
        self.rng = np.random.default_rng(seed)
        self.episode_len, self.lives0 = episode_len, lives

    def get_action_space(self):
        return self.actions

    def _obs(self):
        # image-like: flat background with a few bright blocks that move
        f = np.full((self.h, self.w, 1), 40, np.uint8)
        for _ in range(6):
            y, x = self.rng.integers(0, self.h - 8), self.rng.integers(0, self.w - 8)
            f[y:y + 8, x:x + 8, 0] = self.rng.integers(60, 256)
        f[self.rng.random((self.h, self.w, 1)) < 0.02] = 255
        return f

    def reset(self):
        self.t, self.lives = 0, self.lives0
        return self._obs()

    def step(self, action):
        self.t += 1
        reward = float(self.rng.choice([0, 0, 0, 1, 4, 400]))  # 400 wraps to 144 in the uint8 column
        if self.t % self.episode_len == 0:
            self.lives -= 1
        return self._obs(), reward, False, {'ale.lives': self.lives}


def drive(agent, env, max_agent_calls, frames_per_state=4, on_call=None):
    """rl/game_runner.py:76-233 for a training run, until the agent has been handed max_agent_calls game states (or says
    training is done).  Returns the list of recorded calls: dicts with game_done, old_state, action, reward, cont, state."""
    calls = []

    def train(gs):
        rec = dict(game_done=gs['game_done'], old_state=gs['old_state'], action=gs['action'], reward=gs['reward'],
                   cont=gs['cont'], state=gs['state'])
        calls.append(rec)
        if on_call:
            on_call(rec)
        return agent.train(gs)

    while len(calls) < max_agent_calls and not agent.is_training_done():
        # _reset_game_state (:139-148)
        gs = dict(game_done=False, episode_done=False, old_state=None, state=None, action=0, reward=None, cont=1,
                  num_lives=None, observation=env.reset())
        frames = deque(maxlen=frames_per_state)
        frames.append(gs['observation'])

        def update_frame_state():  # :236-244
            if len(frames) >= frames_per_state:
                gs['old_state'] = gs['state']
                gs['state'] = np.concatenate(frames, axis=2)

        while not gs['game_done'] and len(calls) < max_agent_calls:  # _run_game (:115-136)
            gs['episode_done'] = False
            while not gs['episode_done'] and not gs['game_done'] and len(calls) < max_agent_calls:  # _run_episode (:151-191)
                if len(frames) >= frames_per_state:
                    gs['cont'] = 1
                    update_frame_state()
                    train(gs)
                    gs['action'] = agent.get_action(gs['state'])
                gs['reward'] = 0
                obs, r, done, info = env.step(gs['action'])  # _run_game_step (:194-233), frame_skip 1
                gs['observation'], gs['game_done'] = obs, done
                gs['reward'] += r
                lives = info['ale.lives']
                if gs['num_lives'] is not None and lives != gs['num_lives']:
                    gs['episode_done'] = True
                if lives <= 0:
                    gs['game_done'] = True
                gs['num_lives'] = lives
                frames.append(gs['observation'])
            gs['cont'] = 0
            update_frame_state()
            train(gs)
    return calls
