"""Agent seam called by the runner (rl/agent/game_agent.py:4-55): train / get_action / is_training_done /
open / close + context manager.  RandomAgent is the reference's baseline agent (:58-68)."""
import random


class GameAgent:
    def train(self, game_state):
        return True

    def get_action(self, state):
        raise NotImplementedError

    def is_training_done(self):
        return True

    def open(self):
        pass

    def close(self, ty=None, value=None, tb=None):
        pass

    def __enter__(self):
        return self.open()

    def __exit__(self, ty, value, tb):
        self.close(ty, value, tb)


class RandomAgent(GameAgent):
    def __init__(self, conf):
        self.action_space = conf['action_space']

    def get_action(self, state):
        return random.randint(0, self.action_space - 1)
