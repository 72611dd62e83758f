"""Dotted-path class loader used by the --agent/--model/--env/--modelenv plugin seams
(rl/utils/import_class.py:5-20).  Paths written by the reference ('rl.deep_q_model.BreakoutModel', saved in
existing .conf files) resolve to this package's mirrors of the same names."""
import importlib
import inspect

_ALIASES = {
    'rl.deep_q_model': 'deepqnetwork_b200.deep_q_model',
    'rl.deep_q_agent': 'deepqnetwork_b200.deep_q_agent',
    'rl.agent.game_agent': 'deepqnetwork_b200.agent.game_agent',
    'rl.data.replay_memory': 'deepqnetwork_b200.data.replay_memory',
    'rl.data.sum_tree': 'deepqnetwork_b200.data.sum_tree',
    'rl.data.replay_sampler': 'deepqnetwork_b200.data.replay_sampler',
    'rl.data.replay_sampler_priority': 'deepqnetwork_b200.data.replay_sampler_priority',
}


def import_class(name):
    if inspect.isclass(name):
        return name
    if not isinstance(name, str):
        raise Exception('invalid env class')
    module_path, _, class_name = name.rpartition('.')
    module = importlib.import_module(_ALIASES.get(module_path, module_path))
    return getattr(module, class_name)
