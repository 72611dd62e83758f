"""INTEGRATION.md section 1 ("zero-change route"), executed: the reference's own train.py / play.py entry points are run
with the documented command line up to the point where they hand control to GameRunner.  Needs the reference tree
(/root/reference, build container only) and no GPU: TensorFlow / gym / matplotlib / PIL / h5py are stubbed as empty
modules (never called), DeepQAgent.open()/close() are replaced by recorders (they would create the device learner)."""
import os
import runpy
import sys
import types

import pytest

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present (GPU box)')

# the command INTEGRATION.md documents.  NOTE the first flag: rl train.py:39-40 is missing a comma, so the reference's
# agent option is literally "-a--agent" (SURVEY Q2); "--agent" is rejected by its argparse.
DOCUMENTED = ['-a--agent', 'deepqnetwork_b200.deep_q_agent.DeepQAgent', '--model', 'deepqnetwork_b200.deep_q_model.BreakoutModel',
              '--env', 'tests.stub_env.StubEnvironment', '--double', '--dueling', '--per']


@pytest.fixture()
def stubbed(monkeypatch, tmp_path):
    for name in ('tensorflow', 'gym', 'h5py', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.animation', 'PIL', 'PIL.Image'):
        monkeypatch.setitem(sys.modules, name, types.ModuleType(name))
    sys.modules['matplotlib'].use = lambda *a, **k: None
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.modules['matplotlib'].animation = sys.modules['matplotlib.animation']
    for sub in ('Image', 'ImageFont', 'ImageDraw'):
        setattr(sys.modules['PIL'], sub, types.ModuleType('PIL.' + sub))
    monkeypatch.syspath_prepend(REF)
    for m in [m for m in sys.modules if m == 'rl' or m.startswith('rl.')]:
        monkeypatch.delitem(sys.modules, m)
    yield tmp_path
    for m in [m for m in sys.modules if m == 'rl' or m.startswith('rl.')]:
        sys.modules.pop(m, None)


def _run_reference_train(monkeypatch, argv):
    import deepqnetwork_b200.deep_q_agent as mirror
    seen = {}
    monkeypatch.setattr(mirror.DeepQAgent, 'open', lambda self: seen.setdefault('opened', self))
    monkeypatch.setattr(mirror.DeepQAgent, 'close', lambda self, *a: seen.setdefault('closed', True))
    import rl.game_runner as gr
    monkeypatch.setattr(gr.GameRunner, 'run', lambda self: seen.setdefault('runner', self))
    monkeypatch.setattr(sys, 'argv', ['train.py'] + argv)
    runpy.run_path(os.path.join(REF, 'train.py'), run_name='__main__')
    return seen


def test_documented_train_command_reaches_game_runner_with_the_mirror_agent(stubbed, monkeypatch):
    from deepqnetwork_b200.deep_q_agent import DeepQAgent
    from deepqnetwork_b200.deep_q_model import BreakoutModel
    save = str(stubbed / 'data-breakout')
    # train.py:158-159 opens <save_dir>/<environment>.log before the agent gets to create the directory, so a first run
    # needs the directory to exist; the mirror agent then starts from DEFAULT_OPTIONS (the reference's own agent would
    # fail on `None[key] = value` there, rl/deep_q_agent.py:106-113)
    os.makedirs(save)
    seen = _run_reference_train(monkeypatch, DOCUMENTED + ['-O', save])
    agent = seen['opened']
    assert isinstance(agent, DeepQAgent) and isinstance(agent.model, BreakoutModel) and seen['closed']
    assert seen['runner'].agent is agent
    c = agent.conf
    assert c['use_double'] and c['use_dueling'] and c['use_per'] and c['is_training'] and c['action_space'] == 4
    # defaults come from DEFAULT_OPTIONS (deep_q_agent.py:33-67), flags left at None do not override them
    assert c['batch_size'] == 32 and c['replay_max_memory_length'] == 300000 and c['copy_network_steps'] == 10000
    # model flags reach the model (honor_model_flags)
    assert agent.model.conf['use_dueling'] and agent.model.conf['use_per']
    assert os.path.exists(os.path.join(save, 'tests.stub_env.StubEnvironment.conf'))


def test_modelenv_shorthand_resolves_reference_paths_to_the_mirrors(stubbed, monkeypatch):
    """--modelenv=Breakout writes 'rl.deep_q_model.BreakoutModel' into the conf (train.py:150-155); the mirror agent
    resolves that dotted path to its own BreakoutModel.  The environment stays the reference's class, which needs gym:
    only the path handling is checked."""
    from deepqnetwork_b200.deep_q_model import BreakoutModel
    from deepqnetwork_b200.utils.import_class import import_class
    assert import_class('rl.deep_q_model.BreakoutModel') is BreakoutModel
    import rl.utils.import_class as ref_ic
    from deepqnetwork_b200.deep_q_agent import DeepQAgent
    assert ref_ic.import_class('deepqnetwork_b200.deep_q_agent.DeepQAgent') is DeepQAgent


def test_the_undocumented_spelling_is_rejected_by_the_reference(stubbed, monkeypatch):
    with pytest.raises(SystemExit):
        _run_reference_train(monkeypatch, ['--agent', 'deepqnetwork_b200.deep_q_agent.DeepQAgent', '-O', str(stubbed)])
