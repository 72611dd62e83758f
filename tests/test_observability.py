"""SURVEY 8f-4, host side: the 'rl' logger mirror (deepqnetwork_b200/utils/logging.py) against the reference's own module
executed on the same messages (rl/utils/logging.py:9-48): same logger name, level, line format, `<save_path>.log` file through a
daily TimedRotatingFileHandler, print() fallback before init.  The `[train]` report line itself is checked on the GPU
(tests/test_gpu_agent.py), the per-kernel CUDA-event profile by bench.py's roofline leg."""
import importlib.util
import logging
import os
from logging.handlers import TimedRotatingFileHandler

import pytest

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present (GPU box)')


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _run(mod, save_path, capsys):
    lg = logging.getLogger('rl')
    for h in list(lg.handlers):
        lg.removeHandler(h)
    mod.log('before init', 1, 2.5)                       # no logger yet: plain print
    printed = capsys.readouterr().out
    logger = mod.init_logging(save_path, add_std_err=False)
    assert mod.init_logging('/nonexistent/elsewhere') is logger      # singleton: the second call changes nothing
    mod.log('[train] step', 5, 'game', '3/7')
    mod.log('saving model: ', save_path)
    kinds = [(type(h), getattr(h, 'when', None)) for h in logger.handlers]
    fmts = [h.formatter._fmt for h in logger.handlers]
    for h in list(logger.handlers):
        h.close(); logger.removeHandler(h)
    lines = open(save_path + '.log').read().splitlines()
    return printed, logger.name, logger.level, kinds, fmts, [l[19:] for l in lines], [len(l[:19]) for l in lines]


def test_logging_mirror_behaves_like_the_reference_module(tmp_path, capsys):
    import deepqnetwork_b200.utils.logging as mirror_mod
    ref = _load(os.path.join(REF, 'rl', 'utils', 'logging.py'), 'ref_rl_logging')
    mirror = _load(mirror_mod.__file__, 'mirror_rl_logging')     # a fresh copy: the package's own singleton stays untouched
    a = _run(ref, str(tmp_path / 'ref'), capsys)
    b = _run(mirror, str(tmp_path / 'mirror'), capsys)
    assert a[0] == b[0] == 'before init 1 2.5\n'
    assert a[1:5] == b[1:5] and a[1] == 'rl' and a[2] == logging.DEBUG and a[3] == [(TimedRotatingFileHandler, 'D')]
    assert a[5][0] == b[5][0] == ' [INFO] [train] step 5 game 3/7'
    assert a[5][1].replace(str(tmp_path / 'ref'), 'X') == b[5][1].replace(str(tmp_path / 'mirror'), 'X')
    assert a[6] == b[6] == [19, 19]                               # '%(asctime).19s': the timestamp without milliseconds
