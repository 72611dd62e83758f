import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: test needs a CUDA device (run with -m gpu on a B200)')


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


def pytest_terminal_summary(terminalreporter):
    """ReLU gates that differ between the device step and the fp64 oracle (informational: the oracle backward runs with
    the device's pattern, no tolerance depends on this count)."""
    try:
        from tests.parity_util import FLIP_LOG
    except Exception:
        return
    if not FLIP_LOG:
        return
    steps = len(FLIP_LOG)
    with_flips = [e for e in FLIP_LOG if e[2] > 0]
    terminalreporter.write_line('relu gate flips device vs fp64 oracle: %d of %d compared steps had any (total %d gates of %d)' % (
        len(with_flips), steps, sum(e[2] for e in FLIP_LOG), sum(e[3] for e in FLIP_LOG)))
    for tag, it, flips, total in with_flips[:20]:
        terminalreporter.write_line('  %s step %d: %d / %d' % (tag, it, flips, total))
