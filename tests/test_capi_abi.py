"""CPU: the C-ABI library builds, loads, and exports every symbol include/dqn_b200.h declares.
No compute call is made (there is no GPU here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, 'include', 'dqn_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(dqn_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    from deepqnetwork_b200 import _capi
    lib = ctypes.CDLL(_capi.LIB_PATH)
    declared = _declared()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), 'missing export: ' + name
    assert declared == _capi.exported_symbols(), 'ctypes binding and header disagree'


def test_config_struct_matches_header_layout():
    from deepqnetwork_b200 import _capi
    # field-by-field natural alignment of the header's struct
    assert ctypes.sizeof(_capi.DqnConfig) == 176
    assert _capi.DqnConfig.learning_rate.offset == 48 and _capi.DqnConfig.seed.offset == 152
    assert _capi.DqnConfig.frame_dedup.offset == 160 and _capi.DqnConfig.frame_capacity.offset == 168


def test_create_without_gpu_fails_loudly():
    """No CPU fallback: creating a learner without a device is an error, not a silent slow path."""
    import pytest
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip('GPU present')
    except ImportError:
        pass
    from deepqnetwork_b200.learner import Learner
    with pytest.raises(Exception) as e:
        Learner(89, 80, 4)
    assert 'no CUDA device' in str(e.value) or 'CUDA' in str(e.value)
