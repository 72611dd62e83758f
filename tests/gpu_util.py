"""Helpers shared by the -m gpu tests."""
import numpy as np
import pytest


def need_gpu():
    import ctypes
    try:
        cudart = ctypes.CDLL('libcuda.so.1')
        n = ctypes.c_int()
        if cudart.cuInit(0) != 0 or cudart.cuDeviceGetCount(ctypes.byref(n)) != 0 or n.value == 0:
            pytest.skip('no CUDA device')
    except OSError:
        pytest.skip('no CUDA driver')


def rel_err(a, b):
    """||a-b||inf / max(||b||inf, tiny): the per-tensor measure of SURVEY.md 8c."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(float(np.max(np.abs(b))), 1e-30))


def small_learner(capacity, batch=32, use_per=True, **kw):
    from deepqnetwork_b200.learner import Learner
    args = dict(height=8, width=8, channels=4, num_actions=4, num_hidden=64, batch_size=batch,
                replay_capacity=capacity, use_per=use_per, max_rows=max(batch, 128), math_mode='fp32')
    args.update(kw)
    return Learner(**args)
