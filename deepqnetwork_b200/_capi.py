"""ctypes binding of libdqn_b200.so (C ABI: include/dqn_b200.h)."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libdqn_b200.so')

DQN_CONV_REFERENCE, DQN_CONV_NATURE = 0, 1
DQN_MATH_FP32, DQN_MATH_TC3X, DQN_MATH_TC1X = 0, 1, 2


class DqnConfig(C.Structure):
    _fields_ = [
        ('struct_size', C.c_int32), ('device', C.c_int32),
        ('input_height', C.c_int32), ('input_width', C.c_int32), ('input_channels', C.c_int32),
        ('num_actions', C.c_int32), ('num_hidden', C.c_int32), ('conv_variant', C.c_int32),
        ('use_dueling', C.c_int32), ('use_double', C.c_int32), ('use_per', C.c_int32), ('use_momentum', C.c_int32),
        ('learning_rate', C.c_double), ('momentum', C.c_double), ('discount_rate', C.c_double),
        ('batch_size', C.c_int32), ('replay_capacity', C.c_int64),
        ('per_a', C.c_double), ('per_b_start', C.c_double), ('per_b_end', C.c_double),
        ('per_anneal_steps', C.c_int64), ('use_per_annealing', C.c_int32),
        ('per_calculate_steps', C.c_int64), ('copy_network_steps', C.c_int64),
        ('max_rows', C.c_int32), ('math_mode', C.c_int32), ('seed', C.c_uint64),
        ('frame_dedup', C.c_int32), ('frame_capacity', C.c_int64),
    ]


_P = C.c_void_p
_SIGS = {
    'dqn_create': [C.POINTER(DqnConfig), C.POINTER(_P)],
    'dqn_destroy': [_P],
    'dqn_sync': [_P],
    'dqn_set_stream': [_P, _P],
    'dqn_num_params': [_P, C.POINTER(C.c_int32), C.POINTER(C.c_int64)],
    'dqn_param_info': [_P, C.c_int32, C.POINTER(C.c_char_p), C.POINTER(C.c_int64), C.POINTER(C.c_int64)],
    'dqn_get_param': [_P, C.c_int32, C.c_int32, C.c_char_p, _P, C.c_int64],
    'dqn_set_param': [_P, C.c_int32, C.c_int32, C.c_char_p, _P, C.c_int64],
    'dqn_device_ptr': [_P, C.c_char_p, C.POINTER(_P), C.POINTER(C.c_int64)],
    'dqn_get_step': [_P, C.POINTER(C.c_int64)],
    'dqn_set_step': [_P, C.c_int64],
    'dqn_get_game_count': [_P, C.POINTER(C.c_int64)],
    'dqn_set_game_count': [_P, C.c_int64],
    'dqn_get_beta_powers': [_P, C.POINTER(C.c_float), C.POINTER(C.c_float)],
    'dqn_set_beta_powers': [_P, C.c_float, C.c_float],
    'dqn_copy_network': [_P],
    'dqn_get_stats': [_P, _P],
    'dqn_replay_append': [_P, C.c_int64, _P, _P, _P, _P, _P, _P],
    'dqn_replay_read': [_P, C.c_int64, _P, _P, _P, _P, _P, _P, _P],
    'dqn_replay_info': [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int32)],
    'dqn_replay_clear': [_P],
    'dqn_frames_append': [_P, C.c_int64, _P, C.POINTER(C.c_int64)],
    'dqn_frames_info': [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)],
    'dqn_preprocess': [_P, C.c_int64, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int32, C.POINTER(C.c_int64)],
    'dqn_replay_append_indexed': [_P, C.c_int64, _P, _P, _P, _P, _P],
    'dqn_replay_fill_synthetic': [_P, C.c_int64, C.c_uint64],
    'dqn_tree_add': [_P, C.c_int64, _P, _P],
    'dqn_tree_update': [_P, C.c_int64, _P, _P, C.c_int32],
    'dqn_tree_sample': [_P, C.c_int32, _P, _P, _P, _P, _P],
    'dqn_tree_retrieve': [_P, C.c_int32, _P, _P, _P, _P, _P],
    'dqn_tree_stats': [_P, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)],
    'dqn_tree_read': [_P, _P, _P, C.POINTER(C.c_int64), C.POINTER(C.c_int64)],
    'dqn_tree_info': [_P, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_float)],
    'dqn_tree_get': [_P, C.c_int64, _P, _P, _P],
    'dqn_per_sample': [_P, _P, C.c_int32, C.c_double, _P, _P, _P],
    'dqn_uniform_sample': [_P, _P],
    'dqn_batch_read': [_P, _P, _P, _P, _P, _P, _P],
    'dqn_train_batch': [_P, C.c_int32, _P, _P, _P, _P, _P, _P, C.POINTER(C.c_int64), _P, C.POINTER(C.c_float)],
    'dqn_predict': [_P, C.c_int32, _P, C.c_int32, _P],
    'dqn_get_losses': [_P, C.c_int32, _P, _P, _P, _P, _P, _P],
    'dqn_debug_read': [_P, C.c_char_p, _P, C.c_int64],
    'dqn_train_step': [_P, _P, _P, C.POINTER(C.c_int64), _P, C.POINTER(C.c_float)],
    'dqn_train_step_async': [_P, _P, _P],
    'dqn_train_step_result': [_P, C.POINTER(C.c_int64), _P, C.POINTER(C.c_float)],
    'dqn_train_steps': [_P, C.c_int64],
    'dqn_train_step_phase': [_P, C.c_int32, C.c_double],
    'dqn_phase_wait': [_P, C.c_char_p, _P],
    'dqn_set_per_normaliser': [_P, C.c_float],
    'dqn_set_option': [_P, C.c_char_p, C.c_int32],
    'dqn_launch_count': [_P, C.POINTER(C.c_int64)],
    'dqn_set_profile': [_P, C.c_int32],
    'dqn_get_profile': [_P, C.POINTER(C.c_int32), C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int32,
                        C.POINTER(C.c_int64)],
    'dqn_selftest_gemm': [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P],
    'dqn_bench_gemm': [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_float)],
    'dqn_debug_timeline': [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P],
    'dqn_save': [_P, C.c_char_p],
    'dqn_restore': [_P, C.c_char_p],
    'dqn_save_replay': [_P, C.c_char_p],
    'dqn_restore_replay': [_P, C.c_char_p, C.POINTER(C.c_int64)],
}

_lib = None


def exported_symbols():
    """Names include/dqn_b200.h declares (used by the CPU-side ABI test)."""
    return sorted(list(_SIGS) + ['dqn_last_error'])


def lib():
    """Load the CUDA library.  There is no fallback: a missing build is an error."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError('libdqn_b200.so is not built (run `python -c "import __graft_entry__ as g; g.build()"` '
                               'or `python deepqnetwork_b200/build.py`); deepqnetwork_b200 has no CPU fallback')
        l = C.CDLL(LIB_PATH)
        for name, args in _SIGS.items():
            fn = getattr(l, name)
            fn.argtypes = args
            fn.restype = C.c_int
        l.dqn_last_error.argtypes = [_P]
        l.dqn_last_error.restype = C.c_char_p
        _lib = l
    return _lib


_from_buffer = C.c_char.from_buffer


def ptr(a):
    """Host pointer of a C-contiguous numpy array (or None).

    Through the buffer protocol (0.3 us) rather than ndarray.ctypes (3-4 us per pointer, ~13 pointers per host-fed
    train step); read-only and empty arrays take the slow way."""
    if a is None:
        return None
    assert a.flags.c_contiguous, 'array must be C-contiguous'
    try:
        return C.byref(_from_buffer(a))
    except (TypeError, ValueError, BufferError):
        return a.ctypes.data_as(_P)


def as_u8(a, shape=None):
    a = np.ascontiguousarray(a)
    if a.dtype == np.bool_:
        a = a.view(np.uint8)
    elif a.dtype != np.uint8:
        # NumPy assignment semantics of the reference's uint8 ring: float rewards wrap (400.0 -> 144)
        a = a.astype(np.int64).astype(np.uint8) if a.dtype.kind == 'f' else a.astype(np.uint8)
    return a.reshape(shape) if shape is not None else a


class DqnError(RuntimeError):
    pass


def check(handle, rc):
    if rc != 0:
        msg = lib().dqn_last_error(handle)
        raise DqnError((msg or b'').decode() or 'libdqn_b200 call failed with status %d' % rc)
