"""ctypes wrapper over oracle/sumtree_ref.c (C restatement of the reference sum tree).  TEST INFRASTRUCTURE."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, 'sumtree_ref.c')
_OUT = os.path.join(_HERE, '_build', 'libsumtree_ref.so')
_lib = None


def build(force=False):
    if force or not os.path.exists(_OUT) or os.path.getmtime(_OUT) < os.path.getmtime(_SRC):
        os.makedirs(os.path.dirname(_OUT), exist_ok=True)
        subprocess.check_call(['gcc', '-O2', '-ffp-contract=off', '-fno-fast-math', '-shared', '-fPIC', _SRC, '-o', _OUT])
    return _OUT


def lib():
    global _lib
    if _lib is None:
        l = C.CDLL(build())
        l.st_new.restype = C.c_void_p; l.st_new.argtypes = [C.c_int64]
        l.st_free.argtypes = [C.c_void_p]
        l.st_tree.restype = C.POINTER(C.c_float); l.st_tree.argtypes = [C.c_void_p]
        l.st_data.restype = C.POINTER(C.c_uint32); l.st_data.argtypes = [C.c_void_p]
        l.st_size.restype = C.c_int64; l.st_size.argtypes = [C.c_void_p]
        l.st_write.restype = C.c_int64; l.st_write.argtypes = [C.c_void_p]
        l.st_set_state.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int]
        l.st_add_many.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        l.st_update_many.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]
        l.st_sample.argtypes = [C.c_void_p, C.c_int32] + [C.c_void_p] * 5
        _lib = l
    return _lib


class SumTreeC:
    def __init__(self, capacity):
        self.capacity = int(capacity)
        self._l = lib()
        self._t = C.c_void_p(self._l.st_new(self.capacity))
        self.tree = np.ctypeslib.as_array(self._l.st_tree(self._t), shape=(2 * self.capacity - 1,))
        self.data = np.ctypeslib.as_array(self._l.st_data(self._t), shape=(self.capacity,))

    def __del__(self):
        try:
            self._l.st_free(self._t)
        except Exception:
            pass

    @property
    def size(self):
        return self._l.st_size(self._t)

    @property
    def write(self):
        return self._l.st_write(self._t)

    @property
    def total(self):
        return self.tree[0]

    def load(self, tree, data, size, write):
        self.tree[:] = tree; self.data[:] = data
        self._l.st_set_state(self._t, int(size), int(write), int(size >= self.capacity))

    def add(self, scores, data=None):
        s = np.ascontiguousarray(np.atleast_1d(scores), np.float32)
        d = None if data is None else np.ascontiguousarray(np.atleast_1d(data), np.uint32)
        self._l.st_add_many(self._t, s.size, s.ctypes.data, None if d is None else d.ctypes.data)

    def update(self, idx, scores):
        i = np.ascontiguousarray(np.atleast_1d(idx), np.int64); s = np.ascontiguousarray(np.atleast_1d(scores), np.float32)
        self._l.st_update_many(self._t, i.size, i.ctypes.data, s.ctypes.data)

    def sample(self, uniforms):
        u = np.ascontiguousarray(uniforms, np.float64); b = u.size
        t = np.empty(b, np.int64); d = np.empty(b, np.int64); p = np.empty(b, np.float64); m = np.empty(b, np.int64)
        self._l.st_sample(self._t, b, u.ctypes.data, t.ctypes.data, d.ctypes.data, p.ctypes.data, m.ctypes.data)
        return t, d, p, m
