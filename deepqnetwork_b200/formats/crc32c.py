"""CRC-32C (Castagnoli, reflected polynomial 0x82F63B78) and the masking TensorFlow / leveldb apply to stored CRCs.

Used by the tensor-bundle checkpoint (`tf_bundle.py`): every table block and every tensor in a `tf.train.Saver`
checkpoint (rl/model/model.py:60-83,114) carries a masked CRC-32C.  Host-side file-format code; NumPy only.

The byte-serial recurrence would take minutes on the 128 MB of a Breakout checkpoint in Python, so `value()` runs the same
recurrence on K lanes at once (one NumPy gather per byte position) and joins the lane registers with the CRC's linearity:
advancing a register through m zero bytes is a 32x32 matrix over GF(2), and the matrices for 2^j * L bytes come from
repeated squaring.  Known answers (RFC 3720 B.4 / leveldb crc32c_test.cc) are pinned in tests/test_formats.py.
"""
import numpy as np

_POLY = 0x82F63B78
_MASK_DELTA = 0xA282EAD8


def _make_table():
    t = np.arange(256, dtype=np.uint32)
    for _ in range(8):
        t = np.where(t & 1, (t >> 1) ^ np.uint32(_POLY), t >> 1).astype(np.uint32)
    return t


_TABLE = _make_table()
_TABLE_LIST = [int(x) for x in _TABLE]


def _serial(reg, data):
    """Raw register update (no init / final xor) over a short bytes object."""
    t = _TABLE_LIST
    for b in data:
        reg = t[(reg ^ b) & 0xFF] ^ (reg >> 8)
    return reg


def _zero_byte_matrix():
    """Columns of the GF(2) matrix that advances the register through one zero byte."""
    return np.array([_serial(1 << j, b'\0') for j in range(32)], dtype=np.uint32)


def _mat_vec(cols, v):
    """cols: uint32[32]; v: uint32 array -> M v, element-wise over v."""
    out = np.zeros_like(v)
    for j in range(32):
        out ^= ((v >> np.uint32(j)) & np.uint32(1)) * cols[j]
    return out


def _mat_sq(cols):
    return _mat_vec(cols, cols)


def _mat_pow_bytes(nbytes):
    """Matrix advancing the register through `nbytes` zero bytes."""
    result = np.array([1 << j for j in range(32)], dtype=np.uint32)  # identity
    base = _zero_byte_matrix()
    while nbytes:
        if nbytes & 1:
            result = _mat_vec(base, result)
        base = _mat_sq(base)
        nbytes >>= 1
    return result


def extend(crc, data):
    """CRC-32C of (bytes whose CRC is `crc`) + data; extend(0, data) == value(data)."""
    buf = np.frombuffer(memoryview(data).cast('B'), dtype=np.uint8) if not isinstance(data, np.ndarray) \
        else np.ascontiguousarray(data).reshape(-1).view(np.uint8)
    n = buf.size
    reg = (crc ^ 0xFFFFFFFF) & 0xFFFFFFFF
    if n < 4096:
        return _serial(reg, buf.tobytes()) ^ 0xFFFFFFFF
    # K lanes of L bytes each (K a power of two), the n - K*L leading bytes go through the serial loop first
    K = 1
    while K * 2048 < n and K < (1 << 16):
        K *= 2
    L = n // K
    head = n - K * L
    reg = _serial(reg, buf[:head].tobytes())
    lanes = np.ascontiguousarray(buf[head:].reshape(K, L).T)  # [L, K]: one row per byte position
    regs = np.zeros(K, dtype=np.uint32)
    regs[0] = reg
    for t in range(L):
        regs = _TABLE[(regs ^ lanes[t]) & np.uint32(0xFF)] ^ (regs >> np.uint32(8))
    # join: register(a ++ b) = advance(register(a), len(b)) ^ register_from_zero(b)
    shift = _mat_pow_bytes(L)
    while regs.size > 1:
        regs = _mat_vec(shift, regs[0::2]) ^ regs[1::2]
        shift = _mat_sq(shift)
    return int(regs[0]) ^ 0xFFFFFFFF


def value(data):
    return extend(0, data)


def mask(crc):
    """leveldb/TensorFlow crc32c::Mask: rotate right by 15 bits and add a constant (stored CRCs of data that itself
    contains CRCs would otherwise be degenerate)."""
    return (((crc >> 15) | (crc << 17)) + _MASK_DELTA) & 0xFFFFFFFF


def unmask(masked):
    rot = (masked - _MASK_DELTA) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF
