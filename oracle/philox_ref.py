"""Host restatement of the learner's device random draws.  TEST INFRASTRUCTURE.

The throughput mode of the new build (no reference counterpart: the reference draws from CPython's ``random`` on the
host, rl/data/replay_sampler_priority.py:33, rl/data/replay_sampler.py:30) replaces those draws by a counter-based
Philox-4x32-10 stream on the device: draw i of step k is ``philox(counter=(i, stream, k_lo, k_hi), key=seed)`` with
stream 0 for the prioritised sampler's uniforms and 1 for the uniform sampler's rows, mapped to a 53-bit uniform the way
``random.random()`` builds one.  Restating it here lets a test feed the CPU oracle loop the very draws a
``dqn_train_steps`` replay makes, so the graph-replayed multi-step path can be compared with the oracle as well.
"""
import numpy as np

M32 = 0xFFFFFFFF


def philox4x32(c0, c1, c2, c3, k0, k1):
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
        k0 = (k0 + 0x9E3779B9) & M32
        k1 = (k1 + 0xBB67AE85) & M32
    return c0, c1, c2, c3


def u53(a, b):
    return ((a >> 5) * 67108864.0 + (b >> 6)) * (1.0 / 9007199254740992.0)


def per_uniforms(batch, counter, seed):
    """The B uniforms of the prioritised sampler for the step with device rng counter `counter`."""
    out = np.empty(batch, np.float64)
    for i in range(batch):
        r = philox4x32(i, 0, counter & M32, (counter >> 32) & M32, seed & M32, (seed >> 32) & M32)
        out[i] = u53(r[0], r[1])
    return out


def uniform_rows(batch, length, counter, seed):
    """The B rows of the stratified uniform sampler (replay_sampler.py:27-30 ranges, device draws)."""
    size = float(length) / float(batch)
    out = np.empty(batch, np.int64)
    for i in range(batch):
        lo, hi = int(i * size), int((i + 1) * size - 1.0)
        r = philox4x32(i, 1, counter & M32, (counter >> 32) & M32, seed & M32, (seed >> 32) & M32)
        span = max(hi - lo + 1, 1)
        off = min(int(u53(r[0], r[1]) * span), span - 1)
        out[i] = lo + off
    return out


def philox4x32_np(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox-4x32-10 over uint64-held 32-bit lanes (NumPy arrays or scalars)."""
    c0, c1, c2, c3 = (np.asarray(x, np.uint64) for x in (c0, c1, c2, c3))
    k0, k1 = np.uint64(k0), np.uint64(k1)
    m = np.uint64(M32)
    for _ in range(10):
        p0 = np.uint64(0xD2511F53) * c0
        p1 = np.uint64(0xCD9E8D57) * c2
        c0, c1, c2, c3 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & m, p1 & m, ((p0 >> np.uint64(32)) ^ c3 ^ k1) & m, p0 & m
        k0 = (k0 + np.uint64(0x9E3779B9)) & m
        k1 = (k1 + np.uint64(0xBB67AE85)) & m
    return c0, c1, c2, c3


def synthetic_state_rows(rows, state_bytes, seed, next_states=False):
    """Bytes of ring rows written by the device-side synthetic fill (bench / tests, SURVEY.md 8d "synthetic inputs"):
    16-byte word w of the whole ring = the four Philox outputs for counter (w_lo, w_hi, stream, 0), stream 1 for
    `states`, 2 for `next_states`, little-endian.  Lets a test check the gather at ring offsets beyond 4 GiB without
    shipping the ring to the host."""
    assert state_bytes % 16 == 0
    nvec = state_bytes // 16
    rows = np.asarray(rows, np.int64)
    out = np.empty((rows.size, state_bytes), np.uint8)
    for j, r in enumerate(rows.tolist()):
        c = np.arange(r * nvec, (r + 1) * nvec, dtype=np.uint64)
        o = philox4x32_np(c & np.uint64(M32), c >> np.uint64(32), 2 if next_states else 1, 0, seed & M32, (seed >> 32) & M32)
        words = np.stack([x.astype(np.uint32) for x in o], axis=1)  # [nvec, 4] little-endian uint4
        out[j] = words.reshape(-1).view(np.uint8)
    return out
