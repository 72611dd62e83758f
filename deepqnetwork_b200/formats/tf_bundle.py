"""TensorFlow tensor-bundle ("checkpoint V2") reader / writer, hand-written: TensorFlow is not installable here.

What it replaces: `tf.train.Saver().save / .restore` as the reference's Model uses them (rl/model/model.py:60-83, saver made
at :114) -- `<prefix>.index` + `<prefix>.data-00000-of-00001` (+ the text `checkpoint` state file in the directory).  The
`.meta` MetaGraphDef is *not* written: `Saver.restore` never reads it, and there is no TF graph here to describe.

Format, restated from the published TensorFlow sources (tensorflow/core/util/tensor_bundle/tensor_bundle.{h,cc},
core/lib/io/{table_builder,block_builder,format}.cc, core/protobuf/tensor_bundle.proto; TF 1.12, the version the
reference pins in its Pipfile:13):

* `.data-00000-of-00001`: the tensors' raw little-endian bytes back to back (no alignment padding), in the order of the keys.
* `.index`: a leveldb-format sorted string table.  Key "" -> BundleHeaderProto {num_shards=1, endianness=LITTLE,
  version{producer=1}}; key <variable name> -> BundleEntryProto {dtype, shape, shard_id, offset, size, crc32c (masked CRC-32C
  of the tensor bytes)}.  Table = data blocks, an (empty) metaindex block, an index block, a 48-byte footer (two block handles,
  zero padding, magic 0xdb4775248b80fb57).  Block = prefix-compressed entries (varint shared / unshared / value lengths) with a
  restart array every 16 entries (1 in the index block), followed by a 5-byte trailer: compression type (0: the bundle writer
  turns compression off) and the masked CRC-32C of contents + type.

PARITY UNPINNED: no file written by TensorFlow exists in this environment and the reference holds none, so this module is
checked against itself (round trip), against the format's own redundancy (every CRC, the footer magic, sorted keys, handle
arithmetic) and against hand-assembled byte strings of the protos in tests/test_formats.py -- not against TensorFlow.
"""
import os
import struct

import numpy as np

from . import crc32c

TABLE_MAGIC = 0xDB4775248B80FB57
BLOCK_SIZE = 262144          # TF's table::Options::block_size
RESTART_INTERVAL = 16
FOOTER_LEN = 48

# tensorflow/core/framework/types.proto
DT = {'float32': 1, 'float64': 2, 'int32': 3, 'uint8': 4, 'int16': 5, 'int8': 6, 'int64': 9, 'bool': 10,
      'uint16': 17, 'float16': 19, 'uint32': 22, 'uint64': 23}
DT_INV = {v: k for k, v in DT.items()}


# -- varints / protobuf wire format ------------------------------------------------------------------------------------------
def put_varint(n):
    n &= (1 << 64) - 1
    out = bytearray()
    while n >= 0x80:
        out.append((n & 0x7F) | 0x80)
        n >>= 7
    out.append(n)
    return bytes(out)


def get_varint(buf, pos):
    shift = 0
    val = 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not b & 0x80:
            return val, pos
        shift += 7
        if shift > 63:
            raise ValueError('varint too long')


def _field(num, wire, payload):
    return put_varint((num << 3) | wire) + payload


def _parse_fields(buf):
    """-> list of (field number, wire type, value) of one serialized message."""
    pos, out = 0, []
    while pos < len(buf):
        key, pos = get_varint(buf, pos)
        num, wire = key >> 3, key & 7
        if wire == 0:
            v, pos = get_varint(buf, pos)
        elif wire == 1:
            v = struct.unpack_from('<Q', buf, pos)[0]; pos += 8
        elif wire == 2:
            n, pos = get_varint(buf, pos)
            v = bytes(buf[pos:pos + n]); pos += n
            if len(v) != n:
                raise ValueError('truncated length-delimited field')
        elif wire == 5:
            v = struct.unpack_from('<I', buf, pos)[0]; pos += 4
        else:
            raise ValueError('unsupported wire type %d' % wire)
        out.append((num, wire, v))
    return out


def encode_header(num_shards=1):
    # BundleHeaderProto: num_shards = 1 (int32), endianness = 2 (LITTLE = 0, proto3 default: omitted), version = 3 {producer = 1}
    return _field(1, 0, put_varint(num_shards)) + _field(3, 2, put_varint(2) + _field(1, 0, put_varint(1)))


def decode_header(buf):
    h = {'num_shards': 0, 'endianness': 0, 'producer': 0}
    for num, _, v in _parse_fields(buf):
        if num == 1:
            h['num_shards'] = v
        elif num == 2:
            h['endianness'] = v
        elif num == 3:
            for n2, _, v2 in _parse_fields(v):
                if n2 == 1:
                    h['producer'] = v2
    return h


def encode_entry(dtype, shape, offset, size, crc_masked, shard_id=0):
    # BundleEntryProto: dtype = 1, shape = 2 (TensorShapeProto: repeated Dim dim = 2 {int64 size = 1}), shard_id = 3,
    # offset = 4, size = 5, crc32c = 6 (fixed32); proto3: zero-valued scalars are omitted, the shape message is always present
    dims = b''.join(_field(2, 2, (lambda d: put_varint(len(d)) + d)(_field(1, 0, put_varint(s)) if s else b'')) for s in shape)
    out = _field(1, 0, put_varint(DT[dtype])) + _field(2, 2, put_varint(len(dims)) + dims)
    if shard_id:
        out += _field(3, 0, put_varint(shard_id))
    if offset:
        out += _field(4, 0, put_varint(offset))
    if size:
        out += _field(5, 0, put_varint(size))
    if crc_masked:
        out += _field(6, 5, struct.pack('<I', crc_masked))
    return out


def decode_entry(buf):
    e = {'dtype': 0, 'shape': (), 'shard_id': 0, 'offset': 0, 'size': 0, 'crc32c': 0, 'slices': 0}
    for num, _, v in _parse_fields(buf):
        if num == 1:
            e['dtype'] = v
        elif num == 2:
            shape = []
            for n2, _, v2 in _parse_fields(v):
                if n2 == 2:
                    size = 0
                    for n3, _, v3 in _parse_fields(v2):
                        if n3 == 1:
                            size = v3 - (1 << 64) if v3 >> 63 else v3
                    shape.append(size)
                elif n2 == 3 and v2:
                    raise ValueError('tensor of unknown rank in a checkpoint')
            e['shape'] = tuple(shape)
        elif num == 3:
            e['shard_id'] = v
        elif num == 4:
            e['offset'] = v
        elif num == 5:
            e['size'] = v
        elif num == 6:
            e['crc32c'] = v
        elif num == 7:
            e['slices'] += 1
    return e


# -- leveldb-format table -----------------------------------------------------------------------------------------------------
class _BlockBuilder:
    def __init__(self, restart_interval):
        self.interval = restart_interval
        self.buf = bytearray()
        self.restarts = [0]
        self.count = 0
        self.last_key = b''

    def add(self, key, value):
        shared = 0
        if self.count < self.interval:
            m = min(len(key), len(self.last_key))
            while shared < m and key[shared] == self.last_key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.count = 0
        self.buf += put_varint(shared) + put_varint(len(key) - shared) + put_varint(len(value)) + key[shared:] + value
        self.last_key = key
        self.count += 1

    def size_estimate(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def finish(self):
        return bytes(self.buf) + b''.join(struct.pack('<I', r) for r in self.restarts) + struct.pack('<I', len(self.restarts))


def _shortest_separator(a, b):
    """A key k with a <= k < b, as short as leveldb's BytewiseComparator makes it."""
    m = min(len(a), len(b))
    i = 0
    while i < m and a[i] == b[i]:
        i += 1
    if i < m and a[i] < 0xFF and a[i] + 1 < b[i]:
        return a[:i] + bytes([a[i] + 1])
    return a


def _short_successor(a):
    for i, c in enumerate(a):
        if c != 0xFF:
            return a[:i] + bytes([c + 1])
    return a


def _block_with_trailer(contents):
    trailer_type = b'\0'  # kNoCompression
    return contents + trailer_type + struct.pack('<I', crc32c.mask(crc32c.value(contents + trailer_type)))


def build_table(items):
    """items: iterable of (key bytes, value bytes), keys strictly increasing -> the bytes of the table file."""
    out = bytearray()
    index = _BlockBuilder(1)
    data = _BlockBuilder(RESTART_INTERVAL)
    pending = None  # (last key of the finished block, handle) waiting for the next key to shorten the separator
    last = None

    def flush():
        nonlocal data, pending
        contents = data.finish()
        handle = put_varint(len(out)) + put_varint(len(contents))
        out.extend(_block_with_trailer(contents))
        pending = (data.last_key, handle)
        data = _BlockBuilder(RESTART_INTERVAL)

    for key, value in items:
        if last is not None and not key > last:
            raise ValueError('table keys must be strictly increasing')
        if pending:
            index.add(_shortest_separator(pending[0], key), pending[1])
            pending = None
        data.add(key, value)
        last = key
        if data.size_estimate() >= BLOCK_SIZE:
            flush()
    if data.buf:
        flush()
    if pending:
        index.add(_short_successor(pending[0]), pending[1])
    meta = _BlockBuilder(RESTART_INTERVAL).finish()
    meta_handle = put_varint(len(out)) + put_varint(len(meta))
    out.extend(_block_with_trailer(meta))
    idx = index.finish()
    index_handle = put_varint(len(out)) + put_varint(len(idx))
    out.extend(_block_with_trailer(idx))
    footer = meta_handle + index_handle
    footer += b'\0' * (FOOTER_LEN - 8 - len(footer)) + struct.pack('<Q', TABLE_MAGIC)
    out.extend(footer)
    return bytes(out)


def _read_block(buf, offset, size, verify=True):
    if offset + size + 5 > len(buf):
        raise ValueError('block handle beyond the end of the table')
    contents = buf[offset:offset + size]
    ctype = buf[offset + size]
    stored = struct.unpack_from('<I', buf, offset + size + 1)[0]
    if verify and crc32c.unmask(stored) != crc32c.value(buf[offset:offset + size + 1]):
        raise ValueError('table block checksum mismatch at offset %d' % offset)
    if ctype != 0:
        raise NotImplementedError('compressed table block (type %d): tensor bundles are written uncompressed' % ctype)
    return contents


def _block_entries(contents):
    if len(contents) < 4:
        raise ValueError('table block too short')
    nrestarts = struct.unpack_from('<I', contents, len(contents) - 4)[0]
    end = len(contents) - 4 - 4 * nrestarts
    if end < 0:
        raise ValueError('bad restart count in table block')
    pos, key = 0, b''
    while pos < end:
        shared, pos = get_varint(contents, pos)
        unshared, pos = get_varint(contents, pos)
        vlen, pos = get_varint(contents, pos)
        if shared > len(key) or pos + unshared + vlen > end:
            raise ValueError('corrupt table block entry')
        key = key[:shared] + bytes(contents[pos:pos + unshared]); pos += unshared
        yield key, bytes(contents[pos:pos + vlen])
        pos += vlen


def read_table(buf, verify=True):
    """-> list of (key, value) of a leveldb-format table held in `buf`."""
    if len(buf) < FOOTER_LEN or struct.unpack_from('<Q', buf, len(buf) - 8)[0] != TABLE_MAGIC:
        raise ValueError('not a tensor-bundle index (bad table magic)')
    footer = buf[len(buf) - FOOTER_LEN:]
    _, pos = get_varint(footer, 0); _, pos = get_varint(footer, pos)       # metaindex handle
    ioff, pos = get_varint(footer, pos); isize, pos = get_varint(footer, pos)
    out = []
    for _, handle in _block_entries(_read_block(buf, ioff, isize, verify)):
        off, p = get_varint(handle, 0); size, p = get_varint(handle, p)
        out.extend(_block_entries(_read_block(buf, off, size, verify)))
    for a, b in zip(out, out[1:]):
        if not a[0] < b[0]:
            raise ValueError('table keys out of order')
    return out


# -- bundle ------------------------------------------------------------------------------------------------------------------
def data_path(prefix, shard=0, num_shards=1):
    return '%s.data-%05d-of-%05d' % (prefix, shard, num_shards)


def write_bundle(prefix, tensors):
    """tensors: {variable name: ndarray}.  Writes <prefix>.index and <prefix>.data-00000-of-00001 the way BundleWriter does
    for a single-shard save: names in byte order, data appended in that order."""
    names = sorted(tensors, key=lambda s: s.encode())
    items = [(b'', encode_header(1))]
    offset = 0
    tmp_data, tmp_index = data_path(prefix) + '.tempstate', prefix + '.index.tempstate'
    with open(tmp_data, 'wb') as f:
        for name in names:
            if not name:
                raise ValueError('empty tensor name')
            a = np.asarray(tensors[name])
            dt = a.dtype.name
            if dt not in DT:
                raise TypeError('dtype %s has no checkpoint encoding here' % dt)
            shape = a.shape  # before ascontiguousarray, which turns a scalar into shape (1,)
            raw = np.ascontiguousarray(a.astype(a.dtype.newbyteorder('<'), copy=False)).reshape(-1).view(np.uint8)
            f.write(raw.data)
            items.append((name.encode(), encode_entry(dt, shape, offset, raw.size, crc32c.mask(crc32c.value(raw)))))
            offset += raw.size
    with open(tmp_index, 'wb') as f:
        f.write(build_table(items))
    os.replace(tmp_data, data_path(prefix))
    os.replace(tmp_index, prefix + '.index')


class BundleReader:
    """Read-side of a checkpoint prefix: `names()`, `entry(name)`, `tensor(name)` (CRC-checked unless verify=False)."""

    def __init__(self, prefix, verify=True):
        self.prefix = prefix
        self.verify = verify
        with open(prefix + '.index', 'rb') as f:
            items = read_table(f.read(), verify)
        if not items or items[0][0] != b'':
            raise ValueError('tensor bundle has no header entry')
        self.header = decode_header(items[0][1])
        if self.header['endianness'] != 0:
            raise NotImplementedError('big-endian tensor bundle')
        self.entries = {k.decode(): decode_entry(v) for k, v in items[1:]}

    def names(self):
        return list(self.entries)

    def __contains__(self, name):
        return name in self.entries

    def entry(self, name):
        return self.entries[name]

    def tensor(self, name):
        e = self.entries[name]
        if e['slices']:
            raise NotImplementedError('partitioned variable %s (tensor slices)' % name)
        if e['dtype'] not in DT_INV:
            raise NotImplementedError('dtype enum %d of %s' % (e['dtype'], name))
        dt = np.dtype(DT_INV[e['dtype']])
        count = int(np.prod(e['shape'], dtype=np.int64)) if e['shape'] else 1
        if count * dt.itemsize != e['size']:
            raise ValueError('%s: size %d does not match shape %s of %s' % (name, e['size'], e['shape'], dt))
        with open(data_path(self.prefix, e['shard_id'], self.header['num_shards']), 'rb') as f:
            f.seek(e['offset'])
            raw = np.fromfile(f, dtype=np.uint8, count=e['size'])
        if raw.size != e['size']:
            raise ValueError('%s: data file truncated' % name)
        if self.verify and crc32c.unmask(e['crc32c']) != crc32c.value(raw):
            raise ValueError('%s: tensor checksum mismatch' % name)
        return raw.view(dt.newbyteorder('<')).astype(dt, copy=False).reshape(e['shape'])


def write_checkpoint_state(prefix):
    """The text-format CheckpointState proto `Saver.save` keeps next to the checkpoint (`checkpoint`), paths relative to
    the directory as TF writes them."""
    d, base = os.path.split(prefix)
    with open(os.path.join(d, 'checkpoint'), 'w') as f:
        f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))


# -- the reference model's variable set ---------------------------------------------------------------------------------------
ONLINE, TARGET, TRAIN = 'q_networks/online/', 'q_networks/target/', 'train/'


def model_variables(learner, shapes, use_momentum=False):
    """The variables `tf.train.Saver()` would save for the reference graph (rl/deep_q_model.py:217,243-244,380-391): both
    towers under q_networks/{online,target}/, the optimizer slots of the online tower (created inside variable_scope('train'),
    hence the train/ prefix: train/<var>/Adam, /Adam_1 or /Momentum), train/beta{1,2}_power, train/step, game_count."""
    out = {}
    for tower, scope in ((0, ONLINE), (1, TARGET)):
        for name, v in learner.get_tower(tower=tower, shapes=shapes).items():
            out[scope + name] = v
    slots = ((1, '/Momentum'),) if use_momentum else ((1, '/Adam'), (2, '/Adam_1'))
    for slot, suffix in slots:
        for name, v in learner.get_tower(tower=0, slot=slot, shapes=shapes).items():
            out[TRAIN + ONLINE + name + suffix] = v
    if not use_momentum:
        b1p, b2p = learner.get_beta_powers()
        out[TRAIN + 'beta1_power'] = np.float32(b1p)
        out[TRAIN + 'beta2_power'] = np.float32(b2p)
    out[TRAIN + 'step'] = np.int32(learner.get_step())
    out['game_count'] = np.int32(learner.get_game_count())
    return out


def save_model(learner, prefix, shapes, use_momentum=False):
    write_bundle(prefix, model_variables(learner, shapes, use_momentum))
    write_checkpoint_state(prefix)


def restore_model(learner, prefix, shapes, use_momentum=False):
    """Load a reference save into the device learner; every variable the graph defines must be present with its shape
    (`Saver.restore` fails the same way on a missing key or a shape mismatch)."""
    r = BundleReader(prefix)

    def fetch(name, shape):
        if name not in r:
            raise KeyError('checkpoint %s has no variable %s' % (prefix, name))
        v = r.tensor(name)
        if tuple(v.shape) != tuple(shape):
            raise ValueError('%s: checkpoint shape %s, model shape %s' % (name, v.shape, tuple(shape)))
        return v

    for tower, scope in ((0, ONLINE), (1, TARGET)):
        learner.set_tower({n: fetch(scope + n, s) for n, s in shapes.items()}, tower=tower)
    slots = ((1, '/Momentum'),) if use_momentum else ((1, '/Adam'), (2, '/Adam_1'))
    for slot, suffix in slots:
        for n, s in shapes.items():
            learner.set_param(n, fetch(TRAIN + ONLINE + n + suffix, s), tower=0, slot=slot)
    if not use_momentum:
        learner.set_beta_powers(float(fetch(TRAIN + 'beta1_power', ())), float(fetch(TRAIN + 'beta2_power', ())))
    learner.set_step(int(fetch(TRAIN + 'step', ())))
    learner.set_game_count(int(fetch('game_count', ())))
    return True
