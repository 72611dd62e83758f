"""CPU tests of the hand-written on-disk formats: CRC-32C, the TensorFlow tensor-bundle checkpoint
(`formats/tf_bundle.py`) and the HDF5 replay file (`formats/hdf5.py`, `data/replay_memory_disk.py`).

Neither TensorFlow nor h5py exists here and the reference holds no files written by them: PARITY UNPINNED.  What these tests
do pin: the CRC against its published known answers; protobuf / table / superblock bytes against byte strings assembled by hand
from the format descriptions; every redundancy the formats carry (checksums, magic numbers, sorted keys); reader paths for the
structures libhdf5 produces that the writer itself never emits (continuation blocks, version-3 attributes, version-2
dataspaces, late-allocated and compact datasets); and the ring semantics of ReplayMemoryDisk against the reference's own class
executed over an in-memory stand-in for h5py.
"""
import os
import struct
import sys
import types

import numpy as np
import pytest

from deepqnetwork_b200.formats import crc32c, hdf5, tf_bundle as tb

REF = '/root/reference'


# -- CRC-32C -------------------------------------------------------------------------------------------------------------------
def test_crc32c_known_answers():
    # RFC 3720 B.4 (also leveldb/TensorFlow crc32c_test.cc) and the classic check value
    assert crc32c.value(b'123456789') == 0xE3069283
    assert crc32c.value(bytes(32)) == 0x8A9136AA
    assert crc32c.value(b'\xff' * 32) == 0x62A8AB43
    assert crc32c.value(bytes(range(32))) == 0x46DD794E
    assert crc32c.value(bytes(range(31, -1, -1))) == 0x113FDB5C
    assert crc32c.value(b'') == 0
    # leveldb's Mask: rotate right 15, add 0xa282ead8; a mask is never the identity and unmask inverts it
    crc = crc32c.value(b'foo')
    assert crc32c.mask(crc) != crc and crc32c.mask(crc32c.mask(crc)) != crc
    assert crc32c.unmask(crc32c.mask(crc)) == crc and crc32c.unmask(crc32c.unmask(crc32c.mask(crc32c.mask(crc)))) == crc


def test_crc32c_lane_parallel_path_equals_the_byte_serial_recurrence():
    rng = np.random.default_rng(0)
    for n in (4095, 4096, 4097, 6000, 100003, 1 << 20, (1 << 20) + 12345):
        d = rng.integers(0, 256, n, dtype=np.uint8)
        serial = crc32c._serial(0xFFFFFFFF, d.tobytes()) ^ 0xFFFFFFFF
        assert crc32c.value(d) == serial, n
        assert crc32c.extend(crc32c.value(d[:777]), d[777:]) == serial, n   # Extend(crc(a), b) == crc(a + b)
    f = rng.normal(size=(300, 70)).astype(np.float32)
    assert crc32c.value(f) == crc32c.value(f.tobytes())


# -- tensor bundle ---------------------------------------------------------------------------------------------------------------
def test_bundle_protos_match_hand_assembled_bytes():
    # BundleHeaderProto{num_shards: 1, version{producer: 1}}: 08 01 | 1a 02 (08 01)
    assert tb.encode_header() == bytes.fromhex('0801' '1a02' '0801')
    # BundleEntryProto{dtype: DT_INT32(3), shape{}, size: 4, crc32c: 0x11223344}: offset 0 is the proto3 default and omitted
    assert tb.encode_entry('int32', (), 0, 4, 0x11223344) == bytes.fromhex('0803' '1200' '2804' '3544332211')
    # {dtype: DT_FLOAT(1), shape{dim{size: 2} dim{size: 3}}, offset: 16, size: 24, crc32c: 1}
    assert tb.encode_entry('float32', (2, 3), 16, 24, 1) == bytes.fromhex('0801' '1208' '12020802' '12020803' '2010' '2818' '3501000000')
    e = tb.decode_entry(bytes.fromhex('0801' '1208' '12020802' '12020803' '2010' '2818' '3501000000'))
    assert (e['dtype'], e['shape'], e['offset'], e['size'], e['crc32c']) == (1, (2, 3), 16, 24, 1)
    # an entry as another writer may order it (fields in any order, shard_id present, unknown field 15 skipped)
    e = tb.decode_entry(bytes.fromhex('2804' '1800' '7801' '0809' '1200'))
    assert (e['dtype'], e['shape'], e['size'], e['shard_id']) == (9, (), 4, 0)
    assert tb.decode_header(bytes.fromhex('0801' '1000' '1a04' '0801' '1000')) == {'num_shards': 1, 'endianness': 0, 'producer': 1}


def test_table_bytes_of_a_two_entry_index_by_hand():
    hdr, ent = tb.encode_header(), tb.encode_entry('int32', (), 0, 4, 0x11223344)
    table = tb.build_table([(b'', hdr), (b'a', ent)])
    # data block: (shared 0, unshared 0, value 6, "", header) (shared 0, unshared 1, value 11, "a", entry); restarts [0]; count 1
    block = bytes([0, 0, len(hdr)]) + hdr + bytes([0, 1, len(ent)]) + b'a' + ent + struct.pack('<II', 0, 1)
    def trailer(b):
        return b'\0' + struct.pack('<I', crc32c.mask(crc32c.value(b + b'\0')))
    assert table[:len(block) + 5] == block + trailer(block)
    meta_off = len(block) + 5
    meta = struct.pack('<II', 0, 1)                              # empty metaindex block
    assert table[meta_off:meta_off + 13] == meta + trailer(meta)
    idx_off = meta_off + 13
    # index block: one entry, key = short successor of "a" = "b", value = handle(offset 0, size len(block))
    idx = bytes([0, 1, 2]) + b'b' + bytes([0, len(block)]) + struct.pack('<II', 0, 1)
    assert table[idx_off:idx_off + len(idx) + 5] == idx + trailer(idx)
    footer = table[idx_off + len(idx) + 5:]
    assert len(footer) == 48 and footer[:4] == bytes([meta_off, 8, idx_off, len(idx)]) and footer[4:40] == bytes(36)
    assert footer[40:] == bytes.fromhex('57fb808b247547db')      # kTableMagicNumber 0xdb4775248b80fb57, little endian
    assert tb.read_table(table) == [(b'', hdr), (b'a', ent)]


def test_table_with_many_blocks_and_prefix_compression(monkeypatch):
    monkeypatch.setattr(tb, 'BLOCK_SIZE', 200)
    items = [(b'', b'h')] + [(('scope/layer_%03d/kernel' % i).encode(), bytes([i]) * (i % 7 + 1)) for i in range(120)]
    table = tb.build_table(items)
    assert tb.read_table(table) == items
    # separators in the index block are shortened but still separate the blocks
    footer = table[-48:]
    _, p = tb.get_varint(footer, 0); _, p = tb.get_varint(footer, p)
    ioff, p = tb.get_varint(footer, p); isize, p = tb.get_varint(footer, p)
    index = list(tb._block_entries(tb._read_block(table, ioff, isize)))
    assert len(index) >= 8
    prev_last = None
    for sep, handle in index:
        off, q = tb.get_varint(handle, 0); size, q = tb.get_varint(handle, q)
        keys = [k for k, _ in tb._block_entries(tb._read_block(table, off, size))]
        assert keys[-1] <= sep and (prev_last is None or prev_last < keys[0])
        prev_last = sep
    with pytest.raises(ValueError):
        tb.build_table([(b'b', b''), (b'a', b'')])


def test_bundle_round_trip_and_corruption_is_detected(tmp_path):
    rng = np.random.default_rng(1)
    tensors = {'q_networks/online/conv2d/kernel': rng.normal(size=(8, 8, 4, 32)).astype(np.float32),
               'q_networks/online/dense/kernel': rng.normal(size=(7680, 64)).astype(np.float32),
               'train/step': np.int32(1234), 'game_count': np.int32(7), 'train/beta1_power': np.float32(0.9 ** 5),
               'empty': np.zeros((2, 0, 3), np.float32), 'half': rng.normal(size=5).astype(np.float16),
               'flags': np.array([True, False]), 'big': np.int64(1 << 40)}
    prefix = str(tmp_path / 'rl.game_environment.BreakoutEnvironment')
    tb.write_bundle(prefix, tensors)
    tb.write_checkpoint_state(prefix)
    assert sorted(os.listdir(tmp_path)) == ['checkpoint', 'rl.game_environment.BreakoutEnvironment.data-00000-of-00001',
                                            'rl.game_environment.BreakoutEnvironment.index']
    assert open(tmp_path / 'checkpoint').read() == ('model_checkpoint_path: "rl.game_environment.BreakoutEnvironment"\n'
                                                    'all_model_checkpoint_paths: "rl.game_environment.BreakoutEnvironment"\n')
    r = tb.BundleReader(prefix)
    assert r.names() == sorted(tensors, key=str.encode) and r.header == {'num_shards': 1, 'endianness': 0, 'producer': 1}
    off = 0
    for name in r.names():                       # data file: tensors back to back in key order, no padding
        e, v = r.entry(name), np.asarray(tensors[name])
        assert e['offset'] == off and e['size'] == v.nbytes and e['shape'] == v.shape
        off += v.nbytes
        got = r.tensor(name)
        assert got.dtype == v.dtype and got.shape == v.shape and np.array_equal(got, v)
    assert os.path.getsize(tb.data_path(prefix)) == off
    # flip one bit of a tensor: the entry's CRC catches it
    with open(tb.data_path(prefix), 'r+b') as f:
        f.seek(r.entry('q_networks/online/dense/kernel')['offset'] + 1000); b = f.read(1); f.seek(-1, 1); f.write(bytes([b[0] ^ 4]))
    with pytest.raises(ValueError, match='checksum'):
        tb.BundleReader(prefix).tensor('q_networks/online/dense/kernel')
    assert np.array_equal(tb.BundleReader(prefix).tensor('train/step'), 1234)
    # flip one bit of the index: the block trailer catches it; a truncated index has no magic
    raw = bytearray(open(prefix + '.index', 'rb').read())
    raw[20] ^= 1
    open(prefix + '.index', 'wb').write(raw)
    with pytest.raises(ValueError, match='checksum'):
        tb.BundleReader(prefix)
    open(prefix + '.index', 'wb').write(raw[:-3])
    with pytest.raises(ValueError, match='magic'):
        tb.BundleReader(prefix)


class _HostLearner:
    """The slice of the Learner interface save_model / restore_model use, on host arrays."""

    def __init__(self, shapes, seed):
        rng = np.random.default_rng(seed)
        self.p = {(t, s): {n: rng.normal(size=shp).astype(np.float32) for n, shp in shapes.items()} for t in (0, 1) for s in (0, 1, 2)}
        self.step, self.games, self.betas = int(rng.integers(1, 1000)), int(rng.integers(1, 50)), (np.float32(0.59049), np.float32(0.995))

    def get_tower(self, tower=0, slot=0, shapes=None):
        return {n: v.copy() for n, v in self.p[(tower, slot)].items()}

    def set_tower(self, params, tower=0):
        self.p[(tower, 0)] = {n: np.array(v, np.float32) for n, v in params.items()}

    def set_param(self, name, value, tower=0, slot=0):
        self.p[(tower, slot)][name] = np.array(value, np.float32)

    def get_beta_powers(self): return self.betas
    def set_beta_powers(self, a, b): self.betas = (np.float32(a), np.float32(b))
    def get_step(self): return self.step
    def set_step(self, s): self.step = s
    def get_game_count(self): return self.games
    def set_game_count(self, n): self.games = n


def test_checkpoint_holds_the_reference_graphs_variables(tmp_path):
    """Names as the reference graph creates them: rl/deep_q_model.py:217 game_count; :243-244 + :292-343 the two towers
    (tf.layers default names conv2d, conv2d_1, conv2d_2, dense .. dense_3); :363-391 inside variable_scope('train'): step, the
    Adam slots <scope>/<var>/Adam and /Adam_1 and the two beta powers."""
    from deepqnetwork_b200.deep_q_model import tower_param_shapes
    shapes = tower_param_shapes(89, 80, 4, 4, 512, True)
    layers = ['conv2d', 'conv2d_1', 'conv2d_2', 'dense', 'dense_1', 'dense_2', 'dense_3']
    assert list(shapes) == [l + s for l in layers for s in ('/kernel', '/bias')]
    a = _HostLearner(shapes, 1)
    prefix = str(tmp_path / 'ckpt')
    tb.save_model(a, prefix, shapes)
    r = tb.BundleReader(prefix)
    tower_vars = [l + s for l in layers for s in ('/bias', '/kernel')]
    expect = (['game_count'] + ['q_networks/online/' + v for v in tower_vars] + ['q_networks/target/' + v for v in tower_vars]
              + ['train/beta1_power', 'train/beta2_power']
              + ['train/q_networks/online/' + v + s for v in tower_vars for s in ('/Adam', '/Adam_1')] + ['train/step'])
    assert r.names() == sorted(expect, key=str.encode) and len(expect) == 60
    assert r.entry('train/step')['dtype'] == 3 and r.entry('game_count')['dtype'] == 3 and r.entry('train/beta1_power')['shape'] == ()
    assert r.entry('q_networks/target/dense/kernel')['shape'] == (7680, 512) and r.entry('q_networks/online/conv2d/kernel')['shape'] == (8, 8, 4, 32)
    b = _HostLearner(shapes, 2)
    assert tb.restore_model(b, prefix, shapes) is True
    for key in a.p:
        if key == (1, 1) or key == (1, 2):
            continue  # the target tower has no optimizer slots
        for n in shapes:
            assert np.array_equal(a.p[key][n], b.p[key][n]), (key, n)
    assert (b.step, b.games, b.betas) == (a.step, a.games, a.betas)
    # Momentum graphs carry one slot and no beta powers (deep_q_model.py:384-387)
    tb.save_model(a, prefix + 'm', shapes, use_momentum=True)
    names = tb.BundleReader(prefix + 'm').names()
    assert 'train/q_networks/online/dense/kernel/Momentum' in names and not [n for n in names if 'Adam' in n or 'beta' in n]
    # Saver.restore semantics: a missing variable or another shape is an error, not a partial load
    with pytest.raises(KeyError):
        tb.restore_model(b, prefix + 'm', shapes)
    other = tower_param_shapes(89, 80, 4, 9, 512, True)
    with pytest.raises(ValueError, match='shape'):
        tb.restore_model(_HostLearner(other, 3), prefix, other)


# -- HDF5 ----------------------------------------------------------------------------------------------------------------------
def _replay_file(path, n=20, h=3, w=2, c=4):
    return hdf5.create_file(path, {'cur_idx': 0, 'is_full': False, 'max_size': n, 'input_height': h, 'input_width': w,
                                   'input_channels': c, 'state_type': 'uint8'},
                            {'states': ((n, h, w, c), 'uint8'), 'actions': ((n,), 'uint8'), 'rewards': ((n,), 'uint8'),
                             'next_states': ((n, h, w, c), 'uint8'), 'continues': ((n,), 'bool'), 'losses': ((n,), 'float16')})


def test_hdf5_superblock_and_structures_by_hand(tmp_path):
    path = str(tmp_path / 'r.hdf5')
    specs = _replay_file(path)
    raw = open(path, 'rb').read()
    eof = max(s['address'] + (s['nbytes'] + 7) // 8 * 8 for s in specs.values())
    assert len(raw) == eof
    # version-0 superblock as every libhdf5 "earliest" file starts: signature, five version bytes, 8-byte offsets and lengths,
    # leaf K 4, internal K 16, flags 0, base 0, free-space UNDEF, EOF, driver UNDEF, root entry (header at 96, cached B-tree + heap)
    assert raw[:24] == bytes.fromhex('894844460d0a1a0a' '0000000000080800' '0400100000000000')
    base, free, eof_addr, drv = struct.unpack_from('<QQQQ', raw, 24)
    assert (base, free, eof_addr, drv) == (0, hdf5.UNDEF, eof, hdf5.UNDEF)
    name_off, root, cache, _, btree, heap = struct.unpack_from('<QQIIQQ', raw, 56)
    assert (name_off, root, cache) == (0, 96, 1)
    # root object header: version 1, 8 messages (symbol table + 7 attributes), one reference; first message = symbol table
    ver, nmsg, refs, size = struct.unpack_from('<BxHII', raw, 96)
    assert (ver, nmsg, refs) == (1, 8, 1) and btree == 96 + 16 + size
    assert struct.unpack_from('<HHB3xQQ', raw, 112) == (0x11, 16, 0, btree, heap)
    # B-tree node of 544 bytes (2K+1 keys, 2K children, K = 16), then the local heap
    assert raw[btree:btree + 8] == b'TREE' + bytes([0, 0, 1, 0]) and heap == btree + 544
    left, right, key0, child0, key1 = struct.unpack_from('<QQQQQ', raw, btree + 8)
    assert (left, right, key0) == (hdf5.UNDEF, hdf5.UNDEF, 0)
    hsig, hver, hsize, hfree, hdata = struct.unpack_from('<4sB3xQQQ', raw, heap)
    assert (hsig, hver, hdata) == (b'HEAP', 0, heap + 32)
    names = raw[hdata:hdata + hsize]
    assert names[:8] == bytes(8) and names[key1:].split(b'\0')[0] == b'states'      # key 1 = the largest name of the child
    nxt, fsize = struct.unpack_from('<QQ', names, hfree)
    assert nxt == 1 and hfree + fsize == hsize                                        # free list: one block, H5HL_FREE_NULL ends it
    # symbol node: six entries in name order, each pointing at a version-1 object header
    assert raw[child0:child0 + 8] == b'SNOD' + bytes([1, 0, 6, 0])
    got = []
    for j in range(6):
        off, hdr = struct.unpack_from('<QQ', raw, child0 + 8 + 40 * j)
        got.append(names[off:].split(b'\0')[0].decode())
        assert hdr == specs[got[-1]]['header'] and raw[hdr] == 1
    assert got == ['actions', 'continues', 'losses', 'next_states', 'rewards', 'states']
    # the variable-length string lives in a global heap collection: object 1 = "uint8", object 0 = the free space
    g = raw.index(b'GCOL')
    assert struct.unpack_from('<B3xQ', raw, g + 4) == (1, 4096)
    assert struct.unpack_from('<HH4xQ', raw, g + 16) == (1, 0, 5) and raw[g + 32:g + 37] == b'uint8'
    assert struct.unpack_from('<HH4xQ', raw, g + 40) == (0, 0, 4096 - 40)


def test_hdf5_datatype_messages_by_hand():
    # fixed-point: class 0 version 1; bit 3 of the first flag byte = signed; size; bit offset, precision
    assert hdf5.encode_datatype('int64') == bytes.fromhex('10080000' '08000000' '0000' '4000')
    assert hdf5.encode_datatype('uint8') == bytes.fromhex('10000000' '01000000' '0000' '0800')
    # IEEE half: class 1; mantissa normalisation "implied" (0x20), sign bit 15; exponent at 10 (5 bits), mantissa at 0 (10 bits), bias 15
    assert hdf5.encode_datatype('float16') == bytes.fromhex('11200f00' '02000000' '0000' '1000' '0a05000a' '0f000000')
    assert hdf5.encode_datatype('float32') == bytes.fromhex('11201f00' '04000000' '0000' '2000' '17080017' '7f000000')
    # h5py's numpy-bool: enumeration (class 8) of 2 members over int8, names padded to 8 bytes, values 0 and 1
    assert hdf5.encode_datatype('bool') == (bytes.fromhex('18020000' '01000000') + bytes.fromhex('10080000' '01000000' '0000' '0800')
                                            + b'FALSE\0\0\0' + b'TRUE\0\0\0\0' + bytes([0, 1]))
    # variable-length UTF-8 string: class 9, type 1 (string), character set 1, 16-byte elements, base type unsigned char
    assert hdf5.encode_datatype('vlen_str') == bytes.fromhex('19010100' '10000000') + hdf5.encode_datatype('uint8')
    for name in ('int64', 'uint8', 'float16', 'float32', 'float64', 'int8', 'bool'):
        d, used = hdf5.decode_datatype(hdf5.encode_datatype(name))
        assert d['np'] == np.dtype(name) and used == len(hdf5.encode_datatype(name))
    d, _ = hdf5.decode_datatype(hdf5.encode_datatype('vlen_str'))
    assert d['vlen_str'] and d['utf8']
    d, _ = hdf5.decode_datatype(bytes.fromhex('10090000' '04000000' '0000' '2000'))   # big-endian int32
    assert d['np'] == np.dtype('>i4')


def test_hdf5_round_trip_attrs_and_datasets(tmp_path):
    path = str(tmp_path / 'r.hdf5')
    _replay_file(path)
    with hdf5.File(path, 'r+') as f:
        a = f.attrs
        assert list(a) == ['cur_idx', 'is_full', 'max_size', 'input_height', 'input_width', 'input_channels', 'state_type']
        assert a['cur_idx'] == 0 and a['cur_idx'].dtype == np.int64 and a['is_full'].dtype == np.bool_ and not a['is_full']
        assert a['max_size'] == 20 and a['state_type'] == 'uint8' and isinstance(a['state_type'], str)
        assert {k: (v.shape, v.dtype) for k, v in f.datasets.items()} == {
            'states': ((20, 3, 2, 4), np.uint8), 'next_states': ((20, 3, 2, 4), np.uint8), 'actions': ((20,), np.uint8),
            'rewards': ((20,), np.uint8), 'continues': ((20,), np.bool_), 'losses': ((20,), np.float16)}
        assert not f.array('states').any()                       # allocated storage reads as zeros
        st = f.array('states', 'r+'); st[3] = 7; st.flush()
        ls = f.array('losses', 'r+'); ls[:] = np.arange(20, dtype=np.float16) / 8; ls.flush()
        f.set_attr('cur_idx', 13); f.set_attr('is_full', True)
        with pytest.raises(NotImplementedError):
            f.set_attr('state_type', 'float32')
    with hdf5.File(path) as f:
        assert f.attrs['cur_idx'] == 13 and f.attrs['is_full'] and f.attrs['max_size'] == 20
        assert (f.array('states')[3] == 7).all() and not f.array('states')[4].any()
        assert np.array_equal(f.array('losses'), np.arange(20, dtype=np.float16) / 8)
    with open(path, 'r+b') as fh:
        fh.write(b'\x89HDG')
    with pytest.raises(ValueError, match='not an HDF5'):
        hdf5.File(path)


def _libhdf5_style_file(path):
    """A file laid out the way libhdf5 lays out an h5py-written one, with the structures our writer never emits: the root
    header holds only the symbol-table message and a continuation; the continuation block holds a NIL message, attributes as
    version-3 messages (no padding, character-set byte) with version-2 dataspaces, and a fixed-length string attribute; the
    dataset headers carry old-fill, modification-time messages and maximum dimensions; one dataset is late-allocated (undefined
    address), one is compact.  All offsets computed here, by hand."""
    M = hdf5._message
    def attr3(name, dtype_body, data):
        nm = name.encode() + b'\0'
        space = struct.pack('<BBBB', 2, 0, 0, 0)                       # dataspace version 2: rank 0, flags 0, type 0 = scalar
        return M(hdf5.MSG_ATTRIBUTE, struct.pack('<BBHHHB', 3, 0, len(nm), len(dtype_body), len(space), 0) + nm + dtype_body + space + data)
    cont = b''.join([
        M(hdf5.MSG_NIL, bytes(16)),
        attr3('cur_idx', hdf5.encode_datatype('int64'), struct.pack('<q', 5)),
        attr3('is_full', hdf5.encode_datatype('bool'), b'\x01'),
        attr3('max_size', hdf5.encode_datatype('int32'), struct.pack('<i', 6)),
        attr3('state_type', struct.pack('<BBBBI', 0x13, 0x00, 0, 0, 5), b'uint8'),        # fixed-length ASCII string
    ])
    root_msgs = lambda btree, heap, caddr: [M(hdf5.MSG_SYMBOL_TABLE, struct.pack('<QQ', btree, heap)),
                                            M(hdf5.MSG_CONTINUATION, struct.pack('<QQ', caddr, len(cont)))]
    def header(msgs, total):
        data = b''.join(msgs)
        return struct.pack('<BxHII4x', 1, total, 1, len(data)) + data
    def dset(shape, dt, layout_body, maxdims=True):
        space = struct.pack('<BBB5x', 1, len(shape), 1 if maxdims else 0) + b''.join(struct.pack('<Q', s) for s in shape)
        if maxdims:
            space += b''.join(struct.pack('<Q', s) for s in shape)
        msgs = [M(hdf5.MSG_DATASPACE, space), M(hdf5.MSG_DATATYPE, hdf5.encode_datatype(dt), 1),
                M(hdf5.MSG_FILL_OLD, struct.pack('<I', 0)), M(hdf5.MSG_FILL, struct.pack('<BBBB', 2, 2, 2, 0), 1),
                M(hdf5.MSG_LAYOUT, layout_body), M(0x12, struct.pack('<B3xI', 1, 1600000000)), M(hdf5.MSG_NIL, bytes(24))]
        return header(msgs, len(msgs))
    names = [b'actions', b'losses', b'states']
    heap = bytearray(8)
    offs = []
    for n in names:
        offs.append(len(heap)); heap += hdf5._pad8(n + b'\0')
    heap += struct.pack('<QQ', 1, 16)
    root_len = len(header(root_msgs(0, 0, 0), 7))
    btree = 96 + root_len
    heap_addr = btree + 544
    snod = heap_addr + 32 + len(heap)
    caddr = snod + 328
    pos = caddr + len(cont)
    hdrs = {}
    actions_data = bytes([3, 1, 4, 1, 5, 9])
    bodies = {'actions': lambda addr: struct.pack('<BBH', 3, 0, 6) + actions_data,            # compact
              'losses': lambda addr: struct.pack('<BBQQ', 3, 1, hdf5.UNDEF, 12),               # contiguous, never written
              'states': lambda addr: struct.pack('<BBQQ', 3, 1, addr, 6 * 2 * 2 * 4)}
    shapes = {'actions': ((6,), 'uint8'), 'losses': ((6,), 'float16'), 'states': ((6, 2, 2, 4), 'uint8')}
    for n in ('actions', 'losses', 'states'):
        hdrs[n] = pos
        pos += len(dset(shapes[n][0], shapes[n][1], bodies[n](0)))
    states_addr = pos
    eof = states_addr + 96
    out = bytearray(hdf5.SIGNATURE + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0))
    out += struct.pack('<QQQQ', 0, hdf5.UNDEF, eof, hdf5.UNDEF) + struct.pack('<QQIIQQ', 0, 96, 1, 0, btree, heap_addr)
    out += header(root_msgs(btree, heap_addr, caddr), 7)            # 2 messages here + 5 in the continuation block
    node = b'TREE' + struct.pack('<BBHQQ', 0, 0, 1, hdf5.UNDEF, hdf5.UNDEF) + struct.pack('<QQQ', 0, snod, offs[-1])
    out += node + bytes(544 - len(node))
    out += b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap), len(heap) - 16, heap_addr + 32) + heap
    s = b'SNOD' + struct.pack('<BxH', 1, 3) + b''.join(struct.pack('<QQII16x', o, hdrs[n.decode()], 0, 0) for o, n in zip(offs, names))
    out += s + bytes(328 - len(s))
    assert len(out) == caddr
    out += cont
    for n in ('actions', 'losses', 'states'):
        assert len(out) == hdrs[n]
        out += dset(shapes[n][0], shapes[n][1], bodies[n](states_addr))
    out += bytes(range(96))
    open(path, 'wb').write(out)
    return actions_data


def test_hdf5_reader_handles_libhdf5_style_structures(tmp_path):
    path = str(tmp_path / 'h5py_like.hdf5')
    actions = _libhdf5_style_file(path)
    with hdf5.File(path, 'r+') as f:
        a = f.attrs
        assert a['cur_idx'] == 5 and a['is_full'] == True and a['max_size'] == 6 and a['max_size'].dtype == np.int32  # noqa: E712
        assert a['state_type'] == b'uint8'
        assert f.array('actions').tolist() == list(actions)
        assert f.array('states', 'r').reshape(-1).tolist() == list(range(96))
        assert not f.datasets['losses'].allocated and f.array('losses', 'r').tolist() == [0.0] * 6
        with pytest.raises(NotImplementedError, match='no storage'):
            f.array('losses', 'r+')
        f.set_attr('cur_idx', 2); f.set_attr('max_size', 6)
    with hdf5.File(path) as f:
        assert f.attrs['cur_idx'] == 2 and f.attrs['is_full'] == True  # noqa: E712
    # formats this reader does not cover are refused by name, not misread
    raw = bytearray(open(path, 'rb').read())
    raw[8] = 2
    open(path, 'wb').write(raw)
    with pytest.raises(NotImplementedError, match='superblock version 2'):
        hdf5.File(path)


# -- ReplayMemoryDisk against the reference's class ---------------------------------------------------------------------------
class _FakeH5File:
    """In-memory stand-in for h5py.File with the calls rl/data/replay_memory_disk.py makes; contents survive reopen."""
    store = {}

    def __init__(self, fn, mode):
        assert mode == 'a'
        self._d = _FakeH5File.store.setdefault(fn, {'attrs': {}, 'data': {}})
        self.attrs = self._d['attrs']
        open(fn, 'a').close()   # the reference decides "new file" with os.path.exists before opening

    def create_dataset(self, name, shape, dtype):
        self._d['data'][name] = np.zeros(shape, dtype)

    def __getitem__(self, name):
        return self._d['data'][name]

    def close(self):
        pass


@pytest.mark.skipif(not os.path.isdir(REF), reason='reference tree not present (GPU box)')
def test_replay_memory_disk_ring_semantics_match_the_reference_class(tmp_path, monkeypatch):
    fake = types.ModuleType('h5py'); fake.File = _FakeH5File
    monkeypatch.setitem(sys.modules, 'h5py', fake)
    monkeypatch.syspath_prepend(REF)
    for m in [m for m in sys.modules if m == 'rl' or m.startswith('rl.')]:
        monkeypatch.delitem(sys.modules, m)
    try:
        from rl.data.replay_memory import ReplayMemory as RefMemory
        from rl.data.replay_memory_disk import ReplayMemoryDisk as RefDisk
        from deepqnetwork_b200.data.replay_memory_disk import ReplayMemoryDisk
        H, W, C, N = 3, 2, 4, 20
        rng = np.random.default_rng(4)
        src = RefMemory(H, W, C, max_size=13)
        for i in range(13):
            src.append(state=rng.integers(0, 256, (H, W, C), dtype=np.uint8), action=i % 4, reward=i % 3,
                       next_state=rng.integers(0, 256, (H, W, C), dtype=np.uint8), cont=i % 5 != 0, loss=rng.random())
        ref_fn, my_fn = str(tmp_path / 'ref.hdf5'), str(tmp_path / 'mine.hdf5')

        def same(ref, mine):
            assert (len(ref), ref.cur_idx, bool(ref.is_full), ref.max_size) == (len(mine), mine.cur_idx, mine.is_full, mine.max_size)
            for name in ('states', 'actions', 'rewards', 'next_states', 'continues', 'losses'):
                a, b = np.asarray(getattr(ref, name)), np.asarray(getattr(mine, name))
                assert a.dtype == b.dtype and np.array_equal(a, b), name

        ref, mine = RefDisk(ref_fn, H, W, C, 'uint8', max_size=N, cache_size=0), ReplayMemoryDisk(my_fn, H, W, C, 'uint8', max_size=N, cache_size=0)
        same(ref, mine)
        ref.extend(src); mine.extend(src); same(ref, mine)                       # 13 rows
        ref.close(); mine.close()
        # reopened without geometry (deep_q_agent.py:218): the file dictates it; extending wraps the ring and sets is_full
        ref, mine = RefDisk(ref_fn, cache_size=0), ReplayMemoryDisk(my_fn, cache_size=0)
        same(ref, mine)
        assert (mine.input_height, mine.input_width, mine.input_channels, mine.state_type) == (H, W, C, 'uint8')
        ref.extend(src); mine.extend(src); same(ref, mine)                       # 26 rows into 20: wrapped
        assert mine.is_full and mine.cur_idx == 6 and len(mine) == N
        row = dict(state=np.full((H, W, C), 9, np.uint8), action=3, reward=2, next_state=np.full((H, W, C), 8, np.uint8), cont=True, loss=0.25)
        ref.append(**row); mine.append(**row); same(ref, mine)
        ref.set(2, action=1, loss=0.5); mine.set(2, action=1, loss=0.5); same(ref, mine)
        r, m = ref.get(6), mine.get(6)
        assert set(r) == set(m) and all(np.array_equal(r[k], m[k]) for k in r)
        dst_r, dst_m = RefMemory(H, W, C, max_size=4), RefMemory(H, W, C, max_size=4)
        ref.copy(6, dst_r, 1); mine.copy(6, dst_m, 1)
        assert np.array_equal(dst_r.states, dst_m.states) and np.array_equal(dst_r.losses, dst_m.losses)
        ref.clear(); mine.clear(); same(ref, mine)
        ref.close(); mine.close()
        # an existing file keeps its own geometry whatever the constructor is told (the 'a' open of :37-55)
        mine = ReplayMemoryDisk(my_fn, 84, 84, 4, 'uint8', max_size=1000)
        assert (mine.max_size, mine.input_height, len(mine)) == (N, H, 0)
        mine.close()
    finally:
        for m in [m for m in sys.modules if m == 'rl' or m.startswith('rl.')]:
            sys.modules.pop(m, None)
        _FakeH5File.store.clear()


def test_inspection_tool_lists_a_save_dir_and_flags_damage(tmp_path):
    import io
    from deepqnetwork_b200.data.replay_memory_disk import ReplayMemoryDisk
    from deepqnetwork_b200.formats.__main__ import describe
    prefix = str(tmp_path / 'rl.game_environment.BreakoutEnvironment')
    tb.write_bundle(prefix, {'q_networks/online/dense/kernel': np.ones((6, 4), np.float32), 'train/step': np.int32(5)})
    ReplayMemoryDisk(prefix + '_replay_memory.hdf5', 4, 4, 4, 'uint8', max_size=8).close()
    out = io.StringIO()
    assert describe(prefix, out) is True
    text = out.getvalue()
    assert 'q_networks/online/dense/kernel' in text and 'crc ok  = 5' in text and "attr state_type       = 'uint8'" in text
    assert 'dataset next_states  uint8    (8, 4, 4, 4)' in text
    with open(tb.data_path(prefix), 'r+b') as f:
        f.seek(10); f.write(b'\xff')
    out = io.StringIO()
    assert describe(prefix, out) is False and 'UNREADABLE' in out.getvalue()
