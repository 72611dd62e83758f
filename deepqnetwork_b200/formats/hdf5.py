"""Minimal HDF5 reader / writer for the reference's replay file, hand-written: h5py / libhdf5 are not installable here.

What it replaces: `h5py.File(fn, 'a')`, `attrs[...]`, `create_dataset(name, shape, dtype)` and dataset row access as
rl/data/replay_memory_disk.py:37-102 uses them (the file `<prefix>_replay_memory.hdf5` written at exit by
rl/deep_q_agent.py:410-432 and reloaded by :216-221).

Scope = what h5py with its default `libver` (earliest) produces for that code, restated from the published "HDF5 File Format
Specification Version 2.0" (and the 1.10 library sources where the specification leaves a choice):

* superblock version 0 (or 1), 8-byte offsets and lengths;
* version-1 object headers (16-byte prefix, 8-byte aligned messages, continuation blocks);
* an old-style root group: symbol-table message -> version-1 B-tree ('TREE') -> symbol-table nodes ('SNOD'), names in a local
  heap ('HEAP');
* datasets with contiguous (or compact) layout messages of version 3, simple dataspaces of version 1 or 2;
* datatypes: fixed-point, IEEE floating point, fixed strings, the int8 enumeration {FALSE, TRUE} h5py uses for numpy bool,
  and variable-length strings (h5py's encoding of a Python str attribute; the characters live in a global heap 'GCOL');
* attribute messages of versions 1-3 on the root group.

Not handled (raises NotImplementedError, never guesses): superblock 2/3 and version-2 object headers (`libver='latest'`),
chunked / filtered datasets, compound / array / reference types, dense attribute storage.

A contiguous dataset is a plain byte range of the file, so datasets are exposed as `numpy.memmap`s: row reads and writes need no
further format code.  Datasets h5py created but never wrote have no storage yet (late allocation, undefined address): they
read as zeros and cannot be written in place.

PARITY UNPINNED: no file written by libhdf5 exists in this environment, so reader and writer are checked against each other and
against hand-assembled structures in tests/test_formats.py -- not against libhdf5.
"""
import os
import struct

import numpy as np

SIGNATURE = b'\x89HDF\r\n\x1a\n'
UNDEF = 0xFFFFFFFFFFFFFFFF
LEAF_K, INTERNAL_K = 4, 16
HEAP_FREE_NULL = 1           # H5HL_FREE_NULL: end of a local heap's free list
GCOL_SIZE = 4096             # minimum global heap collection size

MSG_NIL, MSG_DATASPACE, MSG_LINK_INFO, MSG_DATATYPE, MSG_FILL_OLD, MSG_FILL, MSG_LINK = 0x0, 0x1, 0x2, 0x3, 0x4, 0x5, 0x6
MSG_LAYOUT, MSG_ATTRIBUTE, MSG_CONTINUATION, MSG_SYMBOL_TABLE, MSG_ATTR_INFO = 0x8, 0xC, 0x10, 0x11, 0x15


def _pad8(b):
    return b + b'\0' * (-len(b) % 8)


# -- datatypes ---------------------------------------------------------------------------------------------------------------
_FLOAT_PROPS = {2: (15, 10, 5, 0, 10, 15), 4: (31, 23, 8, 0, 23, 127), 8: (63, 52, 11, 0, 52, 1023)}  # sign, exp loc/size, mant loc/size, bias


def encode_datatype(dt):
    """Datatype message body for a numpy dtype (bool -> h5py's {FALSE,TRUE} int8 enum) or the string 'vlen_str'."""
    if isinstance(dt, str) and dt == 'vlen_str':
        # class 9 (variable length), version 1; type = string (1), padding = null terminate (0), character set = UTF-8 (1);
        # 16-byte elements (length, heap collection address, object index); base type = 1-byte unsigned integer
        return struct.pack('<BBBBI', 0x19, 0x01, 0x01, 0x00, 16) + encode_datatype(np.dtype('uint8'))
    dt = np.dtype(dt)
    if dt == np.bool_:
        base = encode_datatype(np.dtype('int8'))
        names = b''.join(n + b'\0' * (8 - len(n) % 8 if (len(n) + 1) % 8 else 1) for n in (b'FALSE', b'TRUE'))
        return struct.pack('<BBBBI', 0x18, 2, 0, 0, 1) + base + names + bytes([0, 1])
    if dt.byteorder == '>':
        raise NotImplementedError('big-endian dataset types')
    if dt.kind in 'iu':
        return struct.pack('<BBBBIHH', 0x10, 0x08 if dt.kind == 'i' else 0x00, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == 'f' and dt.itemsize in _FLOAT_PROPS:
        sign, eloc, esize, mloc, msize, bias = _FLOAT_PROPS[dt.itemsize]
        return struct.pack('<BBBBIHHBBBBI', 0x11, 0x20, sign, 0, dt.itemsize, 0, 8 * dt.itemsize, eloc, esize, mloc, msize, bias)
    if dt.kind == 'S':
        return struct.pack('<BBBBI', 0x13, 0x00, 0, 0, dt.itemsize)
    raise NotImplementedError('no HDF5 encoding here for dtype %s' % dt)


def decode_datatype(buf, pos=0):
    """-> (description dict, bytes consumed).  description: {'np': numpy dtype or None, 'size', 'class', 'vlen_str', 'bool'}."""
    cls_ver, b0, b1, _b2, size = struct.unpack_from('<BBBBI', buf, pos)
    cls, ver = cls_ver & 0x0F, cls_ver >> 4
    d = {'class': cls, 'size': size, 'np': None, 'vlen_str': False, 'bool': False}
    order = '>' if b0 & 1 else '<'
    if cls == 0:
        d['np'] = np.dtype('%s%s%d' % (order, 'i' if b0 & 0x08 else 'u', size))
        return d, 12
    if cls == 1:
        if size not in _FLOAT_PROPS or struct.unpack_from('<BBBBI', buf, pos + 12) != _FLOAT_PROPS[size][1:] or b1 != _FLOAT_PROPS[size][0]:
            raise NotImplementedError('non-IEEE floating-point type')
        d['np'] = np.dtype('%sf%d' % (order, size))
        return d, 20
    if cls == 3:
        d['np'] = np.dtype('S%d' % size)
        d['utf8'] = (b0 >> 4) == 1
        return d, 8
    if cls == 8:
        nmemb = b0 | (b1 << 8)
        base, used = decode_datatype(buf, pos + 8)
        p = pos + 8 + used
        names = []
        for _ in range(nmemb):
            end = buf.index(b'\0', p)
            names.append(bytes(buf[p:end]))
            n = end - p
            p += (n + 8) // 8 * 8 if ver < 3 else n + 1
        values = np.frombuffer(buf, dtype=base['np'], count=nmemb, offset=p)
        p += nmemb * base['size']
        d['np'] = base['np']
        d['enum'] = dict(zip(names, values.tolist()))
        if base['size'] == 1 and d['enum'] == {b'FALSE': 0, b'TRUE': 1}:
            d['np'], d['bool'] = np.dtype(np.bool_), True
        return d, p - pos
    if cls == 9:
        if (b0 & 0x0F) != 1:
            raise NotImplementedError('variable-length sequence type')
        _, used = decode_datatype(buf, pos + 8)
        d['vlen_str'] = True
        d['utf8'] = (b1 & 0x0F) == 1
        return d, 8 + used
    raise NotImplementedError('HDF5 datatype class %d' % cls)


def encode_dataspace(shape):
    # version 1: version, rank, flags (no maximum dimensions), 5 reserved bytes, dimension sizes; rank 0 = scalar
    return struct.pack('<BBB5x', 1, len(shape), 0) + b''.join(struct.pack('<Q', int(s)) for s in shape)


def decode_dataspace(buf):
    ver, rank, flags = struct.unpack_from('<BBB', buf, 0)
    if ver == 1:
        pos = 8
    elif ver == 2:
        if buf[3] == 2:
            raise NotImplementedError('null dataspace')
        pos = 4
    else:
        raise NotImplementedError('dataspace message version %d' % ver)
    return tuple(struct.unpack_from('<%dQ' % rank, buf, pos)) if rank else ()


# -- writer ------------------------------------------------------------------------------------------------------------------
def _message(mtype, body, flags=0):
    body = _pad8(body)
    return struct.pack('<HHB3x', mtype, len(body), flags) + body


def _object_header(messages):
    data = b''.join(messages)
    return struct.pack('<BxHII4x', 1, len(messages), 1, len(data)) + data


def _attribute(name, dtype_body, shape, data):
    nm = name.encode() + b'\0'
    space = encode_dataspace(shape)
    return _message(MSG_ATTRIBUTE, struct.pack('<BxHHH', 1, len(nm), len(dtype_body), len(space)) + _pad8(nm) + _pad8(dtype_body)
                    + _pad8(space) + data)


def create_file(path, attrs, datasets):
    """Write a new file: root-group attributes `attrs` ({name: int | bool | str | numpy scalar}, in creation order) and
    zero-initialised contiguous datasets `datasets` ({name: (shape, dtype)}; at most 2 * LEAF_K = 8 of them, one symbol node).
    The dataset storage is allocated (the file is extended sparsely) but not written."""
    if len(datasets) > 2 * LEAF_K:
        raise NotImplementedError('more than %d datasets need a split symbol-table node' % (2 * LEAF_K))
    names = sorted(datasets, key=lambda s: s.encode())
    # local heap data segment: "" at offset 0 (the root's own name and the B-tree's first key), then the link names
    heap = bytearray(8)
    name_off = {}
    for n in names:
        name_off[n] = len(heap)
        heap += _pad8(n.encode() + b'\0')
    free_off = len(heap)
    heap += struct.pack('<QQ', HEAP_FREE_NULL, 32) + bytes(16)   # one free block closing the segment
    # global heap collection for variable-length string attributes
    gcol_objs = bytearray()
    str_index = {}
    for k, v in attrs.items():
        if isinstance(v, str):
            raw = v.encode()
            str_index[k] = len(str_index) + 1
            gcol_objs += struct.pack('<HH4xQ', str_index[k], 0, len(raw)) + _pad8(raw)
    # file offsets (everything 8-byte aligned).  The root header's length does not depend on the addresses it holds,
    # so lay it out with placeholders first
    def root_header(btree, heap_addr, gcol_addr):
        msgs = [_message(MSG_SYMBOL_TABLE, struct.pack('<QQ', btree, heap_addr))]
        for k, v in attrs.items():
            if isinstance(v, str):
                body = struct.pack('<IQI', len(v.encode()), gcol_addr, str_index[k])
                msgs.append(_attribute(k, encode_datatype('vlen_str'), (), body))
            else:
                a = np.asarray(v)
                if a.dtype.kind in 'iu' and not isinstance(v, np.generic):
                    a = a.astype(np.int64)   # h5py stores a Python int as a 64-bit integer
                if a.shape != ():
                    raise NotImplementedError('array-valued attribute %s' % k)
                msgs.append(_attribute(k, encode_datatype(a.dtype), (), a.tobytes()))
        return _object_header(msgs)

    root_addr = 96
    root_len = len(root_header(0, 0, 0))
    btree_addr = root_addr + root_len
    btree_len = 24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8
    heap_addr = btree_addr + btree_len
    heap_data_addr = heap_addr + 32
    snod_addr = heap_data_addr + len(heap)
    snod_len = 8 + 2 * LEAF_K * 40
    gcol_addr = snod_addr + snod_len
    pos = gcol_addr + (GCOL_SIZE if str_index else 0)
    # dataset object headers, then the raw data
    specs = {}
    for n in names:
        shape, dt = datasets[n]
        dt = np.dtype(dt)
        specs[n] = {'shape': tuple(int(s) for s in shape), 'dtype': dt, 'header': pos,
                    'nbytes': int(np.prod(shape, dtype=np.int64)) * dt.itemsize}
        pos += len(_dataset_header(specs[n], 0))
    for n in names:
        specs[n]['address'] = pos
        pos += (specs[n]['nbytes'] + 7) // 8 * 8
    eof = pos

    out = bytearray()
    # superblock version 0: signature; versions of superblock / free-space storage / root symbol-table entry, reserved,
    # version of shared header messages; sizes of offsets and lengths, reserved; group leaf K, internal K; consistency flags;
    # base address, free-space address, end-of-file address, driver-information address; root group symbol-table entry
    out += SIGNATURE + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, LEAF_K, INTERNAL_K, 0)
    out += struct.pack('<QQQQ', 0, UNDEF, eof, UNDEF)
    out += struct.pack('<QQII', 0, root_addr, 1, 0) + struct.pack('<QQ', btree_addr, heap_addr)  # cache type 1: B-tree + heap
    assert len(out) == root_addr
    out += root_header(btree_addr, heap_addr, gcol_addr)
    # B-tree node: one child (the symbol node); key 0 = "" and key 1 = the largest name in the child
    node = b'TREE' + struct.pack('<BBHQQ', 0, 0, 1, UNDEF, UNDEF) + struct.pack('<QQQ', 0, snod_addr, name_off[names[-1]] if names else 0)
    out += node + bytes(btree_len - len(node))
    out += b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap), free_off, heap_data_addr) + heap
    snod = b'SNOD' + struct.pack('<BxH', 1, len(names))
    for n in names:
        snod += struct.pack('<QQII16x', name_off[n], specs[n]['header'], 0, 0)
    out += snod + bytes(snod_len - len(snod))
    if str_index:
        used = 16 + len(gcol_objs)
        gcol = b'GCOL' + struct.pack('<B3xQ', 1, GCOL_SIZE) + gcol_objs
        gcol += struct.pack('<HH4xQ', 0, 0, GCOL_SIZE - used)   # object 0: the free space, its size includes this header
        out += gcol + bytes(GCOL_SIZE - len(gcol))
    for n in names:
        assert len(out) == specs[n]['header']
        out += _dataset_header(specs[n], specs[n]['address'])
    with open(path, 'wb') as f:
        f.write(out)
        f.truncate(eof)
    return specs


def _dataset_header(spec, address):
    return _object_header([
        _message(MSG_DATASPACE, encode_dataspace(spec['shape']), flags=1),
        _message(MSG_DATATYPE, encode_datatype(spec['dtype']), flags=1),
        # fill value (version 2): allocate early, write the fill value if set, default fill value (defined, size 0)
        _message(MSG_FILL, struct.pack('<BBBBI', 2, 1, 2, 1, 0), flags=1),
        # layout (version 3), class 1 = contiguous: address, size
        _message(MSG_LAYOUT, struct.pack('<BBQQ', 3, 1, address, spec['nbytes'])),
    ])


# -- reader ------------------------------------------------------------------------------------------------------------------
class Dataset:
    def __init__(self, name, shape, dtype, address, nbytes, compact=None):
        self.name, self.shape, self.dtype, self.address, self.nbytes, self.compact = name, shape, dtype, address, nbytes, compact

    @property
    def allocated(self):
        return self.compact is not None or self.address != UNDEF


class File:
    """An open replay-style HDF5 file: `.attrs` (dict of decoded root attributes), `.datasets` ({name: Dataset}),
    `array(name, mode)` (numpy.memmap over a contiguous dataset) and `set_attr(name, value)` for in-place updates of
    fixed-size scalar attributes (what ReplayMemoryDisk's cur_idx / is_full setters do)."""

    def __init__(self, path, mode='r'):
        self.path, self.mode = path, mode
        self.f = open(path, 'rb' if mode == 'r' else 'r+b')
        try:
            self._parse()
        except Exception:
            self.f.close()
            raise

    def close(self):
        if self.f is not None:
            self.f.close()
            self.f = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _read(self, addr, n):
        if addr == UNDEF or addr + n > self.size:
            raise ValueError('HDF5 structure points outside the file (address %#x, %d bytes)' % (addr, n))
        self.f.seek(self.base + addr)
        b = self.f.read(n)
        if len(b) != n:
            raise ValueError('HDF5 file truncated')
        return b

    def _parse(self):
        self.size = os.fstat(self.f.fileno()).st_size
        self.base = 0
        head = self.f.read(16)
        if head[:8] != SIGNATURE:
            raise ValueError('%s is not an HDF5 file' % self.path)
        ver = head[8]
        if ver not in (0, 1):
            raise NotImplementedError('HDF5 superblock version %d (file written with libver="latest")' % ver)
        if head[13] != 8 or head[14] != 8:
            raise NotImplementedError('HDF5 offsets / lengths of %d / %d bytes' % (head[13], head[14]))
        pos = 16 + 4 + 4 + (4 if ver == 1 else 0)   # leaf K, internal K, flags (+ indexed-storage K, reserved)
        self.f.seek(pos)
        base, _free, self.eof, _drv = struct.unpack('<QQQQ', self.f.read(32))
        self.base = base
        _name_off, root_addr, cache, _ = struct.unpack('<QQII', self.f.read(24))
        self.root_addr = root_addr
        msgs = self._object_header(root_addr)
        self.attrs, self._attr_where = {}, {}
        btree = heap = None
        for mtype, body, where in msgs:
            if mtype == MSG_SYMBOL_TABLE:
                btree, heap = struct.unpack_from('<QQ', body, 0)
            elif mtype == MSG_ATTRIBUTE:
                self._attribute(body, where)
            elif mtype in (MSG_LINK_INFO, MSG_LINK, MSG_ATTR_INFO):
                raise NotImplementedError('new-style group / dense attributes (file written with libver="latest")')
        if btree is None:
            raise ValueError('root group has no symbol table')
        self.datasets = {}
        hsig, hver, hsize, _hfree, hdata = struct.unpack('<4sB3xQQQ', self._read(heap, 32))
        if hsig != b'HEAP' or hver != 0:
            raise ValueError('bad local heap')
        names = self._read(hdata, hsize)
        for name_off, header in self._group_entries(btree):
            end = names.index(b'\0', name_off)
            self._dataset(names[name_off:end].decode(), header)

    def _object_header(self, addr):
        """-> [(message type, body bytes, file address of the body)] of a version-1 object header."""
        pre = self._read(addr, 16)
        if pre[:4] == b'OHDR':
            raise NotImplementedError('version-2 object header (file written with libver="latest")')
        ver, nmsgs, _refs, size = struct.unpack_from('<BxHII', pre, 0)
        if ver != 1:
            raise ValueError('object header version %d' % ver)
        chunks = [(addr + 16, size)]
        out = []
        while chunks and len(out) < nmsgs:
            caddr, clen = chunks.pop(0)
            chunk = self._read(caddr, clen)
            pos = 0
            while pos + 8 <= clen and len(out) < nmsgs:
                mtype, msize, _flags = struct.unpack_from('<HHB', chunk, pos)
                body = chunk[pos + 8:pos + 8 + msize]
                if len(body) != msize:
                    raise ValueError('object header message overruns its chunk')
                if mtype == MSG_CONTINUATION:
                    chunks.append(struct.unpack_from('<QQ', body, 0))
                out.append((mtype, body, caddr + pos + 8))
                pos += 8 + msize
        return out

    def _group_entries(self, addr):
        node = self._read(addr, 24)
        sig, ntype, level, used, _left, _right = struct.unpack('<4sBBHQQ', node)
        if sig != b'TREE' or ntype != 0:
            raise ValueError('bad group B-tree node')
        body = self._read(addr + 24, (2 * used + 1) * 8)
        for i in range(used):
            child = struct.unpack_from('<Q', body, (2 * i + 1) * 8)[0]
            if level > 0:
                yield from self._group_entries(child)
            else:
                sig, sver, nsym = struct.unpack('<4sBxH', self._read(child, 8))
                if sig != b'SNOD' or sver != 1:
                    raise ValueError('bad symbol-table node')
                ents = self._read(child + 8, nsym * 40)
                for j in range(nsym):
                    yield struct.unpack_from('<QQ', ents, j * 40)

    def _attribute(self, body, where):
        ver = body[0]
        if ver == 1:
            nlen, tlen, slen = struct.unpack_from('<HHH', body, 2)
            pos, pad = 8, lambda n: (n + 7) // 8 * 8
        elif ver in (2, 3):
            if body[1] & 3:
                raise NotImplementedError('attribute with a shared datatype / dataspace')
            nlen, tlen, slen = struct.unpack_from('<HHH', body, 2)
            pos, pad = (8 if ver == 2 else 9), lambda n: n
        else:
            raise NotImplementedError('attribute message version %d' % ver)
        name = body[pos:pos + nlen].split(b'\0')[0].decode(); pos += pad(nlen)
        dt, _ = decode_datatype(body, pos); pos += pad(tlen)
        shape = decode_dataspace(body[pos:pos + slen]); pos += pad(slen)
        count = int(np.prod(shape, dtype=np.int64)) if shape else 1
        raw = body[pos:pos + count * dt['size']]
        if dt['vlen_str']:
            vals = []
            for i in range(count):
                n, gaddr, idx = struct.unpack_from('<IQI', raw, 16 * i)
                s = self._global_heap_object(gaddr, idx)[:n]
                vals.append(s.decode('utf-8', 'replace'))
            val = vals[0] if shape == () else np.array(vals, dtype=object).reshape(shape)
            self._attr_where[name] = None
        else:
            a = np.frombuffer(raw, dtype=dt['np'] if not dt['bool'] else np.int8, count=count)
            if dt['bool']:
                a = a.astype(np.bool_)
            elif dt['class'] == 3 and dt.get('utf8'):
                a = np.array([x.decode('utf-8', 'replace') for x in a.tolist()], dtype=object)
            val = a.reshape(shape)[()] if shape == () else a.reshape(shape)
            self._attr_where[name] = (where + pos, dt)
        self.attrs[name] = val

    def _global_heap_object(self, addr, index):
        sig, ver, size = struct.unpack('<4sB3xQ', self._read(addr, 16))
        if sig != b'GCOL' or ver != 1:
            raise ValueError('bad global heap collection')
        col = self._read(addr, size)
        pos = 16
        while pos + 16 <= size:
            idx, _refs, osize = struct.unpack_from('<HH4xQ', col, pos)
            if idx == 0:
                break
            if idx == index:
                return col[pos + 16:pos + 16 + osize]
            pos += 16 + (osize + 7) // 8 * 8
        raise ValueError('global heap object %d not found' % index)

    def _dataset(self, name, header):
        shape = dt = None
        layout = None
        for mtype, body, _ in self._object_header(header):
            if mtype == MSG_DATASPACE:
                shape = decode_dataspace(body)
            elif mtype == MSG_DATATYPE:
                dt, _ = decode_datatype(body)
            elif mtype == MSG_LAYOUT:
                if body[0] != 3:
                    raise NotImplementedError('data layout message version %d' % body[0])
                if body[1] == 1:
                    layout = ('contiguous',) + struct.unpack_from('<QQ', body, 2)
                elif body[1] == 0:
                    n = struct.unpack_from('<H', body, 2)[0]
                    layout = ('compact', bytes(body[4:4 + n]))
                else:
                    raise NotImplementedError('chunked dataset %s' % name)
            elif mtype == MSG_SYMBOL_TABLE:
                return  # a sub-group: not part of the replay schema
        if shape is None or dt is None or layout is None:
            raise ValueError('dataset %s lacks a dataspace, datatype or layout message' % name)
        if dt['np'] is None:
            raise NotImplementedError('dataset %s of a variable-length type' % name)
        nbytes = int(np.prod(shape, dtype=np.int64)) * dt['size'] if shape else dt['size']
        if layout[0] == 'contiguous':
            if layout[1] != UNDEF and layout[1] + nbytes > self.size:
                raise ValueError('dataset %s extends beyond the end of the file' % name)
            self.datasets[name] = Dataset(name, shape, dt['np'], layout[1], nbytes)
        else:
            self.datasets[name] = Dataset(name, shape, dt['np'], UNDEF, nbytes, compact=layout[1])

    # -- access -------------------------------------------------------------------------------------------------------
    def array(self, name, mode=None):
        d = self.datasets[name]
        if d.compact is not None:
            return np.frombuffer(d.compact, dtype=d.dtype).reshape(d.shape)
        if d.address == UNDEF:
            if (mode or self.mode) != 'r':
                raise NotImplementedError('dataset %s has no storage allocated yet (never written by h5py): cannot be '
                                          'written in place' % name)
            return np.zeros(d.shape, d.dtype)
        if d.nbytes == 0:
            return np.zeros(d.shape, d.dtype)
        return np.memmap(self.path, dtype=d.dtype, mode='r' if (mode or self.mode) == 'r' else 'r+',
                         offset=self.base + d.address, shape=d.shape)

    def set_attr(self, name, value):
        where = self._attr_where.get(name)
        if where is None:
            raise NotImplementedError('attribute %s cannot be updated in place' % name)
        off, dt = where
        raw = np.asarray(value).astype(np.int8 if dt['bool'] else dt['np']).tobytes()
        assert len(raw) == dt['size']
        self.f.seek(self.base + off)
        self.f.write(raw)
        self.f.flush()
        self.attrs[name] = np.asarray(value).astype(np.bool_ if dt['bool'] else dt['np'])[()]
