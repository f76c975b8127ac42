"""TEST INFRASTRUCTURE: a minimal HDF5 writer for fast5 fixtures (no h5py in this container).

Writes what ONT fast5 files of the reference's era contain and what `STRique_lib/fast5Index.py:62-84` walks:
superblock version 0, version-1 object headers, old-style groups (symbol-table message -> B-tree v1 node -> symbol
node + local heap), 1-D int16 datasets stored contiguous or chunked (B-tree v1, optional deflate, optional unknown
filter id for the VBZ error path), fixed-length string attributes.  Layouts:

    single-read   /Raw/Reads/Read_<n>/Signal        (+ read_id attribute on Read_<n>), /UniqueGlobalKey
    multi-read    /read_<id>/Raw/Signal             (+ read_id attribute on Raw), one top-level group per read

Every structure follows the HDF5 file-format specification 1.x ("Disk Format: Level 0 / 1 / 2"); nothing here is
used by the product.
"""
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


def _pad8(b):
    return b + b'\0' * (-len(b) % 8)


class Writer:
    def __init__(self):
        self.buf = bytearray(b'\0' * 96)      # superblock (56) + root symbol-table entry (40), filled in at the end

    def _alloc(self, data):
        while len(self.buf) % 8:
            self.buf.append(0)
        addr = len(self.buf)
        self.buf += data
        return addr

    # ---- object headers -------------------------------------------------------------------------
    def _object_header(self, messages):
        body = b''
        for mtype, payload in messages:
            payload = _pad8(payload)
            body += struct.pack('<HHB3x', mtype, len(payload), 0) + payload
        hdr = struct.pack('<BxHII4x', 1, len(messages), 1, len(body))       # version, #messages, refcount, size, pad to 16
        return self._alloc(hdr + body)

    @staticmethod
    def _attr_string(name, value):
        name_b = name.encode() + b'\0'
        val = value.encode()
        dtype = struct.pack('<BBBBI', 0x13, 0x00, 0, 0, len(val))            # class 3 (string) v1, null-terminated ASCII
        space = struct.pack('<BBB5x', 1, 0, 0)                               # scalar dataspace v1
        return (0x0C, struct.pack('<BxHHH', 1, len(name_b), len(dtype), len(space)) + _pad8(name_b) + _pad8(dtype) +
                _pad8(space) + val)

    # ---- groups -----------------------------------------------------------------------------------
    def group(self, members, attrs=()):
        """members: {name: object header address} -> object header address of the new group"""
        names = sorted(members)
        heap_data = bytearray(b'\0' * 8)                                     # offset 0: the empty string
        offs = {}
        for n in names:
            offs[n] = len(heap_data)
            heap_data += _pad8(n.encode() + b'\0')
        heap_data_addr = self._alloc(bytes(heap_data))
        heap = self._alloc(b'HEAP' + struct.pack('<B3xQQQ', 0, len(heap_data), UNDEF, heap_data_addr))
        # symbol nodes of <= 8 entries (2K with the default K = 4) under one B-tree leaf
        snods, keys = [], [0]
        for i in range(0, max(len(names), 1), 8):
            part = names[i:i + 8]
            ent = b''.join(struct.pack('<QQII16x', offs[n], members[n], 0, 0) for n in part)
            snods.append(self._alloc(b'SNOD' + struct.pack('<BxH', 1, len(part)) + ent + b'\0' * (40 * (8 - len(part)))))
            keys.append(offs[part[-1]] if part else 0)
        node = b'TREE' + struct.pack('<BBHQQ', 0, 0, len(snods), UNDEF, UNDEF)
        for k, s in enumerate(snods):
            node += struct.pack('<QQ', keys[k], s)
        node += struct.pack('<Q', keys[-1])
        btree = self._alloc(node)
        msgs = [(0x11, struct.pack('<QQ', btree, heap))] + [self._attr_string(k, v) for k, v in attrs]
        return self._object_header(msgs)

    # ---- datasets ---------------------------------------------------------------------------------
    def dataset_i16(self, values, chunk=None, deflate=False, extra_filter=None):
        values = np.ascontiguousarray(values, dtype='<i2')
        n = len(values)
        space = struct.pack('<BBB5xQ', 1, 1, 0, n)
        dtype = struct.pack('<BBBBIHH', 0x10, 0x08, 0, 0, 2, 0, 16)          # fixed point v1, signed, little endian, 16 bits
        msgs = [(0x01, space), (0x03, dtype)]
        if chunk is None:
            data = self._alloc(values.tobytes())
            msgs.append((0x08, struct.pack('<BBQQ', 3, 1, data, n * 2)))
        else:
            filters = []
            if deflate:
                filters.append((1, [6]))
            if extra_filter is not None:
                filters.append((extra_filter, [0, 2, 1, 1]))
            entries = []
            for off in range(0, n, chunk):
                raw = values[off:off + chunk].tobytes().ljust(chunk * 2, b'\0')
                if deflate:
                    raw = zlib.compress(raw, 6)
                entries.append((off, len(raw), self._alloc(raw)))
            node = b'TREE' + struct.pack('<BBHQQ', 1, 0, len(entries), UNDEF, UNDEF)
            for off, size, addr in entries:
                node += struct.pack('<IIQQ', size, 0, off, 0) + struct.pack('<Q', addr)
            node += struct.pack('<IIQQ', 0, 0, n, 0)
            btree = self._alloc(node)
            msgs.append((0x08, struct.pack('<BBBQII', 3, 2, 2, btree, chunk, 2)))
            if filters:
                body = struct.pack('<BB6x', 1, len(filters))
                for fid, cd in filters:
                    body += struct.pack('<HHHH', fid, 0, 0, len(cd)) + b''.join(struct.pack('<I', c) for c in cd)
                    if len(cd) % 2:
                        body += b'\0' * 4
                msgs.append((0x0B, body))
        return self._object_header(msgs)

    def finish(self, root):
        sb = b'\x89HDF\r\n\x1a\n' + struct.pack('<BBBBBBBBHHI', 0, 0, 0, 0, 0, 8, 8, 0, 4, 16, 0)
        sb += struct.pack('<QQQQ', 0, UNDEF, len(self.buf), UNDEF)
        sb += struct.pack('<QQII16x', 0, root, 0, 0)
        assert len(sb) == 96
        self.buf[:96] = sb
        return bytes(self.buf)


def single_read_fast5(path, signal, read_id, read_number=1, **ds):
    w = Writer()
    sig = w.dataset_i16(signal, **ds)
    read = w.group({'Signal': sig}, attrs=[('read_id', read_id)])
    reads = w.group({'Read_%d' % read_number: read})
    raw = w.group({'Reads': reads})
    ugk = w.group({})
    open(path, 'wb').write(w.finish(w.group({'Raw': raw, 'UniqueGlobalKey': ugk})))


def multi_read_fast5(path, reads, **ds):
    """reads: list of (read_id, int16 signal)"""
    w = Writer()
    top = {}
    for rid, signal in reads:
        sig = w.dataset_i16(signal, **ds)
        raw = w.group({'Signal': sig}, attrs=[('read_id', rid)])
        top['read_' + rid] = w.group({'Raw': raw})
    open(path, 'wb').write(w.finish(w.group(top)))
