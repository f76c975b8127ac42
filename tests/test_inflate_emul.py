"""The inflate kernel's per-thread decoder (strique_b200/csrc/inflate_core.h, compiled for the host by
tests/native/Makefile) against zlib itself -- the library behind h5py's deflate filter, through which the reference
reads fast5 Signal chunks (STRique_lib/fast5Index.py:76-84).  Byte-exact output, and every damaged stream refused."""
import ctypes
import os
import subprocess
import zlib

import numpy as np
import pytest

from .conftest import ROOT

NATIVE = os.path.join(ROOT, 'tests', 'native')


@pytest.fixture(scope='module')
def emul():
    subprocess.run(['make', '-C', NATIVE, 'libinflate_emul.so'], check=True, capture_output=True)
    lib = ctypes.CDLL(os.path.join(NATIVE, 'libinflate_emul.so'))
    lib.strique_test_inflate.restype = ctypes.c_int
    lib.strique_test_inflate.argtypes = [ctypes.c_char_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint,
                                         ctypes.c_void_p, ctypes.c_uint, ctypes.POINTER(ctypes.c_uint)]
    return lib


def run(lib, stream, full, keep=None, misalign=0):
    keep = full if keep is None else keep
    out = np.full(max(keep, 1), 0xEE, np.uint8)
    spill = np.full(max(full - keep, 1), 0xEE, np.uint8)
    produced = ctypes.c_uint(0)
    st = lib.strique_test_inflate(stream, len(stream), misalign, out.ctypes.data, keep, spill.ctypes.data, full,
                                  ctypes.byref(produced))
    data = np.concatenate([out[:keep], spill[:full - keep]])[:produced.value].tobytes()
    return st, data


def signal_like(rng, n):
    levels = np.repeat(rng.uniform(450, 750, n // 5 + 2), rng.integers(5, 10, n // 5 + 2))[:n]
    return np.round(levels + rng.normal(0, 12, n)).astype('<i2').tobytes()


def payloads():
    rng = np.random.default_rng(11)
    yield 'signal 8192 samples', signal_like(rng, 8192)
    yield 'signal 100 samples', signal_like(rng, 100)
    yield 'random bytes', rng.integers(0, 256, 20000, dtype=np.uint8).tobytes()
    yield 'constant', bytes(16384)
    yield 'ramp', (np.arange(30000) % 251).astype(np.uint8).tobytes()
    yield 'text', b'the quick brown fox jumps over the lazy dog. ' * 300
    yield 'far matches', rng.integers(0, 256, 32768, dtype=np.uint8).tobytes() * 3       # distances up to the window size
    yield 'one byte', b'x'
    yield 'empty', b''
    yield 'two symbols', bytes(rng.integers(0, 2, 5000, dtype=np.uint8))                  # one-bit literal codes
    yield 'skewed', bytes(np.minimum(rng.geometric(0.02, 40000), 255).astype(np.uint8))   # long code lengths


def streams(data):
    for level in (0, 1, 6, 9):
        yield 'level %d' % level, zlib.compress(data, level)
    for name, strategy in (('fixed', zlib.Z_FIXED), ('huffman only', zlib.Z_HUFFMAN_ONLY), ('rle', zlib.Z_RLE),
                           ('filtered', zlib.Z_FILTERED)):
        c = zlib.compressobj(6, zlib.DEFLATED, 15, 8, strategy)
        yield name, c.compress(data) + c.flush()
    c = zlib.compressobj(6, zlib.DEFLATED, 9, 1, zlib.Z_DEFAULT_STRATEGY)                # small window, many blocks
    yield 'window 512, memLevel 1', c.compress(data) + c.flush()
    c = zlib.compressobj(1)
    parts = [c.compress(data[i:i + 1000]) + c.flush(zlib.Z_FULL_FLUSH) for i in range(0, len(data), 1000)]
    yield 'full flush every 1000 bytes', b''.join(parts) + c.flush()                      # empty stored blocks in between


def test_every_block_type_and_strategy_matches_zlib(emul):
    n = 0
    for pname, data in payloads():
        for sname, stream in streams(data):
            assert zlib.decompress(stream) == data
            for misalign in (0, 1, 3, 4, 7, 12, 15):
                st, got = run(emul, stream, len(data), misalign=misalign)
                assert st == 0, (pname, sname, misalign, st)
                assert got == data, (pname, sname, misalign)
            n += 1
    assert n >= 90


def test_last_chunk_of_a_dataset_keeps_only_its_head(emul):
    """HDF5 pads the last chunk to the chunk size: the decoder keeps `keep` bytes in the read's buffer and spills
    the rest (matches may still reach into either part)."""
    rng = np.random.default_rng(5)
    data = signal_like(rng, 3000) + bytes(16384 - 6000)
    stream = zlib.compress(data, 6)
    for keep in (0, 1, 5999, 6000, 6001, 16383):
        st, got = run(emul, stream, len(data), keep=keep)
        assert st == 0 and got == data, keep


def test_output_larger_than_the_chunk_is_refused(emul):
    data = bytes(range(256)) * 40
    stream = zlib.compress(data, 6)
    st, _ = run(emul, stream, len(data) - 1)
    assert st == 4
    st, got = run(emul, stream, len(data) + 100)           # a shorter stream than the chunk: reported length tells
    assert st == 0 and got == data


def test_damaged_streams_are_refused(emul):
    rng = np.random.default_rng(17)
    data = signal_like(rng, 8192)
    stream = zlib.compress(data, 4)
    for cut in (0, 1, 2, 5, 100, len(stream) // 2, len(stream) - 4, len(stream) - 1):
        st, _ = run(emul, stream[:cut], len(data))
        assert st != 0, cut
    flipped = 0
    for k in rng.integers(0, len(stream), 300):
        bad = bytearray(stream)
        bad[k] ^= 1 << int(rng.integers(0, 8))
        st, got = run(emul, bytes(bad), len(data))
        try:
            want = zlib.decompress(bytes(bad))            # (a flip can survive Adler-32; then both must agree on the bytes)
        except zlib.error:
            want = None
        assert (st == 0) == (want is not None), k
        if want is not None:
            assert got == want, k
        flipped += st != 0
    assert flipped >= 290
    assert run(emul, b'\x78\x9c' + b'\x07' + bytes(20), 100)[0] == 2         # block type 3
    assert run(emul, b'\x78\xbb' + bytes(20), 100)[0] == 1                   # preset dictionary / bad check bits
    assert run(emul, b'\x1f\x8b' + bytes(20), 100)[0] == 1                   # gzip, not zlib
