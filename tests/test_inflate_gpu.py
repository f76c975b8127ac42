"""strique_inflate_batch (csrc/inflate.cu: one thread per zlib stream) against zlib itself -- the library behind
h5py's deflate filter through which the reference reads fast5 Signal chunks (STRique_lib/fast5Index.py:76-84).
Byte-exact samples for every block type and strategy, padded last chunks, misaligned streams, damaged streams
refused with the status the host build of the same decoder gives (tests/test_inflate_emul.py pins that build)."""
import ctypes
import zlib

import numpy as np
import pytest

from strique_b200 import _lib
from . import test_inflate_emul as te

pytestmark = pytest.mark.gpu


def _batch(ctx, streams, fulls, keeps, pad=None, memspace=_lib.HOST):
    """-> (per-stream kept bytes, status)"""
    rng = np.random.default_rng(1)
    comp, recs, dst = bytearray(), [], 0
    for s, full, keep in zip(streams, fulls, keeps):
        comp += bytes(int(rng.integers(0, 4)) if pad is None else pad)        # streams start at any byte
        recs.append((len(comp), dst, len(s), keep, full, 0))
        comp += s
        dst += keep
    comp_np = np.frombuffer(bytes(comp), np.uint8)
    dev, status = ctx.inflate_batch(comp_np, len(comp_np), np.array(recs, dtype=_lib.INFLATE_CHUNK_DTYPE), dst, memspace=memspace)
    import torch

    class _Ext:
        __cuda_array_interface__ = {'shape': (max(dst, 1),), 'typestr': '|u1', 'data': (dev, False), 'version': 2}
    t = torch.as_tensor(_Ext(), device='cuda')
    out = t.cpu().numpy()
    pieces, p = [], 0
    for keep in keeps:
        pieces.append(out[p:p + keep].tobytes())
        p += keep
    return pieces, status


def test_every_block_type_in_one_batch(ctx):
    streams, datas = [], []
    for _, data in te.payloads():
        for _, stream in te.streams(data):
            streams.append(stream)
            datas.append(data)
    assert len(streams) >= 90
    fulls = [len(d) for d in datas]
    for memspace_pad in (None, 0, 3):
        got, status = _batch(ctx, streams, fulls, fulls, pad=memspace_pad)
        assert not status.any(), np.nonzero(status)
        assert got == datas


def test_signal_chunks_at_batch_scale_with_padded_last_chunks(ctx):
    """4096 chunks like a batch of reads: 8192-sample chunks of signal-like int16, every fifth one the padded tail of
    a read (only `keep` bytes belong to the dataset)."""
    rng = np.random.default_rng(2)
    streams, fulls, keeps, want = [], [], [], []
    base = [te.signal_like(rng, 8192) for _ in range(64)]
    for k in range(4096):
        data = base[k % 64]
        if k % 5 == 4:
            keep = int(rng.integers(0, 8192)) * 2
            data = data[:keep] + bytes(16384 - keep)
        else:
            keep = 16384
        streams.append(zlib.compress(data, 1 + k % 9))
        fulls.append(16384)
        keeps.append(keep)
        want.append(data[:keep])
    got, status = _batch(ctx, streams, fulls, keeps)
    assert not status.any()
    assert got == want


def test_damaged_streams_get_the_host_decoders_status_and_spare_the_others(ctx):
    import os
    import subprocess
    subprocess.run(['make', '-C', te.NATIVE, 'libinflate_emul.so'], check=True, capture_output=True)
    lib = ctypes.CDLL(os.path.join(te.NATIVE, 'libinflate_emul.so'))
    lib.strique_test_inflate.restype = ctypes.c_int
    lib.strique_test_inflate.argtypes = [ctypes.c_char_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p, ctypes.c_uint,
                                         ctypes.c_void_p, ctypes.c_uint, ctypes.POINTER(ctypes.c_uint)]
    rng = np.random.default_rng(3)
    data = te.signal_like(rng, 8192)
    good = zlib.compress(data, 4)
    streams = []
    for k in range(400):
        if k % 2 == 0:
            streams.append(good)
            continue
        bad = bytearray(good)
        if k % 4 == 1:
            bad[int(rng.integers(0, len(bad)))] ^= 1 << int(rng.integers(0, 8))
            bad = bytes(bad)
        else:
            bad = bytes(bad[:int(rng.integers(0, len(bad)))])
        streams.append(bad)
    fulls = [len(data)] * len(streams)
    got, status = _batch(ctx, streams, fulls, fulls, pad=0)
    n_bad = 0
    for k, s in enumerate(streams):
        st, out = te.run(lib, s, len(data))
        assert status[k] == (st if st else (8 if len(out) < len(data) else 0)), k      # 8: shorter than the chunk's share
        if st == 0:
            assert got[k] == out
        n_bad += st != 0
    assert n_bad >= 190
    assert all(got[k] == data for k in range(0, 400, 2))


def test_bad_descriptors_are_refused(ctx):
    comp = np.frombuffer(zlib.compress(b'abc'), np.uint8)
    with pytest.raises(_lib.StriqueError):
        ctx.inflate_batch(comp, len(comp), np.array([(0, 0, len(comp) + 1, 3, 3, 0)], dtype=_lib.INFLATE_CHUNK_DTYPE), 3)
    with pytest.raises(_lib.StriqueError):
        ctx.inflate_batch(comp, len(comp), np.array([(0, 2, len(comp), 3, 3, 0)], dtype=_lib.INFLATE_CHUNK_DTYPE), 3)
    dev, status = ctx.inflate_batch(comp, len(comp), np.zeros(0, dtype=_lib.INFLATE_CHUNK_DTYPE), 0)
    assert len(status) == 0
