"""fast5 access beyond the bundled single-read file (reference: STRique_lib/fast5Index.py:62-84 read lookup,
132-179 `index`, 220-233 `get_raw`): multi-read ("bulk") files, tar archives of single-read files, contiguous /
chunked / deflated signal datasets, and the VBZ filter that the built-in reader must name instead of mis-decoding.
Fixtures are written by tests/hdf5_writer.py (old-style groups, version-1 object headers: the layout of ONT files)."""
import os
import tarfile

import numpy as np
import pytest

from strique_b200 import fast5
from . import hdf5_writer as hw


def _signal(seed, n):
    rng = np.random.default_rng(seed)
    return np.round(rng.normal(600, 70, n)).astype(np.int16)


@pytest.mark.parametrize('ds', [dict(), dict(chunk=1000), dict(chunk=4096, deflate=True)])
def test_single_read_layouts(tmp_path, ds):
    sig = _signal(1, 9876)
    path = str(tmp_path / 'read.fast5')
    hw.single_read_fast5(path, sig, 'abc-123', read_number=77, **ds)
    assert np.array_equal(fast5.read_raw_signal(path), sig)
    assert fast5.read_id_of(path) == 'abc-123'
    assert 'Raw' in fast5.top_level_groups(path)


def test_the_writer_reproduces_the_bundled_file_semantics():
    """sanity of the fixture writer itself: the reader returns the same for the real bundled ONT file"""
    from .conftest import ROOT
    raw = fast5.read_raw_signal(os.path.join(ROOT, 'data', 'c9orf72.fast5'))
    assert raw.dtype == np.int16 and len(raw) == 284184
    assert fast5.read_id_of(os.path.join(ROOT, 'data', 'c9orf72.fast5')) == 'ce47b364-ed6e-4409-808a-1041c0b5aac2'


def test_multi_read_file_index_and_lookup(tmp_path):
    reads = [('id-%02d' % k, _signal(10 + k, 3000 + 111 * k)) for k in range(11)]      # > 8: two symbol nodes
    d = tmp_path / 'batch'
    d.mkdir()
    hw.multi_read_fast5(str(d / 'bulk_0.fast5'), reads[:6], chunk=2048, deflate=True)
    hw.multi_read_fast5(str(d / 'bulk_1.fast5'), reads[6:])
    records = list(fast5.fast5Index.index(str(d)))
    assert len(records) == 11
    assert all(r.split('\t')[0].startswith('bulk_') and '.fast5/read_id-' in r for r in records)
    idx_file = d / 'reads.fofn'
    idx_file.write_text('\n'.join(records) + '\n')
    f5 = fast5.fast5Index(str(idx_file))
    for rid, sig in reads:
        assert np.array_equal(f5.get_raw(rid), sig)
    with pytest.raises(RuntimeError):
        f5.get_raw('nope')


def test_tar_of_single_read_files(tmp_path):
    src = tmp_path / 'src' / 'sub'
    src.mkdir(parents=True)
    reads = [('tar-%d' % k, _signal(30 + k, 2500 + 7 * k)) for k in range(3)]
    for k, (rid, sig) in enumerate(reads):
        hw.single_read_fast5(str(src / ('r%d.fast5' % k)), sig, rid, read_number=k, chunk=1024, deflate=True)
    d = tmp_path / 'arch'
    d.mkdir()
    with tarfile.open(str(d / 'reads.tar'), 'w') as tar:
        tar.add(str(tmp_path / 'src'), arcname='.')
    records = list(fast5.fast5Index.index(str(d)))
    assert len(records) == 3 and all(r.startswith('reads.tar/') for r in records)
    idx_file = d / 'reads.fofn'
    idx_file.write_text('\n'.join(records) + '\n')
    f5 = fast5.fast5Index(str(idx_file))
    for rid, sig in reads:
        assert np.array_equal(f5.get_raw(rid), sig)


def test_vbz_is_named_not_misdecoded(tmp_path):
    path = str(tmp_path / 'vbz.fast5')
    hw.single_read_fast5(path, _signal(5, 5000), 'vbz-1', chunk=5000, extra_filter=32020)
    if fast5._h5py is not None:
        pytest.skip('h5py decides about its own plugins')
    with pytest.raises(fast5.HDF5Error, match='VBZ'):
        fast5.read_raw_signal(path)
    idx = tmp_path / 'i.fofn'
    idx.write_text('vbz.fast5\tvbz-1\n')
    with pytest.raises(RuntimeError, match='VBZ'):
        fast5.fast5Index(str(idx)).get_raw('vbz-1')


def test_open_files_are_cached_and_refreshed(tmp_path):
    path = str(tmp_path / 'x.fast5')
    a, b = _signal(1, 1000), _signal(2, 1200)
    hw.multi_read_fast5(path, [('a', a)])
    assert np.array_equal(fast5.read_raw_signal(path, 'read_a'), a)
    assert path in fast5._open_files
    os.remove(path)
    hw.multi_read_fast5(path, [('b', b)])
    os.utime(path, ns=(1, 1))                # different mtime: the cached mapping must not be used
    assert np.array_equal(fast5.read_raw_signal(path, 'read_b'), b)
