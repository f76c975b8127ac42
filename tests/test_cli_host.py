"""Host side of the CLI mirror (strique_b200/cli.py vs scripts/STRique.py of the reference):
config parsing, SAM decoding, locus intersection, fast5 index and the output format."""
import io
import os

import numpy as np
import pytest

from strique_b200 import cli, fast5
from .conftest import ROOT

CONFIG_TSV = os.path.join(ROOT, 'configs', 'repeat_config.tsv')
CONFIG_JSON = os.path.join(ROOT, 'configs', 'STRique.json')
SAM = os.path.join(ROOT, 'data', 'c9orf72.sam')


def test_parse_config_like_the_reference():
    cfg = cli.parse_config(CONFIG_TSV, CONFIG_JSON)
    assert set(cfg['repeat']) == {'c9orf72', 'fmr1'}
    chrom, begin, end, repeat, prefix, suffix = cfg['repeat']['c9orf72']
    assert (chrom, begin, end, repeat) == ('chr9', 27573527, 27573544, 'GGCCCC') and len(prefix) == len(suffix) == 150
    assert cfg['align']['dist_offset'] == 16.0 and 'match_loop' in cfg['HMM'] or isinstance(cfg['HMM'], dict)
    assert cli.parse_config(CONFIG_TSV)['align'] is None


def test_decode_sam_and_target_intersection():
    line = [l for l in open(SAM) if not l.startswith('@')][0]
    sr = cli.decode_sam(line)
    assert sr.QNAME == 'ce47b364-ed6e-4409-808a-1041c0b5aac2' and sr.FLAG == 16 and sr.RNAME == 'chr9'
    assert sr.POS == 27541232 and sr.TLEN > 30000
    assert cli.decode_sam('too\tshort').QNAME == ''
    assert cli.decode_sam('\t'.join(['r', 'x', 'chr9', '1', '0', '5M', '*', '0', '0', 'ACGTA', '*'])).QNAME == ''
    rd = cli.repeatDetector.__new__(cli.repeatDetector)
    from collections import defaultdict
    rd.repeatLoci, rd.repeat_config, rd.is_init = defaultdict(list), cli.parse_config(CONFIG_TSV)['repeat'], False

    class Stub(object):
        targets = {}

        def add_target(self, *a):
            self.targets[a[0]] = a
    rd.repeatCounter = Stub()
    work = rd.plan([line, 'garbage\n'])
    assert len(work) == 1 and work[0][2] == '-' and work[0][3] == ['c9orf72']
    assert set(Stub.targets) == {'c9orf72', 'fmr1'}
    # clipping arithmetic of the overlap test (S.py:677)
    sr2 = cli.decode_sam('\t'.join(['r', '0', 'chr9', '27573560', '60', '20S10M5H', '*', '0', '0', 'A' * 30, '*']))
    assert (sr2.CLIP_BEGIN, sr2.CLIP_END, sr2.TLEN) == (20, 5, 10)
    assert rd.intersect_target(sr2) == []
    sr2.CLIP_BEGIN = 100
    assert rd.intersect_target(sr2) == ['c9orf72']


def test_index_and_raw_lookup(tmp_path):
    records = list(fast5.fast5Index.index(os.path.join(ROOT, 'data'), output_prefix=os.path.join(ROOT, 'data')))
    assert records == [os.path.join(ROOT, 'data', 'c9orf72.fast5') + '\tce47b364-ed6e-4409-808a-1041c0b5aac2']
    idx_file = tmp_path / 'reads.fofn'
    idx_file.write_text('\n'.join(records) + '\n')
    f5 = fast5.fast5Index(str(idx_file))
    raw = f5.get_raw('ce47b364-ed6e-4409-808a-1041c0b5aac2')
    assert raw.dtype == np.int16 and len(raw) == 284184
    with pytest.raises(RuntimeError):
        f5.get_raw('nope')
    with pytest.raises(RuntimeError):
        fast5.fast5Index(str(tmp_path / 'missing.fofn'))


def test_output_writer_format(tmp_path):
    out = tmp_path / 'o.tsv'
    ow = cli.outputWriter(str(out))
    ow.write_line([('id1', 'c9orf72', '-', 735, 6.3155927807600545, 6.031860427335506, -119860.52066647023, 1633, 40758, '-'),
                   ('id2', 'c9orf72', '+', 0, 0.0, 0.0, 0, 12, 0, '-')])
    ow.close()
    lines = out.read_text().split('\n')
    assert lines[0] == 'ID\ttarget\tstrand\tcount\tscore_prefix\tscore_suffix\tlog_p\toffset\tticks\tmod'
    assert lines[1] == 'id1\tc9orf72\t-\t735\t6.3155927807600545\t6.031860427335506\t-119860.52066647023\t1633\t40758\t-'
    assert lines[2] == 'id2\tc9orf72\t+\t0\t0.0\t0.0\t0\t12\t0\t-'


def _stub_detector(batch_samples, fetch_log=None, fail_ids=()):
    """repeatDetector with a stub counter and a stub index (the host logic of `count` without a GPU)"""
    from collections import defaultdict

    class Counter(object):
        def __init__(self):
            self.targets, self.batches = {}, []

        def add_target(self, name, *a):
            self.targets[name] = a

        def detect_batch(self, items):
            self.batches.append(len(items))
            return [(len(sig) % 997, 1.5, 2.5, -1.0, 10, 20, '-') for _, sig, _ in items]

    class Index(object):
        def get_raw(self, ID):
            if fetch_log is not None:
                fetch_log.append(ID)
            if ID in fail_ids:
                raise RuntimeError('[Error] Read {} not found'.format(ID))
            return np.zeros(1000 + int(ID[4:]), dtype=np.int16)

    rd = cli.repeatDetector.__new__(cli.repeatDetector)
    rd.repeatCounter, rd.f5 = Counter(), Index()
    rd.repeatLoci, rd.repeat_config, rd.is_init = defaultdict(list), cli.parse_config(CONFIG_TSV)['repeat'], False
    rd.io_threads, rd.batch_samples = 2, batch_samples
    return rd


def _sam(i):
    return '\t'.join(['read%03d' % i, '16' if i % 2 else '0', 'chr9', '27573000', '60', '1000M', '*', '0', '0', 'A' * 100, '*']) + '\n'


def test_count_streams_input_and_appends_rows_batch_by_batch(tmp_path):
    """scripts/STRique.py:720-727, 936-945: the SAM is consumed as it arrives and rows reach the output while later
    reads are still to come -- partial output survives a crash.  Here: rows of the first batches are in the file
    before the input iterator is exhausted, fetching runs at most about one batch ahead, rows are in input order,
    and a read without a fast5 is dropped with a warning like in the reference."""
    out = tmp_path / 'o.tsv'
    fetched = []
    rd = _stub_detector(batch_samples=10 * 1050, fetch_log=fetched, fail_ids={'read007'})
    seen_at_yield = []

    def lines():
        for i in range(60):
            if i % 5 == 0:
                seen_at_yield.append((i, out.read_text().count('\n') - 1 if out.exists() else 0, len(fetched)))
            yield _sam(i)
        yield 'garbage\n'

    n = cli.run_count(rd, lines(), str(out))
    rows = out.read_text().strip().split('\n')[1:]
    assert n == len(rows) == 59
    ids = [r.split('\t')[0] for r in rows]
    assert ids == sorted(ids) and 'read007' not in ids
    assert len(rd.repeatCounter.batches) >= 5                                 # many GPU batches ...
    assert any(written > 0 for i, written, _ in seen_at_yield if i < 55)      # ... whose rows were on disk mid-stream
    # bounded prefetch: when line i is pulled, at most ~one batch (SEQ_LEN 100 * 9 = 900 estimated samples per read:
    # 12 reads) plus the pool's slack has been fetched beyond what was consumed
    assert all(nf <= i + 1 for i, _, nf in seen_at_yield)
    # rows lag the input by at most the batch being fetched + the one being staged + the one on the GPU
    assert max(i - w for i, w, _ in seen_at_yield) <= 42


def test_failing_batch_is_retried_read_by_read(tmp_path):
    rd = _stub_detector(batch_samples=1 << 30)
    calls = []

    def detect_batch(items):
        calls.append(len(items))
        if len(items) > 1:
            raise RuntimeError('boom')
        if items[0][1].shape[0] == 1003:
            raise RuntimeError('bad read')
        return [(1, 1.0, 1.0, -1.0, 1, 1, '-')]

    rd.repeatCounter.detect_batch = detect_batch
    rd.repeatCounter.detect = lambda *it: detect_batch([it])[0]
    out = tmp_path / 'o.tsv'
    assert cli.run_count(rd, iter([_sam(i) for i in range(6)]), str(out)) == 5   # read003 is dropped, the rest survive
    assert calls[0] == 6 and calls.count(1) == 6


def test_tsv_consumers_of_the_reference_parse_our_rows(tmp_path):
    """the reference's downstream tools work off the unchanged TSV.  The two parse expressions are
    restated from scripts/fast5Masker.py:55-60 (typed record, then mask[offset:offset+ticks]) and
    scripts/STRique.py:972-990 (`plot`: nine positional fields); rows come from the golden sets through our writer,
    including a mod string and a degenerate all-zero row."""
    import json
    from collections import namedtuple
    golden = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'pipeline_golden.json')))['sets']
    rows = []
    for name in ('c2_small', 'c3_mod'):
        for k, r in enumerate(golden[name]['rows'][:20]):
            rows.append(('%s-%d' % (name, k), r['target'], r['strand'], r['count'], r['score_prefix'], r['score_suffix'],
                         r['log_p'], r['offset'], r['ticks'], r['mod']))
    rows.append(('none', 'c9orf72', '+', 0, 0.0, 0.0, 0, 0, 0, '-'))
    out = tmp_path / 'counts.tsv'
    ow = cli.outputWriter(str(out))
    ow.write_line(rows)
    ow.close()
    # fast5Masker
    Record = namedtuple('STRique_record', ['ID', 'target', 'strand', 'count', 'score_prefix', 'score_suffix', 'log_p', 'offset', 'ticks', 'mod'])
    with open(out) as fp:
        fields = (row.strip().split('\t') for row in fp if row and not row.startswith('ID'))
        records = [Record(*row[:3], int(row[3]), *[float(x) for x in row[4:7]], int(row[7]), int(row[8]), row[9]) for row in fields]
    assert len(records) == len(rows)
    assert any(set(r.mod) <= set('01') and len(r.mod) > 10 for r in records if r.mod != '-')
    for rec, row in zip(records, rows):
        assert (rec.ID, rec.count, rec.offset, rec.ticks, rec.mod) == (row[0], row[3], row[7], row[8], row[9])
        assert rec.log_p == float(row[6]) and rec.score_prefix == row[4]
        raw = np.arange(rec.offset + rec.ticks + 100, dtype=np.int16)
        mask = np.ones(raw.shape, dtype=bool)
        mask[rec.offset:rec.offset + rec.ticks] = False
        assert len(raw[mask]) == len(raw) - rec.ticks
    # plot
    with open(out) as fp:
        for line in fp:
            if not line.startswith('ID'):
                ID, target, strand, count, score_prefix, score_suffix, _, offset, ticks = line.strip().split('\t')[:9]
                assert int(offset) >= 0 and int(ticks) >= 0 and float(score_prefix) >= 0.0 and float(score_suffix) >= 0.0


@pytest.mark.parametrize('workers', [1, 2])
def test_stored_reads_reach_the_counter_as_chunk_records(tmp_path, monkeypatch, workers):
    """The GPU-inflate path of `count` without a GPU: reads of deflate-chunked multi-read fast5 files travel as
    stored chunk bytes + strique_inflate_chunk records (worker threads for --t 1, worker processes otherwise); a stub
    counter inflates the records with zlib and must recover exactly the signals get_raw decodes.  A contiguous
    (uncompressed) file in the same run takes the decoded path."""
    import zlib
    from collections import defaultdict
    from strique_b200 import _lib
    from . import hdf5_writer as hw

    class FakePinned(object):
        def __init__(self, n, dtype=np.int16):
            self.array = np.zeros(int(n), dtype)
    monkeypatch.setattr(_lib, 'PinnedBuffer', FakePinned)
    rng = np.random.default_rng(8)
    sigs = {'read%03d' % i: rng.integers(200, 900, int(rng.integers(3000, 40000))).astype(np.int16) for i in range(40)}
    ids = sorted(sigs)
    hw.multi_read_fast5(str(tmp_path / 'a.fast5'), [(i, sigs[i]) for i in ids[:25]], chunk=8192, deflate=True)
    hw.multi_read_fast5(str(tmp_path / 'b.fast5'), [(i, sigs[i]) for i in ids[25:35]], chunk=1000, deflate=True)
    hw.multi_read_fast5(str(tmp_path / 'c.fast5'), [(i, sigs[i]) for i in ids[35:]])                 # contiguous
    index = tmp_path / 'reads.fofn'
    index.write_text(''.join('%s.fast5/read_%s\t%s\n' % ('a' if k < 25 else ('b' if k < 35 else 'c'), i, i)
                             for k, i in enumerate(ids)))
    seen = {}

    class Counter(object):
        targets = {}

        def add_target(self, name, *a):
            self.targets[name] = a

        def detect_deflated(self, targets, comp, comp_bytes, chunks, offsets):
            comp = np.asarray(comp)[:comp_bytes]
            out = np.zeros(int(offsets[-1]), np.int16).view(np.uint8)
            assert chunks.dtype == _lib.INFLATE_CHUNK_DTYPE
            for c in chunks:
                data = zlib.decompress(comp[c['src_off']:c['src_off'] + c['src_len']].tobytes())
                assert len(data) == c['full'] and c['keep'] <= c['full']
                out[c['dst_off']:c['dst_off'] + c['keep']] = np.frombuffer(data, np.uint8)[:c['keep']]
            sam = out.view(np.int16)
            for k in range(len(targets)):
                seen[('stored', len(seen))] = sam[offsets[k]:offsets[k + 1]].copy()
            return [(1, 1.0, 1.0, -1.0, 1, 1, '-')] * len(targets), np.zeros(len(chunks), np.int32)

        def detect_packed(self, targets, raw, offsets):
            for k in range(len(targets)):
                seen[('packed', len(seen))] = np.array(raw[offsets[k]:offsets[k + 1]])
            return [(1, 1.0, 1.0, -1.0, 1, 1, '-')] * len(targets)

    rd = cli.repeatDetector.__new__(cli.repeatDetector)
    rd.repeatCounter, rd.f5 = Counter(), fast5.fast5Index(str(index))
    rd.repeatLoci, rd.repeat_config, rd.is_init = defaultdict(list), cli.parse_config(CONFIG_TSV)['repeat'], False
    rd.io_threads, rd.batch_samples, rd.gpu_inflate = workers, 300000, True
    out = tmp_path / 'o.tsv'
    lines = ['\t'.join([i, '0', 'chr9', '27573000', '60', '1000M', '*', '0', '0', 'A' * 100, '*']) + '\n' for i in ids]
    assert cli.run_count(rd, iter(lines), str(out)) == 40
    assert [l.split('\t')[0] for l in out.read_text().strip().split('\n')[1:]] == ids
    kinds = [k for k, _ in seen]
    assert kinds.count('stored') == 35 and kinds.count('packed') == 5
    got = sorted((a.tobytes() for a in seen.values()))
    assert got == sorted(sigs[i].tobytes() for i in ids)
