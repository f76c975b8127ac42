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
