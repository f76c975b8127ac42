"""End to end through the reference's command line on the bundled data (the reference's own CI smoke,
.travis.yml:99-102: `STRique.py index data/ > reads.fofn; cat c9orf72.sam | STRique.py count ...`)."""
import os
import subprocess
import sys

import pytest

from .conftest import ROOT

pytestmark = pytest.mark.gpu
SCRIPT = os.path.join(ROOT, 'scripts', 'STRique.py')


def test_index_then_count_on_bundled_read(tmp_path):
    data = os.path.join(ROOT, 'data')
    idx = subprocess.run([sys.executable, SCRIPT, 'index', data, '--out_prefix', data], capture_output=True, text=True,
                         check=True).stdout
    fofn = tmp_path / 'reads.fofn'
    fofn.write_text(idx)
    sam = open(os.path.join(data, 'c9orf72.sam')).read()
    out = subprocess.run([sys.executable, SCRIPT, 'count', str(fofn), os.path.join(ROOT, 'models', 'r9_4_450bps.model'),
                          os.path.join(ROOT, 'configs', 'repeat_config.tsv'), '--config',
                          os.path.join(ROOT, 'configs', 'STRique.json'), '--mod_model',
                          os.path.join(ROOT, 'models', 'r9_4_450bps_mCpG.model'), '--log_level', 'debug'],
                         input=sam, capture_output=True, text=True, check=True).stdout
    lines = out.strip().split('\n')
    assert lines[0] == 'ID\ttarget\tstrand\tcount\tscore_prefix\tscore_suffix\tlog_p\toffset\tticks\tmod'
    cols = lines[1].split('\t')
    # docs/installation/test.md:16 -- the reference's documented row ("similar to"):
    #   ce47b364-... c9orf72 - 735 6.3155927807600545 6.031860427335506 -119860.52066647023 1633 40758 -
    # ID, target, strand, offset and ticks are integers of the alignment geometry: exact.  The documented count,
    # scores and log p come from an earlier revision / dependency set that cannot be run here (they back-solve to the
    # same path geometry on slightly different signal values): held to the doc's own precision,
    # +-2 repeats and ~1.5 %.  (That the CUDA path equals the ORACLE on this read to the last bit is asserted in
    # tests/test_pipeline_gpu.py -- an oracle-relative statement, labelled as such.)
    assert cols[:3] == ['ce47b364-ed6e-4409-808a-1041c0b5aac2', 'c9orf72', '-']
    assert cols[7] == '1633' and cols[8] == '40758'
    assert abs(int(cols[3]) - 735) <= 2
    assert float(cols[4]) == pytest.approx(6.3155927807600545, rel=0.015)
    assert float(cols[5]) == pytest.approx(6.031860427335506, rel=0.015)
    assert float(cols[6]) == pytest.approx(-119860.52066647023, rel=0.02)
    assert set(cols[9]) <= {'0', '1'} and abs(len(cols[9]) - int(cols[3])) <= 3
    assert len(lines) == 2
