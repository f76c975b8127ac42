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


def test_count_on_multi_read_fast5_files_with_worker_processes(tmp_path):
    """N1 + N2 through the command line: 48 synthetic panel reads in multi-read fast5 files (deflate-chunked), decoded by
    --t 4 worker PROCESSES into the pinned staging buffer, several GPU batches (STRIQUE_BATCH_SAMPLES), rows in input
    order and equal to what repeatCounter.detect_batch gives for the same signals."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import make_fast5_dataset as mk
    from strique_b200 import workload
    from strique_b200.counter import repeatCounter
    from strique_b200.pore_model import pore_model
    model = os.path.join(ROOT, 'models', 'r9_4_450bps.model')
    reads = workload.make_reads(pore_model(model), 48, seed=99, loci=workload.PANEL, n_lo=2, n_hi=150)
    index_file, sam_file, ids = mk.build(str(tmp_path / 'ds'), reads, per_file=20, procs=2)
    out = tmp_path / 'out.tsv'
    env = dict(os.environ, STRIQUE_BATCH_SAMPLES=str(400000))
    subprocess.run([sys.executable, SCRIPT, 'count', index_file, model, os.path.join(ROOT, 'configs', 'panel_config.tsv'),
                    '--algn', sam_file, '--t', '4', '--out', str(out)], check=True, env=env)
    rows = [l.split('\t') for l in out.read_text().strip().split('\n')[1:]]
    assert [r[0] for r in rows] == ids
    dt = repeatCounter(model)
    for name in workload.PANEL:
        dt.add_target(name, *workload.LOCI[name])
    want = dt.detect_batch([(name, sig, strand) for name, sig, strand, _ in reads])
    for r, (name, _, strand, _), w in zip(rows, reads, want):
        assert r[1:3] == [name, strand]
        assert r[3:] == [str(x) for x in w]
    # the run above inflated the Signal chunks on the GPU (strique_inflate_batch); zlib on the workers and the
    # single-threaded default (--t 1) must give the same file
    out2, out3 = tmp_path / 'out_host_inflate.tsv', tmp_path / 'out_t1.tsv'
    subprocess.run([sys.executable, SCRIPT, 'count', index_file, model, os.path.join(ROOT, 'configs', 'panel_config.tsv'),
                    '--algn', sam_file, '--t', '4', '--out', str(out2)], check=True, env=dict(env, STRIQUE_HOST_INFLATE='1'))
    subprocess.run([sys.executable, SCRIPT, 'count', index_file, model, os.path.join(ROOT, 'configs', 'panel_config.tsv'),
                    '--algn', sam_file, '--out', str(out3)], check=True, env=env)
    assert out2.read_text() == out.read_text() == out3.read_text()
