"""Committed golden vectors of the whole per-read path (tests/golden/pipeline_golden.json, made by
tests/golden/make_pipeline_golden.py from the oracle with the COMPILED reference aligner):
 * CPU: the oracle with its C aligner restatement reproduces them (travels to boxes without
   /root/reference);
 * GPU: the CUDA path behind strique_detect_batch reproduces them -- integers and the fp32-derived
   alignment scores bit-exact, log p within 1e-9 relative."""
import json
import os
import zlib

import pytest

from strique_b200 import workload
from strique_b200.pore_model import pore_model
from .conftest import ROOT

GOLDEN = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'pipeline_golden.json')))


def reads_of(tag, model_file, mod_model_file):
    s = GOLDEN['sets'][tag]
    pm = pore_model(model_file)
    reads = workload.make_reads(pm, pm_mod=pore_model(mod_model_file) if s['use_mod'] else None, **s['kwargs'])
    for (name, sig, strand, n), row in zip(reads, s['rows']):
        assert zlib.crc32(sig.tobytes()) == row['crc'], 'synthetic read generator drifted from the golden file'
        assert (name, strand, n) == (row['target'], row['strand'], row['n_true'])
    return reads, s


def check(row, got):
    assert got[0] == row['count']
    assert got[1] == row['score_prefix'] and got[2] == row['score_suffix']
    assert got[3] == pytest.approx(row['log_p'], rel=1e-9)
    assert (int(got[4]), int(got[5]), got[6]) == (row['offset'], row['ticks'], row['mod'])


@pytest.mark.parametrize('tag,idx', [('c2_small', (11, 5)), ('c3_mod', (6,)), ('c4_panel', (2,)), ('c4_panel4', (3, 5))])
def test_oracle_with_c_aligner_reproduces_golden(tag, idx, model_file, mod_model_file):
    from oracle import reference_path as rp
    reads, s = reads_of(tag, model_file, mod_model_file)
    ref = rp.RefRepeatCounter(model_file, mod_model_file=mod_model_file if s['use_mod'] else None, aligner='c')
    for name in s['kwargs']['loci']:
        ref.add_target(name, *workload.LOCI[name])
    for k in idx:
        name, sig, strand, _ = reads[k]
        check(s['rows'][k], ref.detect(name, sig, strand))


@pytest.mark.gpu
@pytest.mark.parametrize('tag', sorted(GOLDEN['sets']))
def test_cuda_path_reproduces_golden(tag, ctx, model_file, mod_model_file):
    from strique_b200.counter import repeatCounter
    reads, s = reads_of(tag, model_file, mod_model_file)
    dt = repeatCounter(model_file, mod_model_file=mod_model_file if s['use_mod'] else None, context=ctx)
    for name in s['kwargs']['loci']:
        dt.add_target(name, *workload.LOCI[name])
    got = dt.detect_batch([(name, sig, strand) for name, sig, strand, _ in reads])
    for row, g in zip(s['rows'], got):
        check(row, g)
    # counts of noisy reads equal the simulated truth (c5: within one unit in 4000).  ATTCT is the exception the
    # reference itself has: for a unit shorter than k its repeatHMM profile spans 9 k-mers = 1.8 units per pass
    # (scripts/STRique.py:329-334), so it reports ~n/1.8 -- reproduced, not corrected.
    ok = [abs(g[0] - r['n_true']) <= (1 if tag == 'c5_long' else 0) for g, r in zip(got, s['rows']) if r['target'] != 'atxn10']
    assert sum(ok) >= len(ok) - 1
