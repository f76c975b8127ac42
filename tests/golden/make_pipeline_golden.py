"""Generates tests/golden/pipeline_golden.json: outputs of the oracle pipeline
(oracle/reference_path.py with the COMPILED REFERENCE aligner oracle/_ref/pyseqan, which only
exists where /root/reference is mounted) on seeded synthetic reads.  The reads themselves are
regenerated from their seeds (strique_b200/workload.py); a CRC of every signal guards against drift.

    python -m tests.golden.make_pipeline_golden
"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import reference_path as rp          # noqa: E402
from strique_b200 import workload                # noqa: E402
from strique_b200.pore_model import pore_model   # noqa: E402

MODEL = os.path.join(ROOT, 'models', 'r9_4_450bps.model')
MOD_MODEL = os.path.join(ROOT, 'models', 'r9_4_450bps_mCpG.model')

SETS = [
    # tag, make_reads kwargs, use_mod
    ('c2_small', dict(n_reads=16, seed=4242, loci=('c9orf72',), n_lo=2, n_hi=300), False),
    ('c3_mod', dict(n_reads=8, seed=4343, loci=('c9orf72',), n_lo=2, n_hi=200, mod_fraction=0.5), True),
    ('c4_panel', dict(n_reads=8, seed=4444, loci=('c9orf72', 'fmr1'), n_lo=2, n_hi=200), False),
    # all four panel loci: flank templates of 570 / 870 / 1170 / 1770 samples, repeat units of 3, 5 and 6 nt
    ('c4_panel4', dict(n_reads=12, seed=4545, loci=('c9orf72', 'fmr1', 'atxn10', 'dmpk'), n_lo=2, n_hi=150), False),
    # long-expansion stress (C5): ~4000 repeats, ~200 k samples, one read per strand
    ('c5_long', dict(n_reads=2, seed=4646, loci=('c9orf72',), fixed_n=4000, flank=4000), False),
]


def generate(tag, kwargs, use_mod):
    pm = pore_model(MODEL)
    pm_mod = pore_model(MOD_MODEL) if use_mod else None
    return workload.make_reads(pm, pm_mod=pm_mod, **kwargs)


def main():
    assert rp.load_pyseqan() is not None, 'build oracle/_ref first (make -C oracle ref)'
    out = {'note': 'oracle pipeline with the compiled reference aligner; see make_pipeline_golden.py', 'sets': {}}
    for tag, kwargs, use_mod in SETS:
        ref = rp.RefRepeatCounter(MODEL, mod_model_file=MOD_MODEL if use_mod else None, aligner='ref')
        for name in kwargs['loci']:
            ref.add_target(name, *workload.LOCI[name])
        rows = []
        for name, sig, strand, n_true in generate(tag, kwargs, use_mod):
            det = {}
            r = ref.detect(name, sig, strand, details=det)
            rows.append({'target': name, 'strand': strand, 'n_true': n_true, 'len': int(len(sig)),
                         'crc': zlib.crc32(sig.tobytes()),
                         'count': int(r[0]), 'score_prefix': float(r[1]), 'score_suffix': float(r[2]),
                         'log_p': float(r[3]), 'offset': int(r[4]), 'ticks': int(r[5]), 'mod': r[6],
                         'prefix_begin': det['prefix_begin'], 'suffix_end': det['suffix_end']})
            print(tag, rows[-1]['n_true'], rows[-1]['count'], rows[-1]['offset'], rows[-1]['ticks'], flush=True)
        out['sets'][tag] = {'kwargs': kwargs, 'use_mod': use_mod, 'rows': rows}
    with open(os.path.join(ROOT, 'tests', 'golden', 'pipeline_golden.json'), 'w') as fp:
        json.dump(out, fp, indent=1)


if __name__ == '__main__':
    main()
