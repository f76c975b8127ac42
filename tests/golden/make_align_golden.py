"""Generates tests/golden/align_golden.npz by running the REAL reference aligner
(oracle/_ref/pyseqan = /root/reference/src/pyalign.cpp compiled by oracle/build_ref.sh) on the
seeded cases of tests/align_cases.py.  Run from the repo root: python tests/golden/make_align_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reference_path as rp   # noqa: E402
from tests import align_cases as ac       # noqa: E402

ref = rp.make_aligner('ref')
out = {}
cs = ac.cases(seed=2024, n=240)
for k, (ps, a, b) in enumerate(cs):
    ref.gap_open_h, ref.gap_open_v, ref.gap_extension_h, ref.gap_extension_v, ref.dist_offset, ref.dist_min = ps
    score, a_idx, b_idx = ref.align_overlap(a.tolist(), b.tolist())
    out['params_%d' % k] = np.array(ps, dtype=np.float64)
    out['a_%d' % k] = a
    out['b_%d' % k] = b
    out['score_%d' % k] = np.float32(score)
    out['a_idx_%d' % k] = np.array(a_idx, dtype=np.uint64)
    out['b_idx_%d' % k] = np.array(b_idx, dtype=np.uint64)
out['n'] = np.int64(len(cs))
np.savez_compressed(os.path.join(ROOT, 'tests', 'golden', 'align_golden.npz'), **out)
print('wrote', len(cs), 'cases')
