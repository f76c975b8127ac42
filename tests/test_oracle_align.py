"""The C restatement of the aligner (oracle/align_oracle.c) against the REAL reference aligner
(oracle/_ref/pyseqan, compiled from /root/reference/src by oracle/build_ref.sh) and against the
committed golden vectors generated from it (tests/golden/make_align_golden.py)."""
import os

import numpy as np
import pytest

from oracle import reference_path as rp
from strique_b200.align import view_positions, run_length_levels, encode_signal
from . import align_cases as ac

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'align_golden.npz')


def _set(al, ps):
    al.gap_open_h, al.gap_open_v, al.gap_extension_h, al.gap_extension_v, al.dist_offset, al.dist_min = ps


def test_c_oracle_matches_golden_vectors_from_compiled_reference():
    g = np.load(GOLDEN)
    c = rp.CAligner()
    n = int(g['n'])
    assert n >= 200
    for k in range(n):
        _set(c, tuple(g['params_%d' % k]))
        score, a_idx, b_idx = c.align_overlap(g['a_%d' % k], g['b_%d' % k])
        assert np.float32(score) == g['score_%d' % k]
        assert np.array_equal(a_idx, g['a_idx_%d' % k])
        assert np.array_equal(b_idx, g['b_idx_%d' % k])


@pytest.mark.skipif(rp.load_pyseqan() is None, reason='oracle/_ref/pyseqan not built (no /root/reference here)')
def test_c_oracle_matches_compiled_reference_live():
    ref, c = rp.make_aligner('ref'), rp.CAligner()
    for ps, a, b in ac.cases(seed=11, n=240):
        _set(ref, ps)
        _set(c, ps)
        r1 = ref.align_overlap(a.tolist(), b.tolist())
        r2 = c.align_overlap(a, b)
        assert np.float32(r1[0]) == np.float32(r2[0])
        assert np.array_equal(np.array(r1[1], dtype=np.uint64), r2[1])
        assert np.array_equal(np.array(r1[2], dtype=np.uint64), r2[2])


def test_empty_inputs_follow_reference():
    c = rp.CAligner()
    s, a_idx, b_idx = c.align_overlap(np.zeros(0), np.array([1.0, 2.0]))
    assert s == np.finfo(np.float32).tiny and list(b_idx) == [0, 1] and len(a_idx) == 0


def test_rows_roundtrip_rebuilds_view_positions():
    """Host logic of the product: the compact per-flank-sample records carry a_idx/b_idx."""
    c = rp.CAligner()
    for ps, a, b in ac.cases(seed=5, n=150):
        _set(c, ps)
        _, a_idx, b_idx = c.align_overlap(a, b)
        rows = ac.rows_from_view_positions(a_idx, b_idx)
        a2, b2 = view_positions(rows, len(a))
        assert np.array_equal(a2, a_idx) and np.array_equal(b2, b_idx)


def test_run_length_and_codes():
    lev, s = run_length_levels(np.repeat([1.0, 2.0, 2.0, 3.0], 6))
    assert s == 6 and list(lev) == [1.0, 2.0, 2.0, 3.0]
    lev, s = run_length_levels([1.0, 2.0, 3.0])
    assert s == 1 and len(lev) == 3
    codes, vals = encode_signal([3.5, 1.25, 3.5, 2.0])
    assert codes.dtype == np.uint8 and np.array_equal(vals[codes], np.float32([3.5, 1.25, 3.5, 2.0]))
