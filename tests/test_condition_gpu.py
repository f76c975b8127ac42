"""GPU parity of the conditioning kernel (strique_condition_batch) against the oracle's numpy
restatement of scripts/STRique.py:590-597 + normalize2model('minmax'). Bit-exact."""
import os
import warnings

import numpy as np
import pytest

from oracle import reference_path as rp
from strique_b200 import fast5
from . import synth
from .conftest import C9_PREFIX, C9_SUFFIX, ROOT

pytestmark = pytest.mark.gpu


def _reads(pm):
    rng = np.random.default_rng(21)
    out = []
    for n, int16, noise in [(30, True, True), (5, False, True), (80, True, True), (40, False, False), (2, True, True)]:
        seq = synth.read_sequence(rng, C9_PREFIX, 'GGCCCC', C9_SUFFIX, n, flank=400)
        out.append(synth.simulate(pm, seq, rng, noise=noise, int16=int16))
    return out


def _check(ctx, pm, reads, want_raw):
    pore = (lambda q: None)
    means = pm.means
    q_lo, q_hi = np.percentile(means, [1, 99])
    consts = (float(np.median(means[means < q_lo])), float(np.median(means[means > q_hi])), float(pm.model_min), float(pm.model_max))
    flt, codes, vals, stats, off = ctx.condition_batch(consts, reads, want_raw_stats=want_raw)
    m5m, m95m = consts[0], consts[1]
    c3, c4 = (m95m - m5m) / 2, m5m + (m95m - m5m) / 2
    for k, raw in enumerate(reads):
        sl = slice(off[k], off[k + 1])
        raw = raw.astype(flt.dtype)
        f0 = rp.medfilt3(raw)
        assert np.array_equal(flt[sl], f0)
        u8 = rp.open_close_u8(rp.quantise_u8(f0))
        assert np.array_equal(codes[sl], u8.astype(np.uint16))
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            morph, fltn = rp.condition(pm, raw)
        if np.isnan(morph).any() or np.isnan(fltn).any():
            assert stats['status'][k] == 1      # empty percentile tail: NaN in numpy, flagged by the kernel
            continue
        assert np.array_equal(vals[k][codes[sl]], morph.astype(np.float32))
        assert stats['status'][k] == 0
        assert stats['flt_median'][k] == np.median(f0) and stats['flt_mad'][k] == pytest.approx(rp.PoreModel.MAD(f0), rel=1e-14)
        mine = ((f0.astype(np.float64) - stats['flt_c1'][k]) / stats['flt_c2'][k]) * c3 + c4
        np.clip(mine, pm.model_min + .5, pm.model_max - .5, out=mine)
        assert np.array_equal(mine, fltn)
        if want_raw:
            nrm = pm.normalize_minmax(raw.astype(np.float64))
            mine = ((raw.astype(np.float64) - stats['raw_c1'][k]) / stats['raw_c2'][k]) * c3 + c4
            np.clip(mine, pm.model_min + .5, pm.model_max - .5, out=mine)
            assert np.array_equal(mine, nrm)


def test_int16_and_float_reads(ctx, model_file):
    pm = rp.PoreModel(model_file)
    reads = _reads(pm)
    _check(ctx, pm, [r for r in reads if r.dtype == np.int16], want_raw=True)
    _check(ctx, pm, [r for r in reads if r.dtype != np.int16], want_raw=True)


def test_bundled_read(ctx, model_file):
    pm = rp.PoreModel(model_file)
    raw = fast5.read_raw_signal(os.path.join(ROOT, 'data', 'c9orf72.fast5'))
    _check(ctx, pm, [raw], want_raw=False)


def test_short_and_ragged_reads(ctx, model_file):
    pm = rp.PoreModel(model_file)
    rng = np.random.default_rng(4)
    reads = []
    for n in (301, 1000, 4097, 257, 513):
        levels = np.repeat(rng.uniform(450, 750, n // 7 + 2), rng.integers(5, 10, n // 7 + 2))[:n]
        reads.append(np.round(levels + rng.normal(0, 12, n)).astype(np.int16))
    _check(ctx, pm, reads, want_raw=True)


def test_degenerate_read_is_flagged(ctx, model_file):
    """White noise collapses under the 8-wide opening/closing: numpy's 'minmax' tail median is the
    mean of an empty slice (NaN) in the reference; the kernel reports status 1 instead."""
    pm = rp.PoreModel(model_file)
    rng = np.random.default_rng(4)
    raw = np.round(rng.normal(600, 80, 301)).astype(np.int16)
    with np.errstate(all='ignore'):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            morph, _ = rp.condition(pm, raw)
    consts = (61.72140829235028, 117.41709863727257, float(pm.model_min), float(pm.model_max))
    _, _, _, stats, _ = ctx.condition_batch(consts, [raw, np.full(500, 7, np.int16)])
    assert bool(np.isnan(morph).any()) == bool(stats['status'][0] == 1)
    assert stats['status'][1] == 1      # constant read: zero MAD
