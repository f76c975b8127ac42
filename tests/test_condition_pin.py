"""Conditioning against REAL dependencies (tests/golden/condition_pin.npz, made by tests/golden/make_condition_pin.py
with scipy.signal.medfilt, numpy and scipy.ndimage.grey_erosion / grey_dilation behind a restatement of
scikit-image 0.14's 20-line wrapper): the uint8 codes after closing(opening(.)) of the bundled read
(scripts/STRique.py:590-595), byte for byte --
 * the oracle's window restatement (oracle/reference_path.py MORPH_WINDOWS), on the CPU;
 * the CUDA conditioning kernel, on the GPU."""
import os

import numpy as np
import pytest

from oracle import reference_path as rp
from strique_b200 import fast5
from .conftest import ROOT

PIN = np.load(os.path.join(ROOT, 'tests', 'golden', 'condition_pin.npz'))


def _raw():
    raw = fast5.read_raw_signal(os.path.join(ROOT, 'data', 'c9orf72.fast5'))
    assert len(raw) == int(PIN['n'])
    return raw


def test_oracle_windows_equal_scipy_ndimage():
    flt = rp.medfilt3(_raw())
    assert float(np.median(flt)) == float(PIN['median'])
    assert rp.PoreModel.MAD(flt) == pytest.approx(float(PIN['mad']), rel=1e-15)
    assert np.array_equal(rp.open_close_u8(rp.quantise_u8(flt)), PIN['u8'])


def test_oracle_windows_equal_scipy_on_random_signals():
    """The same through the generator script's real scipy calls, on signals with edges and plateaus of every phase."""
    import scipy.signal as sp
    from tests.golden.make_condition_pin import condition_u8
    rng = np.random.default_rng(7)
    for k in range(8):
        n = 500 + 37 * k
        x = np.round(rng.normal(500, 60, n) + 90 * np.sign(np.sin(np.arange(n) / (3.0 + k)))).astype(np.int16)
        flt, _, _, u8 = condition_u8(x)
        assert np.array_equal(flt, rp.medfilt3(x)) and np.array_equal(flt, sp.medfilt(x, 3))
        assert np.array_equal(u8, rp.open_close_u8(rp.quantise_u8(rp.medfilt3(x))))


@pytest.mark.gpu
def test_cuda_conditioning_equals_scipy_pin(ctx, model_file):
    pm = rp.PoreModel(model_file)
    means = pm.means
    q_lo, q_hi = np.percentile(means, [1, 99])
    consts = (float(np.median(means[means < q_lo])), float(np.median(means[means > q_hi])), float(pm.model_min), float(pm.model_max))
    flt, codes, vals, stats, off = ctx.condition_batch(consts, [_raw()])
    assert stats['flt_median'][0] == float(PIN['median'])
    assert stats['flt_mad'][0] == pytest.approx(float(PIN['mad']), rel=1e-14)
    assert np.array_equal(codes.astype(np.uint8), PIN['u8'])
