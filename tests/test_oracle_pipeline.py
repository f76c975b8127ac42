"""Pins the oracle (oracle/reference_path.py) against what the reference's own tests hold for the
path: the `n == i` assertions of scripts/STRique_test.py (recipes replayed with a seeded backbone)
and the documented offset / ticks of the bundled read (docs/installation/test.md:16)."""
import os

import numpy as np
import pytest

from oracle import reference_path as rp
from strique_b200 import fast5
from . import synth
from .conftest import C9_PREFIX, C9_SUFFIX, FMR1_PREFIX, FMR1_SUFFIX, ROOT


@pytest.fixture(scope='module')
def ref(model_file):
    r = rp.RefRepeatCounter(model_file)
    r.add_target('c9orf72', 'GGCCCC', C9_PREFIX, C9_SUFFIX)
    r.add_target('fmr1', 'GCG', FMR1_PREFIX, FMR1_SUFFIX)
    return r


def test_detection_recipe(ref, model_file):
    """scripts/STRique_test.py:45-63 (test_Detection): GGCCCC x i with 1000-nt backbones, samples=8."""
    pm = rp.PoreModel(model_file)
    bb = synth.backbone(np.random.default_rng(7), 2000)
    for i in (100, 200):
        sig = pm.generate_signal(bb[:1000] + C9_PREFIX + 'GGCCCC' * i + C9_SUFFIX + bb[-1000:], samples=8)
        assert ref.detect('c9orf72', sig, '+')[0] == i


def test_interpolation_recipe(ref, model_file):
    """scripts/STRique_test.py:67-83 (test_Interpolation): GCG x i, repeat shorter than k."""
    pm = rp.PoreModel(model_file)
    bb = synth.backbone(np.random.default_rng(8), 2000)
    sig = pm.generate_signal(bb[:1000] + FMR1_PREFIX + 'GCG' * 100 + FMR1_SUFFIX + bb[-1000:], samples=8)
    assert ref.detect('fmr1', sig, '+')[0] == 100


def test_normalization_recipe(ref, model_file):
    """scripts/STRique_test.py:86-101 (test_Normalization): no backbone, i = 10..90."""
    pm = rp.PoreModel(model_file)
    for i in range(10, 100, 20):
        sig = pm.generate_signal(C9_PREFIX + 'GGCCCC' * i + C9_SUFFIX, samples=8)
        assert ref.detect('c9orf72', sig, '+')[0] == i


def test_modification_recipe(model_file, mod_model_file):
    """scripts/STRique_test.py:104-124 (test_Modification): noisy base / mCpG signals, n == i."""
    r = rp.RefRepeatCounter(model_file, mod_model_file=mod_model_file)
    r.add_target('c9orf72', 'GGCCCC', C9_PREFIX, C9_SUFFIX)
    rng = np.random.default_rng(9)
    bb = synth.backbone(rng, 2000)
    seq = bb[:1000] + C9_PREFIX + 'GGCCCC' * 100 + C9_SUFFIX + bb[-1000:]
    frac = {}
    for tag, f in (('base', model_file), ('mod', mod_model_file)):
        sig = rp.PoreModel(f).generate_signal(seq, noise=True, rng=rng)
        out = r.detect('c9orf72', sig, '+')
        assert out[0] == 100
        frac[tag] = out[6].count('1') / max(len(out[6]), 1)
    assert frac['mod'] > 0.8 > 0.2 > frac['base']


@pytest.mark.skipif(rp.load_pyseqan() is None, reason='needs the compiled reference aligner (speed)')
def test_bundled_read_documented_integers(model_file):
    """docs/installation/test.md:16: offset 1633, ticks 40758 for data/c9orf72.fast5 (minus strand)."""
    raw = fast5.read_raw_signal(os.path.join(ROOT, 'data', 'c9orf72.fast5'))
    cols = open(os.path.join(ROOT, 'configs', 'repeat_config.tsv')).read().split('\n')[1].split()
    r = rp.RefRepeatCounter(model_file)
    r.add_target(cols[3], cols[4], cols[5], cols[6])
    out = r.detect('c9orf72', raw, '-')
    assert (int(out[4]), int(out[5])) == (1633, 40758)
    assert abs(out[0] - 735) <= 2


def test_pore_model_constants_and_template_signal(model_file, mod_model_file):
    """SURVEY §8 a1 / a2 (scripts/STRique.py:114-127, 182-195): the four scalars of both shipped models as verified
    against the reference's own arithmetic, and the noise-free
    template signal = k-mer means repeated `samples` times ((len - 5) * samples values)."""
    import numpy as np
    from strique_b200.pore_model import pore_model
    pm, pmm = pore_model(model_file), pore_model(mod_model_file)
    for got, want in zip((pm.model_median, pm.model_MAD, pm.model_min, pm.model_max), (91.1925, 10.6584, 49.5413, 133.1231)):
        assert abs(got - want) < 5e-5
    for got, want in zip((pmm.model_median, pmm.model_MAD, pmm.model_min, pmm.model_max), (91.2150, 10.7271, 47.3526, 131.3527)):
        assert abs(got - want) < 5e-5
    seq = 'ACGTTGCA' * 20
    sig = np.asarray(pm.generate_signal(seq, samples=6))
    assert len(sig) == (len(seq) - 5) * 6
    means = np.asarray(pm.kmer_means(seq))
    assert np.array_equal(sig, np.repeat(means, 6))
    ref = rp.PoreModel(model_file)
    assert np.array_equal(sig, np.asarray(ref.generate_signal(seq, samples=6)))
