"""The C-ABI library loads here (no GPU) and exports every function include/strique_b200.h declares;
the product path fails loudly -- no CPU fallback -- when there is no B200."""
import ctypes
import os
import re
import subprocess

import pytest

from .conftest import ROOT

HEADER = os.path.join(ROOT, 'include', 'strique_b200.h')
LIB = os.path.join(ROOT, 'strique_b200', 'libstrique_b200.so')


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(strique_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib():
    if not os.path.exists(LIB):
        subprocess.check_call(['bash', os.path.join(ROOT, 'build.sh')])
    return ctypes.CDLL(LIB)


def test_header_declares_the_boundary():
    names = declared_functions()
    for must in ('strique_ctx_create', 'strique_align_batch', 'strique_viterbi_batch', 'strique_condition_batch',
                 'strique_detect_batch', 'strique_hmm_create', 'strique_target_create'):
        assert must in names


def test_library_exports_every_declared_symbol(lib):
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.strique_version() >= 100


def test_no_cpu_fallback_without_a_device(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from strique_b200 import _lib
    with pytest.raises(_lib.StriqueError):
        _lib.Context(0)
    from strique_b200.counter import repeatCounter
    from .conftest import C9_PREFIX, C9_SUFFIX
    dt = repeatCounter(os.path.join(ROOT, 'models', 'r9_4_450bps.model'))
    with pytest.raises(_lib.StriqueError):
        dt.add_target('c9orf72', 'GGCCCC', C9_PREFIX, C9_SUFFIX)   # registering the HMMs needs the device


def test_flank_shapes_the_alignment_kernels_hold(lib):
    """host-only query used by repeatCounter.add_target: the limits the header states"""
    assert lib.strique_align_supported(145, 6) == 1          # the reference's default: 150 nt flanks, 6 samples per level
    assert lib.strique_align_supported(341, 6) == 1 and lib.strique_align_supported(342, 6) == 0      # 2048 flank samples
    assert lib.strique_align_supported(256, 8) == 1 and lib.strique_align_supported(257, 8) == 0
    assert lib.strique_align_supported(0, 6) == 0
    from strique_b200.counter import repeatCounter
    from .conftest import C9_SUFFIX
    dt = repeatCounter(os.path.join(ROOT, 'models', 'r9_4_450bps.model'))
    with pytest.raises(ValueError, match='does not fit the alignment kernels'):
        dt.add_target('long', 'GGCCCC', 'ACGT' * 100, C9_SUFFIX)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'strique_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', text, flags=re.M), f
                assert 'liboracle' not in text, f
