import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (run with -m gpu on the GPU box)')


@pytest.fixture(scope='session', autouse=True)
def _build_oracle():
    """The oracle's C restatement is test infrastructure; build it once per session."""
    subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle'), 'liboracle.so'])


@pytest.fixture(scope='session')
def model_file():
    return os.path.join(ROOT, 'models', 'r9_4_450bps.model')


@pytest.fixture(scope='session')
def mod_model_file():
    return os.path.join(ROOT, 'models', 'r9_4_450bps_mCpG.model')


@pytest.fixture(scope='session')
def ctx():
    from strique_b200 import _lib
    return _lib.default_context(0)


C9_PREFIX = ('CGGCAGCCGAACCCCAAACAGCCACCCGCCAGGATGCCGCCTCCTCACTCACCCACTCGCCACCGCCTGCGCCTCCGCCGCCGCGGGCGCAGGCACCGC'
             'AACCGCAGCCCCGCCCCGGGCCCGCCCCCGGGCCCGCCCCGACCACGCCCC')
C9_SUFFIX = ('TAGCGCGCGACTCCTGAGTTCCAGAGCTTGCTACAGGCTGCGGTTGTTTCCCTCCTTGTTTTCTTCTGGTTAATCTTTATCAGGTCTTTTCTTGTTCAC'
             'CCTCAGCGAGTACTGTGAGAGCAAGTAGTGGGGAGAGAGGGTGGGAAAAAC')
FMR1_PREFIX = ('AGCGGGCCGGGGGTTCGGCCTCAGTCAGGCGCTCAGCTCCGTTTCGGTTTCACTTCCGGTGGAGGGCCGCCTCTGAGCGGGCGGCGGGCCGACGGCGAG'
               'CGCGGGCGGCGGCGGTGACGGAGGCGCCGCTGCCAGGGGGCGTGCGGCAGC')
FMR1_SUFFIX = ('GAGGCGGCGGCGGCGGCGGCGGCGGCGGCGGCTGGGCCTCGAGCGCCCGCAGCCCACCTCTCGGGGGCGGGCTCCCGGCGCTAGCAGGGCTGAAGAGAA'
               'GATGGAGGAGCTGGTGGTGGAAGTGCGGGGCTCCAATGGCGCTTTCTACAA')
