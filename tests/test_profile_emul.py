"""CPU check of the profile Viterbi kernel's model packing, lane arithmetic, back-pointer encoding and
traceback: the host emulator (tests/native/profile_emul.cpp, which compiles the kernel's own
profile_core.h / profile_pack.h) against the oracle (pomegranate restatement + C float64 Viterbi) --
same best path state by state, same counts and repeat interval, log p within 1e-12 relative."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import reference_path as rp
from strique_b200 import _lib, hmm
from strique_b200.pore_model import pore_model
from . import synth
from .conftest import C9_PREFIX, C9_SUFFIX, FMR1_PREFIX, FMR1_SUFFIX, ROOT

NATIVE = os.path.join(ROOT, 'tests', 'native')


@pytest.fixture(scope='module')
def emul():
    subprocess.check_call(['make', '-s', '-C', NATIVE])
    lib = ctypes.CDLL(os.path.join(NATIVE, 'libprofile_emul.so'))
    lib.strique_test_profile_fits.restype = ctypes.c_int
    lib.strique_test_profile_fits.argtypes = [ctypes.POINTER(_lib.HmmDesc), ctypes.c_char_p, ctypes.c_int]
    lib.strique_test_profile_emulate.restype = ctypes.c_int
    lib.strique_test_profile_emulate.argtypes = [ctypes.POINTER(_lib.HmmDesc), ctypes.c_void_p, ctypes.c_int64] + \
        [ctypes.c_void_p] * 5
    return lib


def _emulate(lib, c, x):
    d, keep = _lib.hmm_desc(c)
    x = np.ascontiguousarray(x, dtype=np.float64)
    logp = np.zeros(1)
    ints = np.zeros(3, dtype=np.int32)
    path = np.full(max(len(x), 1), -1, dtype=np.int32)
    st = lib.strique_test_profile_emulate(ctypes.byref(d), x.ctypes.data, len(x), logp.ctypes.data, ints[0:].ctypes.data,
                                          ints[1:].ctypes.data, ints[2:].ctypes.data, path.ctypes.data)
    del keep
    return st, float(logp[0]), int(ints[0]), int(ints[1]), int(ints[2]), path[:len(x)]


def _segments(pm_o, prefix, repeat, suffix, counts, seed):
    rng = np.random.default_rng(seed)
    segs = []
    for n in counts:
        seq = prefix[-50:] + repeat * n + suffix[:50]
        raw = synth.simulate(pm_o, seq, rng, noise=True)
        segs.append(pm_o.normalize_minmax(rp.medfilt3(raw).astype(np.float64)))
    return segs


CASES = [('GGCCCC', C9_PREFIX, C9_SUFFIX),
         ('GCG', FMR1_PREFIX, FMR1_SUFFIX),
         (synth.revcomp('GGCCCC'), synth.revcomp(C9_SUFFIX), synth.revcomp(C9_PREFIX)),
         ('ATTCT', C9_PREFIX, FMR1_SUFFIX),
         ('CTG', FMR1_PREFIX, C9_SUFFIX)]


@pytest.mark.parametrize('repeat,prefix,suffix', CASES)
def test_count_models_fit_the_profile_kernel(emul, model_file, repeat, prefix, suffix):
    pm = pore_model(model_file)
    g, _ = hmm.flanked_repeat_graph(repeat, prefix[-50:], suffix[:50], pm)
    c = hmm.compile_graph(g)
    assert c.emit_pos is not None
    d, keep = _lib.hmm_desc(c)
    why = ctypes.create_string_buffer(256)
    np_used = emul.strique_test_profile_fits(ctypes.byref(d), why, 256)
    assert np_used > 0, why.value
    assert np_used <= 128


def test_mod_model_has_no_profile_layout(model_file, mod_model_file):
    g, _, _ = hmm.repeat_mod_graph('GGCCCC', pore_model(model_file), pore_model(mod_model_file))
    assert hmm.compile_graph(g).emit_pos is None


@pytest.mark.parametrize('repeat,prefix,suffix', CASES[:3])
def test_emulated_kernel_decodes_the_oracle_path(emul, model_file, repeat, prefix, suffix):
    pm_o = rp.PoreModel(model_file)
    pm = pore_model(model_file)
    oracle = rp.FlankedRepeatHMM(repeat, prefix[-50:], suffix[:50], pm_o)
    g, off = hmm.flanked_repeat_graph(repeat, prefix[-50:], suffix[:50], pm)
    c = hmm.compile_graph(g)
    segs = _segments(pm_o, prefix, repeat, suffix, [1, 3, 10, 25, 60], seed=5)
    segs.append(segs[1][:40])            # too short to traverse the model cleanly: delete-chain jumps
    segs.append(segs[2][:1])
    segs.append(np.full(300, 1000.0))    # outside every uniform range: impossible
    x_nan = segs[3].copy()
    x_nan[100:103] = np.nan              # NaN samples score log 1 (pomegranate)
    segs.append(x_nan)
    x_out = segs[2].copy()
    x_out[50] = 1000.0                   # one sample outside the uniform ranges: the slow emission path
    segs.append(x_out)
    for k, x in enumerate(segs):
        n0, p0, names0 = oracle.count_repeats(x)
        st, logp, n_count, t_first, t_last, path = _emulate(emul, c, x)
        if not names0:
            assert st == 1, k
            continue
        assert st == 0, k
        assert logp == pytest.approx(p0, rel=1e-12)
        got = [c.names[i] for i in path]
        if got != names0 and (len(x) <= 40 or np.isnan(x).any()):
            # sequences too short to traverse the model jump through the delete chains, where several
            # jumps tie exactly (periodic flanks), and NaN samples score log 1 in every state, so moving a
            # transition across them permutes the same addends: equal log p is all that can be asked of
            # either decoder there (candidate order differs: DESIGN.md section 8)
            continue
        assert got == names0, k
        assert n_count + off == n0
        rep = np.array(['repeat' in s for s in names0])
        idx = np.flatnonzero(rep)
        assert (t_first, t_last) == ((idx[0], idx[-1]) if len(idx) else (-1, -1))


# ---- fixed-point kernel (csrc/profile_q.h, profile_q_pack.h, viterbi_profile_q.cu) --------------------------------
def _emulate_q(lib, c, x):
    lib.strique_test_profile_q_emulate.restype = ctypes.c_int
    lib.strique_test_profile_q_emulate.argtypes = [ctypes.POINTER(_lib.HmmDesc), ctypes.c_void_p, ctypes.c_int64] + \
        [ctypes.c_void_p] * 6
    d, keep = _lib.hmm_desc(c)
    x = np.ascontiguousarray(x, dtype=np.float64)
    logp, vfwd = np.zeros(1), np.zeros(1)
    ints = np.zeros(3, dtype=np.int32)
    path = np.full(max(len(x), 1), -1, dtype=np.int32)
    st = lib.strique_test_profile_q_emulate(ctypes.byref(d), x.ctypes.data, len(x), logp.ctypes.data, ints[0:].ctypes.data,
                                            ints[1:].ctypes.data, ints[2:].ctypes.data, path.ctypes.data, vfwd.ctypes.data)
    del keep
    return st, float(logp[0]), int(ints[0]), int(ints[1]), int(ints[2]), path[:len(x)], float(vfwd[0])


@pytest.mark.parametrize('tag,picks', [('c2', (0, 3, 77, 415, 499)), ('c4', (1, 25, 50, 114))])
def test_fixed_point_emulation_on_golden_reads(emul, model_file, tag, picks):
    """The fixed-point lane arithmetic (tagged int32 scores at 2^-16 nat, renormalisation, kill line) on HMM inputs of
    the big golden set -- long C2 reads, the near-tie reads 415 / 499 (top-two gap ~1e-6 nat) and the atxn10 minus-strand
    reads whose best path runs through emissions below -32 nat: the golden count, the oracle's log p to within the
    read's own top-two gap (the re-score is float64), and a forward value inside the quantisation bound."""
    import json
    from strique_b200 import workload
    G = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'pipeline_golden_big.json')))
    col = {n: i for i, n in enumerate(G['columns'])}
    s = G['sets'][tag]
    pm_o, pm = rp.PoreModel(model_file), pore_model(model_file)
    kw = dict(s['kwargs'])
    kw['n_reads'] = max(picks) + 1
    reads = workload.make_reads(pm, **kw)
    for k in picks:
        name, sig, strand, _ = reads[k]
        row = s['rows'][k]
        rep, pre, suf = workload.LOCI[name]
        pre, suf = pre[-50:], suf[:50]
        if strand == '-':
            rep, pre, suf = synth.revcomp(rep), synth.revcomp(suf), synth.revcomp(pre)
        g, off = hmm.flanked_repeat_graph(rep, pre, suf, pm)
        c = hmm.compile_graph(g)
        _, fltn = rp.condition(pm_o, sig)
        x = fltn[row[col['prefix_begin']]:row[col['suffix_end']]]
        st, logp, n_count, t_first, t_last, path, vfwd = _emulate_q(emul, c, x)
        assert st == 0, (tag, k)
        assert n_count + off == row[col['count']], (tag, k)
        gap = row[col['log_p']] - logp
        assert -1e-9 * abs(logp) <= gap <= max(row[col['margin']], 1e-9 * abs(logp)), (tag, k, gap, row[col['margin']])
        assert abs(vfwd - logp) <= 1e-3 + len(x) * 2.0 ** -15, (tag, k)


def test_fixed_point_declines_what_it_cannot_vouch_for(emul, model_file):
    pm = pore_model(model_file)
    g, _ = hmm.flanked_repeat_graph('GGCCCC', C9_PREFIX[-50:], C9_SUFFIX[:50], pm)
    c = hmm.compile_graph(g)
    pm_o = rp.PoreModel(model_file)
    x = _segments(pm_o, C9_PREFIX, 'GGCCCC', C9_SUFFIX, [10], seed=3)[0]
    assert _emulate_q(emul, c, x)[0] == 0
    x_nan = x.copy()
    x_nan[17] = np.nan
    assert _emulate_q(emul, c, x_nan)[0] == 3            # NaN sample
    x_out = x.copy()
    x_out[40] = 1000.0
    assert _emulate_q(emul, c, x_out)[0] == 3            # outside the Uniform ranges
