"""ORACLE (test infrastructure, NOT product code).

CPU restatement of giesselmann/STRique's per-read repeat-detection hot path
(scripts/STRique.py:113-618 at commit f4ee01b): pore model, read conditioning, flank alignment
glue, profile / repeat / flanked / methylation HMM topologies and `repeatCounter.detect`.
numpy/scipy for the array work, oracle/pomegranate_min.py + oracle/viterbi_oracle.c for the HMM,
and for the alignment either the REAL reference aligner compiled into oracle/_ref/ (preferred,
`aligner='ref'`) or its C restatement oracle/align_oracle.c (`aligner='c'`).

Pinned by: the `n == i` assertions of scripts/STRique_test.py (replayed in
tests/test_oracle_pipeline.py), the documented offset/ticks of the bundled read
(docs/installation/test.md:16) and the compiled reference aligner.  Float outputs of the HMM
stage and the `mod` string are PARITY UNPINNED (no reference test pins them; pomegranate and
scikit-image<0.15 are not installable here).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import ctypes
import importlib.util
import itertools
import math
import os
import sys

import numpy as np

from . import pomegranate_min as pg

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF_DIR = os.path.join(_HERE, '_ref')


# ---------------------------------------------------------------------------------------------
# aligners
# ---------------------------------------------------------------------------------------------
class _AlignParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_float) for n in
                ('gap_ext_h', 'gap_ext_v', 'gap_open_h', 'gap_open_v', 'dist_offset', 'dist_min')]


class CAligner:
    """oracle/align_oracle.c behind the property interface of pyseqan.align_raw
    (src/pyalign.cpp:47-62; defaults src/align_raw.h:52-60)."""

    def __init__(self):
        self.gap_open_h = -2.0
        self.gap_open_v = -2.0
        self.gap_extension_h = -8.0
        self.gap_extension_v = -8.0
        self.dist_offset = 8.0
        self.dist_min = -16.0
        self._lib = pg._liboracle()
        self._lib.strique_oracle_align.restype = ctypes.c_int
        self._lib.strique_oracle_align_score.restype = ctypes.c_int

    def _params(self):
        return _AlignParams(self.gap_extension_h, self.gap_extension_v, self.gap_open_h, self.gap_open_v,
                            self.dist_offset, self.dist_min)

    def align_overlap(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        a_idx = np.empty(len(a), dtype=np.uint64)
        b_idx = np.empty(len(b), dtype=np.uint64)
        score = ctypes.c_float(0)
        p = self._params()
        rc = self._lib.strique_oracle_align(ctypes.c_void_p(a.ctypes.data), ctypes.c_int64(len(a)),
                                            ctypes.c_void_p(b.ctypes.data), ctypes.c_int64(len(b)),
                                            ctypes.byref(p), ctypes.byref(score),
                                            ctypes.c_void_p(a_idx.ctypes.data), ctypes.c_void_p(b_idx.ctypes.data),
                                            None)
        if rc != 0:
            raise MemoryError('oracle alignment failed')
        return score.value, a_idx, b_idx

    def score_only(self, a, b):
        a = np.ascontiguousarray(a, dtype=np.float32)
        b = np.ascontiguousarray(b, dtype=np.float32)
        score = ctypes.c_float(0)
        bj = ctypes.c_int64(0)
        p = self._params()
        rc = self._lib.strique_oracle_align_score(ctypes.c_void_p(a.ctypes.data), ctypes.c_int64(len(a)),
                                                  ctypes.c_void_p(b.ctypes.data), ctypes.c_int64(len(b)),
                                                  ctypes.byref(p), ctypes.byref(score), ctypes.byref(bj))
        if rc != 0:
            raise MemoryError('oracle alignment failed')
        return score.value, bj.value


def load_pyseqan():
    """Import the compiled reference module from oracle/_ref (None if it was not built)."""
    if 'pyseqan' in sys.modules:
        return sys.modules['pyseqan']
    if not os.path.isdir(_REF_DIR):
        return None
    for f in sorted(os.listdir(_REF_DIR)):
        if f.startswith('pyseqan') and f.endswith('.so'):
            spec = importlib.util.spec_from_file_location('pyseqan', os.path.join(_REF_DIR, f))
            try:
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
            except ImportError:
                continue
            sys.modules['pyseqan'] = mod
            return mod
    return None


def make_aligner(kind='auto'):
    if kind in ('auto', 'ref'):
        mod = load_pyseqan()
        if mod is not None:
            return mod.align_raw()
        if kind == 'ref':
            raise RuntimeError('oracle/_ref/pyseqan is not built (run oracle/build_ref.sh where /root/reference exists)')
    return CAligner()


# ---------------------------------------------------------------------------------------------
# pore model  (scripts/STRique.py:113-195)
# ---------------------------------------------------------------------------------------------
class PoreModel:
    def __init__(self, model_file):
        table = {}
        with open(model_file, 'r') as fp:
            for line in fp:
                f = line.strip().split('\t')[:3]
                table[f[0]] = (float(f[1]), float(f[2]))
        self.table = table
        self.kmer = len(next(iter(table)))
        means = np.array([v[0] for v in table.values()])
        self.means = means
        self.model_median = np.median(means)                               # S.py:122
        self.model_MAD = np.mean(np.absolute(np.subtract(means, self.model_median)))  # S.py:123
        lo = min(table.values(), key=lambda v: v[0])
        hi = max(table.values(), key=lambda v: v[0])
        self.model_min = lo[0] - 6 * lo[1]                                  # S.py:126
        self.model_max = hi[0] + 6 * hi[1]                                  # S.py:127

    @staticmethod
    def MAD(signal):                                                        # S.py:142-143
        return np.mean(np.absolute(np.subtract(signal, np.median(signal))))

    def scale2stdv(self, other):                                            # S.py:145-148
        mine = np.median(np.array([v[1] for v in self.table.values()]))
        theirs = np.median(np.array([v[1] for v in other.table.values()]))
        return theirs / mine

    def normalize_minmax(self, signal):                                     # S.py:151-160,178-179
        q_lo, q_hi = np.percentile(signal, [1, 99])
        m_lo, m_hi = np.percentile(self.means, [1, 99])
        s5 = np.median(signal[signal < q_lo])
        s95 = np.median(signal[signal > q_hi])
        m5 = np.median(self.means[self.means < m_lo])
        m95 = np.median(self.means[self.means > m_hi])
        out = (signal - (s5 + (s95 - s5) / 2)) / ((s95 - s5) / 2)
        out = out * ((m95 - m5) / 2) + (m5 + (m95 - m5) / 2)
        np.clip(out, self.model_min + .5, self.model_max - .5, out=out)
        return out

    def generate_signal(self, sequence, samples=10, noise=False, rng=None):  # S.py:182-195
        k = self.kmer
        kmers = [sequence[i:i + k] for i in range(len(sequence) - k + 1)]
        means = np.array([self.table[x][0] for x in kmers])
        rnd = np.random if rng is None else rng
        if samples and not noise:
            return np.repeat(means, samples)
        if not noise:
            return np.repeat(means, rnd.uniform(6, 10, len(means)).astype(int))
        stdvs = np.array([self.table[x][1] for x in kmers])
        dwell = rnd.uniform(6, 10, len(means)).astype(int)
        return rnd.normal(np.repeat(means, dwell), np.repeat(stdvs, dwell))


# ---------------------------------------------------------------------------------------------
# read conditioning  (scripts/STRique.py:590-597)
# ---------------------------------------------------------------------------------------------
def medfilt3(x):
    """scipy.signal.medfilt(x, 3): zero padded ends, dtype preserved (S.py:590)."""
    x = np.asarray(x)
    p = np.concatenate([np.zeros(1, x.dtype), x, np.zeros(1, x.dtype)])
    a, b, c = p[:-2], p[1:-1], p[2:]
    return np.maximum(np.minimum(a, b), np.minimum(np.maximum(a, b), c))


def _window_extreme(x, lo, hi, fn):
    """fn over x[i+lo .. i+hi] with scipy.ndimage 'reflect' borders (d c b a | a b c d | d c b a)."""
    n = len(x)
    left, right = -lo, hi
    p = np.concatenate([x[:left][::-1], x, x[::-1][:right]]) if n >= max(left, right) else \
        np.pad(x, (left, right), mode='symmetric')
    out = p[0:n].copy()
    for k in range(1, left + right + 1):
        out = fn(out, p[k:k + n])
    return out


# Window conventions of scikit-image < 0.15 for the even 1x8 structuring element (the reason the
# reference pins that version, requirements.txt:9): the element is zero-padded to 9 taps, the
# second pass of opening/closing pads on the other side, dilation passes the reversed element to
# scipy.ndimage.grey_dilation (which reverses it again).  Resulting sample windows:
MORPH_WINDOWS = {
    'open_erode': (-3, 4), 'open_dilate': (-4, 3),
    'close_dilate': (-3, 4), 'close_erode': (-4, 3),
}


def open_close_u8(u8, windows=MORPH_WINDOWS):
    """closing(opening(u8, rectangle(1,8)), rectangle(1,8)) (S.py:593-595)."""
    e = _window_extreme(u8, *windows['open_erode'], np.minimum)
    o = _window_extreme(e, *windows['open_dilate'], np.maximum)
    d = _window_extreme(o, *windows['close_dilate'], np.maximum)
    return _window_extreme(d, *windows['close_erode'], np.minimum)


def quantise_u8(flt):
    """(flt - median)/MAD * 24 + 127, clipped and truncated to uint8 (S.py:591-592)."""
    z = (flt - np.median(flt)) / PoreModel.MAD(flt)
    return np.clip(z * 24 + 127, 0, 255).astype(np.uint8)


def condition(pm, raw, windows=MORPH_WINDOWS):
    """-> (morph signal for the aligner, median-filtered signal for the count HMM), both float64."""
    flt = medfilt3(raw)
    u8 = open_close_u8(quantise_u8(flt), windows)
    morph = pm.normalize_minmax(u8.astype(np.float64))
    fltn = pm.normalize_minmax(flt.astype(np.float64))
    return morph, fltn


# ---------------------------------------------------------------------------------------------
# HMM topologies  (scripts/STRique.py:201-500)
# ---------------------------------------------------------------------------------------------
PROFILE_DEFAULTS = {'match_loop': .75, 'match_match': .15, 'match_insert': .09, 'match_delete': .01,
                    'insert_loop': .15, 'insert_match_0': .40, 'insert_match_1': .40, 'insert_delete': .05,
                    'delete_delete': .005, 'delete_insert': .05, 'delete_match': .945}


class Profile:
    """profileHMM (S.py:201-307): states/edges are added to `model`; exposes s1, s2, e1, e2."""

    def __init__(self, model, sequence, pm, probs, prefix, no_silent=False, std_scale=1.0, std_offset=0.0):
        tp = dict(PROFILE_DEFAULTS)
        tp.update(probs or {})
        k = pm.kmer
        n = len(sequence) - k + 1
        digits = int(np.ceil(np.log10(n)))
        M, I, D = [], [], []
        for idx in range(n):
            name = prefix + str(idx).rjust(digits, '0')
            mean, std = pm.table[sequence[idx:idx + k]]
            M.append(pg.State(pg.NormalDistribution(mean, std * std_scale + std_offset), name=name + 'm'))
            if not no_silent:
                D.append(pg.State(None, name=name + 'd'))
            I.append(pg.State(pg.UniformDistribution(pm.model_min, pm.model_max), name=name + 'i'))
        self.s1, self.s2 = pg.State(None, name=prefix + 's1'), pg.State(None, name=prefix + 's2')
        self.e1, self.e2 = pg.State(None, name=prefix + 'e1'), pg.State(None, name=prefix + 'e2')
        model.add_states(M)
        model.add_states(I)
        if not no_silent:
            model.add_states(D)
        model.add_states([self.s1, self.s2, self.e1, self.e2])
        t = model.add_transition
        for i in range(n):
            t(M[i], M[i], tp['match_loop'])
            if i < n - 1:
                t(M[i], M[i + 1], tp['match_match'])
        for i in range(n):
            t(I[i], I[i], tp['insert_loop'])
            t(M[i], I[i], tp['match_insert'])
            t(I[i], M[i], tp['insert_match_1'])
            if i < len(D) - 1 and not no_silent:
                t(I[i], D[i + 1], tp['insert_delete'])
            if i < n - 1:
                t(I[i], M[i + 1], tp['insert_match_0'])
        if not no_silent:
            for i in range(n):
                t(D[i], I[i], tp['delete_insert'])
                if i > 0:
                    t(M[i - 1], D[i], tp['match_delete'])
                if i < n - 1:
                    t(D[i], M[i + 1], tp['delete_match'])
                if i < n - 1:
                    t(D[i], D[i + 1], tp['delete_delete'])
            t(self.s1, D[0], 1)
            t(self.s2, M[0], 1)
            t(D[-1], self.e1, tp['delete_delete'])
            t(D[-1], self.e2, tp['delete_match'])
        else:
            for i in range(n - 2):
                t(M[i], M[i + 2], tp['match_delete'])
            t(self.s1, I[0], 1)
            t(self.s2, M[0], 1)
        t(I[-1], self.e1, tp['insert_delete'])
        t(I[-1], self.e2, tp['insert_match_0'])
        t(M[-1], self.e2, tp['match_match'])
        t(M[-1], self.e1, tp['match_delete'])
        self.M, self.I, self.D = M, I, D


def repeat_unit_string(repeat, k):
    """S.py:329-335 -> (unit string, repeat_offset)."""
    if len(repeat) >= k:
        return repeat + repeat[:k - 1], 0
    ext = k - 1 + (len(repeat) - 1) - ((k - 1) % len(repeat))
    unit = repeat + (repeat * k)[:ext]
    return unit, int(len(unit) / len(repeat)) - 1


class RepeatLoop:
    """repeatHMM (S.py:313-378) added into `model`."""

    def __init__(self, model, repeat, pm, probs, prefix, std_scale=1.0, std_offset=0.0):
        tp = {'skip': .999, 'leave_repeat': .002}
        tp.update(probs or {})
        unit, self.repeat_offset = repeat_unit_string(repeat, pm.kmer)
        prof = Profile(model, unit, pm, tp, prefix, no_silent=True, std_scale=std_scale, std_offset=std_offset)
        self.d1 = pg.State(pg.UniformDistribution(pm.model_min, pm.model_max), name=prefix + 'dummy1')
        self.d2 = pg.State(pg.UniformDistribution(pm.model_min, pm.model_max), name=prefix + 'dummy2')
        self.e1, self.e2 = pg.State(None, name=prefix + 'e1'), pg.State(None, name=prefix + 'e2')
        self.s1, self.s2 = prof.s1, prof.s2
        # NB the reference never add_state()s its own e1/e2; networkx creates them with the edges
        model.add_state(self.d1)
        model.add_state(self.d2)
        t = model.add_transition
        t(prof.e1, self.d1, 1)
        t(prof.e2, self.d2, 1)
        t(self.d1, self.e1, tp['leave_repeat'])
        t(self.d2, self.e2, tp['leave_repeat'])
        t(self.d1, self.s1, 1 - tp['leave_repeat'])
        t(self.d2, self.s2, 1 - tp['leave_repeat'])
        model.add_state(self.e1)
        model.add_state(self.e2)
        self.profile = prof


class FlankedRepeatHMM:
    """flankedRepeatHMM (S.py:384-441)."""

    def __init__(self, repeat, prefix, suffix, pm, config=None):
        tp = {'skip': 1 - 1e-4, 'seq_std_scale': 1.0, 'rep_std_scale': 1.0, 'seq_std_offset': 0.0,
              'rep_std_offset': 0.0, 'e1_ratio': 0.1}
        if config and isinstance(config, dict):
            tp.update(config)
        c = int(np.ceil(pm.kmer / len(repeat)))
        pre = prefix + (repeat * c)[:-1]
        suf = repeat * c + suffix
        self.flanking_count = c * 2 - 1
        m = pg.HiddenMarkovModel()
        P = Profile(m, pre, pm, tp, 'prefix', std_scale=tp['seq_std_scale'], std_offset=tp['seq_std_offset'])
        R = RepeatLoop(m, repeat, pm, tp, 'repeat', std_scale=tp['rep_std_scale'], std_offset=tp['rep_std_offset'])
        S = Profile(m, suf, pm, tp, 'suffix', std_scale=tp['seq_std_scale'], std_offset=tp['seq_std_offset'])
        t = m.add_transition
        t(m.start, P.s1, tp['e1_ratio'])
        t(m.start, P.s2, 1 - tp['e1_ratio'])
        t(P.e1, R.s1, 1)
        t(P.e2, R.s2, 1)
        t(R.e1, S.s1, 1)
        t(R.e2, S.s2, 1)
        t(S.e1, m.end, 1)
        t(S.e2, m.end, 1)
        m.bake(merge='All')
        self.model, self.repeat_loop = m, R

    def count_repeats(self, sequence):
        """-> (n, log p, emitting-state names) (S.py:433-441, 374-378)."""
        p, path = self.model.viterbi(sequence)
        if path is None:
            return 0, 0, []
        n1 = sum(1 for _, s in path if s is self.repeat_loop.d1)
        n2 = sum(1 for _, s in path if s is self.repeat_loop.d2)
        n = n1 + n2 - self.repeat_loop.repeat_offset + self.flanking_count
        return n, p, [s.name for i, s in path if i < self.model.silent_start]


class RepeatModHMM:
    """repeatModHMM (S.py:447-500)."""

    def __init__(self, repeat, pm_base, pm_mod, config=None):
        tp = {'rep_std_scale': 1.5, 'rep_std_offset': 0.0, 'leave_repeat': .002}
        if config and isinstance(config, dict):
            tp.update(config)
        unit, _ = repeat_unit_string(repeat, pm_base.kmer)
        self.model_min = min(pm_base.model_min, pm_mod.model_min)
        self.model_max = max(pm_base.model_max, pm_mod.model_max)
        m = pg.HiddenMarkovModel()
        s0 = pg.State(pg.UniformDistribution(self.model_min, self.model_max), name='s0')
        e0 = pg.State(pg.UniformDistribution(self.model_min, self.model_max), name='e0')
        base = Profile(m, unit, pm_base, tp, 'base', no_silent=True, std_scale=tp['rep_std_scale'],
                       std_offset=tp['rep_std_offset'])
        mod = Profile(m, unit, pm_mod, tp, 'mod', no_silent=True,
                      std_scale=tp['rep_std_scale'] * pm_mod.scale2stdv(pm_base), std_offset=tp['rep_std_offset'])
        m.add_state(s0)
        m.add_state(e0)
        t = m.add_transition
        t(m.start, s0, 1)
        for s in (base.s1, base.s2, mod.s1, mod.s2):
            t(s0, s, 0.25)
        for e in (base.e1, base.e2, mod.e1, mod.e2):
            t(e, e0, 1)
        t(e0, m.end, tp['leave_repeat'])
        t(e0, s0, 1 - tp['leave_repeat'])
        m.bake(merge='All')
        self.model = m

    def mod_repeats(self, signal):
        p, path = self.model.viterbi(np.clip(signal, self.model_min, self.model_max))
        if path is None:
            return '-'
        names = [s.name for i, s in path if i < self.model.silent_start]
        firsts = [next(g) for k, g in itertools.groupby(names, key=lambda x: x not in ('s0', 'e0')) if k]
        return ''.join('1' if 'mod' in x else '0' for x in firsts)


# ---------------------------------------------------------------------------------------------
# repeatCounter  (scripts/STRique.py:505-618)
# ---------------------------------------------------------------------------------------------
_COMPLEMENT = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A'}


def reverse_complement(seq):
    return ''.join(_COMPLEMENT.get(b, b) for b in reversed(seq))


ALIGN_DEFAULTS = {'dist_offset': 16.0, 'dist_min': 0.0, 'gap_open_h': -1.0, 'gap_open_v': -16.0,
                  'gap_extension_h': -1.0, 'gap_extension_v': -16.0, 'samples': 6}


class RefRepeatCounter:
    def __init__(self, model_file, mod_model_file=None, align_config=None, HMM_config=None, aligner='auto',
                 windows=MORPH_WINDOWS):
        cfg = dict(ALIGN_DEFAULTS)
        if align_config and isinstance(align_config, dict):
            cfg.update(align_config)
        self.algn = make_aligner(aligner)
        for key in ('dist_offset', 'dist_min', 'gap_open_h', 'gap_open_v', 'gap_extension_h', 'gap_extension_v'):
            setattr(self.algn, key, cfg[key])
        self.pm = PoreModel(model_file)
        self.pm_mod = PoreModel(mod_model_file) if mod_model_file else self.pm
        self.samples = cfg['samples']
        self.HMM_config = HMM_config
        self.targets = {}
        self.windows = windows

    def detect_range(self, signal, segment, pre_trim=0, post_trim=0):      # S.py:538-548
        score, idx_signal, idx_segment = self.algn.align_overlap(signal, segment)
        idx_signal = np.array(idx_signal)
        begin = np.abs(idx_signal - idx_segment[0]).argmin()
        end = np.abs(idx_signal - idx_segment[-1]).argmin()
        score = score / (end - begin) if end > begin else 0.0
        begin = np.abs(idx_signal - idx_segment[0 + pre_trim]).argmin()
        end = np.abs(idx_signal - idx_segment[-1 - post_trim]).argmin()
        return score, begin, end

    def add_target(self, name, repeat, prefix, suffix):                    # S.py:553-579
        if name in self.targets:
            raise ValueError('RepeatCounter: Target with name ' + str(name) + ' already defined.')
        prefix_ext, suffix_ext = prefix.upper(), suffix.upper()
        prefix, suffix, repeat = prefix[-50:].upper(), suffix[:50].upper(), repeat.upper()
        rc = reverse_complement
        gen = lambda s: self.pm.generate_signal(s, samples=self.samples)
        plus = dict(prefix=gen(prefix), suffix=gen(suffix), prefix_ext=gen(prefix_ext), suffix_ext=gen(suffix_ext),
                    repeatHMM=FlankedRepeatHMM(repeat, prefix, suffix, self.pm, self.HMM_config),
                    modHMM=RepeatModHMM(repeat, self.pm, self.pm_mod, config=self.HMM_config))
        minus = dict(prefix=gen(rc(suffix)), suffix=gen(rc(prefix)), prefix_ext=gen(rc(suffix_ext)),
                     suffix_ext=gen(rc(prefix_ext)),
                     repeatHMM=FlankedRepeatHMM(rc(repeat), rc(suffix), rc(prefix), self.pm, self.HMM_config),
                     modHMM=RepeatModHMM(rc(repeat), self.pm, self.pm_mod, config=self.HMM_config))
        self.targets[name] = (plus, minus)

    def detect(self, name, raw_signal, strand, details=None):              # S.py:581-618
        if name not in self.targets:
            raise ValueError('RepeatCounter: Target with name ' + str(name) + ' not defined.')
        if strand == '+':
            tc = self.targets[name][0]
        elif strand == '-':
            tc = self.targets[name][1]
        else:
            raise ValueError('RepeatCounter: Strand must be + or -.')
        raw_signal = np.asarray(raw_signal)
        morph, fltn = condition(self.pm, raw_signal, self.windows)
        trim_prefix = len(tc['prefix_ext']) - len(tc['prefix'])
        trim_suffix = len(tc['suffix_ext']) - len(tc['suffix'])
        score_prefix, prefix_begin, prefix_end = self.detect_range(morph, tc['prefix_ext'], pre_trim=trim_prefix)
        score_suffix, suffix_begin, suffix_end = self.detect_range(morph, tc['suffix_ext'], post_trim=trim_suffix)
        n, p, states, mod_pattern = 0, 0, [], '-'
        margin = None
        if prefix_begin < suffix_end and score_prefix > 0.0 and score_suffix > 0.0:
            n, p, states = tc['repeatHMM'].count_repeats(fltn[prefix_begin:suffix_end])
            margin = getattr(tc['repeatHMM'].model, 'last_margin', None)
            if self.pm is not self.pm_mod:
                nrm = self.pm.normalize_minmax(raw_signal.astype(np.float64))
                mask = np.array(['repeat' in s for s in states], dtype=bool)
                rep = nrm[prefix_begin:suffix_end][mask] if len(states) else nrm[0:0]
                mod_pattern = tc['modHMM'].mod_repeats(rep)
        if details is not None:
            details.update(prefix_begin=int(prefix_begin), prefix_end=int(prefix_end),
                           suffix_begin=int(suffix_begin), suffix_end=int(suffix_end), states=states,
                           margin=margin)
        return n, score_prefix, score_suffix, p, prefix_end, max(suffix_begin - prefix_end, 0), mod_pattern
