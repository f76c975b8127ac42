"""GPU parity of boundary #2 (strique_viterbi_batch on models compiled by strique_b200.hmm)
against the oracle (pomegranate restatement + C float64 Viterbi): same best path state by state,
same visit counts / repeat interval / methylation pattern, log p within 1e-12 relative -- for the float64
kernels without exception, for the fixed-point profile kernel (the default for count models) except on
sequences whose best and second-best path the oracle itself puts closer than NEAR_TIE."""
import itertools

import numpy as np
import pytest

from oracle import reference_path as rp
from strique_b200 import hmm
from strique_b200.pore_model import pore_model
from . import synth
from .conftest import C9_PREFIX, C9_SUFFIX, FMR1_PREFIX, FMR1_SUFFIX

pytestmark = pytest.mark.gpu

NEAR_TIE = 1e-3      # nat: gap between the two best paths below which the fixed-point decoder may take the other one


@pytest.fixture
def exact(ctx):
    """float64 profile kernel only"""
    ctx.set_viterbi_exact(True)
    yield ctx
    ctx.set_viterbi_exact(False)


def _segments(pm_o, prefix, repeat, suffix, counts, seed):
    """Normalised HMM input segments cut by the oracle pipeline (flank alignment included)."""
    rng = np.random.default_rng(seed)
    dt = rp.RefRepeatCounter.__new__(rp.RefRepeatCounter)
    segs = []
    for n in counts:
        seq = prefix[-50:] + repeat * n + suffix[:50]
        raw = synth.simulate(pm_o, seq, rng, noise=True)
        # scale to the model like the pipeline does (the exact cut is irrelevant for this test)
        segs.append(pm_o.normalize_minmax(rp.medfilt3(raw).astype(np.float64)))
    return segs


@pytest.mark.parametrize('repeat,prefix,suffix', [('GGCCCC', C9_PREFIX, C9_SUFFIX), ('GCG', FMR1_PREFIX, FMR1_SUFFIX)])
def test_count_hmm_paths(exact, model_file, repeat, prefix, suffix):
    ctx = exact
    pm_o = rp.PoreModel(model_file)
    pm = pore_model(model_file)
    oracle = rp.FlankedRepeatHMM(repeat, prefix[-50:], suffix[:50], pm_o)
    g, off = hmm.flanked_repeat_graph(repeat, prefix[-50:], suffix[:50], pm)
    c = hmm.compile_graph(g)
    mid = ctx.hmm_create(c)
    segs = _segments(pm_o, prefix, repeat, suffix, [3, 10, 25, 60, 120], seed=5)
    segs.append(segs[0][:40])           # too short to traverse the model cleanly
    segs.append(np.full(300, 1000.0))   # outside every uniform range: impossible
    res, _, paths = ctx.viterbi_batch(mid, segs, want_path=True)
    for k, x in enumerate(segs):
        n0, p0, names0 = oracle.count_repeats(x)
        if not names0:
            assert res['status'][k] == 1
            continue
        assert res['status'][k] == 0
        assert res['logp'][k] == pytest.approx(p0, rel=1e-12)
        got_names = [c.names[i] for i in paths[k]]
        if k == 5 and got_names != names0:
            # the truncated sequence has to jump through the delete chains; with a periodic flank
            # (fmr1: GCG repeats inside the suffix) several jumps tie EXACTLY -- equal log p is all
            # that can be asked of either decoder there
            assert repeat == 'GCG'
            continue
        assert got_names == names0
        assert res['n_count'][k] + off == n0
        rep = np.array(['repeat' in s for s in names0])
        idx = np.flatnonzero(rep)
        assert (res['t_first'][k], res['t_last'][k]) == ((idx[0], idx[-1]) if len(idx) else (-1, -1))
        assert len(idx) == 0 or rep[idx[0]:idx[-1] + 1].all()
    # the fixed-point kernel on the same sequences: the oracle's path unless the oracle's own two best paths are a
    # near-tie; log p is the float64 score of the path it returns, so it can only fall short by less than that gap
    ctx.set_viterbi_exact(False)
    res_q, _, paths_q = ctx.viterbi_batch(mid, segs, want_path=True)
    assert ctx.last_viterbi_fixed[0] >= len(segs) - 2
    assert np.array_equal(res_q['status'], res['status'])
    for k, x in enumerate(segs):
        if res['status'][k] != 0:
            continue
        oracle.count_repeats(x)
        margin = oracle.model.last_margin
        assert res['logp'][k] - res_q['logp'][k] <= max(margin, 1e-9 * abs(res['logp'][k])) + 1e-12
        if margin > NEAR_TIE:
            assert np.array_equal(paths_q[k], paths[k]), (k, margin)
            for f in ('n_count', 't_first', 't_last'):
                assert res_q[f][k] == res[f][k]


def test_mod_hmm_patterns(ctx, model_file, mod_model_file):
    pm_o, pm_mo = rp.PoreModel(model_file), rp.PoreModel(mod_model_file)
    pm, pm_m = pore_model(model_file), pore_model(mod_model_file)
    oracle = rp.RepeatModHMM('GGCCCC', pm_o, pm_mo)
    g, lo, hi = hmm.repeat_mod_graph('GGCCCC', pm, pm_m)
    c = hmm.compile_graph(g)
    mid = ctx.hmm_create(c)
    rng = np.random.default_rng(8)
    segs = []
    for n, model in [(20, pm_o), (20, pm_mo), (75, pm_mo), (150, pm_o)]:
        raw = synth.simulate(model, 'GGCCCC' * n + 'GGCCC', rng, noise=True)
        segs.append(np.clip(raw, lo, hi))
    res, patterns, paths = ctx.viterbi_batch(mid, segs, want_path=True)
    for k, x in enumerate(segs):
        want = oracle.mod_repeats(x)
        p0, path0 = oracle.model.viterbi(np.clip(x, oracle.model_min, oracle.model_max))
        names0 = [s.name for i, s in path0 if i < oracle.model.silent_start]
        assert res['logp'][k] == pytest.approx(p0, rel=1e-12)
        assert [c.names[i] for i in paths[k]] == names0
        assert patterns[k] == want
    # methylated signal decodes mostly '1', unmethylated mostly '0'
    assert patterns[0].count('1') < 0.3 * len(patterns[0]) and patterns[1].count('1') > 0.7 * len(patterns[1])


def test_viterbi_kernels_agree(ctx, model_file, mod_model_file, monkeypatch):
    """The Viterbi kernels must decode identical paths, counts and log p: the generic kernel (csrc/viterbi.cu, forced
    with STRIQUE_VITERBI_GENERIC; any compiled model), the float64 profile kernel (csrc/viterbi_profile.cu: one warp
    per sequence, 4 positions per lane; count HMMs, forced with strique_set_viterbi_exact) and the small-model kernel
    (csrc/viterbi_small.cu: one state per lane; the methylation HMM) -- count HMMs of both loci / both strands and
    the methylation HMM, sequences from empty to 30 k samples, mixed in one batch.  The fixed-point profile kernel
    (csrc/viterbi_profile_q.cu, the default for count HMMs) must agree wherever the answer is not a tie: sequences it
    declines (NaN, samples outside the uniform ranges) come back from the float64 kernel and are identical."""
    pm_o = rp.PoreModel(model_file)
    pm, pm_m = pore_model(model_file), pore_model(mod_model_file)
    rng = np.random.default_rng(12)
    cases = []
    for repeat, prefix, suffix in (('GGCCCC', C9_PREFIX, C9_SUFFIX), ('GCG', FMR1_PREFIX, FMR1_SUFFIX),
                                   (synth.revcomp('GGCCCC'), synth.revcomp(C9_SUFFIX), synth.revcomp(C9_PREFIX)),
                                   ('ATTCT', C9_PREFIX, FMR1_SUFFIX)):
        g, _ = hmm.flanked_repeat_graph(repeat, prefix[-50:], suffix[:50], pm)
        mid = ctx.hmm_create(hmm.compile_graph(g))
        assert ctx.hmm_kernel_shape(mid) == 4000        # the count HMMs are served by the profile kernels
        segs = _segments(pm_o, prefix, repeat, suffix, [1, 2, 7, 33, 150, 640], seed=int(rng.integers(1 << 30)))
        x_out = segs[3].copy()
        x_out[[10, 200]] = 1000.0                       # outside every uniform range: slow emission path, impossible
        x_nan = segs[3].copy()
        x_nan[300] = np.nan
        segs += [segs[1][:5], segs[2][:1], np.full(50, 1000.0), np.full(64, np.nan), x_out, x_nan, np.zeros(0)]
        cases.append((mid, segs, True))
    g, lo, hi = hmm.repeat_mod_graph('GGCCCC', pm, pm_m)
    mid = ctx.hmm_create(hmm.compile_graph(g))
    assert ctx.hmm_kernel_shape(mid) == 32          # the methylation HMM: small-model kernel (one state per lane)
    mod_segs = [np.clip(synth.simulate(pm_o, 'GGCCCC' * n + 'GGCCC', rng, noise=True), lo, hi) for n in (1, 9, 200)]
    m_nan = mod_segs[1].copy()
    m_nan[7] = np.nan                                    # NaN sample: log 1 in every state
    m_out = mod_segs[1].copy()
    m_out[11] = 1000.0                                   # outside the uniform ranges of s0 / e0 / inserts
    mod_segs += [mod_segs[1][:1], mod_segs[1][:9], mod_segs[2][:257], m_nan, m_out, np.full(40, 1000.0), np.zeros(0)]
    cases.append((mid, mod_segs, False))

    def run(generic=False, exact_only=False):
        monkeypatch.delenv('STRIQUE_VITERBI_GENERIC', raising=False)
        if generic:
            monkeypatch.setenv('STRIQUE_VITERBI_GENERIC', '1')
        ctx.set_viterbi_exact(exact_only)
        try:
            return ctx.viterbi_batch(mid, segs, want_path=True)
        finally:
            ctx.set_viterbi_exact(False)
            monkeypatch.delenv('STRIQUE_VITERBI_GENERIC', raising=False)

    for mid, segs, has_profile in cases:
        r0, p0, path0 = run(generic=True)
        variants = [('float64', run(exact_only=True))]    # float64 profile kernel / small-model kernel
        if has_profile:
            variants.append(('fixed', run()))
            n_fixed, n_declined = ctx.last_viterbi_fixed
            assert n_fixed >= 5 and n_declined >= 5        # NaN / out-of-range / truncated sequences went to float64
        for name, (r1, p1, path1) in variants:
            fixed = name == 'fixed'
            assert np.array_equal(r1['status'], r0['status'])
            # log p to the last few ulps: paths that hop along a delete chain add the hop weights in the
            # association of the kernel's max-plus scan; the fixed-point kernel adds the path's terms in its own order
            # and may return the other path of a near-tie (its log p is then lower by less than the gap)
            assert np.allclose(r1['logp'], r0['logp'], rtol=1e-12 if fixed else 1e-13, atol=2e-3 if fixed else 0, equal_nan=True)
            for k in range(len(segs)):
                if r0['status'][k] != 0:
                    continue
                # exact ties (sequences too short to traverse the model, NaN samples that score log 1 in
                # every state) are broken by candidate order, which the profile kernels do not share
                # ... and so are samples outside every uniform range: only match states can emit them, the
                # path is forced along the delete chains, and flank positions with identical k-mers then tie
                # to the last ulp (hop weights summed in a different association) -- equal log p, asserted
                # above, is all that can be asked of either decoder there
                forced = bool((segs[k] > 500).any())
                degenerate = forced or (has_profile and (len(segs[k]) < 100 or np.isnan(segs[k]).any()))
                near_tie = fixed and abs(r1['logp'][k] - r0['logp'][k]) > 1e-9 * abs(r0['logp'][k])
                if (degenerate or near_tie) and not np.array_equal(path1[k], path0[k]):
                    continue
                if fixed and not np.array_equal(path1[k], path0[k]):
                    # same score to the last digits, different path: only possible if two paths tie
                    assert abs(r1['logp'][k] - r0['logp'][k]) <= 1e-9 * abs(r0['logp'][k])
                    continue
                assert np.array_equal(path1[k], path0[k]), (name, k)
                for f in ('n_count', 't_first', 't_last', 'pattern_len'):
                    assert r1[f][k] == r0[f][k], (f, name, k)
                assert p1[k] == p0[k]
