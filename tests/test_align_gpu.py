"""GPU parity of boundary #1 (strique_align_batch behind strique_b200.align.align_raw) against
the golden vectors of the compiled reference aligner and against the oracle on seeded inputs.
Bit-exact: fp32 score, every view position, and the four __detect_range__ indices."""
import os

import numpy as np
import pytest

from oracle import reference_path as rp
from strique_b200 import align as sa
from . import align_cases as ac
from .conftest import C9_PREFIX, C9_SUFFIX

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'align_golden.npz')


def _set(al, ps):
    al.gap_open_h, al.gap_open_v, al.gap_extension_h, al.gap_extension_v, al.dist_offset, al.dist_min = ps


def test_golden_vectors_of_compiled_reference(ctx):
    g = np.load(GOLDEN)
    n = int(g['n'])
    for ps in ac.PARAM_SETS:
        al = sa.align_raw(ctx)
        _set(al, ps)
        ks = [k for k in range(n) if tuple(g['params_%d' % k]) == ps]
        got = al.align_overlap_batch([(g['a_%d' % k], g['b_%d' % k]) for k in ks])
        for k, (score, a_idx, b_idx) in zip(ks, got):
            assert np.float32(score) == g['score_%d' % k], k
            assert np.array_equal(a_idx, g['a_idx_%d' % k]), k
            assert np.array_equal(b_idx, g['b_idx_%d' % k]), k


def test_seeded_cases_against_oracle_with_trims(ctx):
    c = rp.CAligner()
    for ps in ac.PARAM_SETS:
        al = sa.align_raw(ctx)
        _set(al, ps)
        _set(c, ps)
        cs = [(a, b) for p, a, b in ac.cases(seed=77, n=300, max_L=200, max_N=1500) if p == ps]
        got = al.align_overlap_batch(cs)
        for (a, b), (score, a_idx, b_idx) in zip(cs, got):
            s0, a0, b0 = c.align_overlap(a, b)
            assert np.float32(score) == np.float32(s0)
            assert np.array_equal(a_idx, a0) and np.array_equal(b_idx, b0)


def test_detect_range_indices_on_device(ctx, model_file):
    """begin/end indices computed by the traceback kernel == numpy argmin over the oracle's view positions."""
    pm = rp.PoreModel(model_file)
    c = rp.CAligner()
    ps = ac.PARAM_SETS[0]
    _set(c, ps)
    rng = np.random.default_rng(3)
    flanks = [pm.generate_signal(C9_PREFIX, samples=6), pm.generate_signal(C9_SUFFIX, samples=6)]
    sigs = []
    for n in (20, 60):
        seq = ''.join(rng.choice(list('ACGT'), 300)) + C9_PREFIX + 'GGCCCC' * n + C9_SUFFIX + ''.join(rng.choice(list('ACGT'), 300))
        raw = pm.generate_signal(seq, samples=8, noise=True, rng=rng)
        morph, _ = rp.condition(pm, raw)
        sigs.append(morph)
    codes, vals, off = [], [], [0]
    for s in sigs:
        cc, vv = sa.encode_signal(s)
        assert len(vv) <= 256
        v = np.zeros(256, np.float32); v[:len(vv)] = vv
        codes.append(cc.astype(np.uint8)); vals.append(v); off.append(off[-1] + len(cc))
    levels = [sa.run_length_levels(f) for f in flanks]
    assert all(s == 6 for _, s in levels)
    flank_off = np.cumsum([0] + [len(l) for l, _ in levels])
    tasks = [(si, fi) for si in range(len(sigs)) for fi in range(2)]
    pre = [600 if fi == 0 else 0 for _, fi in tasks]
    post = [0 if fi == 0 else 600 for _, fi in tasks]
    res = ctx.align_batch(ps, np.concatenate(codes), off, np.stack(vals), np.concatenate([l for l, _ in levels]),
                          flank_off, 6, [t[0] for t in tasks], [t[1] for t in tasks], pre, post)
    for k, (si, fi) in enumerate(tasks):
        s0, a0, b0 = c.align_overlap(sigs[si], flanks[fi])
        want = ac.detect_range_indices(a0, b0, pre[k], post[k])
        assert np.float32(res['score'][k]) == np.float32(s0)
        assert (res['begin0'][k], res['end0'][k], res['begin_trim'][k], res['end_trim'][k]) == want
        assert res['n_blocks'][k] >= 2      # 870 flank samples span more than one 512-column block


def test_long_signal_score_only_property(ctx):
    """At a size where the oracle's full trace matrix is heavy: score and end column against the
    oracle's score-only scan, plus the planted flank position."""
    rng = np.random.default_rng(9)
    lev = np.round(rng.uniform(60, 120, 145), 2)
    b = np.repeat(lev, 6)
    N = 120000
    a = np.round(rng.uniform(60, 120, N))
    start = 77777
    planted = np.repeat(lev, rng.integers(6, 10, 145))
    a[start:start + len(planted)] = np.round(planted)
    c = rp.CAligner()
    ps = ac.PARAM_SETS[0]
    _set(c, ps)
    al = sa.align_raw(ctx)
    _set(al, ps)
    cc, vv = sa.encode_signal(a)
    v = np.zeros(256, np.float32); v[:len(vv)] = vv
    res = ctx.align_batch(ps, cc.astype(np.uint8), [0, N], v[None], lev.astype(np.float32), [0, 145], 6, [0], [0], [0], [0])
    s0, bj = c.score_only(a, b)
    assert np.float32(res['score'][0]) == np.float32(s0)
    assert res['best_j'][0] == bj
    assert abs(res['begin0'][0] - start) <= 8 and abs(res['end0'][0] - (start + len(planted))) <= 8


def test_linear_gap_scan_equals_affine_scan_and_oracle(ctx, monkeypatch):
    """gap_open == gap_extension selects the 3-FADD linear-gap scan (csrc/align.cu LinSweep); it must be
    bit-identical to the general affine scan (forced with STRIQUE_NO_LINEAR_SCAN) and to the oracle,
    including a second linear parameter set, column-1 effects (flank matching at the signal start)
    and reads long enough to cross many checkpoint columns."""
    c = rp.CAligner()
    rng = np.random.default_rng(21)
    for ps in [ac.PARAM_SETS[0], (-3.0, -7.0, -3.0, -7.0, 10.0, -2.0), (-0.25, -0.5, -0.25, -0.5, 4.0, 0.0)]:
        _set(c, ps)
        al = sa.align_raw(ctx)
        _set(al, ps)
        cs = [(a, b) for _, a, b in ac.cases(seed=123, n=120, max_L=150, max_N=1300)]
        # flank planted at the very beginning of the signal: exercises H of DP column 0 (SeqAn's denormal infinity)
        for _ in range(10):
            lev = np.round(rng.uniform(60, 120, 20))
            b = np.repeat(lev, 6)
            a = np.concatenate([np.repeat(lev, rng.integers(5, 9, 20)), np.round(rng.uniform(60, 120, 300))])
            cs.append((a[int(rng.integers(0, 4)):], b))
        monkeypatch.delenv('STRIQUE_NO_LINEAR_SCAN', raising=False)
        fast = al.align_overlap_batch(cs)
        monkeypatch.setenv('STRIQUE_NO_LINEAR_SCAN', '1')
        slow = al.align_overlap_batch(cs)
        monkeypatch.delenv('STRIQUE_NO_LINEAR_SCAN', raising=False)
        for (a, b), f, s in zip(cs, fast, slow):
            s0, a0, b0 = c.align_overlap(a, b)
            assert np.float32(f[0]) == np.float32(s[0]) == np.float32(s0)
            assert np.array_equal(f[1], s[1]) and np.array_equal(f[1], a0)
            assert np.array_equal(f[2], s[2]) and np.array_equal(f[2], b0)
    # full-size property: 64 reads of ~60 k samples x 870 flank samples, every result field equal
    ps = ac.PARAM_SETS[0]
    n_sig, N, nlev = 32, 60000, 145
    codes = np.repeat(rng.integers(60, 200, size=(n_sig * N) // 6 + 8).astype(np.uint8),
                      rng.integers(6, 10, size=(n_sig * N) // 6 + 8))[:n_sig * N]
    off = np.arange(n_sig + 1, dtype=np.int64) * N
    vals = np.tile(np.linspace(50, 132, 256, dtype=np.float32), (n_sig, 1))
    levels = rng.uniform(60, 120, 2 * nlev).astype(np.float32)
    ts, tf = np.repeat(np.arange(n_sig), 2), np.tile([0, 1], n_sig)
    args = (ps, codes, off, vals, levels, [0, nlev, 2 * nlev], 6, ts, tf, np.full(len(ts), 600), np.zeros_like(ts))
    fast = ctx.align_batch(*args)
    monkeypatch.setenv('STRIQUE_NO_LINEAR_SCAN', '1')
    slow = ctx.align_batch(*args)
    monkeypatch.delenv('STRIQUE_NO_LINEAR_SCAN', raising=False)
    assert fast.tobytes() == slow.tobytes()


def test_long_flanks_and_other_samples_per_level(ctx):
    """Flank shapes beyond the default 145 levels x 6 samples: 8 samples per level (the one-row-per-register
    kernels, up to 2048 flank samples) and 300 levels x 6.  Bit-exact against the oracle."""
    c = rp.CAligner()
    ps = ac.PARAM_SETS[0]
    _set(c, ps)
    al = sa.align_raw(ctx)
    _set(al, ps)
    rng = np.random.default_rng(31)
    cs = []
    for nlev, s in ((200, 8), (255, 8), (300, 6), (319, 6), (335, 6), (150, 7)):
        lev = np.round(rng.uniform(60, 120, nlev))
        b = np.repeat(lev, s)
        a = np.concatenate([np.round(rng.uniform(60, 120, 400)), np.repeat(lev, rng.integers(s - 1, s + 3, nlev)),
                            np.round(rng.uniform(60, 120, 400))])
        cs.append((a, b))
    got = al.align_overlap_batch(cs)
    for (a, b), (score, a_idx, b_idx) in zip(cs, got):
        s0, a0, b0 = c.align_overlap(a, b)
        assert np.float32(score) == np.float32(s0)
        assert np.array_equal(a_idx, a0) and np.array_equal(b_idx, b0)


def test_flank_beyond_the_kernels_is_refused_when_the_target_is_defined(ctx, model_file):
    from strique_b200 import counter
    rd = counter.repeatCounter(model_file, context=ctx)
    with pytest.raises(ValueError, match='does not fit the alignment kernels'):
        rd.add_target('long', 'GGCCCC', 'ACGT' * 100, C9_SUFFIX)


def test_two_flanks_per_warp_scan_equals_single_task_scan_and_oracle(ctx, monkeypatch):
    """csrc/align.cu LinSweepPair (both flanks of a signal scanned by one warp, checkpoints interleaved) against the
    single-task kernels (STRIQUE_NO_PAIR_SCAN) and the oracle: signals shorter than the lane count, around the
    checkpoint spacing, flanks of different length inside one kernel shape (different last lane / last level), a flank
    that fills its last lane exactly, and an odd task left over without a partner."""
    c = rp.CAligner()
    ps = ac.PARAM_SETS[0]
    _set(c, ps)
    rng = np.random.default_rng(41)
    levels = [np.round(rng.uniform(60, 120, n), 2).astype(np.float32) for n in (145, 129, 160, 97)]   # K = 5, 5, 5(6?), 4
    flank_off = np.cumsum([0] + [len(l) for l in levels])
    sig_lens = [1, 7, 29, 31, 200, 511, 512, 513, 1023, 1025, 2500, 6000]
    sigs, codes, vals, off = [], [], [], [0]
    for n in sig_lens:
        lev = levels[0]
        a = np.round(rng.uniform(60, 120, n))
        if n > 1200:                                         # plant both flanks so that the paths are long
            p = np.round(np.repeat(levels[0], rng.integers(6, 9, len(levels[0]))))[:n // 3]
            a[10:10 + len(p)] = p
        cc, vv = sa.encode_signal(a)
        v = np.zeros(256, np.float32); v[:len(vv)] = vv
        sigs.append(a); codes.append(cc.astype(np.uint8)); vals.append(v); off.append(off[-1] + n)
    tasks = []
    for si in range(len(sigs)):
        tasks += [(si, 0), (si, 1)]                          # a pair (same kernel shape, different L)
        if si % 3 == 0:
            tasks += [(si, 2)]                               # left over: no partner of its shape
        if si % 4 == 1:
            tasks += [(si, 3), (si, 3)]                      # another shape, the same flank twice
    args = (ps, np.concatenate(codes), off, np.stack(vals), np.concatenate(levels), flank_off, 6,
            [t[0] for t in tasks], [t[1] for t in tasks], [0] * len(tasks), [0] * len(tasks))
    monkeypatch.delenv('STRIQUE_NO_PAIR_SCAN', raising=False)
    paired = ctx.align_batch(*args)
    monkeypatch.setenv('STRIQUE_NO_PAIR_SCAN', '1')
    single = ctx.align_batch(*args)
    monkeypatch.delenv('STRIQUE_NO_PAIR_SCAN', raising=False)
    for name in ('score', 'best_j', 'begin0', 'end0', 'begin_trim', 'end_trim'):
        assert np.array_equal(paired[name], single[name]), name
    for k, (si, fi) in enumerate(tasks):
        s0, a0, b0 = c.align_overlap(sigs[si], np.repeat(levels[fi], 6))
        assert np.float32(paired['score'][k]) == np.float32(s0), (sig_lens[si], fi)
        want = ac.detect_range_indices(a0, b0, 0, 0)
        assert (paired['begin0'][k], paired['end0'][k], paired['begin_trim'][k], paired['end_trim'][k]) == want, (sig_lens[si], fi)
