"""Seeded alignment test cases shared by the oracle tests and the GPU parity tests
(recipe of SURVEY.md App. A: random, quantised / tie-heavy, planted flank, three parameter sets)."""
import numpy as np

PARAM_SETS = [  # (gap_open_h, gap_open_v, gap_extension_h, gap_extension_v, dist_offset, dist_min)
    (-1.0, -16.0, -1.0, -16.0, 16.0, 0.0),      # STRique's configuration (scripts/STRique.py:507-512)
    (-2.0, -2.0, -8.0, -8.0, 8.0, -16.0),       # pyseqan defaults (src/align_raw.h:52-60)
    (-2.0, -5.0, -0.5, -3.0, 6.0, -4.0),
]


def make_case(rng, trial, max_L=60, max_N=400):
    mode = trial % 5
    L = int(rng.integers(1, max_L))
    N = int(rng.integers(1, max_N))
    b = rng.uniform(60, 120, L)
    if mode == 0:
        a = rng.uniform(60, 120, N)
    elif mode == 1:
        a = np.round(rng.uniform(60, 120, N))
        b = np.round(b)
    elif mode == 2:
        a = np.round(rng.uniform(60, 120, N), 1)
        b = np.round(b, 1)
        if N > L + 5:
            s = int(rng.integers(0, N - L))
            a[s:s + L] = b
    elif mode == 3:     # flank in runs of 6 like generate_signal(samples=6), read = stretched noisy copy
        nl = (L + 5) // 6
        lev = np.round(rng.uniform(60, 120, nl))
        b = np.repeat(lev, 6)
        a = np.round(rng.uniform(60, 120, N))
        stretched = np.repeat(lev, rng.integers(4, 10, nl))
        if N > len(stretched) + 2:
            s = int(rng.integers(0, N - len(stretched)))
            a[s:s + len(stretched)] = stretched + np.round(rng.normal(0, 1.0, len(stretched)))
    else:               # constant / near-constant signals: maximal ties, degenerate zero scores
        a = np.full(N, 100.0) if trial % 2 else np.round(rng.uniform(99, 101, N))
        b = np.full(L, 100.0 if trial % 3 else 30.0)
    return a.astype(np.float64), b.astype(np.float64)


def cases(seed, n, **kw):
    rng = np.random.default_rng(seed)
    out = []
    for trial in range(n):
        a, b = make_case(rng, trial, **kw)
        out.append((PARAM_SETS[trial % 3], a, b))
    return out


def rows_from_view_positions(a_idx, b_idx):
    """(a_idx, b_idx) of align_overlap -> per-flank-sample records (j << 1) | is_vertical_gap."""
    a_idx = np.asarray(a_idx, dtype=np.int64)
    b_idx = np.asarray(b_idx, dtype=np.int64)
    pos = np.searchsorted(a_idx, b_idx, side='left')
    hit = (pos < len(a_idx)) & (a_idx[np.minimum(pos, len(a_idx) - 1)] == b_idx) if len(a_idx) else np.zeros(len(b_idx), bool)
    j = np.where(hit, pos + 1, pos)
    return (j << 1) | (~hit).astype(np.int64)


def detect_range_indices(a_idx, b_idx, pre_trim, post_trim):
    """The index reduction of repeatCounter.__detect_range__ (scripts/STRique.py:540-547)."""
    a = np.asarray(a_idx, dtype=np.int64)
    b = np.asarray(b_idx, dtype=np.int64)
    f = lambda k: int(np.abs(a - b[k]).argmin())
    return f(0), f(-1), f(0 + pre_trim), f(-1 - post_trim)
