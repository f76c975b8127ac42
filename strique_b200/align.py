"""Host-side mirror of the reference's native aligner module `STRique_lib.pyseqan`
(src/pyalign.cpp:47-62): class `align_raw` with the same eight read/write properties and
`align_overlap(a, b) -> (score, a_idx, b_idx)`, executed by the CUDA kernels behind
`strique_align_batch` (include/strique_b200.h).  Batched entry: `align_overlap_batch`.
"""
import numpy as np

from . import _lib


def run_length_levels(b, max_samples=64):
    """Flank vector -> (levels, samples): the largest s such that b is `levels` repeated s times
    (pore_model.generate_signal(seq, samples=s), scripts/STRique.py:185-186)."""
    b = np.asarray(b, dtype=np.float32)
    n = len(b)
    if n == 0:
        return b, 1
    change = np.flatnonzero(b[1:] != b[:-1]) + 1
    g = n
    for c in change:
        g = np.gcd(g, int(c))
        if g == 1:
            break
    g = int(g)
    if g > max_samples:   # constant vectors etc.: keep the kernel shapes sane
        for s in range(max_samples, 0, -1):
            if g % s == 0:
                g = s
                break
    return b[::g].copy(), g


def encode_signal(a):
    """Signal vector -> (codes, value table) with value_table[codes] == float32(a)."""
    a32 = np.asarray(a, dtype=np.float64).astype(np.float32)
    values, codes = np.unique(a32, return_inverse=True)
    if len(values) > 65536:
        raise _lib.StriqueError('signal has more than 65536 distinct fp32 values; the reference hot path feeds '
                                'the aligner a <=256-valued signal (scripts/STRique.py:592-596)')
    dt = np.uint8 if len(values) <= 256 else np.uint16
    return codes.astype(dt), values


def view_positions(rows, N):
    """Rebuild align_overlap's (a_idx, b_idx) (src/align_raw.h:139-147) from the per-flank-sample
    records written by the traceback kernel: rows[q] = (j << 1) | is_vertical_gap."""
    rows = np.asarray(rows, dtype=np.int64)
    j, is_v = rows >> 1, rows & 1
    v_before = np.cumsum(is_v) - is_v
    b_idx = (j - (1 - is_v)) + v_before
    vj = np.sort(j[is_v == 1])
    a_idx = np.arange(N, dtype=np.int64) + np.searchsorted(vj, np.arange(N), side='right')
    return a_idx.astype(np.uint64), b_idx.astype(np.uint64)


class align_raw(object):
    """Drop-in for `pyseqan.align_raw` (defaults from src/align_raw.h:52-60)."""

    def __init__(self, context=None):
        self.gap_open_h = -2.0
        self.gap_open_v = -2.0
        self.gap_extension_h = -8.0
        self.gap_extension_v = -8.0
        self.dist_offset = 8.0
        self.dist_min = -16.0
        self._ctx = context

    # `gap_open` / `gap_extension` set both directions and read the horizontal one (align_raw.h:77-80)
    @property
    def gap_open(self):
        return self.gap_open_h

    @gap_open.setter
    def gap_open(self, v):
        self.gap_open_h = v
        self.gap_open_v = v

    @property
    def gap_extension(self):
        return self.gap_extension_h

    @gap_extension.setter
    def gap_extension(self, v):
        self.gap_extension_h = v
        self.gap_extension_v = v

    def params(self):
        return _lib.AlignParams(self.gap_open_h, self.gap_open_v, self.gap_extension_h, self.gap_extension_v,
                                self.dist_offset, self.dist_min)

    @property
    def context(self):
        if self._ctx is None:
            self._ctx = _lib.default_context()
        return self._ctx

    def align_overlap_batch(self, pairs, pre_trim=0, post_trim=0):
        """pairs: iterable of (signal vector, flank vector). -> list of (score, a_idx, b_idx)."""
        pairs = list(pairs)
        out = [None] * len(pairs)
        todo = []
        for k, (a, b) in enumerate(pairs):
            if len(a) == 0 or len(b) == 0:
                # SeqAn returns MinValue<float> (FLT_MIN) without running the DP; all gaps
                out[k] = (float(np.finfo(np.float32).tiny), np.arange(len(a), dtype=np.uint64),
                          np.arange(len(b), dtype=np.uint64))
            else:
                todo.append(k)
        by_samples = {}
        for k in todo:
            levels, s = run_length_levels(pairs[k][1])
            by_samples.setdefault(s, []).append((k, levels))
        for s, items in by_samples.items():
            codes, vals, sig_off = [], [], [0]
            levels_all, flank_off = [], [0]
            widest = 0
            enc = []
            for k, levels in items:
                c, v = encode_signal(pairs[k][0])
                enc.append((c, v))
                widest = max(widest, len(v))
            ncv = 256 if widest <= 256 else 65536
            dt = np.uint8 if ncv == 256 else np.uint16
            for (k, levels), (c, v) in zip(items, enc):
                codes.append(c.astype(dt))
                vv = np.zeros(ncv, dtype=np.float32)
                vv[:len(v)] = v
                vals.append(vv)
                sig_off.append(sig_off[-1] + len(c))
                levels_all.append(levels)
                flank_off.append(flank_off[-1] + len(levels))
            n = len(items)
            res, rows = self.context.align_batch(
                self.params(), np.concatenate(codes), sig_off, np.stack(vals), np.concatenate(levels_all), flank_off, s,
                np.arange(n), np.arange(n), np.full(n, pre_trim), np.full(n, post_trim), want_rows=True)
            for idx, (k, levels) in enumerate(items):
                L = len(levels) * s
                a_idx, b_idx = view_positions(rows[idx, :L], len(pairs[k][0]))
                out[k] = (float(res['score'][idx]), a_idx, b_idx)
        return out

    def align_overlap(self, a, b):
        """Semi-global signal alignment; same return convention as the pybind11 module."""
        return self.align_overlap_batch([(a, b)])[0]
