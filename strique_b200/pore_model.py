"""Host-side pore model: k-mer table, model constants and template-signal simulation.

Mirrors `pore_model` of the reference (scripts/STRique.py:113-195) -- same constructor, attribute
and method names -- because `repeatCounter` and the reference's tests use it directly
(scripts/STRique_test.py:49-60).  Everything here runs once per model / target on the host; the
per-read normalisation of the hot path is the CUDA conditioning kernel (csrc/condition.cu), which
only needs the four constants returned by `minmax_constants()`.
"""
import numpy as np


class pore_model(object):
    def __init__(self, model_file):
        table = {}
        with open(model_file, 'r') as fp:
            for line in fp:
                cols = line.strip().split('\t')[:3]
                table[cols[0]] = (float(cols[1]), float(cols[2]))
        self.model_dict = table
        self.kmer = len(next(iter(table.keys())))
        means = np.array([v[0] for v in table.values()])
        self._means = means
        self.model_median = np.median(means)
        self.model_MAD = np.mean(np.absolute(np.subtract(means, self.model_median)))
        low = min(table.values(), key=lambda v: v[0])
        high = max(table.values(), key=lambda v: v[0])
        self.model_min = low[0] - 6 * low[1]
        self.model_max = high[0] + 6 * high[1]

    def MAD(self, signal):
        """Mean absolute deviation around the median (S.py:142-143)."""
        return np.mean(np.absolute(np.subtract(signal, np.median(signal))))

    def scale2stdv(self, other):
        """Ratio of the median k-mer stdv of `other` to this model's (S.py:145-148)."""
        mine = np.median(np.array([v[1] for v in self.model_dict.values()]))
        theirs = np.median(np.array([v[1] for v in other.model_dict.values()]))
        return theirs / mine

    def minmax_constants(self):
        """(m5_mod, m95_mod, model_min, model_max): the model side of normalize2model(mode='minmax')
        (S.py:152-156): medians of the k-mer means below the 1st / above the 99th percentile."""
        q_lo, q_hi = np.percentile(self._means, [1, 99])
        m5 = np.median(self._means[self._means < q_lo])
        m95 = np.median(self._means[self._means > q_hi])
        return float(m5), float(m95), float(self.model_min), float(self.model_max)

    def normalize2model(self, signal, clip=True, mode='median'):
        """Host utility with the reference's semantics for 'minmax' and the default median/MAD mode
        (S.py:150-180).  NOT used by repeatCounter.detect: the hot path normalises on the GPU."""
        signal = np.asarray(signal, dtype=np.float64)
        if mode == 'minmax':
            m5_mod, m95_mod, _, _ = self.minmax_constants()
            q_lo, q_hi = np.percentile(signal, [1, 99])
            m5 = np.median(signal[signal < q_lo])
            m95 = np.median(signal[signal > q_hi])
            out = (signal - (m5 + (m95 - m5) / 2)) / ((m95 - m5) / 2)
            out = out * ((m95_mod - m5_mod) / 2) + (m5_mod + (m95_mod - m5_mod) / 2)
        elif mode == 'entropy':
            raise NotImplementedError("mode='entropy' is dead code on the reference's count path (SURVEY.md section 2)")
        else:
            out = np.divide(np.subtract(signal, np.median(signal)), self.MAD(signal))
            out = np.add(np.multiply(out, self.model_MAD), self.model_median)
        if clip is True:
            np.clip(out, self.model_min + .5, self.model_max - .5, out=out)
        return out

    def kmer_means(self, sequence):
        k = self.kmer
        return np.array([self.model_dict[sequence[i:i + k]][0] for i in range(len(sequence) - k + 1)])

    def generate_signal(self, sequence, samples=10, noise=False):
        """Template signal of a sequence (S.py:182-195): k-mer means repeated `samples` times, or with
        random dwell times in [6, 10) and, with noise=True, Gaussian sample noise."""
        level_means = self.kmer_means(sequence)
        if samples and not noise:
            return np.repeat(level_means, samples)
        if not noise:
            return np.repeat(level_means, np.random.uniform(6, 10, len(level_means)).astype(int))
        k = self.kmer
        level_stdvs = np.array([self.model_dict[sequence[i:i + k]][1] for i in range(len(sequence) - k + 1)])
        level_samples = np.random.uniform(6, 10, len(level_means)).astype(int)
        return np.random.normal(np.repeat(level_means, level_samples), np.repeat(level_stdvs, level_samples))
