"""Seeded synthetic reads following the reference's own test recipe (scripts/STRique_test.py:50-60,
scripts/STRique.py:182-195): backbone + prefix + repeat*n + suffix + backbone through the pore model."""
import numpy as np

_COMP = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A'}


def revcomp(s):
    return ''.join(_COMP[b] for b in reversed(s))


def backbone(rng, n):
    return ''.join(rng.choice(list('ACGT'), n))


def read_sequence(rng, prefix, repeat, suffix, n, flank=1000, strand='+'):
    seq = backbone(rng, flank) + prefix + repeat * n + suffix + backbone(rng, flank)
    return seq if strand == '+' else revcomp(seq)


def simulate(pm, seq, rng, noise=True, samples=8, int16=False):
    """pm: oracle PoreModel or strique_b200 pore_model (same table). Noise: dwell U{6..9}, N(mean, stdv)."""
    k = pm.kmer
    table = getattr(pm, 'table', None) or pm.model_dict
    kmers = [seq[i:i + k] for i in range(len(seq) - k + 1)]
    means = np.array([table[x][0] for x in kmers])
    if not noise:
        sig = np.repeat(means, samples)
    else:
        stdvs = np.array([table[x][1] for x in kmers])
        dwell = rng.uniform(6, 10, len(means)).astype(int)
        sig = rng.normal(np.repeat(means, dwell), np.repeat(stdvs, dwell))
    if int16:
        # DAC-like values: pA -> integer counts (the pipeline is scale free)
        sig = np.round(sig * 8.0 + 100.0).astype(np.int16)
    return sig
