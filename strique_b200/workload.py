"""Synthetic read generator for the benchmark configurations.

Recipe = the reference's own test recipe (scripts/STRique_test.py:50-60) with the noisy branch of
pore_model.generate_signal (scripts/STRique.py:189-194): sequence = backbone[:1000] + prefix +
repeat*n + suffix + backbone[-1000:], every k-mer held for a dwell of U{6..9} samples drawn from
N(mean_kmer, stdv_kmer), minus-strand reads are the reverse complement.  Vectorised (integer k-mer
codes, table lookups) so that thousands of reads are generated in seconds; samples are scaled to
int16 DAC-like values (the whole path is scale free: it z-scores / percentile-normalises the read).
"""
import numpy as np

_BASES = np.frombuffer(b'ACGT', dtype=np.uint8)
_LUT = np.full(256, 255, dtype=np.uint8)
for _i, _b in enumerate(b'ACGT'):
    _LUT[_b] = _i

LOCI = {
    # name: (repeat, prefix, suffix) as in configs/repeat_config.tsv of the reference
    'c9orf72': ('GGCCCC',
                'CGGCAGCCGAACCCCAAACAGCCACCCGCCAGGATGCCGCCTCCTCACTCACCCACTCGCCACCGCCTGCGCCTCCGCCGCCGCGGGCGCAGGCACCGC'
                'AACCGCAGCCCCGCCCCGGGCCCGCCCCCGGGCCCGCCCCGACCACGCCCC',
                'TAGCGCGCGACTCCTGAGTTCCAGAGCTTGCTACAGGCTGCGGTTGTTTCCCTCCTTGTTTTCTTCTGGTTAATCTTTATCAGGTCTTTTCTTGTTCAC'
                'CCTCAGCGAGTACTGTGAGAGCAAGTAGTGGGGAGAGAGGGTGGGAAAAAC'),
    'fmr1': ('CGG',
             'GCGGGCCGGGGGTTCGGCCTCAGTCAGGCGCTCAGCTCCGTTTCGGTTTCACTTCCGGTGGAGGGCCGCCTCTGAGCGGGCGGCGGGCCGACGGCGAGCG'
             'CGGGCGGCGGCGGTGACGGAGGCGCCGCTGCCAGGGGGCGTGCGGCAGCG',
             'AGGCGGCGGCGGCGGCGGCGGCGGCGGCGGCTGGGCCTCGAGCGCCCGCAGCCCACCTCTCGGGGGCGGGCTCCCGGCGCTAGCAGGGCTGAAGAGAAGA'
             'TGGAGGAGCTGGTGGTGGAAGTGCGGGGCTCCAATGGCGCTTTCTACAAG'),
    # Panel loci of configuration C4 (BASELINE.json configs[3]).  The reference ships no entry for them: the
    # repeat units are the real ones (ATXN10 ATTCT, DMPK CTG), the flanks are SYNTHETIC stand-ins of mixed
    # length (100 / 300 nt and 300 / 200 nt -> 570 ... 1770 template samples per flank alignment), drawn once
    # with numpy default_rng(20261017); configs/panel_config.tsv carries the same strings.
    'atxn10': ('ATTCT',
               'TGTTGGCCAGAGTTAGTATCGATACGAAAACTGCGCGCAAATTTAAAGGATTTCGAACGTGCTCTGGCACGTCAAGCGTATGATTTTCTCTGACGTACTG',
               'TTTCTGAATCGTTTAGTCTAAAAGTACTCCATTCTTGGCCTGTTTTCGTATTGAGAGGTCGTACGATGCAACGACCCTTAAATCGTGTTAAGATCGAAGG'
               'TCGTTCTGTAGAATTTTATATTCTCTCCACCTAAATTTTATACCGTCTTGGAGTTTTTCGGGGGCACCGCGCCACCAGTCCGTTAGATTATCTGAGATAC'
               'TCGTCGCAAGCCTAAGAAGCTAATCGGCGCTATTCGAAAGAAGTTCTTGCTTTTTATAAGCAACAGCGCCCATAAGCTTAATACATGTCCGACACAGCAC'),
    'dmpk': ('CTG',
             'CAGACTCGCACGGATGGGGGCCGTGCCCAACGCTCAGGTTTGGGAGAAGAGACCGTGGAAGGCGGGGTGGTATCCGTTTCCTGGGGTCCTGTCACAGTTC'
             'GTCACTCAAGATCTGCGTCACGGCGAAGCAGCTCGGTTCCGAAAATGCAACTCCGGTGAGGAGCAAGCATCTCCGAATTGCGAAGGTTACCTTCCCCTCG'
             'GGTACCAGCCCCGATCCATGTGGGGGCCATCTATTCACGCCGGCTTGGTCCCGTTATACTCGTCAGCAGACGGTAAGGGGACACCGGGAGGGGTGAACCC',
             'GGCGAACGGCGGTTCTGCGTCGTAAAAGGGCGCGTGCCCGACCGGGCTTGGCGTCGCAGGCAGGGTTCCTCCTCTAGCGCTGGAGAAGCCGGAGGCGGTA'
             'TTTGCTATGACAACGGGGTCAGGGCATATCCCGCAGGTGAGTATTCATCCAAAGAGCGTCGCCGCCCTTTCTCAATCCTCATACTCGGGGACTGTCTGCG'),
}
PANEL = ('c9orf72', 'fmr1', 'atxn10', 'dmpk')


def encode(seq):
    """ACGT string -> uint8 base codes 0..3."""
    a = _LUT[np.frombuffer(seq.encode('ascii'), dtype=np.uint8)]
    if (a == 255).any():
        raise ValueError('sequence contains characters other than ACGT')
    return a


def revcomp_codes(a):
    return (3 - a)[::-1]


class KmerTable(object):
    """means / stdvs of a pore_model indexed by the integer code of the k-mer (A=0 .. T=3, first base
    most significant)."""

    def __init__(self, pm):
        k = pm.kmer
        self.k = k
        self.means = np.zeros(4 ** k)
        self.stdvs = np.zeros(4 ** k)
        for kmer, (m, s) in pm.model_dict.items():
            idx = 0
            for ch in kmer:
                idx = idx * 4 + 'ACGT'.index(ch)
            self.means[idx] = m
            self.stdvs[idx] = s

    def kmer_codes(self, bases):
        k = self.k
        n = len(bases) - k + 1
        idx = np.zeros(n, dtype=np.int64)
        for j in range(k):
            idx = idx * 4 + bases[j:j + n]
        return idx


def simulate_read(tab, bases, rng, scale=8.0, shift=100.0):
    """Noisy int16 read of a base-code sequence (S.py:189-194)."""
    idx = tab.kmer_codes(bases)
    dwell = rng.uniform(6, 10, len(idx)).astype(np.int64)
    sig = rng.normal(np.repeat(tab.means[idx], dwell), np.repeat(tab.stdvs[idx], dwell))
    return np.round(sig * scale + shift).astype(np.int16)


def make_reads(pm, n_reads, seed=0, loci=('c9orf72',), n_lo=2, n_hi=1000, flank=1000, pm_mod=None, mod_fraction=0.0,
               fixed_n=None, indices=None):
    """-> list of (target_name, int16 signal, strand, true repeat count).  Reads are independent:
    read r uses numpy default_rng([seed, r]) so any subset can be regenerated on its own (`indices`: the reads to
    make, default range(n_reads))."""
    tab = KmerTable(pm)
    tab_mod = KmerTable(pm_mod) if pm_mod is not None else None
    enc = {name: tuple(encode(s) for s in LOCI[name]) for name in loci}
    out = []
    for r in (range(n_reads) if indices is None else indices):
        rng = np.random.default_rng([seed, r])
        name = loci[int(rng.integers(len(loci)))]
        rep, pre, suf = enc[name]
        n = int(fixed_n) if fixed_n is not None else int(rng.integers(n_lo, n_hi + 1))
        bb = rng.integers(0, 4, 2 * flank).astype(np.uint8)
        bases = np.concatenate([bb[:flank], pre, np.tile(rep, n), suf, bb[flank:]])
        strand = '+' if rng.random() < 0.5 else '-'
        if strand == '-':
            bases = revcomp_codes(bases)
        t = tab_mod if (tab_mod is not None and rng.random() < mod_fraction) else tab
        out.append((name, simulate_read(t, bases, rng), strand, n))
    return out


# ---- the same reads, generated by a pool of processes (the bench's workload: seconds instead of tens of seconds) ----
_pool_args = None


def _pool_init(model_file, mod_model_file, kwargs):
    global _pool_args
    from .pore_model import pore_model
    _pool_args = (pore_model(model_file), pore_model(mod_model_file) if mod_model_file else None, kwargs)


def _pool_chunk(indices):
    pm, pm_mod, kwargs = _pool_args
    return make_reads(pm, 0, pm_mod=pm_mod, indices=indices, **kwargs)


def make_reads_parallel(model_file, mod_model_file, indices, procs=None, chunk=64, **kwargs):
    """make_reads(indices=...) over `procs` worker processes; the result is identical, read for read."""
    import multiprocessing as mp
    import os
    indices = list(indices)
    procs = max(1, min(procs or (os.cpu_count() or 1), (len(indices) + chunk - 1) // chunk))
    if procs == 1:
        _pool_init(model_file, mod_model_file, kwargs)
        return _pool_chunk(indices)
    chunks = [indices[i:i + chunk] for i in range(0, len(indices), chunk)]
    with mp.get_context('fork').Pool(procs, initializer=_pool_init, initargs=(model_file, mod_model_file, kwargs)) as pool:
        parts = pool.map(_pool_chunk, chunks, chunksize=1)
    return [r for part in parts for r in part]
