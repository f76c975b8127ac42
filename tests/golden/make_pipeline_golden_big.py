"""Generates tests/golden/pipeline_golden_big.json: the oracle pipeline (oracle/reference_path.py with the
COMPILED REFERENCE aligner oracle/_ref/pyseqan) on seeded synthetic reads at the shapes BASELINE.json names:

    c2   512 reads  c9orf72 GGCCCC, n ~ U{2..1000}, both strands                 (configs[1])
    c3    64 reads  as c2 (n <= 1000) with the mCpG methylation HMM, half simulated from the mCpG table (configs[2])
    c4   128 reads  four-locus panel, both strands, flank templates 570..1770    (configs[3])
    c5     8 reads  long expansion, n = 4000, ~242 k samples                     (configs[4])

One process per core (about 10 CPU-minutes on 8 cores).  Rows are stored column-wise to keep the file small;
`margin` = gap between the best and the second-best Viterbi path of the count HMM (oracle/viterbi_oracle.c), the
quantity PARITY_EXCEPTIONS.md is stated in.  Reads are regenerated from their seeds (strique_b200/workload.py);
a CRC of every signal guards against drift.

    python -m tests.golden.make_pipeline_golden_big [set ...]
"""
import json
import multiprocessing as mp
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from strique_b200 import workload                # noqa: E402
from strique_b200.pore_model import pore_model   # noqa: E402

MODEL = os.path.join(ROOT, 'models', 'r9_4_450bps.model')
MOD_MODEL = os.path.join(ROOT, 'models', 'r9_4_450bps_mCpG.model')
OUT = os.path.join(ROOT, 'tests', 'golden', 'pipeline_golden_big.json')

SETS = {
    'c2': (dict(n_reads=512, seed=5151, loci=('c9orf72',), n_lo=2, n_hi=1000), False),
    'c3': (dict(n_reads=64, seed=5252, loci=('c9orf72',), n_lo=2, n_hi=1000, mod_fraction=0.5), True),
    'c4': (dict(n_reads=128, seed=5353, loci=workload.PANEL, n_lo=2, n_hi=1000), False),
    'c5': (dict(n_reads=8, seed=5454, loci=('c9orf72',), fixed_n=4000, flank=4000), False),
}
COLUMNS = ('target', 'strand', 'n_true', 'len', 'crc', 'count', 'score_prefix', 'score_suffix', 'log_p', 'offset', 'ticks',
           'mod', 'prefix_begin', 'suffix_end', 'margin')

_ref = {}


def _counter(use_mod, loci):
    from oracle import reference_path as rp
    key = (use_mod, tuple(loci))
    if key not in _ref:
        ref = rp.RefRepeatCounter(MODEL, mod_model_file=MOD_MODEL if use_mod else None, aligner='ref')
        for name in loci:
            ref.add_target(name, *workload.LOCI[name])
        _ref[key] = ref
    return _ref[key]


def _one(job):
    tag, k, name, sig, strand, n_true, use_mod, loci = job
    det = {}
    r = _counter(use_mod, loci).detect(name, sig, strand, details=det)
    return tag, k, [name, strand, n_true, int(len(sig)), zlib.crc32(sig.tobytes()), int(r[0]), float(r[1]), float(r[2]),
                    float(r[3]), int(r[4]), int(r[5]), r[6], det['prefix_begin'], det['suffix_end'], det['margin']]


def main():
    from oracle import reference_path as rp
    assert rp.load_pyseqan() is not None, 'build oracle/_ref first (make -C oracle ref)'
    want = sys.argv[1:] or list(SETS)
    out = json.load(open(OUT)) if os.path.exists(OUT) else {
        'note': 'oracle pipeline with the compiled reference aligner; see make_pipeline_golden_big.py',
        'columns': list(COLUMNS), 'sets': {}}
    pm, pm_mod = pore_model(MODEL), pore_model(MOD_MODEL)
    jobs = []
    for tag in want:
        kwargs, use_mod = SETS[tag]
        reads = workload.make_reads(pm, pm_mod=pm_mod if use_mod else None, **kwargs)
        out['sets'][tag] = {'kwargs': kwargs, 'use_mod': use_mod, 'rows': [None] * len(reads)}
        jobs += [(tag, k, name, sig, strand, n, use_mod, kwargs['loci']) for k, (name, sig, strand, n) in enumerate(reads)]
    jobs.sort(key=lambda j: -len(j[3]))                      # longest first
    done = 0
    with mp.Pool(int(os.environ.get('GOLDEN_PROCS', os.cpu_count()))) as pool:
        for tag, k, row in pool.imap_unordered(_one, jobs):
            out['sets'][tag]['rows'][k] = row
            done += 1
            if done % 16 == 0:
                print(done, '/', len(jobs), flush=True)
    with open(OUT, 'w') as fp:
        json.dump(out, fp, separators=(',', ':'))
    for tag in want:
        rows = out['sets'][tag]['rows']
        ok = sum(1 for r in rows if r[5] == r[2])
        print(tag, len(rows), 'reads, count == simulated truth in', ok)


if __name__ == '__main__':
    main()
