#!/usr/bin/env python
"""Parity at the bench's full sizes through implementation-independent properties: the rows of one full batch must
not depend on which of the equivalent kernels computed them -- two-flanks-per-warp scan vs single-task scans vs the
affine scan, fixed-point vs float64 Viterbi (counts / offsets / ticks identical, log p within 1e-3 nat).
    python tools/gpu_scale_parity.py [c2|c4] [reads]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from strique_b200 import _lib, workload
    from strique_b200.counter import repeatCounter
    wl = sys.argv[1] if len(sys.argv) > 1 else 'c2'
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
    model = os.path.join(ROOT, 'models', 'r9_4_450bps.model')
    loci = ('c9orf72',) if wl == 'c2' else workload.PANEL
    reads = workload.make_reads_parallel(model, None, range(n), seed=4242, loci=loci)
    ctx = _lib.Context(0)
    dt = repeatCounter(model, context=ctx)
    for name in loci:
        dt.add_target(name, *workload.LOCI[name])
    items = [(name, sig, strand) for name, sig, strand, _ in reads]

    def run(env, exact=False):
        for k in ('STRIQUE_NO_PAIR_SCAN', 'STRIQUE_NO_PACKED_SCAN', 'STRIQUE_NO_LINEAR_SCAN'):
            os.environ.pop(k, None)
        os.environ.update(env)
        ctx.set_viterbi_exact(exact)
        return dt.detect_batch(items)

    base = run({})
    ok = True
    for label, env, exact in (('single-task packed scan', {'STRIQUE_NO_PAIR_SCAN': '1'}, False),
                              ('single-task unpacked linear scan', {'STRIQUE_NO_PAIR_SCAN': '1', 'STRIQUE_NO_PACKED_SCAN': '1'}, False),
                              ('affine scan', {'STRIQUE_NO_PAIR_SCAN': '1', 'STRIQUE_NO_LINEAR_SCAN': '1'}, False),
                              ('float64 Viterbi', {}, True)):
        other = run(env, exact)
        ints = sum(1 for a, b in zip(base, other) if (a[0], a[4], a[5], a[6]) != (b[0], b[4], b[5], b[6]))
        scores = sum(1 for a, b in zip(base, other) if (a[1], a[2]) != (b[1], b[2]))
        gap = max(abs(a[3] - b[3]) for a, b in zip(base, other))
        print('%-34s reads %d: integer rows differing %d, alignment scores differing %d, max |log p gap| %.3g' % (label, n, ints, scores, gap))
        ok = ok and ints == 0 and scores == 0 and gap <= 1e-3
    print('OK' if ok else 'MISMATCH')
    sys.exit(0 if ok else 1)


if __name__ == '__main__':
    main()
