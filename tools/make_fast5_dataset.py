#!/usr/bin/env python
"""Synthetic fast5 data set for the CLI end-to-end measurement (bench.py `cli_e2e`) and tests: N reads of a bench
workload written as multi-read fast5 files (deflate-chunked int16 signals, like ONT's), the index file
`STRique.py index` would produce, and a SAM whose records overlap the loci of configs/panel_config.tsv.

    python tools/make_fast5_dataset.py <out_dir> --reads 10240 [--workload c2] [--per-file 256]
"""
import argparse
import multiprocessing as mp
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _write(job):
    from tests import hdf5_writer as hw
    path, reads = job
    hw.multi_read_fast5(path, reads, chunk=8192, deflate=True)
    return path


def build(out_dir, reads, per_file=256, procs=None):
    """reads: list of (target, int16 signal, strand, n_true) -> (index file, sam file, read ids)"""
    os.makedirs(out_dir, exist_ok=True)
    loci = {}
    for line in open(os.path.join(ROOT, 'configs', 'panel_config.tsv')).read().split('\n')[1:]:
        c = line.split()
        if len(c) == 7:
            loci[c[3]] = (c[0], int(c[1]), int(c[2]))
    ids = ['synth-%08d' % k for k in range(len(reads))]
    jobs, index = [], []
    for f0 in range(0, len(reads), per_file):
        name = 'batch_%05d.fast5' % (f0 // per_file)
        jobs.append((os.path.join(out_dir, name), [(ids[k], reads[k][1]) for k in range(f0, min(f0 + per_file, len(reads)))]))
        index += ['%s/read_%s\t%s' % (name, ids[k], ids[k]) for k in range(f0, min(f0 + per_file, len(reads)))]
    with mp.get_context('fork').Pool(procs or os.cpu_count()) as pool:
        pool.map(_write, jobs, chunksize=1)
    index_file = os.path.join(out_dir, 'reads.fofn')
    open(index_file, 'w').write('\n'.join(index) + '\n')
    sam_file = os.path.join(out_dir, 'reads.sam')
    with open(sam_file, 'w') as fp:
        fp.write('@HD\tVN:1.6\n')
        for rid, (target, sig, strand, _) in zip(ids, reads):
            chrom, begin, end = loci[target]
            bases = max(len(sig) // 9, 100)
            fp.write('\t'.join([rid, '16' if strand == '-' else '0', chrom, str(begin - 2000), '60', '%dM' % (end - begin + 4000),
                                '*', '0', '0', 'N' * bases, '*']) + '\n')
    return index_file, sam_file, ids


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('out_dir')
    ap.add_argument('--reads', type=int, default=10240)
    ap.add_argument('--workload', default='c2', choices=['c2', 'c4'])
    ap.add_argument('--per-file', type=int, default=256)
    args = ap.parse_args()
    from strique_b200 import workload
    loci = ('c9orf72',) if args.workload == 'c2' else workload.PANEL
    reads = workload.make_reads_parallel(os.path.join(ROOT, 'models', 'r9_4_450bps.model'), None, range(args.reads), seed=1000,
                                         loci=loci)
    print(*build(args.out_dir, reads, args.per_file)[:2])


if __name__ == '__main__':
    main()
