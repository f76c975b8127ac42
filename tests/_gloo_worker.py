"""Worker of tests/test_sharding_gloo.py: one rank of a world_size-2 `gloo` job running the CLI's
host logic (cli.run_count: rank 0 streams and plans the SAM, broadcasts chunks; shard -> detect -> gather -> append)
with a stub in place of the CUDA repeatCounter."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from strique_b200 import cli, sharding  # noqa: E402


class StubCounter(object):
    """Deterministic stand-in: count = signal length mod 997, so rows depend on the read only."""

    def __init__(self):
        self.targets = {}
        self.calls = 0

    def add_target(self, name, repeat, prefix, suffix):
        self.targets[name] = (repeat, prefix, suffix)

    def detect_batch(self, items):
        self.calls += 1
        return [(len(sig) % 997, 1.5, 2.5, -float(len(sig)), 10, 20, '-') for _, sig, _ in items]


class StubIndex(object):
    def get_raw(self, ID):
        n = 1000 + (sum(ord(c) for c in ID) * 37) % 5000
        if ID.endswith('missing'):
            raise RuntimeError('[Error] Read {} not found'.format(ID))
        return list(range(n))


def sam_lines(n):
    out = []
    for i in range(n):
        name = 'read%03d' % i + ('missing' if i % 17 == 5 else '')
        flag = 16 if i % 2 else 0
        seq = 'A' * (100 + (i * 7919) % 900)
        out.append('\t'.join([name, str(flag), 'chr9', '27573000', '60', '1000M', '*', '0', '0', seq, '*']) + '\n')
    out.append('garbage line\n')
    out.append('\t'.join(['offtarget', '0', 'chr1', '100', '60', '50M', '*', '0', '0', 'ACGT', '*']) + '\n')
    return out


def main():
    out_file = sys.argv[1]
    rank, world = sharding.init_host_group('gloo')
    cfg = cli.parse_config(os.path.join(ROOT, 'configs', 'repeat_config.tsv'))
    rd = cli.repeatDetector.__new__(cli.repeatDetector)
    rd.repeatCounter = StubCounter()
    from collections import defaultdict
    rd.repeatLoci = defaultdict(list)
    rd.repeat_config = cfg['repeat']
    rd.is_init = False
    rd.f5 = StubIndex()
    rd.io_threads = 2
    rd.batch_samples = 20000                # several batches per rank and several chunks per run
    # like a pipe on stdin: ONLY rank 0 has the SAM lines (the others get None), rank 0 broadcasts the planned chunks
    lines = iter(sam_lines(60)) if rank == 0 else None
    cli.run_count(rd, lines, out_file if rank == 0 else None, rank, world)
    sharding.finalize()


if __name__ == '__main__':
    main()
