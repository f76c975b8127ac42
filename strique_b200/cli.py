"""`STRique.py` command line of the reference (scripts/STRique.py:874-946), served by the CUDA path.

    STRique.py index <input> [--recursive] [--out_prefix P] [--tmp_prefix T]
    STRique.py count <f5Index> <model> <repeat> [--out F] [--algn SAM] [--mod_model M] [--config J]
                     [--t N] [--log_level L]

Same arguments, same `repeat_config.tsv` / pore model / JSON inputs, same ten TSV columns.  What is
different underneath: SAM records are decoded and intersected with the loci on the host (as in
repeatDetector, scripts/STRique.py:624-705), the raw signals are fetched by `--t` I/O threads, and
reads go to the GPU in batches through `repeatCounter.detect_batch` (strique_detect_batch).  Rows are
written in input order.  Launched under `torchrun --nproc-per-node N` every rank takes a
cost-balanced share of the reads on its own GPU and rank 0 gathers the rows (strique_b200/sharding.py).
"""
import argparse
import json
import os
import re
import signal
import sys
import time
from collections import defaultdict
from concurrent.futures import ThreadPoolExecutor

from . import fast5, sharding

LOG_LEVELS = {'error': 0, 'warning': 1, 'info': 2, 'debug': 3}
HEADER = ['ID', 'target', 'strand', 'count', 'score_prefix', 'score_suffix', 'log_p', 'offset', 'ticks', 'mod']


class logger(object):
    """stderr logger with the reference's line format (scripts/STRique.py:55-107), in-process."""
    level = 1

    @staticmethod
    def init(log_level='warning'):
        logger.level = LOG_LEVELS.get(log_level, 1)

    @staticmethod
    def log(message, level='info'):
        if LOG_LEVELS[level] <= logger.level:
            print('{time} [PID {pid}] [{level}] {msg}'.format(time=time.strftime('%d.%m.%Y %H:%M:%S'), pid=os.getpid(),
                                                              level=level.capitalize(), msg=message), file=sys.stderr)


def parse_config(repeat_config_file, param_config_file=None):
    """repeat_config.tsv (header skipped, any whitespace, exactly 7 columns) and the optional JSON with
    mandatory 'align' and 'HMM' dictionaries (scripts/STRique.py:836-868)."""
    repeats = {}
    with open(repeat_config_file, 'r') as fp:
        next(fp)
        for line in fp:
            cols = line.rstrip().split()
            if len(cols) == 7:
                repeats[cols[3]] = (cols[0], int(cols[1]), int(cols[2]), cols[4], cols[5], cols[6])
            else:
                logger.log('Config: Repeat config column mismatch while parsing \n{line}'.format(line=line), 'error')
    config = {'repeat': repeats, 'align': None, 'HMM': None}
    if param_config_file:
        with open(param_config_file) as fp:
            ld_conf = json.load(fp)
        if not isinstance(ld_conf, dict) or not isinstance(ld_conf.get('align', {}), dict) \
                or not isinstance(ld_conf.get('HMM', {}), dict):
            logger.log('Config: file format broken', 'error')
            sys.exit(1)
        for key in ('align', 'HMM'):
            if key not in ld_conf:
                logger.log('Config: Error loading HMM config file, missing {}'.format(key), 'error')
                sys.exit(1)
        config['align'], config['HMM'] = ld_conf['align'], ld_conf['HMM']
    return config


class sam_record(object):
    __slots__ = ('QNAME', 'FLAG', 'RNAME', 'POS', 'TLEN', 'CLIP_BEGIN', 'CLIP_END', 'SEQ_LEN')

    def __init__(self):
        self.QNAME, self.FLAG, self.RNAME, self.POS = '', 0, '', 0
        self.TLEN = self.CLIP_BEGIN = self.CLIP_END = self.SEQ_LEN = 0


def decode_sam(sam_line):
    """scripts/STRique.py:656-671: QNAME / FLAG / RNAME / POS, reference length from CIGAR ops MDN=X,
    soft/hard clips from the first and last two CIGAR operations; empty record on any parse error."""
    cols = sam_line.rstrip().split('\t')
    sr = sam_record()
    if len(cols) >= 11:
        try:
            sr.QNAME = cols[0]
            sr.FLAG = int(cols[1])
            sr.RNAME = cols[2]
            sr.POS = int(cols[3])
            ops = [(int(op[:-1]), op[-1]) for op in re.findall(r'(\d*\D)', cols[5])]
            sr.TLEN = sum(n for n, op in ops if op in 'MDN=X')
            sr.CLIP_BEGIN = sum(n for n, op in ops[:2] if op in 'SH')
            sr.CLIP_END = sum(n for n, op in ops[-2:] if op in 'SH')
            sr.SEQ_LEN = len(cols[9])
        except Exception:  # noqa: BLE001 - the reference returns an empty record on any error
            return sam_record()
    return sr


class repeatDetector(object):
    """Multi-locus repeat detection over SAM records (scripts/STRique.py:624-705), batched."""

    def __init__(self, repeat_config, model_file, fast5_index_file, mod_model_file=None, align_config=None,
                 HMM_config=None, device=0, io_threads=1, batch_samples=96 << 20, counter=None):
        from .counter import repeatCounter
        self.repeatCounter = counter or repeatCounter(model_file, mod_model_file=mod_model_file,
                                                      align_config=align_config, HMM_config=HMM_config, device=device)
        self.repeatLoci = defaultdict(list)
        self.repeat_config = repeat_config
        self.is_init = False
        self.f5 = fast5.fast5Index(fast5_index_file)
        self.io_threads = max(int(io_threads), 1)
        self.batch_samples = int(batch_samples)

    def __init_hmm__(self):
        for target_name, (chrom, begin, end, repeat, prefix, suffix) in self.repeat_config.items():
            if target_name not in self.repeatCounter.targets:
                self.repeatCounter.add_target(target_name, repeat, prefix, suffix)
            self.repeatLoci[chrom].append((target_name, begin, end))
        self.is_init = True

    def intersect_target(self, sr):
        names = []
        for target_name, begin, end in self.repeatLoci.get(sr.RNAME, ()):
            if begin > sr.POS - sr.CLIP_BEGIN and end < sr.POS + sr.TLEN + sr.CLIP_END:
                names.append(target_name)
        return names

    def plan(self, sam_lines):
        """-> list of (input index, sam_record, strand, [target names]) for the records that hit a locus."""
        if not self.is_init:
            self.__init_hmm__()
        work = []
        for idx, line in enumerate(sam_lines):
            sr = decode_sam(line)
            if not sr.QNAME:
                logger.log('Detector: Error parsing alignment \n{}'.format(line), 'error')
                continue
            names = self.intersect_target(sr)
            if not names:
                logger.log('Detector: No target for {}'.format(sr.QNAME), 'debug')
                continue
            work.append((idx, sr, '-' if sr.FLAG & 0x10 else '+', names))
        return work

    def _fetch(self, item):
        idx, sr, strand, names = item
        try:
            return item, self.f5.get_raw(sr.QNAME)
        except Exception as e:  # noqa: BLE001 - a bad read must not stop the others (S.py:764-768)
            logger.log('Detector: {}'.format(e), 'warning')
            return item, None

    def _flush(self, batch, rows):
        items = [(name, raw, strand) for (_, _, strand, _), raw, name in batch]
        try:
            results = self.repeatCounter.detect_batch(items)
        except Exception as e:  # noqa: BLE001 - isolate the failing read
            logger.log('Detector: batch of {} reads failed ({}); retrying read by read'.format(len(items), e), 'warning')
            results = []
            for it in items:
                try:
                    results.append(self.repeatCounter.detect(*it))
                except Exception as e2:  # noqa: BLE001
                    logger.log('Detector: read failed: {}'.format(e2), 'warning')
                    results.append(None)
        for ((idx, sr, strand, _), _, name), res in zip(batch, results):
            if res is not None:
                rows.append((idx, (sr.QNAME, name, strand) + tuple(res)))

    def detect_records(self, work):
        """work: output of plan() (possibly one rank's share). -> list of (input index, row tuple)."""
        rows, batch, samples = [], [], 0
        with ThreadPoolExecutor(self.io_threads) as pool:
            for item, raw in pool.map(self._fetch, work):
                if raw is None:
                    logger.log('Detector: No fast5 for ID {id}'.format(id=item[1].QNAME), 'warning')
                    continue
                for name in item[3]:
                    batch.append((item, raw, name))
                    samples += len(raw)
                if samples >= self.batch_samples:
                    self._flush(batch, rows)
                    batch, samples = [], 0
        if batch:
            self._flush(batch, rows)
        return rows

    def detect(self, sam_line=''):
        """Reference-compatible single-record entry (scripts/STRique.py:681-705)."""
        work = self.plan([sam_line])
        if not work:
            return None
        rows = self.detect_records(work)
        return {'target_counts': [r for _, r in rows]} if rows else None


class outputWriter(object):
    """Header + tab separated rows, every value through str() (scripts/STRique.py:711-727)."""

    def __init__(self, output_file=None):
        self.output_file = output_file
        self.fp = open(output_file, 'w') if output_file else sys.stdout
        print('\t'.join(HEADER), file=self.fp)

    def write_line(self, target_counts=()):
        for target_count in target_counts:
            print('\t'.join([str(x) for x in target_count]), file=self.fp)

    def close(self):
        if self.output_file:
            self.fp.close()
        else:
            self.fp.flush()


class main(object):
    def __init__(self, argv=None):
        argv = sys.argv[1:] if argv is None else argv
        parser = argparse.ArgumentParser(
            description='STRique: a nanopore raw signal repeat detection pipeline (B200 build)',
            usage='''STRique.py <command> [<args>]
Available commands are:
   index      Index batch(es) of bulk-fast5 or tar archived single fast5
   count      Count single read repeat expansions
''')
        parser.add_argument('command', help='Subcommand to run')
        args = parser.parse_args(argv[0:1])
        if args.command not in ('index', 'count'):
            print('Unrecognized command', file=sys.stderr)
            parser.print_help(file=sys.stderr)
            sys.exit(1)
        getattr(self, args.command)(argv[1:])

    def index(self, argv):
        parser = argparse.ArgumentParser(description='Fast5 raw data archive indexing')
        parser.add_argument('input', help='Input batch or directory of batches')
        parser.add_argument('--recursive', action='store_true', help='Recursively scan input')
        parser.add_argument('--out_prefix', default='', help='Prefix for file paths in output')
        parser.add_argument('--tmp_prefix', default=None, help='Prefix for temporary data')
        args = parser.parse_args(argv)
        for record in fast5.fast5Index.index(args.input, recursive=args.recursive, output_prefix=args.out_prefix,
                                             tmp_prefix=args.tmp_prefix):
            print(record)

    def count(self, argv):
        parser = argparse.ArgumentParser(description='STR Detection in raw nanopore data')
        parser.add_argument('f5Index', help='Fast5 index')
        parser.add_argument('model', help='Pore model')
        parser.add_argument('repeat', help='Repeat region config file')
        parser.add_argument('--out', default=None, help='Output file name, if not given print to stdout')
        parser.add_argument('--algn', default=None, help='Alignment in sam format, if not given read from stdin')
        parser.add_argument('--mod_model', default=None, help='Base modification pore model')
        parser.add_argument('--config', help='Config file with HMM transition probabilities')
        parser.add_argument('--t', type=int, default=1, help='Number of fast5 decoding threads')
        parser.add_argument('--log_level', default='warning', choices=['error', 'warning', 'info', 'debug'], help='Log level')
        args = parser.parse_args(argv)
        logger.init(log_level=args.log_level)
        config = parse_config(args.repeat, args.config)
        logger.log('Main: Parsed config.', 'debug')
        if not os.path.isfile(args.f5Index):
            logger.log('Main: Fast5 index file does not exist.', 'error')
            sys.exit(1)
        if not os.path.isfile(args.model):
            logger.log('Main: Pore model file does not exist.', 'error')
            sys.exit(1)
        if args.mod_model and not os.path.isfile(args.mod_model):
            logger.log('Main: Modification pore model file does not exist.', 'error')
            sys.exit(1)
        rank, world = sharding.init_host_group()
        device = int(os.environ.get('LOCAL_RANK', 0))
        rd = repeatDetector(config['repeat'], args.model, args.f5Index, mod_model_file=args.mod_model,
                            align_config=config['align'], HMM_config=config['HMM'], device=device, io_threads=args.t)
        if args.algn:
            with open(args.algn, 'r') as fp:
                sam_lines = [line for line in fp if not line.startswith('@')]
        else:
            sam_lines = [line for line in sys.stdin if not line.startswith('@')]
        work = rd.plan(sam_lines)
        # cost of a read ~ its length (2 flank alignments over the whole signal dominate)
        shards = sharding.lpt_partition([w[1].SEQ_LEN * len(w[3]) for w in work], world)
        t0 = time.time()
        rows = rd.detect_records([work[i] for i in shards[rank]])
        logger.log('Main: rank {} processed {} reads in {:.2f} s'.format(rank, len(shards[rank]), time.time() - t0), 'info')
        rows = sharding.gather_rows(rows)
        if rank == 0:
            ow = outputWriter(args.out)
            ow.write_line([r for _, r in rows])
            ow.close()
        sharding.finalize()


def run():
    signal.signal(signal.SIGPIPE, signal.SIG_DFL)
    main()
