#!/usr/bin/env python
"""Benchmark of the per-read repeat-detection hot path (BASELINE.json metric: reads/s and DP GCUPS).

  python bench.py [--gpus N --steps K --warmup W]            our arm (CUDA, sm_100a)
  python bench.py --impl reference [...]                     the reference's CPU path on the host cores

One "step" = one pass of strique_detect_batch (conditioning -> 2 flank alignments per read ->
count-HMM Viterbi) over one batch of synthetic reads of configuration C2:
c9orf72 GGCCCC reads, repeat count n ~ U{2..1000}, r9_4_450bps model, strands 50/50, 1000-nt random
backbone either side, noisy pore-model simulation, int16 samples.  The 100 k-read job of
BASELINE.json configs[1] is 100000/batch such steps; reads/s does not depend on the step count.
Every rank owns its own batch (weak scaling, no data-path collective; results stay on the host).
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import re
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = os.path.join(ROOT, 'models', 'r9_4_450bps.model')
MOD_MODEL = os.path.join(ROOT, 'models', 'r9_4_450bps_mCpG.model')
# Algorithmic lane-ops per unit of the DP kernels as built (DESIGN.md "Roofline"): the score scan
# carries no provenance (the trace kernel recomputes the few 512-column blocks on the path), so an
# affine cell is 5 FADD + 4 max = 9 and, when gap_open == gap_extension (the reference's
# configuration), 3 FADD + one 3-input max = 5.  SURVEY.md section 8d's 13 counts 4 provenance selects
# this design does not execute; it is reported as `frac_survey13` for reference.
ALIGN_LANE_OPS_AFFINE = 9
ALIGN_LANE_OPS_LINEAR = 5
# Viterbi: the fixed-point kernel relaxes an in-edge with ONE add-max instruction (VIADDMNMX: the winner's name rides
# in the low bits of the score) on the integer ALU pipe, 64 lanes per SM.  SURVEY.md section 8d counts 3 lane-ops per
# edge (add, compare, select) against the 128 fp32 lanes per SM: reported as `frac_survey_fp32`; round 1's float64
# kernel was rated 3 ops against the 64 float64 lanes: `frac_fp64_convention`, for continuity.
VITERBI_OPS_PER_EDGE_FIXED = 1
VITERBI_OPS_PER_EDGE_SURVEY = 3


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='c2', choices=sorted(WORKLOADS),
                    help='BASELINE.json configuration: c2 (default, the one the metric is quoted on), c3 methylation, '
                         'c4 four-locus panel, c5 long expansions')
    ap.add_argument('--batch', type=int, default=int(os.environ.get('STRIQUE_BENCH_BATCH', 0)),
                    help='reads per step and per GPU (0: the default of the workload)')
    ap.add_argument('--n-lo', type=int, default=2)
    ap.add_argument('--n-hi', type=int, default=1000)
    ap.add_argument('--mod', action='store_true', help='same as --workload c3')
    ap.add_argument('--cpu-reads', type=int, default=0, help='reads of the CPU baseline sample (0: four per core, <= 128)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                    help='weak (default, the driver\'s contract): every rank owns a batch; strong: ONE dataset of '
                         '--dataset reads partitioned over the ranks by cost (strique_b200.sharding), rows gathered on rank 0')
    ap.add_argument('--dataset', type=int, default=65536, help='reads of the strong-scaling dataset')
    ap.add_argument('--exact', action='store_true', help='float64 Viterbi only (strique_set_viterbi_exact)')
    ap.add_argument('--cli-reads', type=int, default=32768,
                    help='reads of the CLI end-to-end measurement (`scripts/STRique.py count` on a synthetic multi-read '
                         'fast5 data set); 0 skips it')
    args = ap.parse_args()
    if args.mod and args.workload == 'c2':
        args.workload = 'c3'
    w = WORKLOADS[args.workload]
    args.mod = w['mod']
    args.loci = w['loci']
    args.fixed_n = w['fixed_n']
    if not args.batch:
        args.batch = w['batch']
    return args


# BASELINE.json configs[1..4]; batch = default reads per step and per GPU
WORKLOADS = {
    'c2': dict(loci=('c9orf72',), mod=False, fixed_n=None, batch=8192),
    'c3': dict(loci=('c9orf72',), mod=True, fixed_n=None, batch=8192),
    'c4': dict(loci=('c9orf72', 'fmr1', 'atxn10', 'dmpk'), mod=False, fixed_n=None, batch=8192),
    'c5': dict(loci=('c9orf72',), mod=False, fixed_n=4000, batch=2048, flank=4000),
}


def workload_name(args):
    from strique_b200.workload import LOCI
    loci = ' / '.join('%s %s' % (name, LOCI[name][0]) for name in args.loci)
    n = 'n=%d' % args.fixed_n if args.fixed_n else 'n~U{%d..%d}' % (args.n_lo, args.n_hi)
    return ('%s: synthetic %s reads, %s, r9_4_450bps%s, noisy int16, strands 50/50, %d reads per step per GPU'
            % (args.workload.upper(), loci, n, ' + mCpG methylation HMM' if args.mod else '', args.batch))


def make_workload(args, indices, seed):
    """reads `indices` of the workload with this seed (a pool of processes; identical to workload.make_reads)"""
    from strique_b200 import workload
    return workload.make_reads_parallel(MODEL, MOD_MODEL if args.mod else None, indices, seed=seed, loci=args.loci,
                                        n_lo=args.n_lo, n_hi=args.n_hi, mod_fraction=0.5 if args.mod else 0.0,
                                        fixed_n=args.fixed_n, flank=WORKLOADS[args.workload].get('flank', 1000))


def bench_config(args, world, raw_bytes=None):
    """`config` of the JSON line: the same for our arm and for the reference arm"""
    cfg = {'workload': workload_name(args), 'reads_per_step': args.batch * world,
           'parallelism': 'reads sharded over %d GPU(s), no collective' % world}
    if args.scaling == 'strong':
        cfg['reads_per_step'] = args.dataset
        cfg['workload'] = cfg['workload'].replace('%d reads per step per GPU' % args.batch,
                                                  'one dataset of %d reads per step' % args.dataset)
        cfg['parallelism'] = 'one dataset partitioned over %d GPU(s) by cost (LPT), rows gathered on rank 0' % world
    cfg['l2'] = 'inputs larger than L2 (several hundred MB of raw samples per GPU per step)'
    return cfg


def traffic_per_unit(kernel):
    """DRAM bytes per algorithmic unit of a kernel from the committed ncu capture (profiles/traffic.json, written by
    tools/ncu_traffic.py from an `ncu --set full` report: dram__bytes_read.sum + dram__bytes_write.sum divided by
    the units of the captured launch) -> (bytes per unit, source) or (None, reason)."""
    try:
        t = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json')))[kernel]
        return float(t['bytes_per_unit']), '%s (%s, commit %s)' % (t['source'], t['unit'], t.get('commit', '?'))
    except Exception as e:  # noqa: BLE001
        return None, 'no capture of this kernel in profiles/traffic.json (%s)' % type(e).__name__


def flank_cells(items):
    """DP cells of the two flank alignments of every (target, signal, strand) item."""
    from strique_b200.workload import LOCI
    return sum(len(s) * 6 * ((len(LOCI[name][1]) - 5) + (len(LOCI[name][2]) - 5)) for name, s, _ in items)


# ------------------------------------------------------------------------------------------------
# clocks during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.sm_max = None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:  # noqa: BLE001
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {'hw_slowdown': 'nvmlClocksThrottleReasonHwSlowdown',
                 'hw_thermal_slowdown': 'nvmlClocksThrottleReasonHwThermalSlowdown',
                 'sw_thermal_slowdown': 'nvmlClocksThrottleReasonSwThermalSlowdown',
                 'sw_power_cap': 'nvmlClocksThrottleReasonSwPowerCap'}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for key, attr in names.items():
                    if mask & getattr(nv, attr, 0):
                        self.reasons.add(key)
            except Exception:  # noqa: BLE001
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return {'sm_mhz': med, 'sm_max_mhz': self.sm_max, 'reasons': sorted(self.reasons), 'samples': len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU path (oracle: compiled reference aligner when oracle/_ref exists, C restatement otherwise)
# ------------------------------------------------------------------------------------------------
_cpu_counter = None


def _cpu_init(use_mod, loci):
    global _cpu_counter
    from oracle import reference_path as rp
    from strique_b200.workload import LOCI
    _cpu_counter = rp.RefRepeatCounter(MODEL, mod_model_file=MOD_MODEL if use_mod else None)
    for name in loci:
        _cpu_counter.add_target(name, *LOCI[name])


def _cpu_detect(item):
    name, sig, strand = item
    t0 = time.perf_counter()
    out = _cpu_counter.detect(name, sig, strand)
    return out, time.perf_counter() - t0


def cpu_kind():
    """What the CPU arm runs: the reference's own compiled aligner (oracle/_ref, built from /root/reference/src) for the
    two flank alignments -- > 80 % of its time -- and the oracle's restatement of conditioning / HMM build / float64
    Viterbi for the rest (pomegranate and scikit-image are not installable here).  'port' when oracle/_ref is absent."""
    from oracle import reference_path as rp
    return 'reference-aligner+restatement' if rp.load_pyseqan() is not None else 'port'


class CpuPool(object):
    """Pool of worker processes running the CPU path (the reference's mt_dispatcher pattern,
    scripts/STRique.py:733-830: workers pull reads from one queue); the HMMs are built in every worker before
    timing (S.py:682)."""

    def __init__(self, cores, use_mod, warm_item, loci=('c9orf72',)):
        import multiprocessing as mp
        import subprocess
        subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle'), 'liboracle.so'])
        self.cores = cores
        self.pool = mp.get_context('fork').Pool(cores, initializer=_cpu_init, initargs=(use_mod, tuple(loci)))
        self.pool.map(_cpu_detect, [warm_item] * cores, chunksize=1)

    def run(self, items):
        """-> (wall seconds, results in input order, summed per-read CPU seconds).  Longest reads first, workers pull
        one read at a time: no core waits for the slowest read of a fixed share."""
        order = sorted(range(len(items)), key=lambda k: -len(items[k][1]))
        t0 = time.perf_counter()
        res = [None] * len(items)
        cpu_s = 0.0
        for k, (out, dt) in zip(order, self.pool.imap(_cpu_detect, [items[k] for k in order], chunksize=1)):
            res[k] = out
            cpu_s += dt
        return time.perf_counter() - t0, res, cpu_s

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_sample_size(args, cores):
    return args.cpu_reads or min(4 * cores, 128)


def reference_main(args, rank, world):
    """The reference arm: the CPU path on the host cores, on the FIRST reads of rank 0's batch of our arm (same
    generator, same seed), `config` identical to our arm's."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_sample = min(cpu_sample_size(args, cores), args.batch)
    if args.fixed_n:
        n_sample = min(n_sample, 2 * cores)                  # long-expansion reads take ~20 s each
    reads = make_workload(args, range(n_sample), seed=1000)
    items = [(n, s, st) for n, s, st, _ in reads]
    times, cpu_total = [], 0.0
    pool = CpuPool(min(cores, n_sample), args.mod, min(items, key=lambda it: len(it[1])), args.loci)
    for step in range(args.warmup + args.steps):
        wall, _, cpu_s = pool.run(items)
        if step >= args.warmup:
            times.append(wall)
            cpu_total += cpu_s
    pool.close()
    total = sum(times)
    value = n_sample * len(times) / total
    cells = flank_cells(items)
    line = {'impl': 'reference', 'metric': 'reads/s', 'value': value, 'unit': 'reads/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times),
            'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
            'dtype': 'f32 align / f64 viterbi', 'data': 'synthetic', 'config': bench_config(args, max(world, 1)),
            'align_gcups': cells * len(times) / total / 1e9,
            'align_gcups_per_core': cells * len(times) / max(cpu_total, 1e-9) / 1e9,
            'cpu_baseline': {'value': value, 'unit': 'reads/s', 'cores': min(cores, n_sample), 'kind': cpu_kind(),
                             'sample': 'each step = the first %d reads of the workload (rank 0, same seed as the GPU arm), '
                                       'workers pull reads longest first' % n_sample},
            'e2e': {'value': value, 'unit': 'reads/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def cli_e2e(args, n_reads, ctx=None):
    """`scripts/STRique.py count` as a user runs it: index of multi-read fast5 files + SAM on disk -> TSV, wall clock of
    the whole process (interpreter start, CUDA context, HMM build, fast5 decode in --t worker processes, GPU batches,
    row writing).  The data set is written first (untimed).  -> dict for the JSON line"""
    import shutil
    import subprocess
    import tempfile
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import make_fast5_dataset as mk
    tmp = tempfile.mkdtemp(prefix='strique_cli_')
    try:
        t0 = time.time()
        reads = make_workload(args, range(n_reads), seed=1000)
        index_file, sam_file, _ = mk.build(tmp, reads)
        t_make = time.time() - t0
        cores = os.cpu_count() or 1
        cmd = [sys.executable, os.path.join(ROOT, 'scripts', 'STRique.py'), 'count', index_file, MODEL,
               os.path.join(ROOT, 'configs', 'panel_config.tsv'), '--algn', sam_file, '--t', str(min(cores, 32)),
               '--out', os.path.join(tmp, 'out.tsv')]
        if args.mod:
            cmd += ['--mod_model', MOD_MODEL]
        inflate = inflate_bench(ctx, index_file, ['synth-%08d' % k for k in range(min(n_reads, 8192))]) if ctx is not None else None
        # start-up alone (empty SAM): interpreter, CUDA context, HMMs of the panel, worker processes; the first launch
        # also pages the interpreter and the libraries in, so it is run once untimed
        empty = os.path.join(tmp, 'empty.sam')
        open(empty, 'w').write('@HD\tVN:1.6\n')
        cmd_empty = cmd[:cmd.index('--algn') + 1] + [empty] + cmd[cmd.index('--algn') + 2:]
        subprocess.run(cmd_empty, check=True, capture_output=True)
        t0 = time.time()
        subprocess.run(cmd_empty, check=True, capture_output=True)
        t_start = time.time() - t0

        def run(workers, host_inflate, out_name):
            """one timed run -> wall clock of the process, and the rate between its first and its last GPU batch (rows
            and times from the 'rows after' log lines: the process without start-up, first-use allocations and drain)"""
            c = [os.path.join(tmp, out_name) if x == os.path.join(tmp, 'out.tsv') else x for x in cmd] + ['--log_level', 'info']
            c[c.index('--t') + 1] = str(workers)
            env = dict(os.environ, STRIQUE_HOST_INFLATE='1') if host_inflate else dict(os.environ)
            env.pop('STRIQUE_HOST_INFLATE', None) if not host_inflate else None
            t0 = time.time()
            p = subprocess.run(c, check=True, capture_output=True, text=True, env=env)
            wall = time.time() - t0
            marks = [(int(m.group(1)), float(m.group(2))) for m in re.finditer(r'Main: (\d+) rows after ([0-9.]+) s', p.stderr)]
            steady = (marks[-1][0] - marks[0][0]) / max(marks[-1][1] - marks[0][1], 1e-9) if len(marks) >= 3 else None
            return {'value': n_reads / wall, 'wall_s': wall, 'value_after_startup': n_reads / max(wall - t_start, 1e-9),
                    'between_first_and_last_batch': steady, 'batches': len(marks)}

        t16 = min(cores, 32)
        gpu16 = run(t16, False, 'out.tsv')
        host16 = run(t16, True, 'out_host.tsv')
        gpu4 = run(4, False, 'out_gpu4.tsv')
        host4 = run(4, True, 'out_host4.tsv')
        ref_rows = open(os.path.join(tmp, 'out.tsv')).read()
        same = all(open(os.path.join(tmp, f)).read() == ref_rows for f in ('out_host.tsv', 'out_gpu4.tsv', 'out_host4.tsv'))
        rows = ref_rows.strip().split('\n')[1:]
        truth = {('synth-%08d' % k): r[3] for k, r in enumerate(reads)}
        exact = sum(1 for r in rows if int(r.split('\t')[3]) == truth[r.split('\t')[0]])
        fast5_mb = sum(os.path.getsize(os.path.join(tmp, f)) for f in os.listdir(tmp) if f.endswith('.fast5')) / 1e6
        out = dict(gpu16)
        out.update({'unit': 'reads/s', 'reads': n_reads, 'startup_s': t_start, 'io_workers': t16, 'rows': len(rows),
                    'count_exact': exact, 'fast5_mb': fast5_mb, 'dataset_build_s': t_make,
                    'inflate': 'GPU (strique_inflate_batch)', 'inflate_kernel': inflate,
                    'host_inflate': host16, 'io_workers_4': {'gpu_inflate': gpu4, 'host_inflate': host4},
                    'same_rows_in_all_runs': same,
                    'what': 'scripts/STRique.py count <index> <model> <panel_config> --algn <sam> --t <workers> --out <tsv> on '
                            'multi-read fast5 files (deflate), wall clock of the process; host_inflate = the same with zlib '
                            'on the worker processes (STRIQUE_HOST_INFLATE=1), like the reference\'s h5py'})
        return out
    except Exception as e:  # noqa: BLE001 - the hot-path numbers above must survive a failure here
        return {'error': '%s: %s' % (type(e).__name__, str(e)[:300])}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


def inflate_bench(ctx, index_file, ids):
    """strique_inflate_batch alone on the stored Signal chunks of `ids`: compressed bytes in page-locked host memory
    -> samples in HBM (upload + kernel + status read-back, as the CLI calls it), and with the compressed bytes already
    on the device (kernel + status only).  Wall clock around the synchronous entry, best of 3 after a warm-up."""
    import torch
    from strique_b200 import _lib, fast5
    f5 = fast5.fast5Index(index_file)
    recs, parts, pos, base = [], [], 0, 0
    for rid in ids:
        st = f5.get_stored(rid)
        if st[0] != 'chunks':
            return None
        _, buf, n, clen, chunks = st
        for off0, a, cs in chunks:
            recs.append((pos, (base + off0) * 2, cs, min(clen, n - off0) * 2, clen * 2, 0))
            parts.append(np.frombuffer(buf, dtype=np.uint8, count=cs, offset=a))
            pos += cs
        base += n
    chunks = np.array(recs, dtype=_lib.INFLATE_CHUNK_DTYPE)
    pinned = _lib.PinnedBuffer(pos + 16, np.uint8)
    pinned.array[:pos] = np.concatenate(parts)
    dev_comp = torch.from_numpy(pinned.array[:pos].copy()).cuda()
    out = {}
    for name, comp, space in (('from_pinned_host', pinned.array, _lib.HOST), ('from_device', dev_comp.data_ptr(), _lib.DEVICE)):
        best = None
        for k in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            _, status = ctx.inflate_batch(comp, pos, chunks, base * 2, memspace=space)
            dt = time.perf_counter() - t0
            if k and (best is None or dt < best):
                best = dt
        out[name] = {'ms': best * 1e3, 'samples_GBps': base * 2 / best / 1e9, 'reads_per_s': len(ids) / best}
    out.update({'reads': len(ids), 'chunks': len(recs), 'compressed_mb': pos / 1e6, 'samples_mb': base * 2 / 1e6,
                'failed_chunks': int(np.count_nonzero(status))})
    return out


def init_distributed(local_rank, world):
    """NCCL process group for the plumbing (barrier, max-over-ranks of the timings); no collective on the data path."""
    import torch
    import torch.distributed as dist
    if world <= 1:
        return
    # NCCL (NCCL_DEBUG=VERSION on the GPU boxes) prints its version banner on stdout when the communicator
    # is created; the contract is ONE JSON line there, so stdout points at stderr until that has happened
    import ctypes
    sys.stdout.flush()
    saved = os.dup(1)
    os.dup2(2, 1)
    try:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
        dist.barrier()
        torch.cuda.synchronize()
    finally:
        try:
            ctypes.CDLL(None).fflush(None)
        except Exception:  # noqa: BLE001
            pass
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(saved)


def roofline_block(args, dt, res, stages, cells, edges, clocks, n_sm, steps):
    """Roofline of the two DP kernels from the live stage timers (CUDA events on the library's stream)."""
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:  # noqa: BLE001
        pass
    sm_mhz = clocks['sm_mhz'] or peaks.get('sm_max_mhz', 1965.0)
    fp32_peak = n_sm * 128 * sm_mhz * 1e6           # fp32 lane-ops/s at the clock seen under load
    half_peak = n_sm * 64 * sm_mhz * 1e6            # integer-ALU and float64 lane-ops/s: 64 lanes per SM
    ran = res['hmm_ran'] == 1
    t_total = float((res['suffix_end'][ran] - res['prefix_begin'][ran]).sum())      # Viterbi columns per launch
    scan_s = stages['align_scan'] / 1e3
    vit_s = stages['viterbi_count'] / 1e3
    ac = dt.align_config
    linear = ac['gap_open_h'] == ac['gap_extension_h'] and ac['gap_open_v'] == ac['gap_extension_v']
    align_ops = ALIGN_LANE_OPS_LINEAR if linear else ALIGN_LANE_OPS_AFFINE
    scan_cups = cells / scan_s if scan_s > 0 else 0.0
    vit_eups = edges / vit_s if vit_s > 0 else 0.0
    vit_ops = VITERBI_OPS_PER_EDGE_SURVEY if args.exact else VITERBI_OPS_PER_EDGE_FIXED
    kernels = {
        'align_scan': {'bound': 'alu_issue_fp32', 'achieved': scan_cups * align_ops / 1e12,
                       'peak': fp32_peak / 1e12, 'unit': 'Tlaneop/s',
                       'frac': scan_cups * align_ops / fp32_peak, 'ops_per_cell': align_ops,
                       'frac_survey13': scan_cups * 13 / fp32_peak, 'gcups': scan_cups / 1e9,
                       'ms_per_step': stages['align_scan'] / steps, 'units_per_launch': cells / steps},
        'viterbi_count': {'bound': 'alu_issue_fp64' if args.exact else 'alu_issue_int32',
                          'achieved': vit_eups * vit_ops / 1e12, 'peak': half_peak / 1e12, 'unit': 'Tlaneop/s',
                          'frac': vit_eups * vit_ops / half_peak, 'ops_per_edge': vit_ops,
                          'frac_survey_fp32': vit_eups * VITERBI_OPS_PER_EDGE_SURVEY / fp32_peak,
                          'frac_fp64_convention': vit_eups * VITERBI_OPS_PER_EDGE_SURVEY / half_peak,
                          'gcups': vit_eups / 1e9, 'ms_per_step': stages['viterbi_count'] / steps,
                          'units_per_launch': t_total,
                          # back-pointers: 128 B per column written, 128 B read by the traceback, 8 + 8 B sample
                          'hbm_GBps': t_total * steps * 272 / max(vit_s, 1e-9) / 1e9,
                          'hbm_peak_GBps': peaks.get('hbm_gbs')},
    }
    dom = max(kernels, key=lambda k: kernels[k]['ms_per_step'])
    roofline = dict(kernels[dom])
    per_unit, source = traffic_per_unit('viterbi_profile_q' if (dom == 'viterbi_count' and not args.exact) else
                                        ('viterbi_profile' if dom == 'viterbi_count' else 'align_scan'))
    roofline.update({'kernel': dom, 'traffic': per_unit * roofline['units_per_launch'] if per_unit is not None else None,
                     'traffic_source': source,
                     'peak_source': 'issue peak = N_SM x lanes x SM clock sampled under load (fp32: 128 lanes per SM; '
                                    'integer ALU and float64: 64 lanes per SM); HBM peak: ' +
                                    ('measured (MEASURED_PEAKS.json)' if peaks else 'fallback 6650 GB/s')})
    return roofline, kernels


def main():
    args = parse_args()
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        reference_main(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from strique_b200 import _lib, sharding, workload
    from strique_b200.counter import repeatCounter

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the hot path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    init_distributed(local_rank, world)
    ctx = _lib.Context(local_rank)
    ctx.set_viterbi_exact(args.exact)
    dt = repeatCounter(MODEL, mod_model_file=MOD_MODEL if args.mod else None, context=ctx)
    for name in args.loci:
        dt.add_target(name, *workload.LOCI[name])
    cfg = dt._detect_config()
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device('cuda', local_rank))
    n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.scaling == 'strong':
        strong_main(args, rank, local_rank, world, ctx, dt, cfg, stream, barrier)
        return

    # ---- this rank's batch ------------------------------------------------------------------------
    t_gen = time.time()
    reads = make_workload(args, range(args.batch), seed=1000 + rank)
    tids = np.array([dt._target_id(name, strand) for name, _, strand, _ in reads], dtype=np.int32)
    raw_np, off, kind = _lib.Context._pack_raw([s for _, s, _, _ in reads])
    t_gen = time.time() - t_gen
    raw_pinned = torch.empty(len(raw_np), dtype=torch.int16).pin_memory()
    raw_pinned.numpy()[:] = raw_np
    raw_dev = raw_pinned.cuda()
    torch.cuda.synchronize()

    def run_steps(n_steps, host_buffers):
        """-> (device ms, accumulated stage ms, results of the last step, cells, edges)"""
        stages = {}
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record(stream)
        res = None
        cells = edges = 0
        for _ in range(n_steps):
            if host_buffers:
                res, mod = ctx.detect_batch(cfg, raw_pinned.numpy(), off, kind, tids, memspace=_lib.HOST)
            else:
                res, mod = ctx.detect_batch(cfg, raw_dev.data_ptr(), off, kind, tids, memspace=_lib.DEVICE)
            for k, v in ctx.stage_ms().items():
                stages[k] = stages.get(k, 0.0) + v
            cells += ctx.last_align_cells
            edges += ctx.last_viterbi_edges
        ev1.record(stream)
        barrier()
        return ev0.elapsed_time(ev1), stages, res, cells, edges

    run_steps(args.warmup, False)
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = ctx.launches
    ms, stages, res, cells, edges = run_steps(args.steps, False)
    launches = ctx.launches - launches0
    fixed, declined = ctx.last_viterbi_fixed
    clocks = sampler.stop()
    run_steps(1, True)
    ms_e2e, _, res_e2e, _, _ = run_steps(args.steps, True)

    # max over ranks of the timed regions
    if world > 1:
        t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, ms_e2e = float(t[0]), float(t[1])
        w = torch.tensor([float(cells), float(edges)], dtype=torch.float64, device='cuda')
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        cells_all, edges_all = float(w[0]), float(w[1])
    else:
        cells_all, edges_all = float(cells), float(edges)
    total_reads = args.batch * args.steps * world
    value = total_reads / (ms / 1e3)
    e2e_value = total_reads / (ms_e2e / 1e3)

    if rank == 0:
        # ---- sanity of the measured pass: counts against the simulated truth ------------------------
        truth = np.array([n for _, _, _, n in reads])
        got = res['count']
        exact = int((got == truth).sum())
        within1 = int((np.abs(got - truth) <= 1).sum())
        assert np.array_equal(res['count'], res_e2e['count']) and np.array_equal(res['offset'], res_e2e['offset'])
        # ---- the fixed-point Viterbi against the float64 kernel on this very batch (after the timed regions) ----
        viterbi_check = None
        if not args.exact:
            ctx.set_viterbi_exact(True)
            res64, _ = ctx.detect_batch(cfg, raw_dev.data_ptr(), off, kind, tids, memspace=_lib.DEVICE)
            ctx.set_viterbi_exact(False)
            diff = np.flatnonzero((res64['count'] != res['count']) | (res64['hmm_ran'] != res['hmm_ran']))
            lp = np.abs(res64['log_p'] - res['log_p'])
            viterbi_check = {'reads': int(len(truth)), 'decoded_fixed_point': int(fixed), 'handed_to_float64': int(declined),
                             'count_differs_from_float64': int(len(diff)), 'differing_reads': [int(i) for i in diff[:32]],
                             'max_abs_log_p_gap': float(lp.max()) if len(lp) else 0.0,
                             'reads_with_log_p_gap_over_1e-9_rel': int((lp > 1e-9 * np.abs(res64['log_p'])).sum())}
        roofline, kernels = roofline_block(args, dt, res, stages, cells, edges, clocks, n_sm, args.steps)
        line = {'metric': 'reads/s', 'value': value, 'unit': 'reads/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None,
                'dtype': 'f32 align / f64 viterbi' if args.exact else 'f32 align / i32 fixed-point viterbi (f64 re-score)',
                'data': 'synthetic', 'config': bench_config(args, world),
                'dp_gcups': (cells_all + edges_all) / (ms / 1e3) / 1e9,
                'align_gcups': cells_all / (ms / 1e3) / 1e9, 'viterbi_gcups': edges_all / (ms / 1e3) / 1e9,
                'stage_ms_per_step': {k: v / args.steps for k, v in stages.items()},
                'roofline': roofline, 'roofline_kernels': kernels,
                'e2e': {'value': e2e_value, 'unit': 'reads/s',
                        'h2d_bytes_per_step': int(raw_np.nbytes + off.nbytes + tids.nbytes) * world,
                        'd2h_bytes_per_step': int(res.nbytes) * world},
                'gpu_launches': int(launches), 'clocks': clocks,
                'accuracy': {'reads': int(len(truth)), 'count_exact': exact, 'count_within_1': within1},
                'viterbi_check': viterbi_check, 'gen_s': t_gen, 'raw_mb_per_gpu_per_step': raw_np.nbytes / 1e6}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_sample = min(cpu_sample_size(args, cores), args.batch)
            if args.fixed_n:
                n_sample = min(n_sample, cores)
            items = [(n, s, st) for n, s, st, _ in reads[:n_sample]]
            pool = CpuPool(min(cores, n_sample), args.mod, min(items, key=lambda it: len(it[1])), args.loci)
            wall, cpu_res, cpu_s = pool.run(items)
            pool.close()
            mism = sum(1 for k in range(n_sample)
                       if (int(cpu_res[k][0]), int(cpu_res[k][4]), int(cpu_res[k][5])) !=
                       (int(res['count'][k]) if res['hmm_ran'][k] else 0, int(res['offset'][k]), int(res['ticks'][k])))
            if args.cli_reads > 0:
                line['cli_e2e'] = cli_e2e(args, args.cli_reads, ctx)
            line['cpu_baseline'] = {'value': n_sample / wall, 'unit': 'reads/s', 'cores': min(cores, n_sample),
                                    'kind': cpu_kind(),
                                    'sample': 'first %d reads of the step, workers pull reads longest first, %.1f s wall'
                                              % (n_sample, wall),
                                    'align_gcups_per_core': flank_cells(items) / max(cpu_s, 1e-9) / 1e9,
                                    'mismatches_vs_gpu': mism}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def strong_main(args, rank, local_rank, world, ctx, dt, cfg, stream, barrier):
    """Strong scaling through the real sharding path (the reference's dispatcher, scripts/STRique.py:733-830, hands
    reads to workers; here strique_b200.sharding.lpt_partition deals the reads of ONE dataset to the ranks by cost):
    every rank decodes its shard in batches of --batch reads from pinned host buffers and rank 0 gathers the result
    rows of all ranks (44 bytes per read, as one byte tensor per rank through the process group) -- all inside the
    timed region.  Reports per-rank milliseconds and the imbalance."""
    import torch
    import torch.distributed as dist
    from strique_b200 import _lib, sharding
    n = args.dataset
    # ---- the dataset: every rank makes a strided slice to learn the read lengths, the lengths are all-gathered, every
    # rank computes the same partition and then makes the reads it was dealt (reads are individually seeded)
    t_gen = time.time()
    mine = list(range(rank, n, world))
    part = make_workload(args, mine, seed=4000)
    lens = torch.zeros(n, dtype=torch.int64, device='cuda')
    lens[torch.tensor(mine, device='cuda')] = torch.tensor([len(s) for _, s, _, _ in part], device='cuda')
    if world > 1:
        dist.all_reduce(lens, op=dist.ReduceOp.SUM)
    lens = lens.cpu().numpy()
    shards = sharding.lpt_partition([int(x) for x in lens], world)
    have = dict(zip(mine, part))
    need = [i for i in shards[rank] if i not in have]
    have.update(zip(need, make_workload(args, need, seed=4000) if need else []))
    reads = [have[i] for i in shards[rank]]
    del have, part
    t_gen = time.time() - t_gen
    # batches of --batch reads, pinned
    batches = []
    for b0 in range(0, len(reads), args.batch):
        chunk = reads[b0:b0 + args.batch]
        tids = np.array([dt._target_id(name, strand) for name, _, strand, _ in chunk], dtype=np.int32)
        raw_np, off, kind = _lib.Context._pack_raw([s for _, s, _, _ in chunk])
        pinned = torch.empty(len(raw_np), dtype=torch.int16).pin_memory()
        pinned.numpy()[:] = raw_np
        batches.append((pinned, off, kind, tids))
    row_dtype = np.dtype([('index', np.int64), ('count', np.int32), ('offset', np.int32), ('ticks', np.int32),
                          ('score_prefix', np.float64), ('score_suffix', np.float64), ('log_p', np.float64)])

    def one_pass():
        rows = np.zeros(len(reads), dtype=row_dtype)
        rows['index'] = shards[rank]
        k = 0
        for pinned, off, kind, tids in batches:
            res, _ = ctx.detect_batch(cfg, pinned.numpy(), off, kind, tids, memspace=_lib.HOST)
            m = len(tids)
            for f in ('count', 'offset', 'ticks', 'score_prefix', 'score_suffix', 'log_p'):
                rows[f][k:k + m] = res[f]
            k += m
        # host-side gather of the rows on rank 0 (44 bytes per read), in input order
        if world > 1:
            sizes = [len(s) for s in shards]
            cap = max(sizes) * row_dtype.itemsize                 # gather wants equal sizes: pad to the largest shard
            mine_t = torch.zeros(cap, dtype=torch.uint8, device='cuda')
            mine_t[:rows.nbytes] = torch.from_numpy(rows.view(np.uint8).copy()).cuda()
            if rank == 0:
                parts = [torch.empty(cap, dtype=torch.uint8, device='cuda') for _ in sizes]
                dist.gather(mine_t, parts, dst=0)
                allrows = np.concatenate([p.cpu().numpy()[:sz * row_dtype.itemsize].view(row_dtype)
                                          for p, sz in zip(parts, sizes)])
            else:
                dist.gather(mine_t, None, dst=0)
                allrows = None
        else:
            allrows = rows
        if allrows is not None:
            allrows = allrows[np.argsort(allrows['index'], kind='stable')]
        return allrows

    for _ in range(args.warmup):
        one_pass()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0 = time.perf_counter()
    ev0.record(stream)
    launches0 = ctx.launches
    my_ms = 0.0
    for _ in range(args.steps):
        t1 = time.perf_counter()
        allrows = one_pass()
        my_ms += (time.perf_counter() - t1) * 1e3
    ev1.record(stream)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = ctx.launches - launches0
    clocks = sampler.stop()
    dev_ms = ev0.elapsed_time(ev1)
    per_rank = torch.zeros(world, dtype=torch.float64, device='cuda')
    per_rank[rank] = my_ms / args.steps
    t = torch.tensor([max(dev_ms, wall_ms)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
    ms = float(t[0])
    per_rank = [float(x) for x in per_rank.cpu()]
    if rank == 0:
        assert len(allrows) == n and np.array_equal(allrows['index'], np.arange(n))
        value = n * args.steps / (ms / 1e3)
        samples = [int(lens[s].sum()) for s in shards]
        line = {'metric': 'reads/s', 'value': value, 'unit': 'reads/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'strong',
                'vs_baseline': None, 'dtype': 'f32 align / i32 fixed-point viterbi (f64 re-score)', 'data': 'synthetic',
                'config': bench_config(args, world),
                'e2e': {'value': value, 'unit': 'reads/s', 'h2d_bytes_per_step': int(lens.sum()) * 2,
                        'd2h_bytes_per_step': n * row_dtype.itemsize},
                'per_rank_ms': per_rank, 'imbalance_max_over_mean': max(per_rank) / (sum(per_rank) / len(per_rank)),
                'per_rank_samples': samples, 'batches_per_rank': len(batches), 'gpu_launches': int(launches),
                'clocks': clocks, 'gen_s': t_gen,
                'note': 'timed region = every rank: detect_batch over its shard from pinned host buffers + gather of the '
                        'rows on rank 0; value = dataset reads / max over ranks'}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
