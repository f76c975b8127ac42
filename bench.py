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
VITERBI_LANE_OPS_PER_EDGE = 3     # 1 DADD + 1 compare + 1 select (fp64)


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
    ap.add_argument('--cpu-reads', type=int, default=0, help='reads of the CPU baseline sample (0: one per core, <= 32)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
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


def make_workload(args, pm, pm_mod, n_reads, seed):
    from strique_b200 import workload
    return workload.make_reads(pm, n_reads, seed=seed, loci=args.loci, n_lo=args.n_lo, n_hi=args.n_hi,
                               pm_mod=pm_mod if args.mod else None, mod_fraction=0.5 if args.mod else 0.0,
                               fixed_n=args.fixed_n, flank=WORKLOADS[args.workload].get('flank', 1000))


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
    from oracle import reference_path as rp
    return 'reference' if rp.load_pyseqan() is not None else 'port'


class CpuPool(object):
    """Pool of worker processes running the CPU path (the reference's mt_dispatcher pattern,
    scripts/STRique.py:733-830); the HMMs are built in every worker before timing (S.py:682)."""

    def __init__(self, cores, use_mod, warm_item, loci=('c9orf72',)):
        import multiprocessing as mp
        import subprocess
        subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle'), 'liboracle.so'])
        self.cores = cores
        self.pool = mp.get_context('fork').Pool(cores, initializer=_cpu_init, initargs=(use_mod, tuple(loci)))
        self.pool.map(_cpu_detect, [warm_item] * cores, chunksize=1)

    def run(self, items):
        """-> (wall seconds, results)"""
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_detect, items, chunksize=1)
        return time.perf_counter() - t0, [r[0] for r in res]

    def close(self):
        self.pool.close()
        self.pool.join()


def reference_main(args, rank, world):
    from strique_b200 import pore_model as pmod, workload
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    pm = pmod.pore_model(MODEL)
    n_sample = args.cpu_reads or min(cores, 256)
    pool_reads = make_workload(args, pm, pmod.pore_model(MOD_MODEL) if args.mod else None, n_sample, seed=0)
    items = [(n, s, st) for n, s, st, _ in pool_reads]
    times = []
    pool = CpuPool(min(cores, n_sample), args.mod, min(items, key=lambda it: len(it[1])), args.loci)
    for step in range(args.warmup + args.steps):
        wall, _ = pool.run(items)
        if step >= args.warmup:
            times.append(wall)
    pool.close()
    total = sum(times)
    value = n_sample * len(times) / total
    cells = flank_cells(items)
    line = {'impl': 'reference', 'metric': 'reads/s', 'value': value, 'unit': 'reads/s', 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * total / len(times),
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 align / f64 viterbi',
            'data': 'synthetic', 'config': {'workload': workload_name(args), 'sample_reads_per_step': n_sample},
            'align_gcups': cells * len(times) / total / 1e9,
            'cpu_baseline': {'value': value, 'unit': 'reads/s', 'cores': min(cores, n_sample), 'kind': cpu_kind(),
                             'sample': '%d reads of the workload per step, one worker process per core' % n_sample},
            'e2e': {'value': value, 'unit': 'reads/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
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
    from strique_b200 import _lib, workload
    from strique_b200.counter import repeatCounter

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- the hot path has no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
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
    ctx = _lib.Context(local_rank)
    dt = repeatCounter(MODEL, mod_model_file=MOD_MODEL if args.mod else None, context=ctx)
    for name in args.loci:
        dt.add_target(name, *workload.LOCI[name])
    cfg = dt._detect_config()

    # ---- this rank's batch ------------------------------------------------------------------------
    t_gen = time.time()
    reads = make_workload(args, dt.pm, dt.pm_mod, args.batch, seed=1000 + rank)
    tids = np.array([dt._target_id(name, strand) for name, _, strand, _ in reads], dtype=np.int32)
    raw_np, off, kind = _lib.Context._pack_raw([s for _, s, _, _ in reads])
    t_gen = time.time() - t_gen
    raw_pinned = torch.empty(len(raw_np), dtype=torch.int16).pin_memory()
    raw_pinned.numpy()[:] = raw_np
    raw_dev = raw_pinned.cuda()
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device('cuda', local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_steps(n_steps, host_buffers):
        """-> (device ms, accumulated stage ms, results of the last step)"""
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

        # ---- roofline of the DP kernels (this rank) ------------------------------------------------
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:  # noqa: BLE001
            pass
        sm_mhz = clocks['sm_mhz'] or peaks.get('sm_max_mhz', 1965.0)
        n_sm = torch.cuda.get_device_properties(local_rank).multi_processor_count
        alu_peak = n_sm * 128 * sm_mhz * 1e6            # fp32 lane-ops/s at the clock seen under load
        fp64_peak = n_sm * 64 * sm_mhz * 1e6            # fp64 lane-ops/s (B200: 64 DFMA lanes per SM)
        ran = res['hmm_ran'] == 1
        t_total = float((res['suffix_end'][ran] - res['prefix_begin'][ran]).sum())
        scan_s = stages['align_scan'] / 1e3
        trace_s = stages['align_trace'] / 1e3
        vit_s = stages['viterbi_count'] / 1e3
        ac = dt.align_config
        linear = ac['gap_open_h'] == ac['gap_extension_h'] and ac['gap_open_v'] == ac['gap_extension_v']
        align_ops = ALIGN_LANE_OPS_LINEAR if linear else ALIGN_LANE_OPS_AFFINE
        scan_cups = cells / scan_s if scan_s > 0 else 0.0
        vit_eups = edges / vit_s if vit_s > 0 else 0.0
        kernels = {
            'align_scan': {'bound': 'alu_issue_fp32', 'achieved': scan_cups * align_ops / 1e12,
                           'peak': alu_peak / 1e12, 'unit': 'Tlaneop/s',
                           'frac': scan_cups * align_ops / alu_peak, 'ops_per_cell': align_ops,
                           'frac_survey13': scan_cups * 13 / alu_peak, 'gcups': scan_cups / 1e9,
                           'ms_per_step': stages['align_scan'] / args.steps},
            'viterbi_count': {'bound': 'alu_issue_fp64', 'achieved': vit_eups * VITERBI_LANE_OPS_PER_EDGE / 1e12,
                              'peak': fp64_peak / 1e12, 'unit': 'Tlaneop/s',
                              'frac': vit_eups * VITERBI_LANE_OPS_PER_EDGE / fp64_peak,
                              'ops_per_edge': VITERBI_LANE_OPS_PER_EDGE, 'gcups': vit_eups / 1e9,
                              'ms_per_step': stages['viterbi_count'] / args.steps,
                              # back-pointers: 128 B per time step written, 128 B read by the traceback, 8 B sample
                              'hbm_GBps': t_total * args.steps * 264 / max(vit_s, 1e-9) / 1e9,
                              'hbm_peak_GBps': peaks.get('hbm_gbs')},
        }
        dom = max(kernels, key=lambda k: kernels[k]['ms_per_step'])
        roofline = dict(kernels[dom])
        # DRAM traffic per launch of the dominant kernel, from the ncu --set full capture in profiles/
        # (viterbi: 251 B per time step measured vs 264 B algorithmic; scan: 0.043 B per cell)
        traffic = t_total * 251.0 if dom == 'viterbi_count' else (cells / args.steps) * 0.0428
        roofline.update({'kernel': dom, 'traffic': traffic, 'traffic_source': 'profiles/ncu_*_r01v.txt scaled to this launch',
                         'peak_source': 'issue peak = N_SM x lanes x SM clock sampled under load (fp32: 128 lanes/SM, '
                                        'fp64: 64 lanes/SM); HBM peak: ' +
                                        ('of measured (MEASURED_PEAKS.json)' if peaks else 'of fallback 6650 GB/s')})
        line = {'metric': 'reads/s', 'value': value, 'unit': 'reads/s', 'n_gpus': world, 'steps': args.steps,
                'warmup': args.warmup, 'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
                'vs_baseline': None, 'dtype': 'f32 align / f64 viterbi', 'data': 'synthetic',
                'config': {'workload': workload_name(args), 'reads_per_step': args.batch * world,
                           'l2': 'inputs larger than L2 (%.0f MB raw per GPU per step)' % (raw_np.nbytes / 1e6),
                           'parallelism': 'reads sharded over %d GPU(s), no collective' % world},
                'dp_gcups': (cells_all + edges_all) / (ms / 1e3) / 1e9,
                'align_gcups': cells_all / (ms / 1e3) / 1e9, 'viterbi_gcups': edges_all / (ms / 1e3) / 1e9,
                'stage_ms_per_step': {k: v / args.steps for k, v in stages.items()},
                'roofline': roofline, 'roofline_kernels': kernels,
                'e2e': {'value': e2e_value, 'unit': 'reads/s',
                        'h2d_bytes_per_step': int(raw_np.nbytes + off.nbytes + tids.nbytes) * world,
                        'd2h_bytes_per_step': int(res.nbytes) * world},
                'gpu_launches': int(launches), 'clocks': clocks,
                'accuracy': {'reads': int(len(truth)), 'count_exact': exact, 'count_within_1': within1},
                'gen_s': t_gen}
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_sample = args.cpu_reads or min(cores, 32)
            items = [(n, s, st) for n, s, st, _ in reads[:n_sample]]
            pool = CpuPool(min(cores, n_sample), args.mod, min(items, key=lambda it: len(it[1])), args.loci)
            wall, cpu_res = pool.run(items)
            pool.close()
            mism = sum(1 for k in range(n_sample)
                       if (int(cpu_res[k][0]), int(cpu_res[k][4]), int(cpu_res[k][5])) !=
                       (int(res['count'][k]) if res['hmm_ran'][k] else 0, int(res['offset'][k]), int(res['ticks'][k])))
            line['cpu_baseline'] = {'value': n_sample / wall, 'unit': 'reads/s', 'cores': min(cores, n_sample),
                                    'kind': cpu_kind(),
                                    'sample': 'first %d reads of the step, one worker process per core, %.1f s wall'
                                              % (n_sample, wall),
                                    'mismatches_vs_gpu': mism}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()



if __name__ == '__main__':
    main()
