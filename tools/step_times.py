"""Per-step stage times of repeated detect_batch calls over one resident batch (run-to-run variance check).
    python tools/step_times.py [--batch 4096] [--steps 8]"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--batch', type=int, default=4096)
    ap.add_argument('--steps', type=int, default=8)
    ap.add_argument('--seed', type=int, default=1000)
    ap.add_argument('--host', action='store_true', help='pass pinned HOST buffers (the e2e path)')
    ap.add_argument('--sample-ms', type=float, default=2.0, help='NVML sampling period (0: no sampler thread)')
    ap.add_argument('--strand', default='', help="keep only reads of this strand ('+' or '-'): one HMM, no model hand-over")
    a = ap.parse_args()
    import torch
    from strique_b200 import _lib, workload
    from strique_b200.counter import repeatCounter
    model = os.path.join(ROOT, 'models', 'r9_4_450bps.model')
    ctx = _lib.Context(0)
    dt = repeatCounter(model, context=ctx)
    dt.add_target('c9orf72', *workload.LOCI['c9orf72'])
    cfg = dt._detect_config()
    reads = workload.make_reads(dt.pm, a.batch * (2 if a.strand else 1) + (64 if a.strand else 0), seed=a.seed)
    if a.strand:
        reads = [r for r in reads if r[2] == a.strand][:a.batch]
    tids = np.array([dt._target_id(n, s) for n, _, s, _ in reads], dtype=np.int32)
    raw_np, off, kind = _lib.Context._pack_raw([s for _, s, _, _ in reads])
    raw_dev = torch.from_numpy(raw_np).cuda()
    torch.cuda.synchronize()
    import threading
    import time
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    samples = []
    stop = threading.Event()

    def sampler():
        while not stop.is_set():
            samples.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                            pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                            pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
            time.sleep(a.sample_ms / 1e3)
    if a.sample_ms > 0:
        threading.Thread(target=sampler, daemon=True).start()
    raw_pinned = torch.from_numpy(raw_np).pin_memory()
    for i in range(a.steps):
        t0 = time.perf_counter()
        if a.host:
            ctx.detect_batch(cfg, raw_pinned.numpy(), off, kind, tids, memspace=_lib.HOST)
        else:
            ctx.detect_batch(cfg, raw_dev.data_ptr(), off, kind, tids, memspace=_lib.DEVICE)
        wall = (time.perf_counter() - t0) * 1e3
        st = ctx.stage_ms()
        t1 = time.perf_counter()
        mine = [x for x in samples if t0 <= x[0] <= t1]
        clk = [x[1] for x in mine]
        print(i, 'wall %.1f ms, stages %.1f ms' % (wall, sum(st.values())), {k: round(v, 1) for k, v in st.items()},
              'sm MHz min/med/max %d/%d/%d' % (min(clk), sorted(clk)[len(clk) // 2], max(clk)) if clk else '',
              'W max %.0f' % max(x[2] for x in mine) if mine else '', 'reasons %s' % sorted({hex(x[3]) for x in mine}), flush=True)


if __name__ == '__main__':
    main()
