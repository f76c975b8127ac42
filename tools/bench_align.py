"""Micro-benchmark of the alignment stage alone (scan kernel GCUPS). Usage:
python tools/bench_align.py [n_tasks] [N] [nlev]"""
import sys
import os
import time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from strique_b200 import _lib

n_tasks = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
N = int(sys.argv[2]) if len(sys.argv) > 2 else 40000
nlev = int(sys.argv[3]) if len(sys.argv) > 3 else 145
rng = np.random.default_rng(0)
ctx = _lib.default_context(0)
n_sig = n_tasks // 2
# piecewise-constant codes like a conditioned read (dwell 6-9 samples)
codes = np.repeat(rng.integers(60, 200, size=(n_sig * N) // 6 + 8).astype(np.uint8), rng.integers(6, 10, size=(n_sig * N) // 6 + 8))[:n_sig * N]
off = np.arange(n_sig + 1, dtype=np.int64) * N
vals = np.tile(np.linspace(50, 132, 256, dtype=np.float32), (n_sig, 1)) + rng.normal(0, 0.01, (n_sig, 256)).astype(np.float32)
levels = rng.uniform(60, 120, 2 * nlev).astype(np.float32)
flank_off = [0, nlev, 2 * nlev]
ts = np.repeat(np.arange(n_sig), 2)
tf = np.tile([0, 1], n_sig)
ps = (-1.0, -16.0, -1.0, -16.0, 16.0, 0.0)
for it in range(3):
    t0 = time.time()
    res = ctx.align_batch(ps, codes, off, vals, levels, flank_off, 6, ts, tf, np.zeros_like(ts), np.zeros_like(ts))
    wall = time.time() - t0
    cells = ctx.last_align_cells
    ms = ctx.last_scan_ms
    print('iter', it, 'tasks', len(ts), 'cells %.3e' % cells, 'scan_ms %.2f' % ms, 'scan GCUPS %.1f' % (cells / ms / 1e6),
          'wall_s %.2f' % wall, 'e2e GCUPS %.1f' % (cells / wall / 1e9), 'blocks/task %.2f' % res['n_blocks'].mean(), flush=True)
