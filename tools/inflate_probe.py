#!/usr/bin/env python
"""strique_inflate_batch alone on synthetic signal chunks (no fast5 files): timing and zlib parity.
    python tools/inflate_probe.py [reads] [level]"""
import os
import sys
import time
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from strique_b200 import _lib
    n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    level = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    rng = np.random.default_rng(0)
    clen = 8192
    pool = []
    for k in range(96):                                   # distinct chunk payloads, cycled
        n = clen * 2
        levels = np.repeat(rng.uniform(450, 750, n // 5 + 2), rng.integers(5, 10, n // 5 + 2))[:clen]
        pool.append(np.round(levels + rng.normal(0, 12, clen)).astype('<i2'))
    full_streams = [zlib.compress(p.tobytes(), level) for p in pool]
    recs, parts, pos, base, want = [], [], 0, 0, []
    for r in range(n_reads):
        n = int(rng.integers(20000, 60000))
        for off0 in range(0, n, clen):
            k = int(rng.integers(0, len(pool)))
            keep = min(clen, n - off0)
            if keep == clen:
                s = full_streams[k]
            else:
                s = zlib.compress(pool[k][:keep].tobytes() + bytes((clen - keep) * 2), level)
            recs.append((pos, (base + off0) * 2, len(s), keep * 2, clen * 2, 0))
            parts.append(np.frombuffer(s, np.uint8))
            want.append((k, keep))
            pos += len(s)
        base += n
    chunks = np.array(recs, dtype=_lib.INFLATE_CHUNK_DTYPE)
    ctx = _lib.Context(0)
    pinned = _lib.PinnedBuffer(pos + 16, np.uint8)
    pinned.array[:pos] = np.concatenate(parts)
    dev_comp = torch.from_numpy(pinned.array[:pos].copy()).cuda()
    for name, comp, space in (('pinned host', pinned.array, _lib.HOST), ('device', dev_comp.data_ptr(), _lib.DEVICE)):
        best = None
        for k in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            dev, status = ctx.inflate_batch(comp, pos, chunks, base * 2, memspace=space)
            dt = time.perf_counter() - t0
            best = dt if best is None or (k and dt < best) else best
        print('%-12s %8.2f ms  %6.1f GB/s of samples  %8.0f reads/s  failed %d' % (
            name, best * 1e3, base * 2 / best / 1e9, n_reads / best, int(np.count_nonzero(status))))

    class _Ext:
        __cuda_array_interface__ = {'shape': (base,), 'typestr': '<i2', 'data': (dev, False), 'version': 2}
    out = torch.as_tensor(_Ext(), device='cuda').cpu().numpy()
    bad = 0
    for (src, dst, sl, keep, full, _), (k, kp) in zip(recs, want):
        if not np.array_equal(out[dst // 2:dst // 2 + kp], pool[k][:kp]):
            bad += 1
    print('reads', n_reads, 'chunks', len(recs), 'compressed MB %.1f' % (pos / 1e6), 'samples MB %.1f' % (base * 2 / 1e6),
          'chunks differing from zlib:', bad)


if __name__ == '__main__':
    main()
