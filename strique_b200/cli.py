"""`STRique.py` command line of the reference (scripts/STRique.py:874-946), served by the CUDA path.

    STRique.py index <input> [--recursive] [--out_prefix P] [--tmp_prefix T]
    STRique.py count <f5Index> <model> <repeat> [--out F] [--algn SAM] [--mod_model M] [--config J]
                     [--t N] [--log_level L]

Same arguments, same `repeat_config.tsv` / pore model / JSON inputs, same ten TSV columns.  What is
different underneath: SAM records are decoded and intersected with the loci on the host (as in
repeatDetector, scripts/STRique.py:624-705), the raw signals are fetched by `--t` I/O threads running ahead of
the GPU (bounded by the size of one batch), and reads go to the GPU in batches through
`repeatCounter.detect_batch` (strique_detect_batch).

Streaming, like the reference (scripts/STRique.py:720-727, 936-945): the SAM is consumed as it arrives (a pipe on
stdin works) and the rows of a batch are appended to the output as soon as the batch is done, so partial output
survives a crash.  Row order is INPUT order (the reference writes in completion order, which is not reproducible
for --t > 1).  Launched under `torchrun --nproc-per-node N`, rank 0 alone reads the SAM, plans it in chunks and
broadcasts every chunk; each rank decodes a cost-balanced share of the chunk on its own GPU and rank 0 gathers and
writes the chunk's rows (strique_b200/sharding.py).
"""
import argparse
import json
import os
import re
import signal
import sys
import threading
import time
from collections import defaultdict, deque
from concurrent.futures import ProcessPoolExecutor, ThreadPoolExecutor

import numpy as np

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
    cols = sam_line.split('\t', 10)          # (SEQ is field 10; QUAL and the tags stay unsplit in cols[10])
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


# ---- fast5 decoding in worker PROCESSES (--t > 1): the HDF5 walk is Python and holds the GIL, only zlib does not --
# Decoded signals come back through shared-memory slots, not through the result pipe (one parent process unpickling
# 80 KB per read tops out near 200 MB/s, i.e. ~2.5 k reads/s): a task names a slot, the worker writes the signals of
# its reads into it back to back and returns only their lengths.
_worker_index = None
_worker_slots = {}
SLOT_SAMPLES = 8 << 20                        # int16 samples per slot (16 MB): a task of 16 reads of up to 500 k samples


def _worker_init(index_file):
    global _worker_index
    signal.signal(signal.SIGINT, signal.SIG_IGN)
    _worker_index = fast5.fast5Index(index_file)


def _worker_fetch(read_ids, slot_name):
    """-> (lengths: samples written into the slot per read, -1 failed, -2 returned in `spill` instead; spill: signals
    that are not int16 or did not fit the slot; error messages)"""
    from multiprocessing import shared_memory
    shm = _worker_slots.get(slot_name)
    if shm is None:
        shm = _worker_slots[slot_name] = shared_memory.SharedMemory(name=slot_name)

    dst = np.frombuffer(shm.buf, dtype=np.int16)
    lens, spill, errs, pos = [], [], [], 0
    for rid in read_ids:
        try:
            raw = np.asarray(_worker_index.get_raw(rid))
        except Exception as e:  # noqa: BLE001 - a bad read must not stop the others (S.py:764-768)
            lens.append(-1)
            errs.append(str(e))
            continue
        if raw.dtype == np.int16 and pos + len(raw) <= len(dst):
            dst[pos:pos + len(raw)] = raw
            pos += len(raw)
            lens.append(len(raw))
        else:
            lens.append(-2)
            spill.append(raw)
    return lens, spill, errs


class StoredRead(object):
    """A read whose Signal chunks are still deflate-compressed (fast5.stored_raw_signal): samples, samples per chunk,
    [(first sample, stored bytes)] and the stored bytes of all chunks back to back (uint8 view)."""
    __slots__ = ('n', 'clen', 'chunks', 'data')

    def __init__(self, n, clen, chunks, data):
        self.n, self.clen, self.chunks, self.data = n, clen, chunks, data


def _stored_chunks_into(st, dst, pos):
    """copy the stored bytes of the chunks of one read (st = fast5.stored_raw_signal result) to dst[pos:]
    -> (record for the parent, bytes written) or None when they do not fit"""
    _, buf, n, clen, chunks = st
    total = sum(cs for _, _, cs in chunks)
    if pos + total > len(dst):
        return None
    p = pos
    for _, a, cs in chunks:
        dst[p:p + cs] = np.frombuffer(buf, dtype=np.uint8, count=cs, offset=a)
        p += cs
    return ('c', n, clen, [(off0, cs) for off0, _, cs in chunks]), total


def _worker_fetch_stored(read_ids, slot_name):
    """Like _worker_fetch, but deflate-compressed reads are NOT inflated: their stored chunk bytes go into the slot
    and the GPU inflates them (strique_inflate_batch).  -> (records, spill, errors); a record is None (failed),
    ('c', samples, samples per chunk, [(first sample, stored bytes)]) with the bytes in the slot, ('r', samples) with
    int16 samples in the slot (a dataset stored some other way), or ('s',) with the decoded signal in `spill`."""
    from multiprocessing import shared_memory
    shm = _worker_slots.get(slot_name)
    if shm is None:
        shm = _worker_slots[slot_name] = shared_memory.SharedMemory(name=slot_name)
    dst = np.frombuffer(shm.buf, dtype=np.uint8)
    recs, spill, errs, pos = [], [], [], 0
    for rid in read_ids:
        try:
            st = _worker_index.get_stored(rid)
            if st[0] == 'chunks':
                put = _stored_chunks_into(st, dst, pos)
                if put is not None:
                    recs.append(put[0])
                    pos += put[1]
                    continue
                raw = np.asarray(_worker_index.get_raw(rid))
            else:
                raw = np.asarray(st[1])
        except Exception as e:  # noqa: BLE001 - a bad read must not stop the others (S.py:764-768)
            recs.append(None)
            errs.append(str(e))
            continue
        pos += pos & 1
        if raw.dtype == np.int16 and pos + 2 * len(raw) <= len(dst):
            dst[pos:pos + 2 * len(raw)] = raw.view(np.uint8)
            pos += 2 * len(raw)
            recs.append(('r', len(raw)))
        else:
            recs.append(('s',))
            spill.append(raw)
    if recs and all(r is not None and r[0] == 'c' for r in recs):
        # the common case as arrays: the parent places the whole task with one copy and a few vector operations
        counts = [len(r[3]) for r in recs]
        return ('B', np.array([r[1] for r in recs], dtype=np.int64), np.array([r[2] for r in recs], dtype=np.int64),
                np.repeat(np.arange(len(recs), dtype=np.int64), counts),
                np.array([off0 for r in recs for off0, _ in r[3]], dtype=np.int64),
                np.array([cs for r in recs for _, cs in r[3]], dtype=np.int64))
    return recs, spill, errs


class StoredBlock(object):
    """All reads of one worker task as stored chunks: samples and samples per chunk of every read; read, first sample
    and stored bytes of every chunk; the stored bytes back to back (a view into the task's slot)."""
    __slots__ = ('items', 'n', 'clen', 'chunk_read', 'off0', 'csize', 'data')

    def __init__(self, items, n, clen, chunk_read, off0, csize, data):
        self.items, self.n, self.clen, self.chunk_read, self.off0, self.csize, self.data = items, n, clen, chunk_read, off0, csize, data

    def reads(self):
        """the same, read by read"""
        starts = np.concatenate(([0], np.cumsum(self.csize)))
        first = np.searchsorted(self.chunk_read, np.arange(len(self.items) + 1))
        for k, item in enumerate(self.items):
            c0, c1 = int(first[k]), int(first[k + 1])
            yield item, StoredRead(int(self.n[k]), int(self.clen[k]),
                                   list(zip(self.off0[c0:c1].tolist(), self.csize[c0:c1].tolist())),
                                   self.data[int(starts[c0]):int(starts[c1])])


class _Staging(object):
    """Reads of one GPU batch back to back in ONE page-locked int16 buffer: decoded signals are copied in as they
    arrive and the batch goes to strique_detect_batch without another concatenation, at pinned-memory PCIe speed."""

    def __init__(self, capacity):
        from . import _lib
        self._lib = _lib
        self.buf = _lib.PinnedBuffer(capacity, np.int16)
        self.reset()

    def reset(self):
        self.meta, self.offsets, self.pos = [], [0], 0

    def fits(self, n):
        return self.pos + n <= len(self.buf.array)

    def grow(self, n):
        """an empty buffer too small for one read"""
        self.buf = self._lib.PinnedBuffer(n + (n >> 2), np.int16)

    def add(self, item, name, raw):
        n = len(raw)
        self.buf.array[self.pos:self.pos + n] = raw
        self.pos += n
        self.offsets.append(self.pos)
        self.meta.append((item, name))


class _StoredStaging(object):
    """Reads of one GPU batch as fast5 stores them: the zlib streams of their Signal chunks back to back in ONE
    page-locked byte buffer + one strique_inflate_chunk record per chunk.  The samples only ever exist on the device."""

    def __init__(self, capacity_bytes, small=False):
        from . import _lib
        self._lib = _lib
        self.small = small                       # the quarter-size buffer of a run's first batch (see detect_stream)
        self.buf = _lib.PinnedBuffer(capacity_bytes, np.uint8)
        self.reset()

    def reset(self):
        self.meta, self.offsets, self.pos, self._chunks, self._chunk_read = [], [0], 0, [], []

    def fits(self, nbytes):
        return self.pos + nbytes <= len(self.buf.array)

    def grow(self, nbytes):
        self.buf = self._lib.PinnedBuffer(nbytes + (nbytes >> 2), np.uint8)

    def _place(self, n, clen, chunk_read, off0, csize, data):
        """reads with n samples / clen samples per chunk; per chunk its read (0..), first sample, stored bytes"""
        nb = len(data)
        self.buf.array[self.pos:self.pos + nb] = data
        base = self.offsets[-1] + np.cumsum(n) - n                        # first sample of every read in the batch
        rec = np.zeros(len(csize), dtype=self._lib.INFLATE_CHUNK_DTYPE)
        rec['src_off'] = self.pos + np.cumsum(csize) - csize
        rec['dst_off'] = (base[chunk_read] + off0) * 2
        rec['src_len'] = csize
        rec['keep'] = np.minimum(clen[chunk_read], n[chunk_read] - off0) * 2
        rec['full'] = clen[chunk_read] * 2
        self._chunks.append(rec)
        self._chunk_read.append(chunk_read + len(self.meta))
        self.pos += nb
        self.offsets.extend((self.offsets[-1] + np.cumsum(n)).tolist())

    def add(self, item, name, sr):
        self._place(np.array([sr.n], dtype=np.int64), np.array([sr.clen], dtype=np.int64), np.zeros(len(sr.chunks), dtype=np.int64),
                    np.array([c[0] for c in sr.chunks], dtype=np.int64), np.array([c[1] for c in sr.chunks], dtype=np.int64), sr.data)
        self.meta.append((item, name))

    def add_block(self, blk):
        """a whole worker task whose reads all have exactly one target"""
        self._place(blk.n, blk.clen, blk.chunk_read, blk.off0, blk.csize, blk.data)
        self.meta.extend((item, item[3][0]) for item in blk.items)

    @property
    def chunks(self):
        return np.concatenate(self._chunks) if self._chunks else np.zeros(0, dtype=self._lib.INFLATE_CHUNK_DTYPE)

    @property
    def chunk_read(self):
        return np.concatenate(self._chunk_read) if self._chunk_read else np.zeros(0, dtype=np.int64)


class repeatDetector(object):
    """Multi-locus repeat detection over SAM records (scripts/STRique.py:624-705), batched."""

    # samples per signal base, for planning before a read is fetched (450 bases/s at 4 kHz)
    SAMPLES_PER_BASE = 9

    def __init__(self, repeat_config, model_file, fast5_index_file, mod_model_file=None, align_config=None,
                 HMM_config=None, device=0, io_threads=1, batch_samples=None, counter=None):
        from .counter import repeatCounter
        self.f5 = fast5.fast5Index(fast5_index_file)
        self.io_threads = max(int(io_threads), 1)
        self._io = None
        if counter is None:
            # the CUDA context (~0.5 s) comes up on a helper thread while the worker processes are started
            from . import _lib

            def warm():
                try:
                    _lib.default_context(device)
                except Exception:  # noqa: BLE001 - reported by the first real use
                    pass
            threading.Thread(target=warm, daemon=True).start()
        self._get_pool()                       # worker processes start (and load the index) meanwhile
        self.repeatCounter = counter or repeatCounter(model_file, mod_model_file=mod_model_file,
                                                      align_config=align_config, HMM_config=HMM_config, device=device)
        self.repeatLoci = defaultdict(list)
        self.repeat_config = repeat_config
        self.is_init = False
        # samples per GPU batch: ~8 k reads of 45 k samples (0.8 GB of int16 on the host and on the device)
        self.batch_samples = int(batch_samples or os.environ.get('STRIQUE_BATCH_SAMPLES', 384 << 20))
        # deflate-compressed Signal chunks are inflated on the GPU (strique_inflate_batch); STRIQUE_HOST_INFLATE=1:
        # by zlib on the I/O workers, like the reference's h5py
        self.gpu_inflate = (hasattr(self.repeatCounter, 'detect_deflated') and hasattr(self.f5, 'get_stored')
                            and not os.environ.get('STRIQUE_HOST_INFLATE'))

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

    def plan_iter(self, sam_lines):
        """Lazily: (input index, sam_record, strand, [target names]) for the records that hit a locus."""
        if not self.is_init:
            self.__init_hmm__()
        for idx, line in enumerate(sam_lines):
            sr = decode_sam(line)
            if not sr.QNAME:
                logger.log('Detector: Error parsing alignment \n{}'.format(line), 'error')
                continue
            names = self.intersect_target(sr)
            if not names:
                logger.log('Detector: No target for {}'.format(sr.QNAME), 'debug')
                continue
            yield (idx, sr, '-' if sr.FLAG & 0x10 else '+', names)

    def plan(self, sam_lines):
        return list(self.plan_iter(sam_lines))

    def _fetch(self, item):
        idx, sr, strand, names = item
        try:
            return item, self.f5.get_raw(sr.QNAME)
        except Exception as e:  # noqa: BLE001 - a bad read must not stop the others (S.py:764-768)
            logger.log('Detector: {}'.format(e), 'warning')
            return item, None

    def _rows(self, meta, results, rows):
        for ((idx, sr, strand, _), name), res in zip(meta, results):
            if res is not None:
                rows.append((idx, (sr.QNAME, name, strand) + tuple(res)))

    def _flush(self, batch, rows):
        """batch: list of ((item), raw, name) -- the generic path (any signal dtype, any counter)"""
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
        self._rows([(item, name) for item, _, name in batch], results, rows)

    def _flush_staged(self, st, rows):
        """st: _Staging -- the int16 path: no per-batch concatenation, pinned upload"""
        targets = [(name, item[2]) for item, name in st.meta]
        raw = st.buf.array[:st.pos]
        try:
            results = self.repeatCounter.detect_packed(targets, raw, st.offsets)
        except Exception as e:  # noqa: BLE001 - isolate the failing read
            logger.log('Detector: batch of {} reads failed ({}); retrying read by read'.format(len(targets), e), 'warning')
            results = []
            for k, (name, strand) in enumerate(targets):
                try:
                    results.append(self.repeatCounter.detect(name, raw[st.offsets[k]:st.offsets[k + 1]].copy(), strand))
                except Exception as e2:  # noqa: BLE001
                    logger.log('Detector: read failed: {}'.format(e2), 'warning')
                    results.append(None)
        self._rows(st.meta, results, rows)
        st.reset()

    def _flush_stored(self, st, rows):
        """st: _StoredStaging -- compressed chunks to the device, inflated there"""
        from . import _lib
        targets = [(name, item[2]) for item, name in st.meta]
        chunks, chunk_read = st.chunks, st.chunk_read
        try:
            results, status = self.repeatCounter.detect_deflated(targets, st.buf.array, st.pos, chunks, st.offsets)
            results = list(results)
            for c in np.nonzero(status)[0]:
                k = int(chunk_read[c])
                if results[k] is not None:
                    logger.log('Detector: damaged Signal chunk in read {} ({})'.format(
                        st.meta[k][0][1].QNAME, _lib.INFLATE_STATUS.get(int(status[c]), status[c])), 'warning')
                    results[k] = None
        except Exception as e:  # noqa: BLE001 - isolate the failing read (decoded on the host this time)
            logger.log('Detector: batch of {} reads failed ({}); retrying read by read'.format(len(targets), e), 'warning')
            results = []
            for (item, name) in st.meta:
                try:
                    results.append(self.repeatCounter.detect(name, self.f5.get_raw(item[1].QNAME), item[2]))
                except Exception as e2:  # noqa: BLE001
                    logger.log('Detector: read failed: {}'.format(e2), 'warning')
                    results.append(None)
        self._rows(st.meta, results, rows)
        st.reset()

    def _fetch_stored(self, item):
        """thread-pool twin of _worker_fetch_stored"""
        try:
            st = self.f5.get_stored(item[1].QNAME)
            if st[0] != 'chunks':
                return item, np.asarray(st[1])
            total = sum(cs for _, _, cs in st[4])
            data = np.empty(total, np.uint8)
            rec, _ = _stored_chunks_into(st, data, 0)
            return item, StoredRead(rec[1], rec[2], rec[3], data)
        except Exception as e:  # noqa: BLE001
            logger.log('Detector: {}'.format(e), 'warning')
            return item, None

    FETCH_CHUNK = 32                           # reads per task of a worker process

    def _get_pool(self):
        """--t worker processes reading fast5 (each loads the index itself; tasks of FETCH_CHUNK reads), or threads
        for --t 1 / a stub index / STRIQUE_IO_THREADS=1; created once and kept until close().  With `gpu_inflate` the
        workers only locate and copy the stored chunks of deflate-compressed reads; otherwise they decode them.
        -> (submit(items) -> future, result(future) -> [(item, int16 array | StoredRead | None)], chunk size)"""
        if getattr(self, '_io', None) is not None:
            return self._io[1:]
        index_file = getattr(self.f5, 'index_file', None)
        if self.io_threads > 1 and index_file and not os.environ.get('STRIQUE_IO_THREADS'):
            import multiprocessing as mp
            from multiprocessing import shared_memory
            pool = ProcessPoolExecutor(self.io_threads, mp_context=mp.get_context('spawn'), initializer=_worker_init,
                                       initargs=(index_file,))
            pool.submit(int)                   # (spawn starts every worker at the first submit)
            slots, free = [], deque()

            def submit(items):
                if not free:
                    shm = shared_memory.SharedMemory(create=True, size=SLOT_SAMPLES * 2)
                    slots.append(shm)
                    free.append(shm)
                shm = free.popleft()
                stored = getattr(self, 'gpu_inflate', False)
                fut = pool.submit(_worker_fetch_stored if stored else _worker_fetch, [it[1].QNAME for it in items], shm.name)
                fut.items, fut.shm, fut.stored = items, shm, stored
                return fut

            def result(fut):
                res = fut.result()
                fut.release = lambda: free.append(fut.shm)
                if res[0] == 'B':
                    nb = int(res[5].sum())
                    return StoredBlock(fut.items, res[1], res[2], res[3], res[4], res[5],
                                       np.frombuffer(fut.shm.buf, dtype=np.uint8)[:nb])
                recs, spill, errs = res
                for err in errs:
                    logger.log('Detector: {}'.format(err), 'warning')
                out, pos, k = [], 0, 0
                if fut.stored:
                    src = np.frombuffer(fut.shm.buf, dtype=np.uint8)
                    for item, rec in zip(fut.items, recs):
                        if rec is None:
                            out.append((item, None))
                        elif rec[0] == 'c':
                            nb = sum(cs for _, cs in rec[3])
                            out.append((item, StoredRead(rec[1], rec[2], rec[3], src[pos:pos + nb])))
                            pos += nb
                        elif rec[0] == 'r':
                            pos += pos & 1
                            out.append((item, src[pos:pos + 2 * rec[1]].view(np.int16)))
                            pos += 2 * rec[1]
                        else:
                            out.append((item, spill[k]))
                            k += 1
                else:
                    src = np.frombuffer(fut.shm.buf, dtype=np.int16)
                    for item, n in zip(fut.items, recs):
                        if n == -1:
                            out.append((item, None))
                        elif n == -2:
                            out.append((item, spill[k]))
                            k += 1
                        else:
                            out.append((item, src[pos:pos + n]))      # a view: the caller copies it into its batch buffer
                            pos += n
                return out

            def close_all():
                pool.shutdown(wait=True, cancel_futures=True)
                for shm in slots:
                    try:
                        shm.unlink()
                    except Exception:  # noqa: BLE001
                        pass
                    try:
                        shm.close()                 # (refuses while numpy views of the slot are alive; harmless)
                    except Exception:  # noqa: BLE001
                        pass
            self._io = (close_all, submit, result, self.FETCH_CHUNK)
            return self._io[1:]
        pool = ThreadPoolExecutor(self.io_threads)

        def submit_one(items):
            return pool.submit(self._fetch_stored if getattr(self, 'gpu_inflate', False) else self._fetch, items[0])
        self._io = ((lambda: pool.shutdown(wait=True, cancel_futures=True)), submit_one, (lambda fut: [fut.result()]), 1)
        return self._io[1:]

    def close(self):
        """stop the I/O workers and release their shared-memory slots (and the staging buffers kept between calls)"""
        self._kept_stored = []
        io, self._io = getattr(self, '_io', None), None
        if io is not None:
            io[0]()

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def detect_stream(self, work_iter, emit):
        """work_iter: items of plan_iter() (possibly one rank's share); emit(rows) is called once per GPU batch with
        that batch's (input index, row tuple) list, in input order.  Three things overlap: the I/O workers fetch at
        most one batch worth of (estimated) samples ahead -- fetched signals never pile up unbounded; this thread
        copies what they return into the staging buffers of batch k + 1; a GPU thread runs batch k (two sets of
        staging buffers, at most one batch in flight)."""
        pending = deque()                      # (future of a chunk of items, estimated samples)
        ahead = 0
        samples = 0
        can_stage = hasattr(self.repeatCounter, 'detect_packed')
        work_iter = iter(work_iter)
        exhausted = False
        submit, result, chunk = self._get_pool()

        class Buffers(object):                 # staging of one GPU batch
            def __init__(self):
                self.staged = None             # _Staging once the first int16 signal has arrived
                self.stored = None             # _StoredStaging once the first still-compressed read has arrived
                self.batch = []                # the generic path

            def any(self):
                return bool(self.batch or (self.staged is not None and self.staged.meta)
                            or (self.stored is not None and self.stored.meta))

        sets = [Buffers(), Buffers()]
        cur = 0
        clock = {'io': 0.0, 'stage': 0.0, 'gpu_wait': 0.0, 'alloc': 0.0}     # where this thread's time goes (log level info)
        gpu = ThreadPoolExecutor(1)
        in_flight = None                       # future of the batch on the GPU
        # Page-locking memory costs ~0.5 ms per MB, 0.3 s for one batch's buffer.  The first batch of a run is a
        # quarter batch in a quarter-size buffer -- the GPU starts early -- while the two full-size buffers are
        # allocated on a helper thread.
        full_cap = self.batch_samples * 3 // 2 + (16 << 20)
        alloc = ThreadPoolExecutor(1)
        spare = deque()
        # buffers of an earlier call of this detector (the multi-rank path calls once per chunk): no ramp-up then
        kept = [st for st in getattr(self, '_kept_stored', []) if len(st.buf.array) >= full_cap]
        self._kept_stored = []
        for k, st in enumerate(kept[:2]):
            st.reset()
            sets[k].stored = st

        def stored_for(b):
            t0 = time.time()
            try:
                return stored_for_(b)
            finally:
                clock['alloc'] += time.time() - t0

        def stored_for_(b):
            if b.stored is None or (b.stored.small and not b.stored.meta):
                if b.stored is None and not spare and not any(x.stored is not None for x in sets):
                    spare.append(alloc.submit(_StoredStaging, full_cap))
                    spare.append(alloc.submit(_StoredStaging, full_cap))
                    b.stored = _StoredStaging(full_cap // 4, small=True)
                elif spare:
                    b.stored = spare.popleft().result()
                elif b.stored is None:
                    b.stored = _StoredStaging(full_cap)
            return b.stored

        def limit(b):
            return self.batch_samples // 4 if (b.stored is not None and b.stored.small) else self.batch_samples

        def run(b):
            rows = []
            if b.stored is not None and b.stored.meta:
                self._flush_stored(b.stored, rows)
            if b.staged is not None and b.staged.meta:
                self._flush_staged(b.staged, rows)
            if b.batch:
                self._flush(b.batch, rows)
                b.batch = []
            rows.sort(key=lambda r: r[0])
            return rows

        def settle():
            nonlocal in_flight
            if in_flight is not None:
                fut, in_flight = in_flight, None
                t0 = time.time()
                rows = fut.result()
                clock['gpu_wait'] += time.time() - t0
                emit(rows)

        def flush():
            nonlocal cur, samples, in_flight
            settle()                           # the other set of buffers is free again
            in_flight = gpu.submit(run, sets[cur])
            cur ^= 1
            samples = 0

        try:
            while True:
                while not exhausted and (ahead < self.batch_samples or not pending):
                    items, est = [], 0
                    for item in work_iter:
                        items.append(item)
                        est += max(item[1].SEQ_LEN, 1) * self.SAMPLES_PER_BASE * len(item[3])
                        if len(items) >= chunk:
                            break
                    if not items:
                        exhausted = True
                        break
                    pending.append((submit(items), est))
                    ahead += est
                if not pending:
                    break
                fut, est = pending.popleft()
                ahead -= est
                t0 = time.time()
                res = result(fut)
                t1 = time.time()
                clock['io'] += t1 - t0
                if isinstance(res, StoredBlock):
                    if all(len(item[3]) == 1 for item in res.items):
                        b = sets[cur]
                        total = int(res.n.sum())
                        if b.stored is not None and b.stored.meta and (samples + total > limit(b) or
                                                                       not b.stored.fits(len(res.data))):
                            flush()
                            b = sets[cur]
                        stored_for(b)
                        if not b.stored.fits(len(res.data)):
                            b.stored.grow(len(res.data))
                        b.stored.add_block(res)
                        samples += total
                        fut.release()
                        continue
                    res = res.reads()
                for item, raw in res:
                    b = sets[cur]
                    if raw is None:
                        logger.log('Detector: No fast5 for ID {id}'.format(id=item[1].QNAME), 'warning')
                        continue
                    if isinstance(raw, StoredRead):
                        for name in item[3]:
                            stored_for(b)
                            if not b.stored.fits(len(raw.data)):
                                if b.stored.meta:
                                    flush()         # the I/O workers keep fetching the next batch meanwhile
                                    b = sets[cur]
                                    stored_for(b)
                                if not b.stored.fits(len(raw.data)):
                                    b.stored.grow(len(raw.data))
                            b.stored.add(item, name, raw)
                            samples += raw.n
                        if samples >= limit(b):
                            flush()
                        continue
                    raw = np.asarray(raw)
                    stage_it = can_stage and raw.dtype == np.int16
                    if not stage_it and raw.base is not None:
                        raw = raw.copy()                # (a view into a worker's slot, which is about to be reused)
                    for name in item[3]:
                        if stage_it:
                            if b.staged is None:
                                b.staged = _Staging(self.batch_samples + (8 << 20))
                            if not b.staged.fits(len(raw)):
                                if b.staged.meta:
                                    flush()
                                    b = sets[cur]
                                    if b.staged is None:
                                        b.staged = _Staging(self.batch_samples + (8 << 20))
                                if not b.staged.fits(len(raw)):
                                    b.staged.grow(len(raw))
                            b.staged.add(item, name, raw)
                        else:
                            b.batch.append((item, raw, name))
                        samples += len(raw)
                    if samples >= self.batch_samples:
                        flush()
                if hasattr(fut, 'release'):
                    fut.release()                       # the slot's signals have been copied out
            if sets[cur].any():
                flush()
            settle()
            logger.log('Detector: waited {io:.2f} s for the I/O workers, {gpu_wait:.2f} s for the GPU, {alloc:.2f} s for '
                       'staging buffers'.format(**clock), 'info')
        finally:
            gpu.shutdown(wait=True)
            alloc.shutdown(wait=True)
            # keep the full-size page-locked buffers for the next call
            keep = [b.stored for b in sets if b.stored is not None and not b.stored.small]
            for fut in spare:
                try:
                    keep.append(fut.result())
                except Exception:  # noqa: BLE001
                    pass
            self._kept_stored = keep[:2]

    def detect_records(self, work):
        """work: output of plan() (possibly one rank's share). -> list of (input index, row tuple)."""
        rows = []
        self.detect_stream(work, rows.extend)
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
        self.fp.flush()                        # rows reach the file batch by batch (S.py:720-724 appends per read)

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
        # under torchrun only rank 0 reads the SAM: a pipe on stdin reaches one process
        lines = None
        if rank == 0:
            lines = (line for line in (open(args.algn, 'r') if args.algn else sys.stdin) if not line.startswith('@'))
        run_count(rd, lines, args.out, rank, world)
        sharding.finalize()


def run_count(rd, lines, out, rank=0, world=1):
    """The body of `count`: SAM lines (rank 0's iterator; None on the other ranks) -> TSV rows, streamed batch by
    batch in input order.  One process: plan -> fetch ahead -> GPU batch -> append rows.  Several ranks: rank 0 plans
    chunks of about one batch per rank and broadcasts each; the ranks decode their cost-balanced shares and rank 0
    gathers and writes the chunk's rows before the next chunk is dealt."""
    t0 = time.time()
    n_rows = 0
    if not rd.is_init:
        rd.__init_hmm__()
    if world == 1:
        ow = outputWriter(out)

        def emit(rows):
            nonlocal n_rows
            n_rows += len(rows)
            ow.write_line([r for _, r in rows])
            logger.log('Main: {} rows after {:.2f} s'.format(n_rows, time.time() - t0), 'info')

        try:
            rd.detect_stream(rd.plan_iter(lines), emit)
        finally:
            rd.close()
        ow.close()
        return n_rows
    ow = outputWriter(out) if rank == 0 else None
    plan = rd.plan_iter(lines) if rank == 0 else None
    chunk_samples = rd.batch_samples * world
    try:
        while True:
            work = None
            if rank == 0:
                work, est = [], 0
                for item in plan:
                    work.append(item)
                    est += max(item[1].SEQ_LEN, 1) * rd.SAMPLES_PER_BASE * len(item[3])
                    if est >= chunk_samples:
                        break
                if not work:
                    work = None
            work = sharding.broadcast_object(work)
            if work is None:
                break
            # cost of a read ~ its length (2 flank alignments over the whole signal dominate)
            shards = sharding.lpt_partition([w[1].SEQ_LEN * len(w[3]) for w in work], world)
            rows = rd.detect_records([work[i] for i in shards[rank]])
            rows = sharding.gather_rows(rows)
            if rank == 0:
                n_rows += len(rows)
                ow.write_line([r for _, r in rows])
                logger.log('Main: {} rows after {:.2f} s'.format(n_rows, time.time() - t0), 'info')
    finally:
        rd.close()                               # (the I/O workers live across the chunks)
    if rank == 0:
        ow.close()
    logger.log('Main: rank {} done in {:.2f} s'.format(rank, time.time() - t0), 'info')
    return n_rows


def run():
    signal.signal(signal.SIGPIPE, signal.SIG_DFL)
    main()
