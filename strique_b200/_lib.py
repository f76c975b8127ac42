"""ctypes binding of libstrique_b200.so (the C ABI declared in include/strique_b200.h).

There is no CPU fallback: importing works anywhere (so host logic can be tested without a GPU),
but creating a `Context` without the compiled library or without a B200 raises `StriqueError`.
"""
import ctypes
import threading
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# STRIQUE_LIB: another build of the same library (kernel A/B experiments, tools/gpu_ab.sh)
LIB_PATH = os.environ.get('STRIQUE_LIB') or os.path.join(_HERE, 'libstrique_b200.so')

HOST, DEVICE = 0, 1
ENOSPC = -5      # STRIQUE_ENOSPC: an output buffer was too small


class StriqueError(RuntimeError):
    pass


class AlignParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_float) for n in
                ('gap_open_h', 'gap_open_v', 'gap_extension_h', 'gap_extension_v', 'dist_offset', 'dist_min')]


class PoreConstants(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in ('m5_mod', 'm95_mod', 'model_min', 'model_max')]


class HmmDesc(ctypes.Structure):
    _fields_ = [('n_emit', ctypes.c_int32), ('n_chain', ctypes.c_int32),
                ('in_ptr', ctypes.c_void_p), ('in_src', ctypes.c_void_p), ('in_logw', ctypes.c_void_p),
                ('emit_kind', ctypes.c_void_p), ('emit_a', ctypes.c_void_p), ('emit_b', ctypes.c_void_p),
                ('emit_flags', ctypes.c_void_p), ('chain_pred_logw', ctypes.c_void_p),
                ('chain_in_ptr', ctypes.c_void_p), ('chain_in_src', ctypes.c_void_p), ('chain_in_logw', ctypes.c_void_p),
                ('n_end', ctypes.c_int32), ('end_src', ctypes.c_void_p), ('end_logw', ctypes.c_void_p),
                ('emit_pos', ctypes.c_void_p), ('emit_slot', ctypes.c_void_p), ('chain_pos', ctypes.c_void_p)]


def hmm_desc(c):
    """strique_hmm_desc of a CompiledHMM -> (desc, arrays that must stay alive while it is used)."""
    keep = [np.ascontiguousarray(a) for a in (c.in_ptr, c.in_src, c.in_logw, c.emit_kind, c.emit_a, c.emit_b,
                                              c.emit_flags, c.chain_pred_logw, c.chain_in_ptr, c.chain_in_src,
                                              c.chain_in_logw, c.end_src, c.end_logw)]
    hints = [None, None, None]
    if getattr(c, 'emit_pos', None) is not None:
        hints = [np.ascontiguousarray(c.emit_pos, dtype=np.int32), np.ascontiguousarray(c.emit_slot, dtype=np.uint8),
                 np.ascontiguousarray(c.chain_pos, dtype=np.int32)]
        keep += hints
    d = HmmDesc(c.n_emit, c.n_chain, *[a.ctypes.data for a in keep[:11]], len(c.end_src),
                keep[11].ctypes.data, keep[12].ctypes.data,
                *[(a.ctypes.data if a is not None else None) for a in hints])
    return d, keep


class TargetDesc(ctypes.Structure):
    _fields_ = [('prefix_levels', ctypes.c_void_p), ('n_prefix_levels', ctypes.c_int32),
                ('suffix_levels', ctypes.c_void_p), ('n_suffix_levels', ctypes.c_int32),
                ('pre_trim', ctypes.c_int32), ('post_trim', ctypes.c_int32),
                ('count_model', ctypes.c_int32), ('mod_model', ctypes.c_int32), ('count_offset', ctypes.c_int32)]


class DetectConfig(ctypes.Structure):
    _fields_ = [('align', AlignParams), ('samples', ctypes.c_int32), ('use_mod', ctypes.c_int32),
                ('pore', PoreConstants), ('mod_clip_lo', ctypes.c_double), ('mod_clip_hi', ctypes.c_double)]


VITERBI_RESULT_DTYPE = np.dtype([('logp', np.float64), ('n_count', np.int32), ('t_first', np.int32),
                                 ('t_last', np.int32), ('pattern_len', np.int32), ('status', np.int32),
                                 ('reserved', np.int32)])
CONDITION_STATS_DTYPE = np.dtype([(n, np.float64) for n in
                                  ('flt_median', 'flt_mad', 'flt_c1', 'flt_c2', 'raw_c1', 'raw_c2', 'u8_c1', 'u8_c2',
                                   'status', 'r0', 'r1', 'r2')])
DETECT_RESULT_DTYPE = np.dtype([('score_prefix', np.float64), ('score_suffix', np.float64), ('log_p', np.float64),
                                ('count', np.int32), ('offset', np.int32), ('ticks', np.int32),
                                ('prefix_begin', np.int32), ('prefix_end', np.int32), ('suffix_begin', np.int32),
                                ('suffix_end', np.int32), ('hmm_ran', np.int32), ('mod_len', np.int32),
                                ('status', np.int32), ('mod_off', np.int64)], align=True)
STAGES = ('condition', 'align_table', 'align_scan', 'align_trace', 'viterbi_count', 'viterbi_mod', 'h2d')

ALIGN_RESULT_DTYPE = np.dtype([('score', np.float32), ('best_j', np.int32), ('begin0', np.int32), ('end0', np.int32),
                               ('begin_trim', np.int32), ('end_trim', np.int32), ('n_blocks', np.int32),
                               ('status', np.int32)])

_lib = None


def load():
    """Load the shared library (once). Raises StriqueError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise StriqueError('libstrique_b200.so is not built: run ./build.sh (or __graft_entry__.build()); '
                           'there is no CPU fallback for the CUDA hot path')
    lib = ctypes.CDLL(LIB_PATH)
    c_void_p, c_int, c_int64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    lib.strique_ctx_create.restype = c_int
    lib.strique_ctx_create.argtypes = [c_int, ctypes.POINTER(c_void_p)]
    lib.strique_ctx_destroy.restype = None
    lib.strique_ctx_destroy.argtypes = [c_void_p]
    lib.strique_last_error.restype = ctypes.c_char_p
    lib.strique_last_error.argtypes = [c_void_p]
    lib.strique_version.restype = c_int
    lib.strique_launch_count.restype = c_int64
    lib.strique_launch_count.argtypes = [c_void_p]
    lib.strique_ctx_stream.restype = c_void_p
    lib.strique_ctx_stream.argtypes = [c_void_p]
    lib.strique_ctx_synchronize.restype = c_int
    lib.strique_ctx_synchronize.argtypes = [c_void_p]
    lib.strique_last_align_cells.restype = c_int64
    lib.strique_last_align_cells.argtypes = [c_void_p]
    lib.strique_last_scan_ms.restype = ctypes.c_float
    lib.strique_last_scan_ms.argtypes = [c_void_p]
    lib.strique_align_batch.restype = c_int
    lib.strique_align_batch.argtypes = [c_void_p, ctypes.POINTER(AlignParams),
                                        c_int, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                        c_int, c_void_p, c_void_p, c_int,
                                        c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_int, c_void_p, c_void_p, c_int64]
    lib.strique_last_viterbi_edges.restype = c_int64
    lib.strique_last_viterbi_edges.argtypes = [c_void_p]
    lib.strique_last_viterbi_fixed.restype = c_int64
    lib.strique_last_viterbi_fixed.argtypes = [c_void_p]
    lib.strique_last_viterbi_declined.restype = c_int64
    lib.strique_last_viterbi_declined.argtypes = [c_void_p]
    lib.strique_set_viterbi_exact.restype = c_int
    lib.strique_set_viterbi_exact.argtypes = [c_void_p, c_int]
    lib.strique_inflate_batch.restype = c_int
    lib.strique_inflate_batch.argtypes = [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int, c_int64, c_void_p,
                                          ctypes.POINTER(c_void_p)]
    lib.strique_align_supported.restype = c_int
    lib.strique_align_supported.argtypes = [c_int, c_int]
    lib.strique_host_alloc.restype = c_void_p
    lib.strique_host_alloc.argtypes = [ctypes.c_size_t]
    lib.strique_host_free.restype = None
    lib.strique_host_free.argtypes = [c_void_p]
    lib.strique_last_mod_bytes.restype = c_int64
    lib.strique_last_mod_bytes.argtypes = [c_void_p]
    lib.strique_last_stage_ms.restype = ctypes.c_float
    lib.strique_last_stage_ms.argtypes = [c_void_p, c_int]
    lib.strique_hmm_create.restype = c_int
    lib.strique_hmm_create.argtypes = [c_void_p, ctypes.POINTER(HmmDesc), ctypes.POINTER(ctypes.c_int32)]
    lib.strique_hmm_kernel_shape.restype = c_int
    lib.strique_hmm_kernel_shape.argtypes = [c_void_p, ctypes.c_int32]
    lib.strique_viterbi_batch.restype = c_int
    lib.strique_viterbi_batch.argtypes = [c_void_p, ctypes.c_int32, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                          c_void_p, c_void_p]
    lib.strique_condition_batch.restype = c_int
    lib.strique_condition_batch.argtypes = [c_void_p, ctypes.POINTER(PoreConstants), c_int, c_void_p, c_int, c_void_p,
                                            c_int, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.strique_target_create.restype = c_int
    lib.strique_target_create.argtypes = [c_void_p, ctypes.POINTER(TargetDesc), ctypes.POINTER(ctypes.c_int32)]
    lib.strique_detect_batch.restype = c_int
    lib.strique_detect_batch.argtypes = [c_void_p, ctypes.POINTER(DetectConfig), c_int, c_void_p, c_int, c_void_p,
                                         c_void_p, c_int, c_void_p, c_void_p, c_int64]
    _lib = lib
    return lib


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, int):
        return ctypes.c_void_p(a)
    return ctypes.c_void_p(a.ctypes.data)


class Context:
    """One CUDA context/stream + scratch arena on one device (strique_ctx)."""

    def __init__(self, device=0):
        self.lib = load()
        h = ctypes.c_void_p()
        rc = self.lib.strique_ctx_create(int(device), ctypes.byref(h))
        if rc != 0:
            raise StriqueError('strique_ctx_create failed ({}): {}'.format(
                rc, self.lib.strique_last_error(None).decode()))
        self.handle = h
        self.device = device

    def close(self):
        if getattr(self, 'handle', None):
            self.lib.strique_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def check(self, rc, what):
        if rc != 0:
            raise StriqueError('{} failed ({}): {}'.format(what, rc, self.lib.strique_last_error(self.handle).decode()))

    @property
    def launches(self):
        return int(self.lib.strique_launch_count(self.handle))

    @property
    def stream(self):
        return self.lib.strique_ctx_stream(self.handle)

    def synchronize(self):
        self.check(self.lib.strique_ctx_synchronize(self.handle), 'strique_ctx_synchronize')

    @property
    def last_align_cells(self):
        return int(self.lib.strique_last_align_cells(self.handle))

    @property
    def last_scan_ms(self):
        return float(self.lib.strique_last_scan_ms(self.handle))

    # -- boundary #1 ------------------------------------------------------------------------------
    def align_batch(self, params, codes, sig_offsets, code_values, flank_levels, flank_offsets, samples,
                    task_signal, task_flank, task_pre_trim, task_post_trim, want_rows=False, memspace=HOST,
                    n_code_values=None, code_bytes=None):
        """Thin wrapper over strique_align_batch. With memspace=HOST all arrays are numpy arrays;
        with DEVICE `codes`, `code_values`, `flank_levels` are integer device addresses."""
        sig_offsets = np.ascontiguousarray(sig_offsets, dtype=np.int64)
        flank_offsets = np.ascontiguousarray(flank_offsets, dtype=np.int32)
        task_signal = np.ascontiguousarray(task_signal, dtype=np.int32)
        task_flank = np.ascontiguousarray(task_flank, dtype=np.int32)
        task_pre_trim = np.ascontiguousarray(task_pre_trim, dtype=np.int32)
        task_post_trim = np.ascontiguousarray(task_post_trim, dtype=np.int32)
        n_signals, n_flanks, n_tasks = len(sig_offsets) - 1, len(flank_offsets) - 1, len(task_signal)
        if memspace == HOST:
            codes = np.ascontiguousarray(codes)
            if codes.dtype not in (np.uint8, np.uint16):
                raise ValueError('codes must be uint8 or uint16')
            code_bytes = codes.dtype.itemsize
            code_values = np.ascontiguousarray(code_values, dtype=np.float32).reshape(n_signals, -1)
            n_code_values = code_values.shape[1]
            flank_levels = np.ascontiguousarray(flank_levels, dtype=np.float32)
        results = np.zeros(n_tasks, dtype=ALIGN_RESULT_DTYPE)
        rows, stride = None, 0
        if want_rows:
            stride = int(np.max(np.diff(flank_offsets))) * int(samples) if n_flanks else 0
            rows = np.zeros((n_tasks, max(stride, 1)), dtype=np.int32)
        p = params if isinstance(params, AlignParams) else AlignParams(*params)
        rc = self.lib.strique_align_batch(self.handle, ctypes.byref(p), n_signals, _ptr(codes), code_bytes,
                                          _ptr(sig_offsets), _ptr(code_values), n_code_values, n_flanks,
                                          _ptr(flank_levels), _ptr(flank_offsets), int(samples), n_tasks,
                                          _ptr(task_signal), _ptr(task_flank), _ptr(task_pre_trim),
                                          _ptr(task_post_trim), memspace, _ptr(results), _ptr(rows), stride)
        self.check(rc, 'strique_align_batch')
        return (results, rows) if want_rows else results


    # -- boundary #2 ------------------------------------------------------------------------------
    def hmm_create(self, c):
        """Register a compiled HMM (strique_b200.hmm.CompiledHMM) -> model id."""
        d, keep = hmm_desc(c)
        mid = ctypes.c_int32(-1)
        self.check(self.lib.strique_hmm_create(self.handle, ctypes.byref(d), ctypes.byref(mid)), 'strique_hmm_create')
        return mid.value

    def hmm_kernel_shape(self, model_id):
        """0: generic Viterbi kernel; 4000: profile kernel; else team-kernel shape wps*1000 + nh*100 + nl*10 + qc."""
        return int(self.lib.strique_hmm_kernel_shape(self.handle, model_id))

    def viterbi_batch(self, model_id, sequences, want_path=False):
        """sequences: list of float64 vectors -> (results, patterns[, paths])."""
        seqs = [np.ascontiguousarray(x, dtype=np.float64) for x in sequences]
        off = np.zeros(len(seqs) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(x) for x in seqs])
        x = np.concatenate(seqs) if seqs else np.zeros(0)
        res = np.zeros(len(seqs), dtype=VITERBI_RESULT_DTYPE)
        pat = np.zeros(max(1, len(x)), dtype=np.uint8)
        path = np.zeros(max(1, len(x)), dtype=np.uint16) if want_path else None
        self.check(self.lib.strique_viterbi_batch(self.handle, model_id, len(seqs), _ptr(x), _ptr(off), HOST,
                                                  _ptr(res), _ptr(pat), _ptr(path)), 'strique_viterbi_batch')
        patterns = []
        for k in range(len(seqs)):
            n = int(res['pattern_len'][k])
            patterns.append(pat[off[k + 1] - n:off[k + 1]].tobytes().decode('ascii') if res['status'][k] == 0 else None)
        if want_path:
            return res, patterns, [path[off[k]:off[k + 1]].copy() for k in range(len(seqs))]
        return res, patterns

    # -- conditioning ------------------------------------------------------------------------------
    @staticmethod
    def _pack_raw(signals):
        sigs = [np.asarray(s) for s in signals]
        kind = 0 if all(s.dtype == np.int16 for s in sigs) else 1
        dt = np.int16 if kind == 0 else np.float64
        off = np.zeros(len(sigs) + 1, dtype=np.int64)
        off[1:] = np.cumsum([len(s) for s in sigs])
        raw = np.concatenate([s.astype(dt, copy=False) for s in sigs]) if sigs else np.zeros(0, dt)
        return np.ascontiguousarray(raw), off, kind

    def condition_batch(self, pore, signals, want_raw_stats=False):
        raw, off, kind = self._pack_raw(signals)
        n = len(off) - 1
        flt = np.zeros_like(raw)
        codes = np.zeros(len(raw), dtype=np.uint16)
        vals = np.zeros((n, 256), dtype=np.float32)
        stats = np.zeros(n, dtype=CONDITION_STATS_DTYPE)
        pc = PoreConstants(*pore)
        self.check(self.lib.strique_condition_batch(self.handle, ctypes.byref(pc), n, _ptr(raw), kind, _ptr(off),
                                                    1 if want_raw_stats else 0, _ptr(flt), _ptr(codes), _ptr(vals),
                                                    _ptr(stats)), 'strique_condition_batch')
        return flt, codes, vals, stats, off

    # -- whole path ----------------------------------------------------------------------------------
    def target_create(self, prefix_levels, suffix_levels, pre_trim, post_trim, count_model, mod_model, count_offset):
        pl = np.ascontiguousarray(prefix_levels, dtype=np.float32)
        sl = np.ascontiguousarray(suffix_levels, dtype=np.float32)
        d = TargetDesc(pl.ctypes.data, len(pl), sl.ctypes.data, len(sl), pre_trim, post_trim, count_model, mod_model,
                       count_offset)
        tid = ctypes.c_int32(-1)
        self.check(self.lib.strique_target_create(self.handle, ctypes.byref(d), ctypes.byref(tid)), 'strique_target_create')
        return tid.value

    def detect_batch(self, cfg, raw, raw_offsets, raw_kind, read_target, memspace=HOST):
        raw_offsets = np.ascontiguousarray(raw_offsets, dtype=np.int64)
        read_target = np.ascontiguousarray(read_target, dtype=np.int32)
        n = len(read_target)
        res = np.zeros(n, dtype=DETECT_RESULT_DTYPE)
        # one pattern character per repeat pass; a pass decodes at least ~3 samples per k-mer of the unit, so this is
        # generous for real reads -- and when it is not, the library says how much it needs (STRIQUE_ENOSPC)
        cap = int(raw_offsets[-1] // 8 + 64 * n + 64) if cfg.use_mod else 0
        for attempt in (0, 1):
            mod = np.zeros(max(cap, 1), dtype=np.uint8)
            rc = self.lib.strique_detect_batch(self.handle, ctypes.byref(cfg), n, _ptr(raw), raw_kind, _ptr(raw_offsets),
                                               _ptr(read_target), memspace, _ptr(res), _ptr(mod), cap)
            if rc == ENOSPC and attempt == 0:
                cap = int(self.lib.strique_last_mod_bytes(self.handle)) + 64
                continue
            self.check(rc, 'strique_detect_batch')
            break
        return res, mod

    def inflate_batch(self, comp, comp_bytes, chunks, out_bytes, memspace=HOST):
        """fast5 Signal chunks (zlib streams packed in `comp`) -> samples in a device buffer of the context.
        chunks: INFLATE_CHUNK_DTYPE array.  -> (device pointer, per-chunk status)"""
        chunks = np.ascontiguousarray(chunks, dtype=INFLATE_CHUNK_DTYPE)
        status = np.zeros(max(len(chunks), 1), dtype=np.int32)
        out = ctypes.c_void_p()
        self.check(self.lib.strique_inflate_batch(self.handle, _ptr(comp), int(comp_bytes), memspace, _ptr(chunks), len(chunks),
                                                  int(out_bytes), _ptr(status), ctypes.byref(out)), 'strique_inflate_batch')
        return int(out.value or 0), status[:len(chunks)]

    def stage_ms(self):
        return {name: float(self.lib.strique_last_stage_ms(self.handle, i)) for i, name in enumerate(STAGES)}

    @property
    def last_viterbi_edges(self):
        return int(self.lib.strique_last_viterbi_edges(self.handle))

    def set_viterbi_exact(self, exact):
        """True: profile models are decoded by the float64 kernel only (pomegranate's arithmetic); False (default):
        by the fixed-point kernel, with float64 for the sequences it declines."""
        self.check(self.lib.strique_set_viterbi_exact(self.handle, 1 if exact else 0), 'strique_set_viterbi_exact')

    @property
    def last_viterbi_fixed(self):
        """(sequences decoded in fixed point, sequences handed on to the float64 kernel) of the last call"""
        return (int(self.lib.strique_last_viterbi_fixed(self.handle)), int(self.lib.strique_last_viterbi_declined(self.handle)))


INFLATE_CHUNK_DTYPE = np.dtype([('src_off', np.int64), ('dst_off', np.int64), ('src_len', np.int32), ('keep', np.int32),
                                ('full', np.int32), ('reserved', np.int32)])
INFLATE_STATUS = {1: 'not a zlib stream', 2: 'bad block header', 3: 'bad Huffman code', 4: 'more output than the chunk holds',
                  5: 'match before the start of the chunk', 6: 'stream longer than its stored size', 7: 'Adler-32 mismatch',
                  8: 'chunk shorter than the dataset needs'}


class PinnedBuffer(object):
    """numpy view over page-locked host memory (strique_host_alloc); freed with the object."""

    def __init__(self, n, dtype=np.int16):
        self.lib = load()
        self.nbytes = int(n) * np.dtype(dtype).itemsize
        self.ptr = self.lib.strique_host_alloc(self.nbytes)
        if not self.ptr:
            raise StriqueError('strique_host_alloc of {} bytes failed'.format(self.nbytes))
        self.array = np.frombuffer((ctypes.c_char * self.nbytes).from_address(self.ptr), dtype=dtype)

    def __del__(self):
        try:
            if getattr(self, 'ptr', None):
                self.array = None
                self.lib.strique_host_free(self.ptr)
                self.ptr = None
        except Exception:  # noqa: BLE001
            pass


_default_ctx = {}


_default_ctx_lock = threading.Lock()


def default_context(device=0):
    with _default_ctx_lock:                      # (the CLI creates it on a helper thread while its I/O workers start)
        if device not in _default_ctx:
            _default_ctx[device] = Context(device)
        return _default_ctx[device]
