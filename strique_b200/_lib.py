"""ctypes binding of libstrique_b200.so (the C ABI declared in include/strique_b200.h).

There is no CPU fallback: importing works anywhere (so host logic can be tested without a GPU),
but creating a `Context` without the compiled library or without a B200 raises `StriqueError`.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libstrique_b200.so')

HOST, DEVICE = 0, 1


class StriqueError(RuntimeError):
    pass


class AlignParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_float) for n in
                ('gap_open_h', 'gap_open_v', 'gap_extension_h', 'gap_extension_v', 'dist_offset', 'dist_min')]


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


_default_ctx = {}


def default_context(device=0):
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
