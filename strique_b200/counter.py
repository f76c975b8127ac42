"""`repeatCounter`: the per-read repeat detection of the reference, as a read-batching GPU driver.

Mirrors `repeatCounter` of the reference (scripts/STRique.py:505-618): same constructor
arguments, `add_target(name, repeat, prefix, suffix)` and `detect(name, raw_signal, strand)` with the
same 7-tuple result and the same exceptions, so the reference's tests run unchanged
(scripts/STRique_test.py:54-61).  New: `detect_batch`, which is what `detect` calls with one read --
all compute happens in libstrique_b200 (`strique_detect_batch`), there is no CPU path.
"""
from collections import namedtuple

import numpy as np

from . import _lib, hmm
from .pore_model import pore_model

_COMPLEMENT = {'A': 'T', 'C': 'G', 'G': 'C', 'T': 'A'}

ALIGN_DEFAULTS = {'dist_offset': 16.0, 'dist_min': 0.0, 'gap_open_h': -1.0, 'gap_open_v': -16.0,
                  'gap_extension_h': -1.0, 'gap_extension_v': -16.0, 'samples': 6}     # S.py:507-513

DetectRecord = namedtuple('DetectRecord', ['count', 'score_prefix', 'score_suffix', 'log_p', 'offset', 'ticks', 'mod',
                                           'prefix_begin', 'prefix_end', 'suffix_begin', 'suffix_end'])


class repeatCounter(object):
    def __init__(self, model_file, mod_model_file=None, align_config=None, HMM_config=None, device=0, context=None):
        cfg = dict(ALIGN_DEFAULTS)
        if align_config and isinstance(align_config, dict):
            cfg.update(align_config)
        self.align_config = cfg
        self.pm = pore_model(model_file)
        self.pm_mod = pore_model(mod_model_file) if mod_model_file else self.pm
        self.samples = cfg['samples']
        self.HMM_config = HMM_config
        self.targets = {}
        self.target_ids = {}
        self._ctx = context
        self._device = device
        self._cfg = None
        self.target_classifier = namedtuple('target_classifier', field_names=['prefix', 'suffix', 'prefix_ext',
                                                                             'suffix_ext', 'repeatHMM', 'modHMM'])

    # ---------------------------------------------------------------------------------------------
    @property
    def context(self):
        if self._ctx is None:
            self._ctx = _lib.default_context(self._device)
        return self._ctx

    @property
    def use_mod(self):
        return self.pm is not self.pm_mod            # object identity, like `self.pm != self.pm_mod` (S.py:605)

    def __reverse_complement__(self, sequence):
        return ''.join(_COMPLEMENT.get(base, base) for base in reversed(sequence))

    def _detect_config(self):
        if self._cfg is None:
            c = self.align_config
            lo = min(self.pm.model_min, self.pm_mod.model_min)
            hi = max(self.pm.model_max, self.pm_mod.model_max)
            self._cfg = _lib.DetectConfig(
                _lib.AlignParams(c['gap_open_h'], c['gap_open_v'], c['gap_extension_h'], c['gap_extension_v'],
                                 c['dist_offset'], c['dist_min']),
                int(self.samples), 1 if self.use_mod else 0, _lib.PoreConstants(*self.pm.minmax_constants()), lo, hi)
        return self._cfg

    def _make_classifier(self, repeat, prefix, suffix, prefix_ext, suffix_ext):
        """One strand's classifier: flank templates + the two compiled HMMs registered on the device."""
        graph, count_offset = hmm.flanked_repeat_graph(repeat, prefix, suffix, self.pm, self.HMM_config)
        count_hmm = hmm.compile_graph(graph)
        count_hmm.count_offset = count_offset
        mod_graph, _, _ = hmm.repeat_mod_graph(repeat, self.pm, self.pm_mod, config=self.HMM_config)
        mod_hmm = hmm.compile_graph(mod_graph)
        tc = self.target_classifier(self.pm.generate_signal(prefix, samples=self.samples),
                                    self.pm.generate_signal(suffix, samples=self.samples),
                                    self.pm.generate_signal(prefix_ext, samples=self.samples),
                                    self.pm.generate_signal(suffix_ext, samples=self.samples),
                                    count_hmm, mod_hmm)
        ctx = self.context
        count_id = ctx.hmm_create(count_hmm)
        mod_id = ctx.hmm_create(mod_hmm) if self.use_mod else -1
        tid = ctx.target_create(self.pm.kmer_means(prefix_ext), self.pm.kmer_means(suffix_ext),
                                len(tc.prefix_ext) - len(tc.prefix), len(tc.suffix_ext) - len(tc.suffix),
                                count_id, mod_id, count_offset)
        return tc, tid

    def add_target(self, target_name, repeat, prefix, suffix):
        if target_name in self.targets:
            raise ValueError("RepeatCounter: Target with name " + str(target_name) + " already defined.")
        prefix_ext = prefix.upper()
        prefix = prefix[-50:].upper()
        suffix_ext = suffix.upper()
        suffix = suffix[:50].upper()
        repeat = repeat.upper()
        # the reference aligns flanks of any length; the kernels here hold a flank in the registers of one warp
        for flank in (prefix_ext, suffix_ext):
            levels = len(flank) - self.pm.kmer + 1
            if levels < 1 or not _lib.load().strique_align_supported(levels, int(self.samples)):
                raise ValueError("RepeatCounter: flank of {} nt x {} samples does not fit the alignment kernels "
                                 "(at most 2048 flank samples = (nt - k + 1) x samples).".format(len(flank), self.samples))
        rc = self.__reverse_complement__
        tc_plus, id_plus = self._make_classifier(repeat, prefix, suffix, prefix_ext, suffix_ext)
        tc_minus, id_minus = self._make_classifier(rc(repeat), rc(suffix), rc(prefix), rc(suffix_ext), rc(prefix_ext))
        self.targets[target_name] = (tc_plus, tc_minus)
        self.target_ids[target_name] = (id_plus, id_minus)

    def _target_id(self, target_name, strand):
        if target_name not in self.targets:
            raise ValueError("RepeatCounter: Target with name " + str(target_name) + " not defined.")
        if strand == '+':
            return self.target_ids[target_name][0]
        if strand == '-':
            return self.target_ids[target_name][1]
        raise ValueError("RepeatCounter: Strand must be + or -.")

    # ---------------------------------------------------------------------------------------------
    def detect_batch(self, items, details=False):
        """items: iterable of (target_name, raw_signal, strand).  Returns one 7-tuple per item, in
        order: (n, score_prefix, score_suffix, log_p, offset, ticks, mod_pattern) (S.py:616);
        with details=True a DetectRecord carrying the flank coordinates as well."""
        items = list(items)
        if not items:
            return []
        tids = np.array([self._target_id(name, strand) for name, _, strand in items], dtype=np.int32)
        raw, off, kind = _lib.Context._pack_raw([np.asarray(sig) for _, sig, _ in items])
        return self._detect_packed(tids, raw, off, kind, details)

    def detect_packed(self, targets, raw, offsets, details=False):
        """targets: [(target_name, strand)] per read; raw: ONE int16 (or float64) array holding the reads back to back
        (e.g. a _lib.PinnedBuffer the fast5 signals were decoded into), offsets: [n + 1] sample offsets.  Same result
        as detect_batch without the concatenation."""
        if not targets:
            return []
        tids = np.array([self._target_id(name, strand) for name, strand in targets], dtype=np.int32)
        raw = np.asarray(raw)
        if raw.dtype not in (np.int16, np.float64):
            raise ValueError('RepeatCounter: packed signals must be int16 or float64')
        return self._detect_packed(tids, raw, np.ascontiguousarray(offsets, dtype=np.int64), 0 if raw.dtype == np.int16 else 1,
                                   details)

    def detect_deflated(self, targets, comp, comp_bytes, chunks, offsets, details=False):
        """Reads still compressed as fast5 stores them: `comp` holds the zlib streams of the Signal chunks (e.g. a
        uint8 _lib.PinnedBuffer), `chunks` (_lib.INFLATE_CHUNK_DTYPE) says where each chunk's samples belong in the
        batch, offsets: [n + 1] sample offsets of the reads.  The chunks are inflated on the device
        (strique_inflate_batch) and never exist on the host.  -> (rows like detect_packed, per-chunk status)"""
        if not targets:
            return [], np.zeros(0, np.int32)
        tids = np.array([self._target_id(name, strand) for name, strand in targets], dtype=np.int32)
        off = np.ascontiguousarray(offsets, dtype=np.int64)
        dev, status = self.context.inflate_batch(comp, comp_bytes, chunks, int(off[-1]) * 2)
        return self._detect_packed(tids, dev, off, 0, details, memspace=_lib.DEVICE), status

    def _detect_packed(self, tids, raw, off, kind, details, memspace=_lib.HOST):
        res, mod = self.context.detect_batch(self._detect_config(), raw, off, kind, tids, memspace=memspace)
        # column-wise to Python objects: per-row access to a structured array costs ~10 us per read, as much as the
        # GPU needs for the read
        ran = res['hmm_ran'].astype(bool)
        # the reference leaves n and p at the int 0 when the HMM stage is skipped or finds no path
        counts = np.where(ran, res['count'], 0).tolist()
        logp = res['log_p'].tolist()
        cols = [res[f].tolist() for f in ('score_prefix', 'score_suffix', 'offset', 'ticks', 'mod_off', 'mod_len')]
        patterns = None
        if (res['mod_len'] >= 0).any():
            patterns = mod.tobytes()
        out = []
        for k, (sp, ss, offset, ticks, mo, ml) in enumerate(zip(*cols)):
            pattern = patterns[mo:mo + ml].decode('ascii') if ml >= 0 else '-'
            out.append((counts[k], sp, ss, logp[k] if ran[k] else 0, offset, ticks, pattern))
        if details:
            extra = [res[f].tolist() for f in ('prefix_begin', 'prefix_end', 'suffix_begin', 'suffix_end')]
            out = [DetectRecord(*rec, *e) for rec, e in zip(out, zip(*extra))]
        return out

    def detect(self, target_name, raw_signal, strand):
        return self.detect_batch([(target_name, raw_signal, strand)])[0]
