"""ORACLE (test infrastructure, NOT product code).

Minimal restatement of the parts of pomegranate 0.10.0 that giesselmann/STRique's hot path uses
(`HiddenMarkovModel.add_state/add_states/add_transition/add_model/bake(merge='All')/viterbi`,
`State`, `NormalDistribution`, `UniformDistribution`; call sites scripts/STRique.py:201-500,
pinned by requirements.txt:10-11 together with networkx<2.0, requirements.txt:5).

pomegranate is NOT vendored under /root/reference and cannot be installed here (no network), so
this follows the published behaviour of that version as recalled in SURVEY.md App. C:
  bake(merge='All'): (1) repeatedly drop non start/end states without in- or out-edges,
  (2) renormalise out-edges of every state whose probabilities do not sum to 1 (8 decimals),
  (3) merge silent states having a probability-1 out-edge into the edge's target (never the start,
  never into the end), (4) order states [emitting sorted by name | silent, topological],
  (5) store in-edges as CSR.  Decoding is oracle/viterbi_oracle.c.
PARITY UNPINNED for float outputs (no reference test pins log_p); integer repeat counts are pinned
by the assertions of scripts/STRique_test.py, replayed in tests/test_oracle_pipeline.py.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.
"""
import ctypes
import math
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class NormalDistribution:
    kind = 0

    def __init__(self, mean, std):
        self.a, self.b = float(mean), float(std)


class UniformDistribution:
    kind = 1

    def __init__(self, lo, hi):
        self.a, self.b = float(lo), float(hi)


class State:
    def __init__(self, distribution, name):
        self.distribution = distribution
        self.name = name

    def is_silent(self):
        return self.distribution is None


class _OracleHMM(ctypes.Structure):
    _fields_ = [('n_states', ctypes.c_int32), ('silent_start', ctypes.c_int32),
                ('start_index', ctypes.c_int32), ('end_index', ctypes.c_int32),
                ('in_ptr', ctypes.c_void_p), ('in_src', ctypes.c_void_p), ('in_logp', ctypes.c_void_p),
                ('dist_kind', ctypes.c_void_p), ('dist_a', ctypes.c_void_p), ('dist_b', ctypes.c_void_p)]


_lib = None


def _liboracle():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, 'liboracle.so')
        if not os.path.exists(path):
            import subprocess
            subprocess.check_call(['make', '-C', _HERE, 'liboracle.so'])
        _lib = ctypes.CDLL(path)
        _lib.strique_oracle_viterbi.restype = ctypes.c_int64
        _lib.strique_oracle_viterbi.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64,
                                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64]
        _lib.strique_oracle_viterbi_margin.restype = ctypes.c_int64
        _lib.strique_oracle_viterbi_margin.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p,
                                                        ctypes.c_void_p, ctypes.c_int64, ctypes.c_void_p]
    return _lib


class HiddenMarkovModel:
    """Directed graph of states with log-probability edges (insertion ordered, like networkx 1.x
    on CPython >= 3.7 dicts)."""

    _counter = 0

    def __init__(self, name=None):
        HiddenMarkovModel._counter += 1
        self.name = name or 'model{}'.format(HiddenMarkovModel._counter)
        self.start = State(None, name=self.name + '-start')
        self.end = State(None, name=self.name + '-end')
        self.nodes = {self.start: None, self.end: None}   # ordered set
        self.edges = {}                                    # (a, b) -> log p, insertion ordered
        self.states = None

    # -- construction ---------------------------------------------------------------------------
    def add_state(self, s):
        self.nodes[s] = None

    def add_states(self, states):
        for s in states:
            self.add_state(s)

    def add_transition(self, a, b, probability, group=None):
        self.edges[(a, b)] = math.log(probability)

    def add_model(self, other):
        for s in other.nodes:
            self.nodes[s] = None
        for k, v in other.edges.items():
            self.edges[k] = v

    # -- bake -----------------------------------------------------------------------------------
    def bake(self, merge='All'):
        merge = merge.lower() if merge else None
        nodes, edges = self.nodes, self.edges
        while merge == 'all':
            indeg = {s: 0 for s in nodes}
            outdeg = {s: 0 for s in nodes}
            for (a, b) in edges:
                outdeg[a] += 1
                indeg[b] += 1
            drop = [s for s in nodes if s is not self.start and s is not self.end
                    and (indeg[s] == 0 or outdeg[s] == 0)]
            if not drop:
                break
            for s in drop:
                del nodes[s]
            dropset = set(drop)
            for k in [k for k in edges if k[0] in dropset or k[1] in dropset]:
                del edges[k]
        if merge in ('all', 'partial'):
            out = {}
            for (a, b), lp in edges.items():
                out.setdefault(a, []).append((a, b))
            for s in nodes:
                tot = round(sum(math.e ** edges[k] for k in out.get(s, [])), 8)
                if tot != 1.0 and s is not self.end:
                    for k in out.get(s, []):
                        edges[k] = edges[k] - math.log(tot)
        while merge in ('all', 'partial'):
            merged = 0
            for (a, b), lp in list(edges.items()):
                if a not in nodes or b not in nodes or (a, b) not in edges:
                    continue
                if a is self.start or b is self.end:
                    continue
                if lp == 0.0 and a.is_silent() and (merge == 'all' or b.is_silent()):
                    for (x, y), d in list(edges.items()):
                        if y is a:
                            merged += 1
                            del edges[(x, y)]
                            edges[(x, b)] = d
                    del nodes[a]
                    for k in [k for k in edges if k[0] is a or k[1] is a]:
                        del edges[k]
            if merged == 0:
                break
        emitting = sorted([s for s in nodes if not s.is_silent()], key=lambda s: s.name)
        silent = sorted([s for s in nodes if s.is_silent()], key=lambda s: s.name)
        # topological order of the silent sub-graph (Kahn, name order among ready states)
        sil_set = set(silent)
        indeg = {s: 0 for s in silent}
        succ = {s: [] for s in silent}
        for (a, b) in edges:
            if a in sil_set and b in sil_set:
                indeg[b] += 1
                succ[a].append(b)
        order, ready = [], [s for s in silent if indeg[s] == 0]
        while ready:
            s = ready.pop(0)
            order.append(s)
            for t in succ[s]:
                indeg[t] -= 1
                if indeg[t] == 0:
                    ready.append(t)
        if len(order) != len(silent):
            raise ValueError('loop of silent states')
        self.states = emitting + order
        self.silent_start = len(emitting)
        index = {s: i for i, s in enumerate(self.states)}
        self.start_index, self.end_index = index[self.start], index[self.end]
        m = len(self.states)
        ins = [[] for _ in range(m)]
        for (a, b), lp in edges.items():
            ins[index[b]].append((index[a], lp))
        self.in_ptr = np.zeros(m + 1, dtype=np.int32)
        for i in range(m):
            self.in_ptr[i + 1] = self.in_ptr[i] + len(ins[i])
        self.in_src = np.array([k for row in ins for (k, _) in row], dtype=np.int32)
        self.in_logp = np.array([lp for row in ins for (_, lp) in row], dtype=np.float64)
        self.dist_kind = np.array([s.distribution.kind for s in emitting], dtype=np.int32)
        self.dist_a = np.array([s.distribution.a for s in emitting], dtype=np.float64)
        self.dist_b = np.array([s.distribution.b for s in emitting], dtype=np.float64)
        self.n_edges = len(self.in_src)

    # -- decoding -------------------------------------------------------------------------------
    def viterbi(self, sequence):
        """-> (log p, [(state_index, State), ...]) or (-inf, None), like pomegranate."""
        x = np.ascontiguousarray(sequence, dtype=np.float64)
        T = len(x)
        h = _OracleHMM(len(self.states), self.silent_start, self.start_index, self.end_index,
                       self.in_ptr.ctypes.data, self.in_src.ctypes.data, self.in_logp.ctypes.data,
                       self.dist_kind.ctypes.data, self.dist_a.ctypes.data, self.dist_b.ctypes.data)
        cap = (T + 2) * (len(self.states) - self.silent_start + 2)
        path = np.empty(cap, dtype=np.int32)
        logp = ctypes.c_double(0.0)
        margin = ctypes.c_double(float('inf'))
        n = _liboracle().strique_oracle_viterbi_margin(ctypes.byref(h), x.ctypes.data, T, ctypes.byref(logp),
                                                       path.ctypes.data, cap, ctypes.byref(margin))
        # gap between the best and the second-best path (test infrastructure: lists the near-tie reads on which a
        # decoder with coarser arithmetic may legitimately differ)
        self.last_margin = margin.value
        if n < 0:
            raise MemoryError('oracle viterbi failed ({})'.format(n))
        if n == 0:
            return float('-inf'), None
        return logp.value, [(int(i), self.states[i]) for i in path[:n]]
