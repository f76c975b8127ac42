"""HMM topologies of the reference and their compilation for the CUDA Viterbi kernel.

The reference describes its models as pomegranate graphs (scripts/STRique.py:201-500):
profileHMM (match / insert / silent delete states per k-mer), repeatHMM (one repeat unit closed
into a loop through two emitting "dummy" states), flankedRepeatHMM (prefix profile -> repeat loop
-> suffix profile) and repeatModHMM (unmethylated and mCpG repeat profiles side by side).  Here the
same graphs are written down as plain state / edge tables (`Graph`) and then *compiled*
(`compile_graph`) into the form `strique_hmm_create` takes (include/strique_b200.h):

  * states that can never be entered or left are dropped and out-going probabilities that do not
    sum to one are renormalised, exactly what pomegranate's bake(merge='All') does before decoding;
  * silent glue states (s1/s2/e1/e2, sub-model starts and ends) are composed away -- their edges
    carry probability 1, so no path score changes;
  * the silent delete states remain, ordered as chains that the kernel evaluates with a max-plus
    scan inside each time step.
"""
import math

import numpy as np

FLAG_COUNT, FLAG_REPEAT, FLAG_SEP, FLAG_MOD = 1, 2, 4, 8

PROFILE_PROBS = {'match_loop': .75, 'match_match': .15, 'match_insert': .09, 'match_delete': .01,
                 'insert_loop': .15, 'insert_match_0': .40, 'insert_match_1': .40, 'insert_delete': .05,
                 'delete_delete': .005, 'delete_insert': .05, 'delete_match': .945}   # S.py:214-227


class Graph(object):
    """States (silent unless they carry a distribution) and probability-weighted edges."""

    def __init__(self):
        self.names, self.dist = [], []
        self.edges = {}            # (src, dst) -> log probability, insertion ordered
        self.start = self.state('start')
        self.end = self.state('end')
        self.counted = set()
        self.hint = {}             # state -> (segment name, index in segment, slot 0 M / 1 I / 2 D)
        self.layout = None         # segment names in chain order: enables the profile-kernel layout hints

    def state(self, name, dist=None):
        self.names.append(name)
        self.dist.append(dist)
        return len(self.names) - 1

    def normal(self, name, mean, std):
        return self.state(name, (0, float(mean), float(std)))

    def uniform(self, name, lo, hi):
        return self.state(name, (1, float(lo), float(hi)))

    def link(self, a, b, p):
        self.edges[(a, b)] = math.log(p)

    def silent(self, s):
        return self.dist[s] is None


def add_profile(g, sequence, pm, probs, prefix, no_silent=False, std_scale=1.0, std_offset=0.0):
    """profileHMM (S.py:201-300). Returns the ids of its s1, s2, e1, e2 junctions."""
    tp = dict(PROFILE_PROBS)
    tp.update(probs or {})
    k = pm.kmer
    n = len(sequence) - k + 1
    digits = int(np.ceil(np.log10(n)))
    tag = [prefix + str(i).rjust(digits, '0') for i in range(n)]
    M = [g.normal(tag[i] + 'm', pm.model_dict[sequence[i:i + k]][0],
                  pm.model_dict[sequence[i:i + k]][1] * std_scale + std_offset) for i in range(n)]
    I = [g.uniform(tag[i] + 'i', pm.model_min, pm.model_max) for i in range(n)]
    D = [] if no_silent else [g.state(tag[i] + 'd') for i in range(n)]
    s1, s2, e1, e2 = (g.state(prefix + x) for x in ('s1', 's2', 'e1', 'e2'))
    for i in range(n):
        g.hint[M[i]] = (prefix, i, 0)
        g.hint[I[i]] = (prefix, i, 1)
        if D:
            g.hint[D[i]] = (prefix, i, 2)
    last = n - 1
    for i in range(n):
        g.link(M[i], M[i], tp['match_loop'])
        g.link(I[i], I[i], tp['insert_loop'])
        g.link(M[i], I[i], tp['match_insert'])
        g.link(I[i], M[i], tp['insert_match_1'])
        if i < last:
            g.link(M[i], M[i + 1], tp['match_match'])
            g.link(I[i], M[i + 1], tp['insert_match_0'])
    if no_silent:
        for i in range(n - 2):
            g.link(M[i], M[i + 2], tp['match_delete'])     # the "delete" of a loop profile skips one match
        g.link(s1, I[0], 1)
        g.link(s2, M[0], 1)
    else:
        for i in range(n):
            g.link(D[i], I[i], tp['delete_insert'])
            if i > 0:
                g.link(M[i - 1], D[i], tp['match_delete'])
            if i < last:
                g.link(I[i], D[i + 1], tp['insert_delete'])
                g.link(D[i], M[i + 1], tp['delete_match'])
                g.link(D[i], D[i + 1], tp['delete_delete'])
        g.link(s1, D[0], 1)
        g.link(s2, M[0], 1)
        g.link(D[last], e1, tp['delete_delete'])
        g.link(D[last], e2, tp['delete_match'])
    g.link(I[last], e1, tp['insert_delete'])
    g.link(I[last], e2, tp['insert_match_0'])
    g.link(M[last], e2, tp['match_match'])
    g.link(M[last], e1, tp['match_delete'])
    return s1, s2, e1, e2


def repeat_unit(repeat, k):
    """Repeat string covering every k-mer of the periodic sequence once (S.py:329-335)."""
    if len(repeat) >= k:
        return repeat + repeat[:k - 1], 0
    ext = k - 1 + (len(repeat) - 1) - ((k - 1) % len(repeat))
    unit = repeat + (repeat * k)[:ext]
    return unit, int(len(unit) / len(repeat)) - 1


def add_repeat_loop(g, repeat, pm, probs, prefix, std_scale=1.0, std_offset=0.0):
    """repeatHMM (S.py:328-354). Returns (s1, s2, e1, e2, repeat_offset)."""
    tp = {'skip': .999, 'leave_repeat': .002}
    tp.update(probs or {})
    unit, repeat_offset = repeat_unit(repeat, pm.kmer)
    s1, s2, pe1, pe2 = add_profile(g, unit, pm, tp, prefix, no_silent=True, std_scale=std_scale, std_offset=std_offset)
    d1 = g.uniform(prefix + 'dummy1', pm.model_min, pm.model_max)
    d2 = g.uniform(prefix + 'dummy2', pm.model_min, pm.model_max)
    e1, e2 = g.state(prefix + 'e1'), g.state(prefix + 'e2')
    g.link(pe1, d1, 1)
    g.link(pe2, d2, 1)
    g.link(d1, e1, tp['leave_repeat'])
    g.link(d2, e2, tp['leave_repeat'])
    g.link(d1, s1, 1 - tp['leave_repeat'])
    g.link(d2, s2, 1 - tp['leave_repeat'])
    g.counted.update((d1, d2))
    g.hint[d2] = (prefix + 'dummy', 0, 0)       # one extra position after the unit: d2 match-like, d1 insert-like
    g.hint[d1] = (prefix + 'dummy', 0, 1)
    return s1, s2, e1, e2, repeat_offset


def flanked_repeat_graph(repeat, prefix, suffix, pm, config=None):
    """flankedRepeatHMM (S.py:384-431). Returns (graph, count_offset) with
    count_offset = flanking_count - repeat_offset (S.py:378, 437)."""
    tp = {'skip': 1 - 1e-4, 'seq_std_scale': 1.0, 'rep_std_scale': 1.0, 'seq_std_offset': 0.0,
          'rep_std_offset': 0.0, 'e1_ratio': 0.1}
    if config and isinstance(config, dict):
        tp.update(config)
    c = int(np.ceil(pm.kmer / len(repeat)))
    g = Graph()
    p = add_profile(g, prefix + (repeat * c)[:-1], pm, tp, 'prefix', std_scale=tp['seq_std_scale'],
                    std_offset=tp['seq_std_offset'])
    r = add_repeat_loop(g, repeat, pm, tp, 'repeat', std_scale=tp['rep_std_scale'], std_offset=tp['rep_std_offset'])
    s = add_profile(g, repeat * c + suffix, pm, tp, 'suffix', std_scale=tp['seq_std_scale'],
                    std_offset=tp['seq_std_offset'])
    g.link(g.start, p[0], tp['e1_ratio'])
    g.link(g.start, p[1], 1 - tp['e1_ratio'])
    g.link(p[2], r[0], 1)
    g.link(p[3], r[1], 1)
    g.link(r[2], s[0], 1)
    g.link(r[3], s[1], 1)
    g.link(s[2], g.end, 1)
    g.link(s[3], g.end, 1)
    g.layout = ['prefix', 'repeat', 'repeatdummy', 'suffix']
    return g, (c * 2 - 1) - r[4]


def repeat_mod_graph(repeat, pm_base, pm_mod, config=None):
    """repeatModHMM (S.py:447-490). Returns (graph, clip_lo, clip_hi)."""
    tp = {'rep_std_scale': 1.5, 'rep_std_offset': 0.0, 'leave_repeat': .002}
    if config and isinstance(config, dict):
        tp.update(config)
    unit, _ = repeat_unit(repeat, pm_base.kmer)
    lo = min(pm_base.model_min, pm_mod.model_min)
    hi = max(pm_base.model_max, pm_mod.model_max)
    g = Graph()
    s0 = g.uniform('s0', lo, hi)
    e0 = g.uniform('e0', lo, hi)
    base = add_profile(g, unit, pm_base, tp, 'base', no_silent=True, std_scale=tp['rep_std_scale'],
                       std_offset=tp['rep_std_offset'])
    mod = add_profile(g, unit, pm_mod, tp, 'mod', no_silent=True,
                      std_scale=tp['rep_std_scale'] * pm_mod.scale2stdv(pm_base), std_offset=tp['rep_std_offset'])
    g.link(g.start, s0, 1)
    for j in (base[0], base[1], mod[0], mod[1]):
        g.link(s0, j, 0.25)
    for j in (base[2], base[3], mod[2], mod[3]):
        g.link(j, e0, 1)
    g.link(e0, g.end, tp['leave_repeat'])
    g.link(e0, s0, 1 - tp['leave_repeat'])
    return g, lo, hi


class CompiledHMM(object):
    """Arrays of `strique_hmm_desc` (include/strique_b200.h) plus bookkeeping for tests."""

    def __init__(self):
        self.names = []          # emitting state names, index = state id on the device side
        self.n_emit = self.n_chain = 0


def compile_graph(g):
    n = len(g.names)
    alive = [True] * n
    edges = dict(g.edges)
    # 1. states that cannot be entered or left (sub-model starts/ends, unused junctions)
    while True:
        indeg, outdeg = [0] * n, [0] * n
        for (a, b) in edges:
            outdeg[a] += 1
            indeg[b] += 1
        drop = [s for s in range(n) if alive[s] and s not in (g.start, g.end) and (indeg[s] == 0 or outdeg[s] == 0)]
        if not drop:
            break
        for s in drop:
            alive[s] = False
        edges = {k: v for k, v in edges.items() if alive[k[0]] and alive[k[1]]}
    # 2. out-going probabilities must sum to one (8 decimals), else renormalise
    out_of = {}
    for k in edges:
        out_of.setdefault(k[0], []).append(k)
    for s, ks in out_of.items():
        tot = round(sum(math.e ** edges[k] for k in ks), 8)
        if tot != 1.0 and s != g.end:
            for k in ks:
                edges[k] = edges[k] - math.log(tot)
    # 3. compose silent glue states away
    def silent_glue(s):
        return alive[s] and g.silent(s) and s not in (g.start, g.end)

    def ins(s):
        return [(k, w) for k, w in edges.items() if k[1] == s]

    def outs(s):
        return [(k, w) for k, w in edges.items() if k[0] == s]

    end_edges = []           # (src, logw): edges into END, parallel edges allowed
    for (a, b), w in list(edges.items()):
        if b == g.end:
            end_edges.append((a, w))
            del edges[(a, b)]
    changed = True
    while changed:
        changed = False
        for s in range(n):
            if not silent_glue(s):
                continue
            o, i = outs(s), ins(s)
            to_end = [(a, w) for (a, w) in end_edges if a == s]
            sil_pred = [k for k, _ in i if silent_glue(k[0])]
            sil_succ = [k for k, _ in o if silent_glue(k[1])]
            if any(k[0] == s and k[1] == s for k, _ in o):
                raise ValueError('silent self loop at ' + g.names[s])
            single_certain = len(o) == 1 and not to_end and o[0][1] == 0.0
            only_end = not o and to_end
            isolated = not sil_pred and not sil_succ
            if not (single_certain or only_end or isolated):
                continue
            for (x, _), w1 in i:
                for (_, y), w2 in o:
                    # pomegranate merges a certain (p = 1) silent hop by re-pointing the in-edges at
                    # the target; its graph keeps ONE edge per state pair, so a second merge onto the
                    # same pair replaces the first (repeatModHMM: e1 and e2 both lead to e0)
                    if (x, y) in edges and not single_certain:
                        raise ValueError('parallel edge while composing ' + g.names[s])
                    edges[(x, y)] = w1 + w2
                for (_, w2) in to_end:
                    end_edges.append((x, w1 + w2))
            for k, _ in i + o:
                edges.pop(k, None)
            end_edges = [(a, w) for (a, w) in end_edges if a != s]
            alive[s] = False
            changed = True
    # 4. what is left of the silent states must form chains
    chain_states = [s for s in range(n) if silent_glue(s)]
    pred, succ = {}, {}
    for (a, b), w in edges.items():
        if silent_glue(a) and silent_glue(b):
            if a in succ or b in pred:
                raise ValueError('silent states do not form simple chains')
            succ[a], pred[b] = b, a
    order = []
    for s in chain_states:
        if s not in pred:
            while s is not None:
                order.append(s)
                s = succ.get(s)
    if len(order) != len(chain_states):
        raise ValueError('loop of silent states')
    emitting = [s for s in range(n) if alive[s] and not g.silent(s)]
    E, C = len(emitting), len(order)
    eid = {s: i for i, s in enumerate(emitting)}
    cid = {s: E + i for i, s in enumerate(order)}
    START = E + C

    def vid(s):
        if s == g.start:
            return START
        return eid[s] if s in eid else cid[s]

    c = CompiledHMM()
    c.n_emit, c.n_chain = E, C
    c.names = [g.names[s] for s in emitting]
    in_lists = [[] for _ in range(E)]
    chain_in = [[] for _ in range(C)]
    chain_pred = np.full(max(C, 1), -np.inf)
    for (a, b), w in edges.items():
        if b in eid:
            in_lists[eid[b]].append((vid(a), w))
        elif b in cid:
            if a in cid:
                if cid[a] != cid[b] - 1:
                    raise ValueError('chain order broken')
                chain_pred[cid[b] - E] = w
            else:
                chain_in[cid[b] - E].append((vid(a), w))
        else:
            raise ValueError('edge into removed state ' + g.names[b])
    c.in_ptr = np.zeros(E + 1, dtype=np.int32)
    c.in_ptr[1:] = np.cumsum([len(x) for x in in_lists])
    c.in_src = np.array([s for x in in_lists for s, _ in x], dtype=np.int32)
    c.in_logw = np.array([w for x in in_lists for _, w in x], dtype=np.float64)
    c.emit_kind = np.array([g.dist[s][0] for s in emitting], dtype=np.int32)
    c.emit_a = np.array([g.dist[s][1] for s in emitting], dtype=np.float64)
    c.emit_b = np.array([g.dist[s][2] for s in emitting], dtype=np.float64)
    flags = []
    for s in emitting:
        name = g.names[s]
        f = (FLAG_COUNT if s in g.counted else 0) | (FLAG_REPEAT if 'repeat' in name else 0)
        f |= (FLAG_SEP if name in ('s0', 'e0') else 0) | (FLAG_MOD if 'mod' in name else 0)
        flags.append(f)
    c.emit_flags = np.array(flags, dtype=np.uint8)
    c.chain_pred_logw = chain_pred.astype(np.float64)
    c.chain_in_ptr = np.zeros(C + 1, dtype=np.int32)
    c.chain_in_ptr[1:] = np.cumsum([len(x) for x in chain_in])
    c.chain_in_src = np.array([s for x in chain_in for s, _ in x] or [0], dtype=np.int32)
    c.chain_in_logw = np.array([w for x in chain_in for _, w in x] or [0.0], dtype=np.float64)
    c.end_src = np.array([vid(a) for a, _ in end_edges], dtype=np.int32)
    c.end_logw = np.array([w for _, w in end_edges], dtype=np.float64)
    c.n_edges = int(len(c.in_src) + sum(len(x) for x in chain_in) + int(np.sum(np.isfinite(chain_pred[:C]))))
    # layout hints for the profile kernel (include/strique_b200.h): position = offset of the segment + index
    c.emit_pos = c.emit_slot = c.chain_pos = None
    if g.layout and all(s in g.hint for s in emitting + order):
        size = {}
        for s in emitting + order:
            seg, i, _ = g.hint[s]
            size[seg] = max(size.get(seg, 0), i + 1)
        if set(size) <= set(g.layout):
            base, at = {}, 0
            for seg in g.layout:
                base[seg] = at
                at += size.get(seg, 0)
            c.emit_pos = np.array([base[g.hint[s][0]] + g.hint[s][1] for s in emitting], dtype=np.int32)
            c.emit_slot = np.array([g.hint[s][2] for s in emitting], dtype=np.uint8)
            c.chain_pos = np.array([base[g.hint[s][0]] + g.hint[s][1] for s in order] or [0], dtype=np.int32)
    return c
