// Host-only: recognises a linear profile HMM in a compiled model (strique_hmm_desc + layout hints) and
// packs it into the per-lane constant table of the profile Viterbi kernel (profile_core.h).
//
// The compiled models of the reference's flankedRepeatHMM (scripts/STRique.py:384-431) fit after one
// rewrite: an insert-like state with in-edges from the PREVIOUS position (the repeat profile's I_0, fed
// by prefix.e1 -> repeat.s1, S.py:81,147; the dummy state d1, fed by the repeat profile's e1, S.py:342)
// gets a virtual delete state at its own position that collects those edges and feeds it with log 1 = 0 --
// the silent junction the graph had before it was composed away (hmm.py), so no path score changes.
#pragma once
#include <math.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/strique_b200.h"
#include "profile_core.h"

namespace strique {

struct ProfileImage {
    int np = 0;                          // positions in use (including the leading offset)
    int p_off = 0;                       // hint position h lives at p_off + h; START = M slot of p_off - 1
    std::vector<double> tab;             // [pf::K_TOTAL][32]
    pf::TraceCfg trace{-1, 0, -1, 0};
    double lo = -INFINITY, hi = INFINITY;   // fast-path emission range (inside every Uniform range)
    // per (position, slot in {M, I}) tables, index p * 2 + slot
    std::vector<uint8_t> em_kind;        // 0 Normal, 1 Uniform, 2 dead
    std::vector<double> em_a, em_b, em_c;   // Normal: mu, c0, c2; Uniform: lo, hi, -log(hi - lo)
    std::vector<uint8_t> flags;
    std::vector<int32_t> state_id;       // caller's emitting state id, -1 dead
    int n_end = 0;
    int32_t end_p[16], end_slot[16];
    double end_w[16];
};

// Returns true and fills `img` when the model fits; otherwise false with a reason in `why`.
inline bool profile_pack(const strique_hmm_desc *d, ProfileImage *img, std::string *why) {
    static const double SQRT_2_PI = 2.50662827463;   // pomegranate's truncated constant (distributions.pyx)
    auto fail = [&](const char *msg) { if (why) *why = msg; return false; };
    if (!d->emit_pos || !d->emit_slot || (d->n_chain > 0 && !d->chain_pos)) return fail("no layout hints");
    const int E = d->n_emit, C = d->n_chain, START = E + C;
    int nph = 0;
    for (int l = 0; l < E; ++l) {
        if (d->emit_pos[l] < 0 || d->emit_slot[l] > 1) return fail("bad emitting hint");
        nph = std::max(nph, d->emit_pos[l] + 1);
    }
    for (int c = 0; c < C; ++c) {
        if (d->chain_pos[c] < 0) return fail("bad chain hint");
        nph = std::max(nph, d->chain_pos[c] + 1);
    }
    if (nph + 1 > pf::NPOS) return fail("more positions than the profile kernel holds");
    if (d->n_end > 16 || d->n_end <= 0) return fail("too many END edges");
    const double NINF = -INFINITY;
    // ---- working tables in hint positions --------------------------------------------------------------
    struct Pos {
        int m = -1, i = -1, dchain = -1;     // emitting ids / chain id
        bool dvirt = false;
        double wM[6], wI[3], wD[3], wXM, wXD;
        bool hasXM = false, hasXD = false;
        Pos() {
            for (double &w : wM) w = -INFINITY;
            for (double &w : wI) w = -INFINITY;
            for (double &w : wD) w = -INFINITY;
            wXM = wXD = -INFINITY;
        }
    };
    std::vector<Pos> pos(nph);
    for (int l = 0; l < E; ++l) {
        Pos &p = pos[d->emit_pos[l]];
        int &slot = d->emit_slot[l] == 0 ? p.m : p.i;
        if (slot >= 0) return fail("two states in one slot");
        slot = l;
    }
    for (int c = 0; c < C; ++c) {
        Pos &p = pos[d->chain_pos[c]];
        if (p.dchain >= 0) return fail("two chain states at one position");
        p.dchain = c;
    }
    int xm_src = -1, xd_src = -1, x_tgt = -1;   // emitting ids of the long-range sources, hint position of the target
    bool ok = true;
    const char *msg = "";
    auto set = [&](double &slot, double w) {
        if (slot != NINF || !(w > NINF)) { ok = false; msg = "edge does not fit the position template"; return; }
        slot = w;
    };
    // entry edge into the (real or virtual) delete state at hint position hp
    auto d_entry_edge = [&](int hp, int src, double w) {
        Pos &p = pos[hp];
        if (src == START) {
            if (hp != 0) { ok = false; msg = "START edge beyond the first position"; return; }
            set(p.wD[0], w);
        } else if (src < E) {
            const int sp = d->emit_pos[src], ss = d->emit_slot[src];
            if (sp == hp - 1) set(p.wD[ss == 0 ? 0 : 1], w);
            else {
                if (p.hasXD || (x_tgt >= 0 && x_tgt != hp)) { ok = false; msg = "second long-range edge"; return; }
                p.hasXD = true; p.wXD = w; xd_src = src; x_tgt = hp;
            }
        } else {
            const int sp = d->chain_pos[src - E];
            if (sp != hp - 1) { ok = false; msg = "delete hop over more than one position"; return; }
            set(p.wD[2], w);
        }
    };
    for (int l = 0; l < E && ok; ++l) {
        const int hp = d->emit_pos[l];
        Pos &p = pos[hp];
        for (int e = d->in_ptr[l]; e < d->in_ptr[l + 1] && ok; ++e) {
            const int src = d->in_src[e];
            const double w = d->in_logw[e];
            if (src < 0 || src > START) { ok = false; msg = "in-edge source out of range"; break; }
            if (d->emit_slot[l] == 0) {                        // match-like
                if (src == l) set(p.wM[0], w);
                else if (src == START) {
                    if (hp != 0) { ok = false; msg = "START edge beyond the first position"; break; }
                    set(p.wM[1], w);
                } else if (src < E) {
                    const int sp = d->emit_pos[src], ss = d->emit_slot[src];
                    if (sp == hp - 1 && ss == 0) set(p.wM[1], w);
                    else if (sp == hp - 1 && ss == 1) set(p.wM[2], w);
                    else if (sp == hp && ss == 1) set(p.wM[4], w);
                    else if (sp == hp - 2 && ss == 0) set(p.wM[5], w);
                    else {
                        if (p.hasXM || (x_tgt >= 0 && x_tgt != hp)) { ok = false; msg = "second long-range edge"; break; }
                        p.hasXM = true; p.wXM = w; xm_src = src; x_tgt = hp;
                    }
                } else {
                    if (d->chain_pos[src - E] != hp - 1) { ok = false; msg = "delete -> match over more than one position"; break; }
                    set(p.wM[3], w);
                }
            } else {                                           // insert-like
                if (src == l) set(p.wI[0], w);
                else if (src < E && d->emit_pos[src] == hp && d->emit_slot[src] == 0) set(p.wI[1], w);
                else if (src >= E && src < START && d->chain_pos[src - E] == hp) set(p.wI[2], w);
                else {
                    // junction rewrite: route the edge through a virtual delete state at this position
                    if (p.dchain >= 0) { ok = false; msg = "insert state with foreign in-edges next to a delete state"; break; }
                    if (!p.dvirt) { p.dvirt = true; set(p.wI[2], 0.0); }
                    if (src == START) { ok = false; msg = "START -> insert"; break; }
                    d_entry_edge(hp, src, w);
                }
            }
        }
    }
    for (int c = 0; c < C && ok; ++c) {
        const int hp = d->chain_pos[c];
        if (c > 0 && d->chain_pred_logw[c] > NINF) {
            if (d->chain_pos[c - 1] != hp - 1) { ok = false; msg = "chain hop over more than one position"; break; }
            set(pos[hp].wD[2], d->chain_pred_logw[c]);
        }
        for (int e = d->chain_in_ptr[c]; e < d->chain_in_ptr[c + 1] && ok; ++e) {
            const int src = d->chain_in_src[e];
            if (!((src >= 0 && src < E) || src == START)) { ok = false; msg = "chain entry from a silent state"; break; }
            d_entry_edge(hp, src, d->chain_in_logw[e]);
        }
    }
    if (!ok) return fail(msg);
    // I-slot states must be Uniform (their fast-path emission is one constant)
    for (int l = 0; l < E; ++l)
        if (d->emit_slot[l] == 1 && d->emit_kind[l] != 1) return fail("insert-like state with a Normal emission");
    // ---- placement: the long-range target at an in-lane index 0, START one position before hint 0 ------
    int p_off = 1;
    if (x_tgt >= 0) while ((x_tgt + p_off) % pf::P != 0) ++p_off;
    const int np = nph + p_off;
    if (np > pf::NPOS) return fail("more positions than the profile kernel holds");
    img->np = np;
    img->p_off = p_off;
    img->tab.assign((size_t)pf::K_TOTAL * 32, 0.0);
    auto T = [&](int k, int p) -> double & { return img->tab[(size_t)(k) * 32 + p / pf::P]; };
    // table index of in-edge k of the M / I / D slot of in-lane position q (k as in struct Pos)
    auto kM = [](int q, int k) {
        return k == 3 ? pf::K_E2 + q * 2 : (k == 5 ? pf::K_WM2 + q : pf::K_WMR + q * 4 + (k == 4 ? 3 : k));
    };
    auto kI = [](int q, int k) { return k == 2 ? pf::K_E2 + q * 2 + 1 : pf::K_WI + q * 2 + k; };
    auto kD = [](int q, int k) { return k == 2 ? pf::K_WH + q : pf::K_WD + q * 2 + k; };
    // defaults: every weight -inf, emissions 0
    for (int p = 0; p < pf::NPOS; ++p) {
        const int q = p % pf::P;
        for (int k = 0; k < 6; ++k) T(kM(q, k), p) = NINF;
        for (int k = 0; k < 3; ++k) T(kI(q, k), p) = NINF;
        for (int k = 0; k < 3; ++k) T(kD(q, k), p) = NINF;
    }
    for (int lane = 0; lane < 32; ++lane) {
        img->tab[(size_t)pf::K_WX * 32 + lane] = NINF;
        img->tab[(size_t)(pf::K_WX + 1) * 32 + lane] = NINF;
    }
    img->em_kind.assign(pf::NPOS * 2, 2);
    img->em_a.assign(pf::NPOS * 2, 0.0);
    img->em_b.assign(pf::NPOS * 2, 0.0);
    img->em_c.assign(pf::NPOS * 2, 0.0);
    img->flags.assign(pf::NPOS * 2, 0);
    img->state_id.assign(pf::NPOS * 2, -1);
    double lo = -INFINITY, hi = INFINITY;
    for (int hp = 0; hp < nph; ++hp) {
        const Pos &ps = pos[hp];
        const int p = hp + p_off, q = p % pf::P;
        for (int k = 0; k < 6; ++k) T(kM(q, k), p) = ps.wM[k];
        for (int k = 0; k < 3; ++k) T(kI(q, k), p) = ps.wI[k];
        for (int k = 0; k < 3; ++k) T(kD(q, k), p) = ps.wD[k];
        if (ps.hasXM) T(pf::K_WX, p) = ps.wXM;
        if (ps.hasXD) T(pf::K_WX + 1, p) = ps.wXD;
        for (int slot = 0; slot < 2; ++slot) {
            const int l = slot == 0 ? ps.m : ps.i;
            if (l < 0) continue;
            const int idx = p * 2 + slot;
            const double a = d->emit_a[l], b = d->emit_b[l];
            img->state_id[idx] = l;
            img->flags[idx] = d->emit_flags ? (d->emit_flags[l] & 0x7f) : 0;
            if (d->emit_kind[l] == 0) {
                img->em_kind[idx] = 0;
                img->em_a[idx] = a;
                img->em_b[idx] = -log(b * SQRT_2_PI);
                img->em_c[idx] = b > 0 ? 1.0 / (2.0 * (b * b)) : 0.0;
                T(pf::K_EM + q * 2, p) = img->em_a[idx];
                T(pf::K_EC + q * 2, p) = img->em_b[idx];
                T(pf::K_EM + q * 2 + 1, p) = img->em_c[idx];
            } else {
                img->em_kind[idx] = 1;
                img->em_a[idx] = a;
                img->em_b[idx] = b;
                img->em_c[idx] = -log(b - a);
                lo = std::max(lo, a);
                hi = std::min(hi, b);
                if (slot == 0) {
                    T(pf::K_EC + q * 2, p) = img->em_c[idx];    // mu = 0, c2 = 0: c0 - (x*x)*0 = c0 exactly
                } else {
                    T(pf::K_EC + q * 2 + 1, p) = img->em_c[idx];
                }
            }
        }
    }
    img->lo = lo;
    img->hi = hi;
    auto slot_of = [&](int l, int *p, int *slot) { *p = d->emit_pos[l] + p_off; *slot = d->emit_slot[l]; };
    if (xm_src >= 0) slot_of(xm_src, &img->trace.xm_src_p, &img->trace.xm_src_slot);
    if (xd_src >= 0) slot_of(xd_src, &img->trace.xd_src_p, &img->trace.xd_src_slot);
    // both long-range sources are read with one in-lane index per kernel instance: require one position
    if (xm_src >= 0 && xd_src >= 0 && img->trace.xm_src_p != img->trace.xd_src_p)
        return fail("long-range sources at different positions");
    // ---- summed hop weights of the cross-lane scan (same association as the kernel's recurrence) --------
    double W[32];
    for (int lane = 0; lane < 32; ++lane) {
        W[lane] = img->tab[(size_t)pf::K_WH * 32 + lane];
        for (int q = 1; q < pf::P; ++q) W[lane] += img->tab[(size_t)(pf::K_WH + q) * 32 + lane];
    }
    for (int r = 0; r < 5; ++r) {
        const int off = 1 << r;
        double Wn[32];
        for (int lane = 0; lane < 32; ++lane) {
            img->tab[(size_t)(pf::K_CWR + r) * 32 + lane] = lane >= off ? W[lane] : NINF;
            Wn[lane] = lane >= off ? W[lane - off] + W[lane] : W[lane];
        }
        memcpy(W, Wn, sizeof(W));
    }
    // ---- END edges --------------------------------------------------------------------------------------
    img->n_end = d->n_end;
    for (int e = 0; e < d->n_end; ++e) {
        const int src = d->end_src[e];
        if (src < 0 || src >= START) return fail("END edge source out of range");
        if (src < E) { img->end_p[e] = d->emit_pos[src] + p_off; img->end_slot[e] = d->emit_slot[src]; }
        else { img->end_p[e] = d->chain_pos[src - E] + p_off; img->end_slot[e] = 2; }
        img->end_w[e] = d->end_logw[e];
    }
    return true;
}

}  // namespace strique
