// Lane arithmetic of the FIXED-POINT profile Viterbi kernel (viterbi_profile_q.cu), written once for device and host.
//
// Same model layout and recurrences as profile_core.h (positions x {M, I, D} slots, lane l owns positions
// 4l..4l+3; replaces pomegranate 0.10.0 `HiddenMarkovModel.viterbi` for the linear profile topologies of the
// reference, scripts/STRique.py:201-441), but the forward pass runs on TAGGED 32-bit FIXED-POINT scores:
//
//   * a score is an int32 in units of 2^-18 nat relative to a running column maximum; its low 3 bits are zero
//     ("clean"), so the resolution of a value is 2^-15 nat = 3e-5 (fp32 at |log p| ~ 1e4 resolves 1e-3) and the
//     range 8192 nat;
//   * every in-edge weight carries the NAME of its edge in its low 3 bits ("tag", larger = earlier in the
//     reference's candidate order).  One `max(v + w, best)` -- a single VIADDMNMX on sm_100a -- therefore relaxes
//     the edge AND keeps the winner's name: the arg-max costs no instruction, and the first candidate wins exact
//     ties like the strict '>' of the float64 decoder;
//   * integer sums do not round: a path's fixed-point score differs from its float64 score only by the rounding
//     of the weights (once per model) and of the emissions (2^-16 each), never by the order of operations;
//   * every R_NORM columns the column maximum is subtracted (renormalisation) and values are clamped FROM BELOW at
//     Q_FLOOR (unreachable states start there too).  Clamping only ever RAISES a value, and max-plus is monotone, so
//     every forward value is an upper bound of the state's true fixed-point score, exact wherever no clamped value
//     won on the way.  The traceback therefore re-adds the quantised weights and emissions along the decoded path in
//     integers: if the sum equals the forward value, no clamp (floor or emission clamp) touched the winning path,
//     and because every other path's forward value bounds its score from above, the decoded path is EXACTLY the
//     optimum of the quantised model.  If not -- a read whose best path dips more than |Q_FLOOR| nat below the column
//     maximum -- the sequence is decoded in float64.  (Reads whose flank alignment went wrong make the best path run
//     700 ... 900 nat below the column maximum for ~2000 columns, 4 of 8192 C2 reads, and "teleporting" paths out of
//     clamped states then need the floor well below that: hence 2^-15 nat with the floor at -2000 rather than 2^-16
//     with -1024, which costs 1 more near-tie path difference in 512 golden reads and no integer output.)  (The first version declared such
//     states unreachable instead; that is not conservative, and those 4 reads came back with a worse path.)
//     The constants below make int32 overflow impossible.
//
// log p is NOT taken from the fixed-point pass: the traceback also re-scores the decoded path in float64 with the
// model's original weights and emissions, so log p is the exact float64 score of the returned path.  Samples
// outside the fast emission range, models outside the bounds and paths that touch a clamp go to the float64 kernel.
//
// Back-pointer byte of position q of a lane (4 per 32-bit word, one word per lane per column):
//     bits 0-2  M: 7 self, 6 M_{p-1}, 5 I_{p-1}, 4 I_p, 3 M_{p-2}, 2 X_M, 1 D_{p-1}
//     bits 3-4  I: 3 self, 2 M_p, 1 D_p
//     bits 5-6  D: 3 M_{p-1}, 2 I_{p-1}, 1 X_D, 0 hop from D_{p-1}
#pragma once
#include <math.h>
#include <stdint.h>

#include "profile_core.h"

namespace strique {
namespace pq {

constexpr int P = pf::P;
#ifndef PQ_FRAC
#define PQ_FRAC 15
#endif
constexpr int FRAC = PQ_FRAC;                    // value resolution 2^-FRAC nat
constexpr int TAG_BITS = 3;
constexpr int32_t TAG_MASK = 7;
constexpr int32_t Q_ONE = 1 << (FRAC + TAG_BITS);   // one nat
#ifndef PQ_RNORM
#define PQ_RNORM 4
#endif
constexpr int R_NORM = PQ_RNORM;                 // columns between renormalisations
// bounds in nat (see the overflow argument in DESIGN.md 4.3): finite weights >= -W_MAX, emissions clamped to
// >= -E_MAX, emission constants <= C0_MAX, so a value drops by at most S = W_MAX + E_MAX per column.
constexpr int W_MAX = 12, E_MAX = 100, C0_MAX = 2, S_STEP = W_MAX + E_MAX;
#ifndef PQ_FLOOR
#define PQ_FLOOR 2000
#define PQ_ABSENT 2500
#endif
constexpr int Q_FLOOR_NAT = -PQ_FLOOR;           // values are clamped from below here (relative to the column maximum)
constexpr int Q_ABSENT_NAT = -PQ_ABSENT;         // weight of an edge the model does not have
// a candidate through an absent edge (best source: the column maximum, grown for R_NORM columns) stays below every
// candidate through a real edge (worst source: the floor, decayed for R_NORM columns) -- it can never win
static_assert(-Q_ABSENT_NAT >= -Q_FLOOR_NAT + R_NORM * S_STEP + W_MAX + R_NORM * C0_MAX, "an absent edge must lose against every real candidate");
// lowest intermediate: floor, decayed, through an absent entry and an absent hop of the delete chain
static_assert(-Q_FLOOR_NAT + R_NORM * S_STEP + 2 * -Q_ABSENT_NAT + W_MAX < (1 << (31 - FRAC - TAG_BITS)), "int32 range");
constexpr int32_t Q_FLOOR = Q_FLOOR_NAT * Q_ONE, Q_ABSENT = Q_ABSENT_NAT * Q_ONE;
constexpr int32_t E_MIN16 = -E_MAX * (1 << FRAC);   // emission clamp in units of 2^-FRAC

// Per-lane table of a quantised model: int32 groups of four (fetched as one 16-byte word) ...
enum : int {
    G_WM = 0,                 // [q] {self, M_{p-1}, I_{p-1}, I_p} -> M_p                  (kept in registers)
    G_WB = G_WM + P,          // [q] {M_{p-2} -> M_p, D_{p-1} -> M_p, I_p self, M_p -> I_p}
    G_WC = G_WB + P,          // [q] {D_p -> I_p, M_{p-1} -> D_p, I_{p-1} -> D_p, hop D_{p-1} -> D_p}
    G_X = G_WC + P,           // {X_M, X_D, scan round 4, 0}
    G_CWR = G_X + 1,          // {scan rounds 0..3}: summed hop weights seen by the cross-lane scan
    G_EI = G_CWR + 1,         // [q] I-slot emission (Uniform), already shifted: multiple of 8
    G_TOTAL = G_EI + 1
};
// ... and float64 emission constants of the M slots.  log N(x; mu, sigma) = c0 - c (x - mu)^2 with c = 1/(2 sigma^2) is
// evaluated as  C + A x - c x^2  (x^2 once per column): pairs {A = 2 c mu, c}[q], then {C[0], C[1]}, {C[2], C[3]} with
// C = c0 - c mu^2.  The cancellation costs ~1e-13 nat, eleven orders below the fixed-point resolution.
enum : int { E_MU = 0, E_C0 = P, E_TOTAL = P + P / 2 };

struct I4 {
    int32_t x, y, z, w;
};

PF_HD int32_t imax(int32_t a, int32_t b) { return a > b ? a : b; }
// a * b + c on the FMA pipe (IMAD).  Opaque to the optimiser on the device: it would otherwise fold the
// "dirty - clean" tag extraction below back into logic operations on the ALU pipe, the pipe that bounds the kernel.
PF_HD uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    return a * b + c;
#endif
}

struct RegsQ {
    int32_t wM[P][4];
};

struct StateQ {
    int32_t M[P], I[P], D[P];       // last finished column (clean)
    int32_t partM[P], partI[P];     // tagged E1 maxima of the upcoming column
    int32_t Dprev;                  // D of the last position of the previous lane (same column as D[], clean)
};

// magic-number rounding of a float64 emission to units of 2^-FRAC (round to nearest even, exact for |e| < 2^(31 - FRAC))
PF_HD int32_t to_q16(double e) {
    const double t = e + 1.5 * (double)(1ull << (52 - FRAC));   // FRAC = 16: 1.5 * 2^36, ulp 2^-16 in [2^36, 2^37)
#ifdef __CUDA_ARCH__
    return __double2loint(t);
#else
    union { double d; uint64_t u; } c;
    c.d = t;
    return (int32_t)(uint32_t)c.u;
#endif
}

PF_HD double fma_rn(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}

// Emission from the constants {A, c, C} of a state (see E_MU), x2 = x * x: units of 2^-16, clamped at E_MIN16.
PF_HD int32_t emission_q16(double A, double c, double C, double x, double x2) {
    return imax(to_q16(fma_rn(-c, x2, fma_rn(A, x, C))), E_MIN16);
}
// Emission of the M slot of in-lane position q for sample x (inside every Uniform range, not NaN), x2 = x * x:
// units of 2^-16, clamped.  Two fused multiply-adds and the rounding add.
template <class Tab>
PF_HD int32_t emission_q(const Tab &tab, double x, double x2, int q) {
    const pf::Pair em = tab.dpair(E_MU + q), cc = tab.dpair(E_C0 + q / 2);
    return emission_q16(em.a, em.b, (q & 1) ? cc.b : cc.a, x, x2);
}
template <class Tab>
PF_HD void emissions_q(const Tab &tab, double x, int32_t eM[P]) {
    const double x2 = x * x;
#pragma unroll
    for (int q = 0; q < P; ++q) eM[q] = emission_q(tab, x, x2, q);
}

// E2 + emission: finishes column t from the tagged E1 maxima and the delete states of column t-1.
// Returns the M / I back-pointer bits of the column.
template <class Tab>
PF_HD uint32_t e2_emit(const Tab &tab, StateQ &s, const int32_t eM[P]) {
    const I4 ei = tab.grp(G_EI);
    uint32_t dirty = 0u, clean = 0u;
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const I4 wb = tab.grp(G_WB + q), wc = tab.grp(G_WC + q);
        const int32_t dsrc = q == 0 ? s.Dprev : s.D[q > 0 ? q - 1 : 0];
        const int32_t bm = imax(dsrc + wb.y, s.partM[q]);
        const int32_t bi = imax(s.D[q] + wc.x, s.partI[q]);
        const int32_t eiq = q == 0 ? ei.x : (q == 1 ? ei.y : (q == 2 ? ei.z : ei.w));
        // the emission is a multiple of 8, so the winner's tag survives the add; tag = dirty - clean.  The byte
        // packing is sum(dirty * 2^k) - sum(clean * 2^k), multiply-adds on the FMA pipe (wrap-around cancels): the
        // ALU pipe (add-max, logic) is the one this kernel is bound by.
        const int32_t md = bm + eM[q] * 8, id = bi + eiq;
        s.M[q] = md & ~TAG_MASK;
        s.I[q] = id & ~TAG_MASK;
        dirty = mad_u32((uint32_t)md, 1u << (8 * q), dirty);
        clean = mad_u32((uint32_t)s.M[q], 1u << (8 * q), clean);
        dirty = mad_u32((uint32_t)id, 8u << (8 * q), dirty);
        clean = mad_u32((uint32_t)s.I[q], 8u << (8 * q), clean);
    }
    return dirty - clean;
}

// E1 of the next column from the emitting values of the column just finished (all clean).
// pM3 / pI3 / pM2: M, I of the last and M of the last-but-one position of the previous lane; xm: X_M source.
template <class Tab>
PF_HD void e1_q(const RegsQ &r, const Tab &tab, StateQ &s, int32_t pM3, int32_t pI3, int32_t pM2, int32_t xm, int q) {
    const I4 wb = tab.grp(G_WB + q);
    const int32_t m1 = q == 0 ? pM3 : s.M[q > 0 ? q - 1 : 0];
    const int32_t i1 = q == 0 ? pI3 : s.I[q > 0 ? q - 1 : 0];
    const int32_t m2 = q == 0 ? pM2 : (q == 1 ? pM3 : s.M[q > 1 ? q - 2 : 0]);
    int32_t c = s.M[q] + r.wM[q][0];
    c = imax(m1 + r.wM[q][1], c);
    c = imax(i1 + r.wM[q][2], c);
    c = imax(s.I[q] + r.wM[q][3], c);
    c = imax(m2 + wb.x, c);
    if (q == 0) c = imax(xm + tab.grp(G_X).x, c);
    s.partM[q] = c;
    s.partI[q] = imax(s.M[q] + wb.w, s.I[q] + wb.z);
}
template <class Tab>
PF_HD void e1(const RegsQ &r, const Tab &tab, StateQ &s, int32_t pM3, int32_t pI3, int32_t pM2, int32_t xm) {
#pragma unroll
    for (int q = 0; q < P; ++q) e1_q(r, tab, s, pM3, pI3, pM2, xm, q);
}

// Delete chain, part 1: tagged entry maxima a[q] of this lane's D states and the lane composite A
// (D of the lane's last position when nothing enters from the previous lane; its low bits are not a tag).
template <class Tab>
PF_HD void d_entry(const Tab &tab, const StateQ &s, int32_t pM3, int32_t pI3, int32_t xd, int32_t a[P], int32_t &A) {
    const I4 wx = tab.grp(G_X);
    int32_t h[P];
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const I4 wc = tab.grp(G_WC + q);
        const int32_t m1 = q == 0 ? pM3 : s.M[q > 0 ? q - 1 : 0];
        const int32_t i1 = q == 0 ? pI3 : s.I[q > 0 ? q - 1 : 0];
        int32_t c = imax(i1 + wc.z, m1 + wc.y);
        if (q == 0) c = imax(xd + wx.y, c);
        a[q] = c;
        h[q] = wc.w;
    }
    A = a[0];
#pragma unroll
    for (int q = 1; q < P; ++q) A = imax(A + h[q], a[q]);
}

// One round of the cross-lane max-plus scan: Al = A of lane - 2^r.
template <class Tab>
PF_HD int32_t d_round(const Tab &tab, int32_t A, int32_t Al, int r) {
    const I4 w = r < 4 ? tab.grp(G_CWR) : tab.grp(G_X);
    const int32_t wr = r == 0 ? w.x : (r == 1 ? w.y : (r == 2 ? w.z : (r == 3 ? w.w : w.z)));
    return imax(Al + wr, A);
}

// Delete chain, part 2: Din = scanned composite of the previous lane (its last D of this column).
// Returns the D back-pointer bits of the column.
template <class Tab>
PF_HD uint32_t d_final(const Tab &tab, StateQ &s, const int32_t a[P], int32_t Din) {
    uint32_t dirty = 0u, clean = 0u;
    int32_t D = Din & ~TAG_MASK;
    s.Dprev = D;
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const I4 wc = tab.grp(G_WC + q);
        const int32_t c = imax(D + wc.w, a[q]);
        D = c & ~TAG_MASK;
        dirty = mad_u32((uint32_t)c, 32u << (8 * q), dirty);
        clean = mad_u32((uint32_t)D, 32u << (8 * q), clean);
        s.D[q] = D;
    }
    return dirty - clean;
}

// Renormalisation: largest emitting value of the lane ...
PF_HD int32_t lane_max(const StateQ &s) {
    int32_t m = imax(s.M[0], s.I[0]);
#pragma unroll
    for (int q = 1; q < P; ++q) m = imax(imax(s.M[q], s.I[q]), m);
    return m;
}
// ... and the shift by the column maximum mx, clamped from below at the floor (one add-max per value).
PF_HD void renorm(StateQ &s, int32_t mx) {
#pragma unroll
    for (int q = 0; q < P; ++q) {
        s.M[q] = imax(s.M[q] - mx, Q_FLOOR);
        s.I[q] = imax(s.I[q] - mx, Q_FLOOR);
    }
}

// Traceback of one step. slot: 0 M, 1 I, 2 D.  Emitting states step back one column.  One table entry per
// (slot, tag): bits 0-1 positions to step back, 2-3 new slot, 4-5 long-range source (1 X_M, 2 X_D), 6 valid,
// 8-15 / 16-19 index of the edge's weight in the float64 table of the model (pf::K_*): base + q * stride, at the
// lane (= position / 4) of the TARGET state.
#define PQ_BACK(dp, ns, sp, base, stride) ((dp) | ((ns) << 2) | ((sp) << 4) | 64 | ((base) << 8) | ((stride) << 16))
#define PQ_BACK_TABLE                                                                                                  \
    {   /* M: tag 0 cannot occur */ 0,                                                                                  \
        PQ_BACK(1, 2, 0, pf::K_E2, 2), PQ_BACK(0, 0, 1, pf::K_WX, 0), PQ_BACK(2, 0, 0, pf::K_WM2, 1),                   \
        PQ_BACK(0, 1, 0, pf::K_WMR + 3, 4), PQ_BACK(1, 1, 0, pf::K_WMR + 2, 4), PQ_BACK(1, 0, 0, pf::K_WMR + 1, 4),     \
        PQ_BACK(0, 0, 0, pf::K_WMR, 4),                                                                                 \
        /* I */ 0, PQ_BACK(0, 2, 0, pf::K_E2 + 1, 2), PQ_BACK(0, 0, 0, pf::K_WI + 1, 2), PQ_BACK(0, 1, 0, pf::K_WI, 2), \
        0, 0, 0, 0,                                                                                                     \
        /* D */ PQ_BACK(1, 2, 0, pf::K_WH, 1), PQ_BACK(0, 0, 2, pf::K_WX + 1, 0), PQ_BACK(1, 1, 0, pf::K_WD + 1, 2),    \
        PQ_BACK(1, 0, 0, pf::K_WD, 2), 0, 0, 0, 0 }

// `entry` = table[slot * 8 + tag of (word, p, slot)].  Returns false on a pointer that cannot occur.
PF_HD int back_index(uint32_t word, int p, int slot) {
    const uint32_t f = word >> (8 * (p & 3));
    return slot * 8 + (int)(slot == 0 ? (f & 7u) : ((f >> (slot == 1 ? 3 : 5)) & 3u));
}
PF_HD bool back_apply(int entry, const pf::TraceCfg &c, int &p, int &slot, int &t, int &wk) {
    if (!(entry & 64)) return false;
    wk = ((entry >> 8) & 255) + (p & 3) * ((entry >> 16) & 15);
    if (slot < 2) --t;
    const int sp = (entry >> 4) & 3;
    const int ns = (entry >> 2) & 3;
    if (sp == 0) { p -= entry & 3; slot = ns; }
    else if (sp == 1) { p = c.xm_src_p; slot = c.xm_src_slot; }
    else { p = c.xd_src_p; slot = c.xd_src_slot; }
    return true;
}
PF_HD bool back(uint32_t word, const pf::TraceCfg &c, int &p, int &slot, int &t, int &wk) {
    const int table[24] = PQ_BACK_TABLE;
    return back_apply(table[back_index(word, p, slot)], c, p, slot, t, wk);
}

}  // namespace pq
}  // namespace strique
