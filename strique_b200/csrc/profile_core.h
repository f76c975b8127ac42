// Lane arithmetic of the PROFILE Viterbi kernel (viterbi_profile.cu), written once for device and host.
//
// Replaces pomegranate 0.10.0 `HiddenMarkovModel.viterbi` for the linear profile topologies of the
// reference (profileHMM / repeatHMM / flankedRepeatHMM, scripts/STRique.py:201-431) -- float64, one
// add per edge in the reference's association (value + edge weight, then + emission), strict '>' maxima.
//
// Layout: the model is a chain of POSITIONS p = 0..127; position p has an M slot (emitting), an I slot
// (emitting, uniform) and a D slot (silent).  Lane l of the decoding warp owns positions 4l..4l+3, so
// every regular edge is either inside the lane (registers) or comes from the last positions of lane l-1
// (one shuffle).  In-edge template of position p (weights are per position; -inf where absent):
//     M_p(t) = max{ M_p, M_{p-1}, I_{p-1}, I_p, M_{p-2}, X_M | D_{p-1} }(t-1) + w, + emission
//     I_p(t) = max{ I_p, M_p | D_p }(t-1) + w, + emission
//     D_p(t) = max{ M_{p-1}(t), I_{p-1}(t), X_D(t), D_{p-1}(t) } + w                (silent: same column)
// X_M / X_D: one long-range edge each (the repeat loop d2 -> M_0, d1 -> s1), target at an in-lane index 0.
// The part left of '|' only needs emitting values ("E1", computed one step ahead, overlapping the delete
// chain scan), the D sources are added after the scan ("E2").
//
// Back-pointer byte of position q of a lane (4 per 32-bit word, one word per lane per column):
//     bits 0-2  M: E1 winner 0 self, 1 M_{p-1}, 2 I_{p-1}, 3 I_p, 4 M_{p-2}, 5 X_M     bit 3: D_{p-1} won
//     bit  4    I: M_p beat the self loop                                               bit 5: D_p won
//     bits 6-7  D: 0 M_{p-1}, 1 I_{p-1}, 2 X_D, 3 D_{p-1}
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define PF_HD __host__ __device__ __forceinline__
#else
#define PF_HD inline
#endif

namespace strique {
namespace pf {

constexpr int P = 4;            // positions per lane
constexpr int NPOS = 128;       // positions per model (32 lanes x P)

// Per-lane constant table of a packed model: logical layout tab[k * 32 + lane], k below.  Entries the kernel
// reads together sit at (even, odd) indices so that it can fetch them as one 16-byte pair (Tab::pair).
enum : int {
    K_WMR = 0,                  // [q][4]: in-edges of M_p from  M_p (self), M_{p-1}, I_{p-1}, I_p
    K_NREG = K_WMR + P * 4,     // ---- everything above is held in registers by the kernel ----
    K_WI = K_NREG,              // [q] pair: in-edges of I_p from  I_p (self), M_p
    K_E2 = K_WI + P * 2,        // [q] pair: D_{p-1} -> M_p, D_p -> I_p
    K_WM2 = K_E2 + P * 2,       // [q]: M_{p-2} -> M_p
    K_WX = K_WM2 + P,           // pair: X_M weight, X_D weight (targets: in-lane position 0)
    K_WD = K_WX + 2,            // [q] pair: entries of D_p from  M_{p-1}, I_{p-1}
    K_WH = K_WD + P * 2,        // [q]: chain hop D_{p-1} -> D_p
    K_EM = K_WH + P,            // [q] pair: Normal mean (0 for Uniform / dead), 1 / (2 sigma^2) | 0
    K_EC = K_EM + P * 2,        // [q] pair: -log(sigma sqrt(2 pi)) | -log(hi - lo) | 0, I-slot emission -log(hi - lo)
    K_CWR = K_EC + P * 2,       // [5] summed hop weights seen by the rounds of the cross-lane scan (+ 1 pad)
    K_TOTAL = K_CWR + 6,
    K_NAUX = K_TOTAL - K_NREG
};
static_assert(K_NREG % 2 == 0 && K_TOTAL % 2 == 0 && P % 2 == 0, "pairs must start at even indices");

struct Pair {
    double a, b;
};

struct Regs {                   // constants kept in registers
    double wM[P][4];
};

struct State {
    double M[P], I[P], D[P];    // last finished column
    double partM[P], partI[P];  // E1 maxima of the upcoming column
    double Dprev;               // D of the last position of the previous lane (same column as D[])
    uint32_t pbits;             // back-pointer bits of the E1 part
};

PF_HD double ninf() {
#ifdef __CUDA_ARCH__
    return __longlong_as_double(0xfff0000000000000ll);
#else
    union { uint64_t u; double d; } c;
    c.u = 0xfff0000000000000ull;
    return c.d;
#endif
}

template <class Tab>
PF_HD void load_regs(const Tab &tab, Regs &r) {
#pragma unroll
    for (int q = 0; q < P; ++q) {
#pragma unroll
        for (int d = 0; d < 4; ++d) r.wM[q][d] = tab(K_WMR + q * 4 + d);
    }
}

// Fast-path emissions (x inside every Uniform range, not NaN).
template <class Aux>
PF_HD void emissions_fast(const Aux &aux, double x, double eM[P], double eI[P]) {
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const Pair em = aux.pair(K_EM + q * 2), ec = aux.pair(K_EC + q * 2);
        const double dx = x - em.a;
        eM[q] = ec.a - (dx * dx) * em.b;
        eI[q] = ec.b;
    }
}

// General emission of one state (slow path: sample outside a Uniform range, or NaN -> log 1 like pomegranate).
// kind 0 Normal (a = mu, b = c0, c = c2), 1 Uniform (a = lo, b = hi, c = -log(hi - lo)), 2 dead slot.
PF_HD double emission_slow(int kind, double a, double b, double c, double x) {
    if (x != x || kind == 2) return 0.0;
    if (kind == 0) {
        const double dx = x - a;
        return b - (dx * dx) * c;
    }
    return (x >= a && x <= b) ? c : ninf();
}

// E2 + emission: finishes column t from the E1 maxima and the delete states of column t-1.
template <class Aux>
PF_HD uint32_t e2_emit(const Aux &aux, State &s, const double eM[P], const double eI[P]) {
    uint32_t word = s.pbits;
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const Pair w = aux.pair(K_E2 + q * 2);
        const double dsrc = q == 0 ? s.Dprev : s.D[q > 0 ? q - 1 : 0];
        const double cm = dsrc + w.a;
        const double ci = s.D[q] + w.b;
        double bm = s.partM[q], bi = s.partI[q];
        if (cm > bm) { bm = cm; word |= 8u << (8 * q); }
        if (ci > bi) { bi = ci; word |= 32u << (8 * q); }
        s.M[q] = bm + eM[q];
        s.I[q] = bi + eI[q];
    }
    return word;
}

// One node of a first-maximum tournament: the right candidate wins only when strictly greater, so the
// earliest candidate of the sequential order wins ties whatever the shape of the tree.
PF_HD void duel(double &v, uint32_t &arg, double vr, uint32_t argr) {
    const bool gt = vr > v;
    v = gt ? vr : v;
    arg = gt ? argr : arg;
}

// E1 of the next column from the emitting values of the column just finished.
// pM3 / pI3 / pM2: M, I of the last and M of the last-but-one position of the previous lane; xm: X_M source.
// Candidate order (ties: first wins): 0 self, 1 M_{p-1}, 2 I_{p-1}, 3 I_p, 4 M_{p-2}, 5 X_M.
template <class Aux>
PF_HD void e1(const Regs &r, const Aux &aux, State &s, double pM3, double pI3, double pM2, double xm) {
    uint32_t bits = 0u;
    const Pair wx = aux.pair(K_WX);
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const double m1 = q == 0 ? pM3 : s.M[q > 0 ? q - 1 : 0];
        const double i1 = q == 0 ? pI3 : s.I[q > 0 ? q - 1 : 0];
        const double m2 = q == 0 ? pM2 : (q == 1 ? pM3 : s.M[q > 1 ? q - 2 : 0]);
        const Pair wm2 = aux.pair(K_WM2 + (q & ~1));
        const Pair wi = aux.pair(K_WI + q * 2);
        double v01 = s.M[q] + r.wM[q][0];
        uint32_t a01 = 0u;
        duel(v01, a01, m1 + r.wM[q][1], 1u);
        double v23 = i1 + r.wM[q][2];
        uint32_t a23 = 2u;
        duel(v23, a23, s.I[q] + r.wM[q][3], 3u);
        double v45 = m2 + ((q & 1) ? wm2.b : wm2.a);
        uint32_t a45 = 4u;
        if (q == 0) duel(v45, a45, xm + wx.a, 5u);
        duel(v01, a01, v23, a23);
        duel(v01, a01, v45, a45);
        s.partM[q] = v01;
        bits |= a01 << (8 * q);
        double bi = s.I[q] + wi.a;
        const double c = s.M[q] + wi.b;
        if (c > bi) { bi = c; bits |= 16u << (8 * q); }
        s.partI[q] = bi;
    }
    s.pbits = bits;
}

// Delete chain, part 1: entry maxima a[q] of this lane's D states and the lane composite A
// (D of the lane's last position when nothing enters from the previous lane).
template <class Aux>
PF_HD uint32_t d_entry(const Aux &aux, const State &s, double pM3, double pI3, double xd, double a[P], double &A) {
    uint32_t bits = 0u;
    const Pair wx = aux.pair(K_WX);
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const double m1 = q == 0 ? pM3 : s.M[q > 0 ? q - 1 : 0];
        const double i1 = q == 0 ? pI3 : s.I[q > 0 ? q - 1 : 0];
        const Pair wd = aux.pair(K_WD + q * 2);
        double best = m1 + wd.a;
        uint32_t arg = 0u;
        duel(best, arg, i1 + wd.b, 1u);
        if (q == 0) duel(best, arg, xd + wx.b, 2u);
        a[q] = best;
        bits |= arg << (8 * q + 6);
    }
    const Pair h01 = aux.pair(K_WH), h23 = aux.pair(K_WH + 2);
    A = a[0];
#pragma unroll
    for (int q = 1; q < P; ++q) {
        const double t0 = A + (q == 1 ? h01.b : (q == 2 ? h23.a : h23.b));
        A = a[q] >= t0 ? a[q] : t0;
    }
    return bits;
}

// One round of the cross-lane max-plus scan: Al = A of lane - 2^r.
template <class Aux>
PF_HD double d_round(const Aux &aux, double A, double Al, int r) {
    const Pair w = aux.pair(K_CWR + (r & ~1));
    const double t0 = Al + ((r & 1) ? w.b : w.a);
    return A >= t0 ? A : t0;
}

// Delete chain, part 2: Din = scanned composite of the previous lane (its last D of this column).
template <class Aux>
PF_HD uint32_t d_final(const Aux &aux, State &s, const double a[P], double Din) {
    uint32_t bits = 0u;
    double D = Din;
    s.Dprev = Din;
    const Pair h01 = aux.pair(K_WH), h23 = aux.pair(K_WH + 2);
#pragma unroll
    for (int q = 0; q < P; ++q) {
        const double t0 = D + (q == 0 ? h01.a : (q == 1 ? h01.b : (q == 2 ? h23.a : h23.b)));
        if (a[q] >= t0) {
            D = a[q];
        } else {
            D = t0;
            bits |= 3u << (8 * q + 6);
        }
        s.D[q] = D;
    }
    return bits;
}

// Traceback of one step. slot: 0 M, 1 I, 2 D.  Emitting states step back one column.
struct TraceCfg {
    int xm_src_p, xm_src_slot, xd_src_p, xd_src_slot;
};
PF_HD void back(uint32_t word, const TraceCfg &c, int &p, int &slot, int &t) {
    const uint32_t f = word >> (8 * (p & 3));
    if (slot == 0) {
        --t;
        if (f & 8u) { p -= 1; slot = 2; return; }
        switch (f & 7u) {
            case 0: break;
            case 1: p -= 1; break;
            case 2: p -= 1; slot = 1; break;
            case 3: slot = 1; break;
            case 4: p -= 2; break;
            default: p = c.xm_src_p; slot = c.xm_src_slot; break;
        }
    } else if (slot == 1) {
        --t;
        if (f & 32u) slot = 2;
        else if (f & 16u) slot = 0;
    } else {
        switch ((f >> 6) & 3u) {
            case 0: p -= 1; slot = 0; break;
            case 1: p -= 1; slot = 1; break;
            case 2: p = c.xd_src_p; slot = c.xd_src_slot; break;
            default: p -= 1; break;
        }
    }
}

}  // namespace pf
}  // namespace strique
