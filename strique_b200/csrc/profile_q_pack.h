// Host-only: quantises a packed profile model (profile_pack.h, float64 per-lane table) into the tagged fixed-point
// tables of the fixed-point profile Viterbi kernel (profile_q.h).  Refuses models outside the bounds under which
// the int32 arithmetic cannot overflow; those stay on the float64 kernel.
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "profile_pack.h"
#include "profile_q.h"

namespace strique {

struct ProfileQImage {
    std::vector<int32_t> grp;        // [pq::G_TOTAL][32][4]
    std::vector<double> em;          // [pq::E_TOTAL][32][2]
    // for the traceback (see VitProfModelDev)
    std::vector<double> trec;        // [pf::NPOS * 2][8]
    std::vector<uint32_t> tmeta;     // [pf::NPOS * 2]
    std::vector<int32_t> tq;         // [pf::NPOS * 2][2]
    std::vector<int32_t> qtab;       // [pf::K_TOTAL][32]
};

inline bool profile_quantise(const ProfileImage &img, ProfileQImage *out, std::string *why) {
    auto fail = [&](const char *msg) { if (why) *why = msg; return false; };
    const double NINF = -INFINITY;
    bool ok = true;
    // weight with the name of its edge in the low bits
    auto qw = [&](double w, int tag) -> int32_t {
        if (!(w > NINF)) return pq::Q_ABSENT + tag;
        if (w > 0.0 || w < -(double)pq::W_MAX) { ok = false; return pq::Q_ABSENT + tag; }
        return (int32_t)llround(w * (double)(1 << pq::FRAC)) * 8 + tag;
    };
    auto T = [&](int k, int lane) { return img.tab[(size_t)k * 32 + lane]; };
    out->grp.assign((size_t)pq::G_TOTAL * 32 * 4, 0);
    out->em.assign((size_t)pq::E_TOTAL * 32 * 2, 0.0);
    auto G = [&](int g, int lane) { return &out->grp[((size_t)g * 32 + lane) * 4]; };
    int32_t hop_sum[32];
    for (int lane = 0; lane < 32; ++lane) {
        int64_t hs = 0;
        for (int q = 0; q < pq::P; ++q) {
            const int p = lane * pq::P + q;
            const bool dead_m = img.state_id[p * 2] < 0, dead_i = img.state_id[p * 2 + 1] < 0;
            int32_t *wm = G(pq::G_WM + q, lane), *wb = G(pq::G_WB + q, lane), *wc = G(pq::G_WC + q, lane);
            // an unused slot (and START, which the kernel clears after the first column) gets a self loop that decays
            // as fast as anything can (S_STEP per column): nothing else enters it, so it sinks to the floor at every
            // renormalisation and can never drift up relative to the live states, nor overflow
            const int32_t dead_self = -pq::S_STEP * pq::Q_ONE;
            wm[0] = dead_m ? dead_self + 7 : qw(T(pf::K_WMR + q * 4 + 0, lane), 7);
            wm[1] = qw(T(pf::K_WMR + q * 4 + 1, lane), 6);
            wm[2] = qw(T(pf::K_WMR + q * 4 + 2, lane), 5);
            wm[3] = qw(T(pf::K_WMR + q * 4 + 3, lane), 4);
            wb[0] = qw(T(pf::K_WM2 + q, lane), 3);
            wb[1] = qw(T(pf::K_E2 + q * 2, lane), 1);
            wb[2] = dead_i ? dead_self + 3 : qw(T(pf::K_WI + q * 2, lane), 3);
            wb[3] = qw(T(pf::K_WI + q * 2 + 1, lane), 2);
            wc[0] = qw(T(pf::K_E2 + q * 2 + 1, lane), 1);
            wc[1] = qw(T(pf::K_WD + q * 2, lane), 3);
            wc[2] = qw(T(pf::K_WD + q * 2 + 1, lane), 2);
            wc[3] = qw(T(pf::K_WH + q, lane), 0);
            hs += wc[3];
            // emissions: M slot Normal {mu, 1/(2 sigma^2), c0} (Uniform and unused slots: mu = 0, c = 0, c0), I slot Uniform
            const double c0 = T(pf::K_EC + q * 2, lane), ei = T(pf::K_EC + q * 2 + 1, lane);
            if (c0 > (double)pq::C0_MAX || ei > (double)pq::C0_MAX || c0 < -(double)pq::E_MAX || ei < -(double)pq::E_MAX) ok = false;
            const double mu = T(pf::K_EM + q * 2, lane), c = T(pf::K_EM + q * 2 + 1, lane);
            out->em[((size_t)(pq::E_MU + q) * 32 + lane) * 2 + 0] = 2.0 * c * mu;
            out->em[((size_t)(pq::E_MU + q) * 32 + lane) * 2 + 1] = c;
            out->em[((size_t)(pq::E_C0 + q / 2) * 32 + lane) * 2 + (q & 1)] = c0 - c * mu * mu;
            G(pq::G_EI, lane)[q] = pq::to_q16(ei) * 8;
        }
        hop_sum[lane] = (int32_t)std::max<int64_t>(hs, pq::Q_ABSENT);
        G(pq::G_X, lane)[0] = qw(T(pf::K_WX, lane), 2);
        G(pq::G_X, lane)[1] = qw(T(pf::K_WX + 1, lane), 1);
    }
    if (!ok) return fail("weights or emission constants outside the fixed-point bounds");
    // summed hop weights of the cross-lane scan: integer sums are exact, clamped at the weight of an absent edge
    int32_t W[32];
    for (int lane = 0; lane < 32; ++lane) W[lane] = hop_sum[lane];
    for (int r = 0; r < 5; ++r) {
        const int off = 1 << r;
        int32_t Wn[32];
        for (int lane = 0; lane < 32; ++lane) {
            const int32_t w = lane >= off ? W[lane] : pq::Q_ABSENT;
            if (r < 4) G(pq::G_CWR, lane)[r] = w; else G(pq::G_X, lane)[2] = w;
            Wn[lane] = lane >= off ? (int32_t)std::max<int64_t>((int64_t)W[lane - off] + W[lane], pq::Q_ABSENT) : W[lane];
        }
        memcpy(W, Wn, sizeof(W));
    }
    // ---- traceback tables -----------------------------------------------------------------------------------------
    out->trec.assign((size_t)pf::NPOS * 2 * 8, 0.0);
    out->tmeta.assign((size_t)pf::NPOS * 2, 0u);
    out->tq.assign((size_t)pf::NPOS * 2 * 2, 0);
    out->qtab.assign((size_t)pf::K_TOTAL * 32, INT32_MIN);
    auto clean = [&](int32_t w, double orig) { return orig > NINF ? (w & ~pq::TAG_MASK) : INT32_MIN; };
    for (int lane = 0; lane < 32; ++lane) {
        for (int q = 0; q < pq::P; ++q) {
            const int32_t *wm = G(pq::G_WM + q, lane), *wb = G(pq::G_WB + q, lane), *wc = G(pq::G_WC + q, lane);
            auto put = [&](int k, int32_t w) { out->qtab[(size_t)k * 32 + lane] = clean(w, T(k, lane)); };
            for (int d = 0; d < 4; ++d) put(pf::K_WMR + q * 4 + d, wm[d]);
            put(pf::K_WM2 + q, wb[0]);
            put(pf::K_E2 + q * 2, wb[1]);
            put(pf::K_WI + q * 2, wb[2]);
            put(pf::K_WI + q * 2 + 1, wb[3]);
            put(pf::K_E2 + q * 2 + 1, wc[0]);
            put(pf::K_WD + q * 2, wc[1]);
            put(pf::K_WD + q * 2 + 1, wc[2]);
            put(pf::K_WH + q, wc[3]);
            for (int slot = 0; slot < 2; ++slot) {
                const int p = lane * pq::P + q, idx = p * 2 + slot;
                const bool normal = img.em_kind[idx] == 0, uniform = img.em_kind[idx] == 1;
                double *r = &out->trec[(size_t)idx * 8];
                r[0] = normal ? img.em_a[idx] : 0.0;
                r[1] = normal ? img.em_b[idx] : (uniform ? img.em_c[idx] : 0.0);
                r[2] = normal ? img.em_c[idx] : 0.0;
                r[3] = T(slot == 0 ? pf::K_WMR + q * 4 : pf::K_WI + q * 2, lane);
                if (slot == 0) {
                    r[4] = out->em[((size_t)(pq::E_MU + q) * 32 + lane) * 2 + 0];
                    r[5] = out->em[((size_t)(pq::E_MU + q) * 32 + lane) * 2 + 1];
                    r[6] = out->em[((size_t)(pq::E_C0 + q / 2) * 32 + lane) * 2 + (q & 1)];
                }
                out->tmeta[idx] = (uint32_t)img.flags[idx] | ((uint32_t)(img.state_id[idx] & 0xffff) << 16);
                // (an unused slot's self loop carries weight 0 in the forward pass; it is never on a path)
                out->tq[(size_t)idx * 2] = (slot == 0 ? wm[0] : wb[2]) & ~pq::TAG_MASK;
                out->tq[(size_t)idx * 2 + 1] = slot == 1 ? G(pq::G_EI, lane)[q] : 0;
            }
        }
        out->qtab[(size_t)pf::K_WX * 32 + lane] = clean(G(pq::G_X, lane)[0], T(pf::K_WX, lane));
        out->qtab[(size_t)(pf::K_WX + 1) * 32 + lane] = clean(G(pq::G_X, lane)[1], T(pf::K_WX + 1, lane));
    }
    return true;
}

}  // namespace strique
