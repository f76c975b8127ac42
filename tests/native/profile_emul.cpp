// TEST INFRASTRUCTURE (not product code): runs the profile Viterbi kernel's lane arithmetic
// (strique_b200/csrc/profile_core.h) and model packer (profile_pack.h) on the host, 32 simulated lanes in
// lock step, shuffles replaced by array reads -- the same phases in the same order as
// strique_b200/csrc/viterbi_profile.cu.  Lets the CPU test suite check packing, back-pointer encoding and
// traceback against the oracle without a GPU.  Nothing under strique_b200/ links or calls this.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../strique_b200/csrc/profile_pack.h"

using namespace strique;

namespace {
struct TabLane {
    const double *tab;
    int lane;
    double operator()(int k) const { return tab[(size_t)k * 32 + lane]; }
    pf::Pair pair(int k) const { return pf::Pair{(*this)(k), (*this)(k + 1)}; }
};
}  // namespace

extern "C" int strique_test_profile_fits(const strique_hmm_desc *d, char *why, int why_cap) {
    ProfileImage img;
    std::string w;
    const bool ok = profile_pack(d, &img, &w);
    if (why && why_cap > 0) { strncpy(why, w.c_str(), why_cap - 1); why[why_cap - 1] = 0; }
    return ok ? img.np : 0;
}

// status: 0 ok, 1 impossible, 2 internal error, -1 model does not fit
extern "C" int strique_test_profile_emulate(const strique_hmm_desc *d, const double *x, int64_t T, double *logp,
                                            int32_t *n_count, int32_t *t_first, int32_t *t_last, int32_t *path) {
    ProfileImage img;
    if (!profile_pack(d, &img, nullptr)) return -1;
    const double NINF = pf::ninf();
    pf::Regs regs[32];
    pf::State st[32];
    TabLane aux[32];
    for (int l = 0; l < 32; ++l) {
        aux[l] = TabLane{img.tab.data(), l};
        pf::load_regs(aux[l], regs[l]);
        for (int q = 0; q < pf::P; ++q) st[l].M[q] = st[l].I[q] = st[l].D[q] = st[l].partM[q] = st[l].partI[q] = NINF;
        st[l].Dprev = NINF;
        st[l].pbits = 0;
    }
    const int p_start = img.p_off - 1;
    st[p_start / pf::P].M[p_start % pf::P] = 0.0;
    std::vector<uint32_t> bp((size_t)(T + 1) * 32);
    auto val = [&](int p, int slot) {
        const pf::State &s = st[p / pf::P];
        return slot == 0 ? s.M[p % pf::P] : (slot == 1 ? s.I[p % pf::P] : s.D[p % pf::P]);
    };
    auto block = [&](uint32_t *dbits) {     // E1 of the next column + delete chain of this column
        double pM3[32], pI3[32], pM2[32], a[32][pf::P], A[32];
        const double xm = img.trace.xm_src_p >= 0 ? val(img.trace.xm_src_p, img.trace.xm_src_slot) : NINF;
        const double xd = img.trace.xd_src_p >= 0 ? val(img.trace.xd_src_p, img.trace.xd_src_slot) : NINF;
        for (int l = 0; l < 32; ++l) {      // __shfl_up(.., 1): lane 0 keeps its own value
            const int s = l > 0 ? l - 1 : 0;
            pM3[l] = st[s].M[3]; pI3[l] = st[s].I[3]; pM2[l] = st[s].M[2];
        }
        for (int l = 0; l < 32; ++l) {
            pf::e1(regs[l], aux[l], st[l], pM3[l], pI3[l], pM2[l], xm);
            dbits[l] = pf::d_entry(aux[l], st[l], pM3[l], pI3[l], xd, a[l], A[l]);
        }
        for (int r = 0; r < 5; ++r) {
            double An[32];
            for (int l = 0; l < 32; ++l) An[l] = pf::d_round(aux[l], A[l], A[l >= (1 << r) ? l - (1 << r) : l], r);
            memcpy(A, An, sizeof(A));
        }
        double Din[32];
        for (int l = 0; l < 32; ++l) Din[l] = A[l > 0 ? l - 1 : 0];
        for (int l = 0; l < 32; ++l) dbits[l] |= pf::d_final(aux[l], st[l], a[l], Din[l]);
    };
    uint32_t dbits[32];
    block(dbits);
    for (int l = 0; l < 32; ++l) bp[l] = dbits[l];
    for (int64_t t = 1; t <= T; ++t) {
        const double xt = x[t - 1];
        uint32_t word[32];
        const bool fast = xt >= img.lo && xt <= img.hi;
        for (int l = 0; l < 32; ++l) {
            double eM[pf::P], eI[pf::P];
            if (fast) {
                pf::emissions_fast(aux[l], xt, eM, eI);
            } else {
                for (int q = 0; q < pf::P; ++q) {
                    const int i0 = (l * pf::P + q) * 2;
                    eM[q] = pf::emission_slow(img.em_kind[i0], img.em_a[i0], img.em_b[i0], img.em_c[i0], xt);
                    eI[q] = pf::emission_slow(img.em_kind[i0 + 1], img.em_a[i0 + 1], img.em_b[i0 + 1], img.em_c[i0 + 1], xt);
                }
            }
            word[l] = pf::e2_emit(aux[l], st[l], eM, eI);
        }
        block(dbits);
        for (int l = 0; l < 32; ++l) bp[(size_t)t * 32 + l] = word[l] | dbits[l];
    }
    double best = NINF;
    int barg = -1;
    for (int e = 0; e < img.n_end; ++e) {
        const double cand = val(img.end_p[e], img.end_slot[e]) + img.end_w[e];
        if (cand > best) { best = cand; barg = e; }
    }
    *logp = best;
    *n_count = 0; *t_first = -1; *t_last = -1;
    if (!(best > NINF) || barg < 0) return 1;
    int p = img.end_p[barg], slot = img.end_slot[barg], t = (int)T;
    long long guard = (long long)(T + 2) * (pf::NPOS + 2);
    while (!(slot == 0 && p == p_start)) {
        if (--guard < 0 || p < 0 || p >= pf::NPOS || t < 0) return 2;
        if (slot < 2) {
            if (t < 1) return 2;
            const int idx = p * 2 + slot;
            if (img.state_id[idx] < 0) return 2;
            if (img.flags[idx] & 1) ++*n_count;
            if (img.flags[idx] & 2) { if (*t_last < 0) *t_last = t - 1; *t_first = t - 1; }
            if (path) path[t - 1] = img.state_id[idx];
        }
        pf::back(bp[(size_t)t * 32 + p / pf::P], img.trace, p, slot, t);
    }
    return t == 0 ? 0 : 2;
}

// Diagnostic: how far below the column maximum the best path runs (float64), per column -- the quantity the
// fixed-point kernel's floor (profile_q.h: Q_FLOOR_NAT) must stay clear of for the kernel to keep the read.  deficit[t - 1] for columns 1..T.
extern "C" int strique_test_profile_deficit(const strique_hmm_desc *d, const double *x, int64_t T, double *deficit) {
    ProfileImage img;
    if (!profile_pack(d, &img, nullptr)) return -1;
    const double NINF = pf::ninf();
    pf::Regs regs[32];
    pf::State st[32];
    TabLane aux[32];
    for (int l = 0; l < 32; ++l) {
        aux[l] = TabLane{img.tab.data(), l};
        pf::load_regs(aux[l], regs[l]);
        for (int q = 0; q < pf::P; ++q) st[l].M[q] = st[l].I[q] = st[l].D[q] = st[l].partM[q] = st[l].partI[q] = NINF;
        st[l].Dprev = NINF;
        st[l].pbits = 0;
    }
    const int p_start = img.p_off - 1;
    st[p_start / pf::P].M[p_start % pf::P] = 0.0;
    std::vector<uint32_t> bp((size_t)(T + 1) * 32);
    std::vector<double> vals((size_t)(T + 1) * pf::NPOS * 3, NINF);
    auto val = [&](int p, int slot) {
        const pf::State &s = st[p / pf::P];
        return slot == 0 ? s.M[p % pf::P] : (slot == 1 ? s.I[p % pf::P] : s.D[p % pf::P]);
    };
    auto block = [&](uint32_t *dbits) {
        double pM3[32], pI3[32], pM2[32], a[32][pf::P], A[32];
        const double xm = img.trace.xm_src_p >= 0 ? val(img.trace.xm_src_p, img.trace.xm_src_slot) : NINF;
        const double xd = img.trace.xd_src_p >= 0 ? val(img.trace.xd_src_p, img.trace.xd_src_slot) : NINF;
        for (int l = 0; l < 32; ++l) {
            const int s = l > 0 ? l - 1 : 0;
            pM3[l] = st[s].M[3]; pI3[l] = st[s].I[3]; pM2[l] = st[s].M[2];
        }
        for (int l = 0; l < 32; ++l) {
            pf::e1(regs[l], aux[l], st[l], pM3[l], pI3[l], pM2[l], xm);
            dbits[l] = pf::d_entry(aux[l], st[l], pM3[l], pI3[l], xd, a[l], A[l]);
        }
        for (int r = 0; r < 5; ++r) {
            double An[32];
            for (int l = 0; l < 32; ++l) An[l] = pf::d_round(aux[l], A[l], A[l >= (1 << r) ? l - (1 << r) : l], r);
            memcpy(A, An, sizeof(A));
        }
        double Din[32];
        for (int l = 0; l < 32; ++l) Din[l] = A[l > 0 ? l - 1 : 0];
        for (int l = 0; l < 32; ++l) dbits[l] |= pf::d_final(aux[l], st[l], a[l], Din[l]);
    };
    auto snapshot = [&](int64_t t) {
        for (int p = 0; p < pf::NPOS; ++p)
            for (int slot = 0; slot < 3; ++slot) vals[((size_t)t * pf::NPOS + p) * 3 + slot] = val(p, slot);
    };
    uint32_t dbits[32];
    block(dbits);
    for (int l = 0; l < 32; ++l) bp[l] = dbits[l];
    snapshot(0);
    for (int64_t t = 1; t <= T; ++t) {
        uint32_t word[32];
        for (int l = 0; l < 32; ++l) {
            double eM[pf::P], eI[pf::P];
            pf::emissions_fast(aux[l], x[t - 1], eM, eI);
            word[l] = pf::e2_emit(aux[l], st[l], eM, eI);
        }
        block(dbits);
        for (int l = 0; l < 32; ++l) bp[(size_t)t * 32 + l] = word[l] | dbits[l];
        snapshot(t);
    }
    double best = NINF;
    int barg = -1;
    for (int e = 0; e < img.n_end; ++e) {
        const double cand = val(img.end_p[e], img.end_slot[e]) + img.end_w[e];
        if (cand > best) { best = cand; barg = e; }
    }
    if (barg < 0) return 1;
    int p = img.end_p[barg], slot = img.end_slot[barg], t = (int)T;
    while (!(slot == 0 && p == p_start)) {
        if (p < 0 || p >= pf::NPOS || t < 0) return 2;
        if (slot < 2 && t >= 1) {
            double mx = NINF;
            for (int q = 0; q < pf::NPOS; ++q)
                for (int sl = 0; sl < 2; ++sl) mx = std::max(mx, vals[((size_t)t * pf::NPOS + q) * 3 + sl]);
            deficit[t - 1] = mx - vals[((size_t)t * pf::NPOS + p) * 3 + slot];
        }
        pf::back(bp[(size_t)t * 32 + p / pf::P], img.trace, p, slot, t);
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------------
// Fixed-point kernel (strique_b200/csrc/profile_q.h, profile_q_pack.h; kernel viterbi_profile_q.cu): same phases in
// the same order, 32 simulated lanes.  log p is the float64 re-score of the decoded path; `vfwd` receives the
// forward pass's own value of that path (fixed point, converted) for the consistency check the kernel makes.
// status: 0 ok, 1 impossible, 2 internal error, 3 declined (sample outside the fast emission range), -1 / -2 model
// does not fit the profile layout / the fixed-point bounds
#include "../../strique_b200/csrc/profile_q_pack.h"

namespace {
struct QTabLane {
    const int32_t *grp_;
    const double *em_;
    int lane;
    pq::I4 grp(int g) const {
        const int32_t *p = grp_ + ((size_t)g * 32 + lane) * 4;
        return pq::I4{p[0], p[1], p[2], p[3]};
    }
    pf::Pair dpair(int k) const {
        const double *p = em_ + ((size_t)k * 32 + lane) * 2;
        return pf::Pair{p[0], p[1]};
    }
};
}  // namespace

extern "C" int strique_test_profile_q_emulate(const strique_hmm_desc *d, const double *x, int64_t T, double *logp,
                                              int32_t *n_count, int32_t *t_first, int32_t *t_last, int32_t *path,
                                              double *vfwd) {
    ProfileImage img;
    if (!profile_pack(d, &img, nullptr)) return -1;
    ProfileQImage qi;
    if (!profile_quantise(img, &qi, nullptr)) return -2;
    pq::RegsQ regs[32];
    pq::StateQ st[32];
    QTabLane tab[32];
    for (int l = 0; l < 32; ++l) {
        tab[l] = QTabLane{qi.grp.data(), qi.em.data(), l};
        for (int q = 0; q < pq::P; ++q) {
            const pq::I4 g = tab[l].grp(pq::G_WM + q);
            regs[l].wM[q][0] = g.x; regs[l].wM[q][1] = g.y; regs[l].wM[q][2] = g.z; regs[l].wM[q][3] = g.w;
            st[l].M[q] = st[l].I[q] = st[l].D[q] = st[l].partM[q] = st[l].partI[q] = pq::Q_FLOOR;
        }
        st[l].Dprev = pq::Q_FLOOR;
    }
    const int p_start = img.p_off - 1;
    st[p_start / pq::P].M[p_start % pq::P] = 0;
    std::vector<uint32_t> bp((size_t)(T + 1) * 32);
    auto val = [&](int p, int slot) {
        const pq::StateQ &s = st[p / pq::P];
        return slot == 0 ? s.M[p % pq::P] : (slot == 1 ? s.I[p % pq::P] : s.D[p % pq::P]);
    };
    auto block = [&](uint32_t *dbits) {
        int32_t pM3[32], pI3[32], pM2[32], a[32][pq::P], A[32];
        const int32_t xm = img.trace.xm_src_p >= 0 ? val(img.trace.xm_src_p, img.trace.xm_src_slot) : pq::Q_FLOOR;
        const int32_t xd = img.trace.xd_src_p >= 0 ? val(img.trace.xd_src_p, img.trace.xd_src_slot) : pq::Q_FLOOR;
        for (int l = 0; l < 32; ++l) {
            const int s = l > 0 ? l - 1 : 0;
            pM3[l] = st[s].M[3]; pI3[l] = st[s].I[3]; pM2[l] = st[s].M[2];
        }
        for (int l = 0; l < 32; ++l) {
            pq::e1(regs[l], tab[l], st[l], pM3[l], pI3[l], pM2[l], xm);
            pq::d_entry(tab[l], st[l], pM3[l], pI3[l], xd, a[l], A[l]);
        }
        for (int r = 0; r < 5; ++r) {
            int32_t An[32];
            for (int l = 0; l < 32; ++l) An[l] = pq::d_round(tab[l], A[l], A[l >= (1 << r) ? l - (1 << r) : l], r);
            memcpy(A, An, sizeof(A));
        }
        for (int l = 31; l >= 0; --l) dbits[l] = pq::d_final(tab[l], st[l], a[l], A[l > 0 ? l - 1 : 0]);
    };
    uint32_t dbits[32];
    block(dbits);
    for (int l = 0; l < 32; ++l) bp[l] = dbits[l];
    int64_t off = 0;                                       // sum of the subtracted column maxima
    for (int64_t t = 1; t <= T; ++t) {
        const double xt = x[t - 1];
        if (!(xt >= img.lo && xt <= img.hi)) return 3;
        uint32_t word[32];
        for (int l = 0; l < 32; ++l) {
            int32_t eM[pq::P];
            pq::emissions_q(tab[l], xt, eM);
            word[l] = pq::e2_emit(tab[l], st[l], eM);
        }
        if ((t % pq::R_NORM) == 0 || t == 1) {
            if (t == 1) st[p_start / pq::P].M[p_start % pq::P] = pq::Q_FLOOR;
            int32_t mx = pq::lane_max(st[0]);
            for (int l = 1; l < 32; ++l) mx = pq::imax(mx, pq::lane_max(st[l]));
            for (int l = 0; l < 32; ++l) pq::renorm(st[l], mx);
            off += mx;
        }
        block(dbits);
        for (int l = 0; l < 32; ++l) bp[(size_t)t * 32 + l] = word[l] | dbits[l];
    }
    const double UNIT = 1.0 / (double)pq::Q_ONE;
    double best = -INFINITY;
    int barg = -1;
    int64_t vend = 0;
    for (int e = 0; e < img.n_end; ++e) {
        const int32_t v = val(img.end_p[e], img.end_slot[e]);
        const double cand = (double)v * UNIT + img.end_w[e];
        if (cand > best) { best = cand; barg = e; vend = v; }
    }
    *n_count = 0; *t_first = -1; *t_last = -1;
    *logp = -INFINITY;
    if (barg < 0 || !(best > -INFINITY)) return 1;
    *vfwd = best + (double)off * UNIT;
    double acc = img.end_w[barg];
    int64_t qacc = 0;                     // the path's fixed-point score, re-added term by term
    bool clamped = false;
    int p = img.end_p[barg], slot = img.end_slot[barg], t = (int)T;
    long long guard = (long long)(T + 2) * (pf::NPOS + 2);
    while (!(slot == 0 && p == p_start)) {
        if (--guard < 0 || p < 0 || p >= pf::NPOS || t < 0) return 2;
        if (slot < 2) {
            if (t < 1) return 2;
            const int idx = p * 2 + slot;
            if (img.state_id[idx] < 0) return 2;
            if (img.flags[idx] & 1) ++*n_count;
            if (img.flags[idx] & 2) { if (*t_last < 0) *t_last = t - 1; *t_first = t - 1; }
            if (path) path[t - 1] = img.state_id[idx];
            acc += pf::emission_slow(img.em_kind[idx], img.em_a[idx], img.em_b[idx], img.em_c[idx], x[t - 1]);
            const double *rec = &qi.trec[(size_t)idx * 8];
            if (slot == 0) {
                const int32_t e16 = pq::emission_q16(rec[4], rec[5], rec[6], x[t - 1], x[t - 1] * x[t - 1]);
                clamped |= e16 == pq::E_MIN16;
                qacc += (int64_t)e16 * 8;
            } else {
                qacc += qi.tq[(size_t)idx * 2 + 1];
            }
        }
        int wk = -1;
        const int lane = p / pq::P;
        if (!pq::back(bp[(size_t)t * 32 + lane], img.trace, p, slot, t, wk)) return 2;
        acc += img.tab[(size_t)wk * 32 + lane];
        const int32_t qw = qi.qtab[(size_t)wk * 32 + lane];
        clamped |= qw == INT32_MIN;
        qacc += qw;
    }
    if (t != 0) return 2;
    *logp = acc;
    // the forward value bounds the path's fixed-point score from above; equality = no clamp touched the path
    if (clamped || qacc != vend + off) return 3;
    return 0;
}
