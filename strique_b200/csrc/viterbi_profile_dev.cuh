// Device code of the float64 profile Viterbi (forward pass, END edges, traceback), shared by its own kernel
// (viterbi_profile.cu: model table in shared memory) and by the fixed-point kernel (viterbi_profile_q.cu), which
// decodes the few sequences it cannot vouch for in float64 on the spot (model table read from global memory / L1).
//
// Replaces pomegranate 0.10.0 `HiddenMarkovModel.viterbi` behind flankedRepeatHMM.count_repeats (reference
// scripts/STRique.py:433-441, 374-378; topology 201-431): float64 throughout, one add per edge, strict-'>' maxima
// (profile_core.h holds the lane arithmetic).
#pragma once
#include <math.h>

#include "profile_core.h"
#include "viterbi.cuh"

namespace strique {
namespace f64 {

constexpr int PROF_STAGE_ROWS = 32;
constexpr int PROF_AUX_BYTES = pf::K_NAUX * 32 * 8; // table of the CTA's current model (shared by its warps)
constexpr int PROF_STAGE_BYTES = PROF_STAGE_ROWS * 32 * 4;   // per warp: back-pointer rows of the traceback
static_assert(PROF_STAGE_BYTES >= 3 * pf::NPOS * 8, "the END gather reuses the stage area");

struct AuxShared {                                  // entries k >= K_NREG of this lane, pair-interleaved
    const double2 *base;                            // &aux[lane]; pair j of lane l at aux[j * 32 + l]
    __device__ __forceinline__ pf::Pair pair(int k) const {
        const double2 v = base[((k - pf::K_NREG) >> 1) * 32];
        return pf::Pair{v.x, v.y};
    }
};
struct TabGlobal {
    const double *base;                             // &tab[lane]
    __device__ __forceinline__ double operator()(int k) const { return __ldg(base + k * 32); }
};
struct AuxGlobal {                                  // the same entries straight from the model's table in global memory
    const double *base;                             // &tab[lane]
    __device__ __forceinline__ pf::Pair pair(int k) const { return pf::Pair{__ldg(base + k * 32), __ldg(base + (k + 1) * 32)}; }
};

constexpr unsigned FULL = 0xffffffffu;

struct ModelScalars {                               // warp-uniform
    int p_start, xlane, xq, xm_slot, xd_slot;
    double lo, hi;
};

struct SeqCtx {
    int seq, T;
    int64_t xo;
    const double *x;
    uint32_t *bp;
};

__device__ __forceinline__ SeqCtx seq_ctx(const VitProfBatch &b, int seq) {
    SeqCtx c;
    c.seq = seq;
    c.xo = b.x_off[seq];
    c.T = (int)(b.x_off[seq + 1] - c.xo);
    c.x = b.x + c.xo;
    c.bp = b.bp + b.bp_off[seq];
    return c;
}

__device__ __forceinline__ void init_state(pf::State &S, int lane, int p_start) {
    const double NINF = pf::ninf();
#pragma unroll
    for (int q = 0; q < pf::P; ++q) {
        S.M[q] = (lane * pf::P + q == p_start) ? 0.0 : NINF;   // START: value 0 before the first sample only
        S.I[q] = S.D[q] = S.partM[q] = S.partI[q] = NINF;
    }
    S.Dprev = NINF;
    S.pbits = 0u;
}

// E1 of the next column + delete chain of the column just finished, for NS sequences decoded side by side (one
// basic block: the constants of the model are fetched once and the NS dependency chains interleave).
// XQ = in-lane index of the position that feeds the repeat loop (compile time: no select chain per column).
template <int NS, int XQ, class Aux>
__device__ __forceinline__ void block(const pf::Regs &R, const Aux &aux, const ModelScalars &ms,
                                      pf::State (&S)[NS], uint32_t (&bits)[NS]) {
    double pM3[NS], pI3[NS], pM2[NS], xm[NS], xd[NS], a[NS][pf::P], A[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) {
        pM3[i] = __shfl_up_sync(FULL, S[i].M[3], 1);
        pI3[i] = __shfl_up_sync(FULL, S[i].I[3], 1);
        pM2[i] = __shfl_up_sync(FULL, S[i].M[2], 1);
        const double vm = S[i].M[XQ], vi = S[i].I[XQ];
        xm[i] = __shfl_sync(FULL, ms.xm_slot ? vi : vm, ms.xlane);
        xd[i] = __shfl_sync(FULL, ms.xd_slot ? vi : vm, ms.xlane);
    }
#pragma unroll
    for (int i = 0; i < NS; ++i) pf::e1(R, aux, S[i], pM3[i], pI3[i], pM2[i], xm[i]);
#pragma unroll
    for (int i = 0; i < NS; ++i) bits[i] = pf::d_entry(aux, S[i], pM3[i], pI3[i], xd[i], a[i], A[i]);
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        double Al[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) Al[i] = __shfl_up_sync(FULL, A[i], 1 << r);
#pragma unroll
        for (int i = 0; i < NS; ++i) A[i] = pf::d_round(aux, A[i], Al[i], r);
    }
    double Din[NS];
#pragma unroll
    for (int i = 0; i < NS; ++i) Din[i] = __shfl_up_sync(FULL, A[i], 1);
#pragma unroll
    for (int i = 0; i < NS; ++i) bits[i] |= pf::d_final(aux, S[i], a[i], Din[i]);
}

// columns t0 .. t1 of NS sequences (t1 <= T of each); xcur[i] = sample t0 - 1 on entry, sample t1 on exit
template <int NS, int XQ, class Aux>
__device__ __forceinline__ void forward(const pf::Regs &R, const Aux &aux, const ModelScalars &ms,
                                        const VitProfModelDev &m, const int lane, pf::State (&S)[NS],
                                        const SeqCtx (&c)[NS], double (&xcur)[NS], const int t0, const int t1) {
#pragma unroll 1
    for (int t = t0; t <= t1; ++t) {
        double xnext[NS], eM[NS][pf::P], eI[NS][pf::P];
        uint32_t word[NS], dbits[NS];
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            xnext[i] = t < c[i].T ? __ldg(c[i].x + t) : 0.0;
            pf::emissions_fast(aux, xcur[i], eM[i], eI[i]);
            if (!(xcur[i] >= ms.lo && xcur[i] <= ms.hi)) {   // outside a Uniform range or NaN: general form (rare)
#pragma unroll
                for (int q = 0; q < pf::P; ++q) {
                    const int i0 = (lane * pf::P + q) * 2;
                    eM[i][q] = pf::emission_slow(m.em_kind[i0], m.em_a[i0], m.em_b[i0], m.em_c[i0], xcur[i]);
                    eI[i][q] = pf::emission_slow(m.em_kind[i0 + 1], m.em_a[i0 + 1], m.em_b[i0 + 1], m.em_c[i0 + 1], xcur[i]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < NS; ++i) word[i] = pf::e2_emit(aux, S[i], eM[i], eI[i]);
        block<NS, XQ, Aux>(R, aux, ms, S, dbits);
#pragma unroll
        for (int i = 0; i < NS; ++i) {
            c[i].bp[(size_t)t * 32 + lane] = word[i] | dbits[i];
            xcur[i] = xnext[i];
        }
    }
}

// END edges: log p = max(v[T][src] + w), first maximum
__device__ __forceinline__ void end_edges(const VitProfModelDev &m, const pf::State &S, uint32_t *stage, const int lane,
                                          double &best_out, int &barg_out) {
    const double NINF = pf::ninf();
    __syncwarp();
    double *vals = reinterpret_cast<double *>(stage);
#pragma unroll
    for (int q = 0; q < pf::P; ++q) {
        vals[lane * pf::P + q] = S.M[q];
        vals[pf::NPOS + lane * pf::P + q] = S.I[q];
        vals[2 * pf::NPOS + lane * pf::P + q] = S.D[q];
    }
    __syncwarp();
    double best = NINF;
    int barg = -1;
    if (lane < m.n_end) {
        const double cand = vals[m.end_slot[lane] * pf::NPOS + m.end_p[lane]] + m.end_w[lane];
        if (cand > best) { best = cand; barg = lane; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ob = __shfl_down_sync(FULL, best, off);
        const int oa = __shfl_down_sync(FULL, barg, off);
        if (ob > best || (ob == best && oa >= 0 && (barg < 0 || oa < barg))) { best = ob; barg = oa; }
    }
    best_out = __shfl_sync(FULL, best, 0);
    barg_out = __shfl_sync(FULL, barg, 0);
    __syncwarp();
}

// traceback (all lanes walk in lock step; lane 0 / lane i write) and the result record of one sequence
__device__ __noinline__ inline void traceback(const VitProfBatch &b, const VitProfModelDev &m, const SeqCtx &c, uint32_t *stage,
                                       const int lane, const int p_start, const double best, const int barg) {
    const double NINF = pf::ninf();
    const int T = c.T;
    const uint32_t *bp = c.bp;
    VitResult r;
    r.logp = best; r.n_count = 0; r.t_first = -1; r.t_last = -1; r.pattern_len = 0; r.status = 0; r.reserved = 0;
    if (!(best > NINF) || barg < 0) {
        r.status = 1;
    } else {
        const pf::TraceCfg tc = m.trace;
        int p = m.end_p[barg], slot = m.end_slot[barg], t = T;
        uint8_t *pat = b.pattern ? b.pattern + c.xo : nullptr;
        uint16_t *path = b.path ? b.path + c.xo : nullptr;
        bool in_group = false;
        uint8_t last_mod = '0';
        int plen = 0;
        int stage_lo = T + 1;                 // rows [stage_lo, stage_lo + 32) are staged
        long long guard = (long long)(T + 2) * (pf::NPOS + 2);
        while (!(slot == 0 && p == p_start)) {
            if (--guard < 0 || p < 0 || p >= pf::NPOS || t < 0) { r.status = 2; break; }
            if (t < stage_lo) {
                // stage the next rows: 16-byte async copies, all in flight at once
                __syncwarp();
                stage_lo = t - (PROF_STAGE_ROWS - 1) > 0 ? t - (PROF_STAGE_ROWS - 1) : 0;
                const uint4 *src = reinterpret_cast<const uint4 *>(bp + (size_t)stage_lo * 32);
                const int nvec = (t - stage_lo + 1) * 8;
                const uint32_t sdst = (uint32_t)__cvta_generic_to_shared(stage);
#pragma unroll
                for (int i = 0; i < PROF_STAGE_ROWS * 8 / 32; ++i)
                    if (lane + 32 * i < nvec)
                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst + (lane + 32 * i) * 16),
                                     "l"(src + lane + 32 * i)
                                     : "memory");
                asm volatile("cp.async.wait_all;" ::: "memory");
                __syncwarp();
            }
            if (slot == 2) {                  // silent delete state: same column
                pf::back(stage[(t - stage_lo) * 32 + (p >> 2)], tc, p, slot, t);
                continue;
            }
            if (t < 1) { r.status = 2; break; }
            // Emitting state (p, slot) at column t.  Samples dwell in a state, so most pointers are self loops:
            // lane i looks at column t - i, the warp skips the whole run of self loops at once and then takes
            // the first other pointer (all lanes keep the same cursor; lane 0 / lane i write the outputs).
            const int ti = t - lane;
            const bool valid = ti >= stage_lo && ti >= 1;
            const uint32_t wfull = valid ? stage[(ti - stage_lo) * 32 + (p >> 2)] : 0u;
            const uint32_t f = wfull >> (8 * (p & 3));
            const bool self = valid && (slot == 0 ? (f & 0xfu) == 0u : (f & 0x30u) == 0u);
            const unsigned other = ~__ballot_sync(FULL, self);
            const int k = other ? __ffs(other) - 1 : 32;               // columns t .. t-k+1 are self loops
            const bool step = k < 32 && ((__ballot_sync(FULL, valid) >> k) & 1u);   // column t-k is staged
            const int visits = k + (step ? 1 : 0);                     // >= 1: column t itself is staged
            const int idx = p * 2 + slot;
            const unsigned fl = m.flags[idx];
            if (fl & HMM_FLAG_COUNT) r.n_count += visits;
            if (fl & HMM_FLAG_REPEAT) { if (r.t_last < 0) r.t_last = t - 1; r.t_first = t - visits; }
            if (fl & HMM_FLAG_SEP) {
                if (in_group) { if (pat && lane == 0) pat[T - 1 - plen] = last_mod; ++plen; in_group = false; }
            } else {
                in_group = true;
                last_mod = (fl & HMM_FLAG_MOD) ? '1' : '0';
            }
            if (path && lane < visits) path[t - 1 - lane] = (uint16_t)m.state_id[idx];
            t -= k;
            if (step) {
                const uint32_t w = __shfl_sync(FULL, wfull, k);
                pf::back(w, tc, p, slot, t);
            }
        }
        if (in_group) { if (pat && lane == 0) pat[T - 1 - plen] = last_mod; ++plen; }
        if (r.status == 0 && t != 0) r.status = 2;
        r.pattern_len = plen;
    }
    if (lane == 0) b.res[c.seq] = r;
    __syncwarp();
}

}  // namespace f64
}  // namespace strique
