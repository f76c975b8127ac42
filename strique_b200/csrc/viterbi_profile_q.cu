// Fixed-point profile Viterbi: the throughput kernel of boundary #2 for the reference's linear profile HMMs.
//
// Replaces pomegranate 0.10.0 `HiddenMarkovModel.viterbi` behind flankedRepeatHMM.count_repeats (reference
// scripts/STRique.py:433-441, 374-378; topology 201-431).  Same mapping as the float64 kernel
// (viterbi_profile.cu: one warp per sequence, lane l owns profile positions 4l..4l+3, neighbours by shuffle,
// delete chain as a max-plus scan, one back-pointer word per lane per column), different arithmetic
// (profile_q.h): tagged int32 fixed-point scores, so that ONE VIADDMNMX per in-edge relaxes the edge and carries
// the winner's name -- no compare / select chains, no 64-bit shuffles, half the registers.  The decoded path is
// re-scored in float64 during the traceback (log p = exact score of the returned path).  The floor clamp of the
// renormalisation only ever raises values, so every forward value is an upper bound of the true fixed-point score;
// the traceback re-adds the quantised weights along the decoded path in int64, and equality with the forward value
// (plus: no clamped emission, no absent edge on the path) PROVES the path optimal for the quantised model.  A
// sequence the pass cannot vouch for (that proof fails, a sample outside the fast emission range, END unreachable)
// is decoded again in float64 by the same warp before it takes its next sequence (decode_float64; reserved = 1).
#include <math.h>

#include <algorithm>

#include "profile_q_pack.h"
#include "viterbi_profile_dev.cuh"

namespace strique {

namespace {

#ifndef PROFQ_WARPS_PER_CTA
#define PROFQ_WARPS_PER_CTA 4
#endif
#ifndef PROFQ_CTAS
#define PROFQ_CTAS 3
#endif
constexpr int PROFQ_WARPS = PROFQ_WARPS_PER_CTA;
constexpr int PROFQ_CTAS_PER_SM = PROFQ_CTAS;
constexpr int PROFQ_STAGE_ROWS = 32;
constexpr int PROFQ_BUF_BYTES = PROFQ_STAGE_ROWS * 32 * 4 + PROFQ_STAGE_ROWS * 8;   // 32 back-pointer rows + their samples
// per warp: two buffers (the traceback prefetches), or -- while the warp decodes a sequence in float64 -- that
// decoder's model table and its staging area
constexpr int PROFQ_STAGE_BYTES = 2 * PROFQ_BUF_BYTES > f64::PROF_AUX_BYTES + f64::PROF_STAGE_BYTES
                                      ? 2 * PROFQ_BUF_BYTES : f64::PROF_AUX_BYTES + f64::PROF_STAGE_BYTES;
static_assert(PROFQ_STAGE_BYTES >= 3 * pf::NPOS * 4, "the END gather reuses the stage area");

// This lane's constants of the warp's current model, ALL in registers (59 weights, 12 float64 emission constants):
// the first version kept them in a shared-memory table like the float64 kernel and was bound by the shared-memory
// pipe (17 LDS.128 = 68 wavefronts per column and warp, 73 % of the pipe with 16 warps per SM, ncu vq_r02a).
struct TabQ {
    int32_t g[pq::G_TOTAL][4];
    double e[pq::E_TOTAL][2];
    __device__ __forceinline__ pq::I4 grp(int k) const { return pq::I4{g[k][0], g[k][1], g[k][2], g[k][3]}; }
    __device__ __forceinline__ pf::Pair dpair(int k) const { return pf::Pair{e[k][0], e[k][1]}; }
};

constexpr unsigned FULL = 0xffffffffu;

__constant__ int c_back[24] = PQ_BACK_TABLE;        // traceback decode (profile_q.h)

struct ModelScalars {                                              // warp-uniform
    int p_start, xlane, xm_slot, xd_slot;
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

// E1 of the next column + delete chain of the column just finished + the emissions of the next column.
// XQ = in-lane index of the position that feeds the repeat loop (compile time).
// A warp issues in order, so what follows a shuffle in program order waits for it: the E1 relaxations and the
// float64 emissions -- neither depends on the scan -- are written BETWEEN the rounds of the cross-lane scan, where
// they cover the shuffle latencies (ptxas keeps this order; it did not hoist them there by itself).
template <int XQ>
__device__ __forceinline__ uint32_t block(const TabQ &tab, const ModelScalars &ms, pq::StateQ &S, const double x,
                                          int32_t (&eM)[pq::P]) {
    pq::RegsQ R;
#pragma unroll
    for (int q = 0; q < pq::P; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k) R.wM[q][k] = tab.g[pq::G_WM + q][k];
    const int32_t pM3 = __shfl_up_sync(FULL, S.M[3], 1);
    const int32_t pI3 = __shfl_up_sync(FULL, S.I[3], 1);
    const int32_t vm = S.M[XQ], vi = S.I[XQ];
    const int32_t xd = __shfl_sync(FULL, ms.xd_slot ? vi : vm, ms.xlane);
    const int32_t pM2 = __shfl_up_sync(FULL, S.M[2], 1);
    const int32_t xm = __shfl_sync(FULL, ms.xm_slot ? vi : vm, ms.xlane);
    const double x2 = x * x;
    eM[0] = pq::emission_q(tab, x, x2, 0);
    int32_t a[pq::P], A;
    pq::d_entry(tab, S, pM3, pI3, xd, a, A);
    int32_t Al = __shfl_up_sync(FULL, A, 1);
    pq::e1_q(R, tab, S, pM3, pI3, pM2, xm, 0);
    A = pq::d_round(tab, A, Al, 0);
    Al = __shfl_up_sync(FULL, A, 2);
    pq::e1_q(R, tab, S, pM3, pI3, pM2, xm, 1);
    A = pq::d_round(tab, A, Al, 1);
    Al = __shfl_up_sync(FULL, A, 4);
    pq::e1_q(R, tab, S, pM3, pI3, pM2, xm, 2);
    A = pq::d_round(tab, A, Al, 2);
    Al = __shfl_up_sync(FULL, A, 8);
    pq::e1_q(R, tab, S, pM3, pI3, pM2, xm, 3);
    A = pq::d_round(tab, A, Al, 3);
    Al = __shfl_up_sync(FULL, A, 16);
    eM[1] = pq::emission_q(tab, x, x2, 1);
    eM[2] = pq::emission_q(tab, x, x2, 2);
    A = pq::d_round(tab, A, Al, 4);
    const int32_t Din = __shfl_up_sync(FULL, A, 1);
    eM[3] = pq::emission_q(tab, x, x2, 3);
    return pq::d_final(tab, S, a, Din);
}

__device__ __forceinline__ int32_t warp_max(int32_t v) { return __reduce_max_sync(FULL, v); }   // one REDUX

// One column: delete chain of column t - 1, E1 + emissions + E2 of column t (x = sample t - 1); stores the
// back-pointer word of column t - 1.
template <int XQ>
__device__ __forceinline__ void column(const TabQ &tab, const ModelScalars &ms, pq::StateQ &S, const double x, bool &ok,
                                       uint32_t &word, uint32_t *&bp) {
    // outside a Uniform range or NaN: the sequence will be declined; the pass runs on (integer arithmetic cannot
    // trap, nothing of it is used) so that the loop keeps one exit and the warp stays converged
    ok = ok && (x >= ms.lo && x <= ms.hi);
    int32_t eM[pq::P];
    const uint32_t dbits = block<XQ>(tab, ms, S, x, eM);
    *bp = word | dbits;
    bp += 32;
    word = pq::e2_emit(tab, S, eM);
}

__device__ __forceinline__ void renormalise(pq::StateQ &S, long long &off) {
    const int32_t mx = warp_max(pq::lane_max(S));
    pq::renorm(S, mx);
    off += mx;
}

// Forward pass of one sequence.  Returns false when a sample lies outside the fast emission range (declined).
// Columns 1..4 and the last T mod 4 go one at a time; in between, groups of four columns are unrolled with the
// renormalisations in place (no branch), their samples fetched one group ahead: the tag packing, the
// stores and the emissions of neighbouring columns then overlap with the shuffle latencies of the scan.
template <int XQ>
__device__ __forceinline__ bool forward(const TabQ &tab, const ModelScalars &ms, const int lane,
                                        pq::StateQ &S, const SeqCtx &c, long long &off) {
#pragma unroll
    for (int q = 0; q < pq::P; ++q) {
        S.M[q] = (lane * pq::P + q == ms.p_start) ? 0 : pq::Q_FLOOR;    // START: value 0 before the first sample only
        S.I[q] = S.D[q] = S.partM[q] = S.partI[q] = pq::Q_FLOOR;
    }
    S.Dprev = pq::Q_FLOOR;
    off = 0;
    // the trip count through a warp reduction: its result lives in a uniform register (see the kernel)
    const int T = __reduce_max_sync(FULL, c.T);
    const double *__restrict__ x = c.x;
    bool ok = true;
    uint32_t word = 0u;                                               // M / I back-pointers of the column just finished
    uint32_t *bp = c.bp + lane;
    int t = 1;
#pragma unroll 1
    for (; t <= T && t <= 4; ++t) {
        column<XQ>(tab, ms, S, __ldg(x + t - 1), ok, word, bp);
        if (t == 1) {                                                 // START does not outlive the first column
#pragma unroll
            for (int q = 0; q < pq::P; ++q)
                if (lane * pq::P + q == ms.p_start) S.M[q] = pq::Q_FLOOR;
            renormalise(S, off);
        }
        if ((t & (pq::R_NORM - 1)) == 0) renormalise(S, off);
    }
    static_assert(pq::R_NORM == 2 || pq::R_NORM == 4, "the group loop below renormalises after two or four columns");
    if (t + 3 <= T) {
        double x0 = __ldg(x + t - 1), x1 = __ldg(x + t), x2 = __ldg(x + t + 1), x3 = __ldg(x + t + 2);
#pragma unroll 1
        for (; t + 3 <= T; t += 4) {                                  // t = 5, 9, ...: t + 3 is a multiple of four
            const bool more = t + 7 <= T;
            const double n0 = more ? __ldg(x + t + 3) : 0.0, n1 = more ? __ldg(x + t + 4) : 0.0,
                         n2 = more ? __ldg(x + t + 5) : 0.0, n3 = more ? __ldg(x + t + 6) : 0.0;
            column<XQ>(tab, ms, S, x0, ok, word, bp);
            column<XQ>(tab, ms, S, x1, ok, word, bp);
            if (pq::R_NORM == 2) renormalise(S, off);
            column<XQ>(tab, ms, S, x2, ok, word, bp);
            column<XQ>(tab, ms, S, x3, ok, word, bp);
            renormalise(S, off);
            x0 = n0; x1 = n1; x2 = n2; x3 = n3;
        }
    }
#pragma unroll 1
    for (; t <= T; ++t) {
        column<XQ>(tab, ms, S, __ldg(x + t - 1), ok, word, bp);
        if ((t & (pq::R_NORM - 1)) == 0) renormalise(S, off);
    }
    int32_t unused[pq::P];
    *bp = word | block<XQ>(tab, ms, S, 0.0, unused);                  // column T: delete chain (END edges may leave it)
    return ok;
}

// END edges: best (v[T][src] + w), first maximum; unreachable sources do not count
__device__ __forceinline__ void end_edges(const VitProfModelDev &m, const pq::StateQ &S, uint32_t *stage, const int lane,
                                          double &best_out, int &barg_out, int32_t &vend_out) {
    const double NINF = pf::ninf();
    __syncwarp();
    int32_t *vals = reinterpret_cast<int32_t *>(stage);
#pragma unroll
    for (int q = 0; q < pq::P; ++q) {
        vals[lane * pq::P + q] = S.M[q];
        vals[pf::NPOS + lane * pq::P + q] = S.I[q];
        vals[2 * pf::NPOS + lane * pq::P + q] = S.D[q];
    }
    __syncwarp();
    double best = NINF;
    int barg = -1;
    int32_t v = 0;
    if (lane < m.n_end) {
        v = vals[m.end_slot[lane] * pf::NPOS + m.end_p[lane]];
        best = (double)v * (1.0 / (double)pq::Q_ONE) + m.end_w[lane];
        barg = lane;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double ob = __shfl_down_sync(FULL, best, off);
        const int oa = __shfl_down_sync(FULL, barg, off);
        if (ob > best || (ob == best && oa >= 0 && (barg < 0 || oa < barg))) { best = ob; barg = oa; }
    }
    best_out = __shfl_sync(FULL, best, 0);
    barg_out = __shfl_sync(FULL, barg, 0);
    vend_out = __shfl_sync(FULL, v, barg_out < 0 ? 0 : barg_out);
    __syncwarp();
}

// Traceback (all lanes walk in lock step; lane 0 / lane i write) with the float64 re-score of the path, and the
// result record of one sequence.  vfwd_q = the forward pass's value of the END source state (fixed point, with the
// subtracted column maxima added back): the path's quantised terms must add up to exactly this.
__device__ __noinline__ bool traceback(const VitProfBatch &b, const VitProfModelDev &m, const SeqCtx &c, uint32_t *stage,
                                       const int lane, const int p_start, const long long vfwd_q, const int barg) {
    const int T = c.T;
    const uint32_t *bp = c.bp;
    // the model's tables: pointers read once (the loop below stores through other pointers, so the compiler would
    // re-read them from the model record in every iteration -- a second dependent load per table access)
    const double *__restrict__ tab = m.tab;
    const double2 *__restrict__ trec = reinterpret_cast<const double2 *>(m.trec);
    const uint32_t *__restrict__ tmeta = m.tmeta;
    const int2 *__restrict__ tq = reinterpret_cast<const int2 *>(m.tq);
    const int32_t *__restrict__ qtab = m.qtab;
    long long qacc = 0;                   // this lane's share of the path's fixed-point score
    bool clamped = false;                 // an emission at the clamp or an edge the model does not have
    const double *__restrict__ xs = c.x;
    VitResult r;
    r.logp = 0.0; r.n_count = 0; r.t_first = -1; r.t_last = -1; r.pattern_len = 0; r.status = 0; r.reserved = 0;
    double acc = lane == 0 ? m.end_w[barg] : 0.0;
    const pf::TraceCfg tc = m.trace;
    int p = m.end_p[barg], slot = m.end_slot[barg], t = T;
    uint8_t *pat = b.pattern ? b.pattern + c.xo : nullptr;
    uint16_t *path = b.path ? b.path + c.xo : nullptr;
    bool in_group = false;
    uint8_t last_mod = '0';
    int plen = 0;
    // Back-pointer rows are staged in aligned blocks of 32 (block B = rows 32 B .. 32 B + 31) with their samples, two
    // buffers: while the walk is inside block B, block B - 1 is already on its way from HBM (the rows were written
    // tens of milliseconds ago and are long gone from L2; waiting ~1 us per block was a third of the traceback).
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(stage);
    auto issue = [&](int B) {
        if (B >= 0) {
            const int r0 = B * PROFQ_STAGE_ROWS;
            const uint32_t sdst = sbase + (B & 1) * PROFQ_BUF_BYTES;
            const uint4 *src = reinterpret_cast<const uint4 *>(bp + (size_t)r0 * 32);
            const int nvec = min(PROFQ_STAGE_ROWS, T + 1 - r0) * 8;
#pragma unroll
            for (int i = 0; i < PROFQ_STAGE_ROWS * 8 / 32; ++i)
                if (lane + 32 * i < nvec)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst + (lane + 32 * i) * 16),
                                 "l"(src + lane + 32 * i)
                                 : "memory");
            // the sample of every staged row (row ti decodes sample ti - 1): the re-score reads it from here
            if (r0 + lane >= 1 && r0 + lane <= T)
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sdst + PROFQ_STAGE_ROWS * 128 + lane * 8),
                             "l"(xs + r0 + lane - 1)
                             : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int cur = T / PROFQ_STAGE_ROWS;
    __syncwarp();
    issue(cur);
    issue(cur - 1);
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncwarp();
    int stage_lo = cur * PROFQ_STAGE_ROWS;    // rows [stage_lo, stage_lo + 32) are readable
    const uint32_t *rows = stage + (cur & 1) * (PROFQ_BUF_BYTES / 4);
    const double *xstage = reinterpret_cast<const double *>(rows + PROFQ_STAGE_ROWS * 32);
    long long guard = (long long)(T + 2) * (pf::NPOS + 2);
    while (!(slot == 0 && p == p_start)) {
        if (--guard < 0 || p < 0 || p >= pf::NPOS || t < 0) { r.status = 2; break; }
        if (t < stage_lo) {
            __syncwarp();                     // everybody is done with the block above
            --cur;
            issue(cur - 1);                   // into the buffer just left
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp();
            stage_lo = cur * PROFQ_STAGE_ROWS;
            rows = stage + (cur & 1) * (PROFQ_BUF_BYTES / 4);
            xstage = reinterpret_cast<const double *>(rows + PROFQ_STAGE_ROWS * 32);
        }
        const int tl = p >> 2;            // lane that owns the current state: its table column holds the in-edge weights
        if (slot == 2) {                  // silent delete state: same column
            int wk = 0;
            if (!pq::back_apply(c_back[pq::back_index(rows[(t - stage_lo) * 32 + tl], p, slot)], tc, p, slot, t, wk)) { r.status = 2; break; }
            if (lane == 0) {
                acc += __ldg(tab + wk * 32 + tl);
                const int32_t qw = __ldg(qtab + wk * 32 + tl);
                clamped |= qw == INT32_MIN;
                qacc += qw;
            }
            continue;
        }
        if (t < 1) { r.status = 2; break; }
        // Emitting state (p, slot) at column t.  Samples dwell in a state, so most pointers are self loops:
        // lane i looks at column t - i, the warp skips the whole run of self loops at once and then takes
        // the first other pointer (all lanes keep the same cursor; lane 0 / lane i write the outputs).
        const int ti = t - lane;
        const bool valid = ti >= stage_lo && ti >= 1;
        const uint32_t wfull = valid ? rows[(ti - stage_lo) * 32 + tl] : 0u;
        const uint32_t f = wfull >> (8 * (p & 3));
        const bool self = valid && (slot == 0 ? (f & 7u) == 7u : ((f >> 3) & 3u) == 3u);
        const unsigned other = ~__ballot_sync(FULL, self);
        const int k = other ? __ffs(other) - 1 : 32;               // columns t .. t-k+1 are self loops
        const bool step = k < 32 && ((__ballot_sync(FULL, valid) >> k) & 1u);   // column t-k is staged
        const int visits = k + (step ? 1 : 0);                     // >= 1: column t itself is staged
        const int idx = p * 2 + slot;
        // everything that depends on the state only: loads issued together
        const double2 ab = __ldg(trec + idx * 4), cw = __ldg(trec + idx * 4 + 1), Ac = __ldg(trec + idx * 4 + 2),
                      Cz = __ldg(trec + idx * 4 + 3);
        const int2 qs = __ldg(tq + idx);
        const unsigned fl = __ldg(tmeta + idx);
        const int sid = (int)(fl >> 16);
        if (fl & HMM_FLAG_COUNT) r.n_count += visits;
        if (fl & HMM_FLAG_REPEAT) { if (r.t_last < 0) r.t_last = t - 1; r.t_first = t - visits; }
        if (fl & HMM_FLAG_SEP) {
            if (in_group) { if (pat && lane == 0) pat[T - 1 - plen] = last_mod; ++plen; in_group = false; }
        } else {
            in_group = true;
            last_mod = (fl & HMM_FLAG_MOD) ? '1' : '0';
        }
        if (path && lane < visits) path[t - 1 - lane] = (uint16_t)sid;
        // re-score: lane i adds the emission of column t - i and, inside the run, the self-loop weight
        if (lane < visits) {
            // the forward pass has checked that every sample is a number inside all Uniform ranges
            const double x = xstage[ti - stage_lo];
            const double dx = x - ab.x;
            acc += ab.y - (dx * dx) * cw.x;
            // ... and the same terms as the forward pass added them
            int32_t qe = qs.y;
            if (slot == 0) {
                const int32_t e16 = pq::emission_q16(Ac.x, Ac.y, Cz.x, x, x * x);
                clamped |= e16 == pq::E_MIN16;
                qe = e16 * 8;
            }
            qacc += qe;
            if (lane < k) { acc += cw.y; qacc += qs.x; }
        }
        t -= k;
        if (step) {
            const uint32_t w = __shfl_sync(FULL, wfull, k);
            int wk = 0;
            if (!pq::back_apply(c_back[pq::back_index(w, p, slot)], tc, p, slot, t, wk)) { r.status = 2; break; }
            if (lane == 0) {
                acc += __ldg(tab + wk * 32 + tl);
                const int32_t qw = __ldg(qtab + wk * 32 + tl);
                clamped |= qw == INT32_MIN;
                qacc += qw;
            }
        }
    }
    if (in_group) { if (pat && lane == 0) pat[T - 1 - plen] = last_mod; ++plen; }
    if (r.status == 0 && t != 0) r.status = 2;
    r.pattern_len = plen;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        acc += __shfl_xor_sync(FULL, acc, off);
        qacc += __shfl_xor_sync(FULL, qacc, off);
    }
    r.logp = acc;
    // The forward value is an upper bound of the path's fixed-point score, exact unless a clamped value won on the
    // way (profile_q.h): equality proves that the path is the optimum of the quantised model.
    if (r.status == 0 && (qacc != vfwd_q || __any_sync(FULL, clamped))) r.status = 3;
    // ... and its float64 score differs from that by the rounding of <= 2 addends per column only (sanity)
    if (r.status == 0 && !(fabs((double)vfwd_q * (1.0 / (double)pq::Q_ONE) + m.end_w[barg] - acc) <=
                           1.0e-3 + (double)T * (2.0 / (double)(1 << pq::FRAC)))) r.status = 3;
    if (r.status == 2) r.status = 3;      // let the float64 decoder have the last word
    if (r.status == 0 && lane == 0) b.res[c.seq] = r;
    __syncwarp();
    return r.status == 0;
}

// A sequence the fixed-point pass cannot vouch for: decoded again in float64, right here, by the same warp (the
// float64 kernel's device code; the warp's staging area holds the model table meanwhile).  Such sequences are rare
// (4 of 8192 C2 reads) and long; a second launch for them would be a serial tail of T x (latency of one column) --
// 18 ms for 45 k samples -- while inside this launch it hides behind the other warps' sequences.  The result carries
// status | 4 on the way out so that the host can count them.
template <int XQ>
__device__ __noinline__ void decode_float64(const VitProfBatch &b, const VitProfModelDev &m, const SeqCtx &c, uint32_t *stage,
                                            const int lane) {
    f64::ModelScalars fs;
    fs.p_start = m.p_start;
    const int xp = m.trace.xm_src_p >= 0 ? m.trace.xm_src_p : (m.trace.xd_src_p >= 0 ? m.trace.xd_src_p : 0);
    fs.xlane = xp / pf::P; fs.xq = xp % pf::P;
    fs.xm_slot = m.trace.xm_src_slot; fs.xd_slot = m.trace.xd_src_slot;
    fs.lo = m.lo; fs.hi = m.hi;
    pf::Regs R;
    pf::load_regs(f64::TabGlobal{m.tab + lane}, R);
    // the model's table: logical [k][lane] in global memory -> pair-interleaved in this warp's shared memory
    __syncwarp();
    double *aux_s = reinterpret_cast<double *>(stage);
    for (int k = 0; k < pf::K_NAUX; ++k) aux_s[((k >> 1) * 32 + lane) * 2 + (k & 1)] = __ldg(m.tab + (pf::K_NREG + k) * 32 + lane);
    __syncwarp();
    const f64::AuxShared aux{reinterpret_cast<const double2 *>(aux_s) + lane};
    stage += f64::PROF_AUX_BYTES / 4;
    f64::SeqCtx fc[1];
    fc[0].seq = c.seq; fc[0].T = c.T; fc[0].xo = c.xo; fc[0].x = c.x; fc[0].bp = c.bp;
    pf::State S[1];
    uint32_t bits[1];
    f64::init_state(S[0], lane, fs.p_start);
    double x0[1] = {c.T > 0 ? __ldg(c.x) : 0.0};
    f64::block<1, XQ, f64::AuxShared>(R, aux, fs, S, bits);
    c.bp[lane] = bits[0];
    f64::forward<1, XQ, f64::AuxShared>(R, aux, fs, m, lane, S, fc, x0, 1, c.T);
    double best;
    int barg;
    f64::end_edges(m, S[0], stage, lane, best, barg);
    f64::traceback(b, m, fc[0], stage, lane, fs.p_start, best, barg);
    __syncwarp();
    if (lane == 0) b.res[c.seq].reserved = 1;             // decoded in float64
    __syncwarp();
}

// one sequence, start to finish, by one warp
template <int XQ>
__device__ __forceinline__ void run_seq(const VitProfBatch &b, const VitProfModelDev &m, const SeqCtx &c, const TabQ &tab,
                                        const ModelScalars &ms, uint32_t *stage, const int lane) {
    pq::StateQ S;
    long long off;
    bool ok = forward<XQ>(tab, ms, lane, S, c, off);
    if (ok) {
        double best;
        int barg;
        int32_t vend;
        end_edges(m, S, stage, lane, best, barg, vend);
        ok = barg >= 0 && traceback(b, m, c, stage, lane, ms.p_start, (long long)vend + off, barg);
    }
    if (!ok) decode_float64<XQ>(b, m, c, stage, lane);
}

// Persistent warps: every warp pulls whole sequences (longest first, all models of the batch in one queue) and
// keeps its model's constants in registers until a sequence of another model comes up.  No CTA-level state: the
// warps of a CTA never wait for each other.
__global__ void __launch_bounds__(PROFQ_WARPS * 32, PROFQ_CTAS_PER_SM) viterbi_profile_q_kernel(VitProfBatch b) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t *stage = reinterpret_cast<uint32_t *>(smem + (size_t)warp * PROFQ_STAGE_BYTES);
    int model = -1;
    TabQ tab;
    ModelScalars ms{0, 0, 0, 0, 0.0, 0.0};
    int xq = 0;
    for (;;) {
        // warp-uniform values go through a warp reduction (result in a uniform register): ptxas then knows that the
        // control flow below does not diverge and emits the shuffles of the column loop without divergence checks
        int task = 0;
        if (lane == 0) task = atomicAdd(b.counters, 1);
        task = __reduce_max_sync(FULL, task);
        if (task >= b.n_tasks) return;
        const int seq = __reduce_max_sync(FULL, b.order[task]);
        const int mi = __reduce_max_sync(FULL, b.seq_model[seq]);
        const VitProfModelDev &m = b.models[mi];
        if (mi != model) {
            model = mi;
            const int4 *gsrc = reinterpret_cast<const int4 *>(m.qgrp);
            const double2 *esrc = reinterpret_cast<const double2 *>(m.qem);
#pragma unroll
            for (int k = 0; k < pq::G_TOTAL; ++k) {
                const int4 g = __ldg(gsrc + k * 32 + lane);
                tab.g[k][0] = g.x; tab.g[k][1] = g.y; tab.g[k][2] = g.z; tab.g[k][3] = g.w;
            }
#pragma unroll
            for (int k = 0; k < pq::E_TOTAL; ++k) {
                const double2 e = __ldg(esrc + k * 32 + lane);
                tab.e[k][0] = e.x; tab.e[k][1] = e.y;
            }
            ms.p_start = m.p_start;
            const int xp = m.trace.xm_src_p >= 0 ? m.trace.xm_src_p : (m.trace.xd_src_p >= 0 ? m.trace.xd_src_p : 0);
            ms.xlane = xp / pq::P; xq = __reduce_max_sync(FULL, xp % pq::P);
            ms.xm_slot = m.trace.xm_src_slot; ms.xd_slot = m.trace.xd_src_slot;
            ms.lo = m.lo; ms.hi = m.hi;
        }
        const SeqCtx c = seq_ctx(b, seq);
        switch (xq) {
            case 0: run_seq<0>(b, m, c, tab, ms, stage, lane); break;
            case 1: run_seq<1>(b, m, c, tab, ms, stage, lane); break;
            case 2: run_seq<2>(b, m, c, tab, ms, stage, lane); break;
            default: run_seq<3>(b, m, c, tab, ms, stage, lane); break;
        }
    }
}

}  // namespace

size_t viterbi_profile_q_smem_bytes() { return (size_t)PROFQ_WARPS * PROFQ_STAGE_BYTES; }

// Largest useful grid: every resident CTA slot of the device (persistent warps pulling from the queue).
int viterbi_profile_q_max_grid(strique_ctx *ctx, int *warps_per_cta) {
    static int cached[64] = {0};
    int &per_sm_cached = cached[ctx->device & 63];
    if (warps_per_cta) *warps_per_cta = PROFQ_WARPS;
    if (per_sm_cached == 0) {
        const size_t smem = viterbi_profile_q_smem_bytes();
        if (cudaFuncSetAttribute(viterbi_profile_q_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, viterbi_profile_q_kernel, PROFQ_WARPS * 32, smem) != cudaSuccess) return 0;
        per_sm_cached = per_sm < 1 ? 1 : per_sm;
    }
    return ctx->num_sms * per_sm_cached;
}

int viterbi_profile_q_launch(strique_ctx *ctx, const VitProfBatch &b, int grid) {
    if (grid <= 0) return STRIQUE_OK;
    viterbi_profile_q_kernel<<<grid, PROFQ_WARPS * 32, viterbi_profile_q_smem_bytes(), ctx->stream>>>(b);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return STRIQUE_OK;
}

// Quantises the packed profile image for the fixed-point kernel when it fits the bounds (sets f->qgrp / f->qem).
int viterbi_profile_q_pack(strique_ctx *ctx, const ProfileImage &img, VitProfModelDev *f) {
    f->qgrp = nullptr;
    f->qem = nullptr;
    f->trec = nullptr;
    f->tmeta = nullptr;
    f->tq = nullptr;
    f->qtab = nullptr;
    ProfileQImage qi;
    std::string why;
    if (!profile_quantise(img, &qi, &why)) return STRIQUE_OK;
    void *pg = nullptr, *pe = nullptr;
    auto upload = [&](const void *src, size_t bytes, const void **dst) -> int {
        void *p = nullptr;
        CUDA_TRY(ctx, cudaMalloc(&p, bytes));
        ctx->owned.push_back(p);
        CUDA_TRY(ctx, cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice));
        *dst = p;
        return STRIQUE_OK;
    };
    TRY(upload(qi.trec.data(), qi.trec.size() * 8, (const void **)&f->trec));
    TRY(upload(qi.tmeta.data(), qi.tmeta.size() * 4, (const void **)&f->tmeta));
    TRY(upload(qi.tq.data(), qi.tq.size() * 4, (const void **)&f->tq));
    TRY(upload(qi.qtab.data(), qi.qtab.size() * 4, (const void **)&f->qtab));
    CUDA_TRY(ctx, cudaMalloc(&pg, qi.grp.size() * 4));
    ctx->owned.push_back(pg);
    CUDA_TRY(ctx, cudaMalloc(&pe, qi.em.size() * 8));
    ctx->owned.push_back(pe);
    CUDA_TRY(ctx, cudaMemcpy(pg, qi.grp.data(), qi.grp.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(pe, qi.em.data(), qi.em.size() * 8, cudaMemcpyHostToDevice));
    f->qgrp = (const int32_t *)pg;
    f->qem = (const double *)pe;
    return STRIQUE_OK;
}

}  // namespace strique
