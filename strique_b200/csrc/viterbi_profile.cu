// Profile Viterbi: the throughput kernel of boundary #2 for the reference's linear profile HMMs.
//
// Replaces pomegranate 0.10.0 `HiddenMarkovModel.viterbi` behind flankedRepeatHMM.count_repeats
// (reference scripts/STRique.py:433-441, 374-378; topology 201-431).  float64 throughout, one add per
// edge, strict-'>' maxima (profile_core.h holds the lane arithmetic, shared with the host emulator of
// the test suite; profile_pack.h recognises and packs the model).
//
// Mapping: ONE WARP decodes one sequence; lane l owns profile positions 4l..4l+3 (match, insert and
// delete state each), so
//   * every regular edge reads a register of the same lane or one of three values shuffled up from
//     lane l-1 -- no shared-memory value columns, no barriers;
//   * the repeat loop (d2 -> M_0, d1 -> s1) is one indexed shuffle pair;
//   * the four hottest in-edge weights of every match state live in registers, the other per-position
//     constants in a shared-memory table of the CTA's current model, laid out in (even, odd) pairs and read
//     with conflict-free LDS.128 [lane*16 + const] -- 128 registers per thread, 16 warps per SM;
//   * the delete chain of a column is a max-plus scan: sequential inside the lane, Kogge-Stone across
//     lanes; the E1 part of the NEXT column (all edges from emitting states) sits in the same basic
//     block and fills the shuffle / fp64 latencies of the scan;
//   * back-pointers: one byte per position -> ONE 32-bit word per lane per column = 128 B per time step,
//     coalesced; traceback by the same warp over rows staged 32 at a time with cp.async.
// A CTA (4 independent warps, no barrier inside a sequence) serves one model at a time: its warps pull
// sequences (longest first) from that model's queue; when the queue runs dry the CTA moves on to the next
// model that still has work (one barrier + table reload), so both strands / all loci share one launch.
#include <math.h>

#include <algorithm>

#include "profile_pack.h"
#include "viterbi.cuh"

namespace strique {

namespace {

#ifndef PROF_WARPS_PER_CTA
#define PROF_WARPS_PER_CTA 4
#endif
constexpr int PROF_WARPS = PROF_WARPS_PER_CTA;       // warps per CTA
#ifndef PROF_SEQS
#define PROF_SEQS 1                                 // sequences decoded side by side by one warp (1 or 2; 2 measured no faster)
#endif
#ifndef PROF_CTAS
#define PROF_CTAS (PROF_SEQS == 2 ? 2 : 4)
#endif
constexpr int PROF_CTAS_PER_SM = PROF_CTAS;
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
template <int NS, int XQ>
__device__ __forceinline__ void block(const pf::Regs &R, const AuxShared &aux, const ModelScalars &ms,
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
template <int NS, int XQ>
__device__ __forceinline__ void forward(const pf::Regs &R, const AuxShared &aux, const ModelScalars &ms,
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
        block<NS, XQ>(R, aux, ms, S, dbits);
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
__device__ __noinline__ void traceback(const VitProfBatch &b, const VitProfModelDev &m, const SeqCtx &c, uint32_t *stage,
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

// the sequences of one warp in one CTA task
template <int XQ>
__device__ __forceinline__ void run_task(const VitProfBatch &b, const VitProfModelDev &m, const VitCtaTask &ct,
                                         const pf::Regs &R, const AuxShared &aux, const ModelScalars &ms,
                                         uint32_t *stage, const int lane, const int warp) {
    // this warp's sequences: neighbours in the length order of the task (first one is the longer)
    const int first = ct.first + warp * PROF_SEQS;
    const int mine = min(PROF_SEQS, ct.count - warp * PROF_SEQS);
    if (mine <= 0) return;
    if (PROF_SEQS == 2 && mine == 2) {
        SeqCtx c[2] = {seq_ctx(b, b.order[first]), seq_ctx(b, b.order[first + 1])};
        if (c[1].T > c[0].T) { const SeqCtx t = c[0]; c[0] = c[1]; c[1] = t; }
        pf::State S[2];
        uint32_t bits[2];
        double xcur[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            init_state(S[i], lane, ms.p_start);
            xcur[i] = c[i].T > 0 ? __ldg(c[i].x) : 0.0;
        }
        block<2, XQ>(R, aux, ms, S, bits);                     // column 0: delete chain from START
        c[0].bp[lane] = bits[0];
        c[1].bp[lane] = bits[1];
        forward<2, XQ>(R, aux, ms, m, lane, S, c, xcur, 1, c[1].T);
        double best1, best0;
        int barg1, barg0;
        end_edges(m, S[1], stage, lane, best1, barg1);
        {
            pf::State S0[1] = {S[0]};
            const SeqCtx c0[1] = {c[0]};
            double x0[1] = {xcur[0]};
            forward<1, XQ>(R, aux, ms, m, lane, S0, c0, x0, c[1].T + 1, c[0].T);
            end_edges(m, S0[0], stage, lane, best0, barg0);
        }
        traceback(b, m, c[0], stage, lane, ms.p_start, best0, barg0);
        traceback(b, m, c[1], stage, lane, ms.p_start, best1, barg1);
    } else {
#pragma unroll 1
        for (int k = 0; k < mine; ++k) {
            const SeqCtx c0[1] = {seq_ctx(b, b.order[first + k])};
            pf::State S0[1];
            uint32_t bits[1];
            init_state(S0[0], lane, ms.p_start);
            double x0[1] = {c0[0].T > 0 ? __ldg(c0[0].x) : 0.0};
            block<1, XQ>(R, aux, ms, S0, bits);
            c0[0].bp[lane] = bits[0];
            forward<1, XQ>(R, aux, ms, m, lane, S0, c0, x0, 1, c0[0].T);
            double best0;
            int barg0;
            end_edges(m, S0[0], stage, lane, best0, barg0);
            traceback(b, m, c0[0], stage, lane, ms.p_start, best0, barg0);
        }
    }
}

__global__ void __launch_bounds__(PROF_WARPS * 32, PROF_CTAS_PER_SM) viterbi_profile_kernel(VitProfBatch b) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_task;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *aux_s = reinterpret_cast<double *>(smem);
    uint32_t *stage = reinterpret_cast<uint32_t *>(smem + PROF_AUX_BYTES + (size_t)warp * PROF_STAGE_BYTES);
    const AuxShared aux{reinterpret_cast<const double2 *>(smem) + lane};

    int model = -1;
    pf::Regs R;
    ModelScalars ms{0, 0, 0, 0, 0, 0.0, 0.0};
  for (;;) {                            // ---- one CTA task: <= PROF_WARPS * PROF_SEQS sequences of one model ----
    __syncthreads();                                    // everybody is done with s_task and the table
    if (threadIdx.x == 0) s_task = atomicAdd(b.counters, 1);
    __syncthreads();
    if (s_task >= b.n_tasks) return;
    const VitCtaTask ct = b.tasks[s_task];
    const VitProfModelDev &m = b.models[ct.model];
    if (ct.model != model) {
        model = ct.model;
        // table of the model: logical [k][lane] in global memory -> pair-interleaved in shared memory
        for (int i = threadIdx.x; i < pf::K_NAUX * 32; i += PROF_WARPS * 32) {
            const int k = i >> 5, l = i & 31;
            aux_s[((k >> 1) * 32 + l) * 2 + (k & 1)] = __ldg(m.tab + (pf::K_NREG + k) * 32 + l);
        }
        pf::load_regs(TabGlobal{m.tab + lane}, R);
        ms.p_start = m.p_start;
        const int xp = m.trace.xm_src_p >= 0 ? m.trace.xm_src_p : (m.trace.xd_src_p >= 0 ? m.trace.xd_src_p : 0);
        ms.xlane = xp / pf::P; ms.xq = xp % pf::P;
        ms.xm_slot = m.trace.xm_src_slot; ms.xd_slot = m.trace.xd_src_slot;
        ms.lo = m.lo; ms.hi = m.hi;
        __syncthreads();
    }
    switch (ms.xq) {
        case 0: run_task<0>(b, m, ct, R, aux, ms, stage, lane, warp); break;
        case 1: run_task<1>(b, m, ct, R, aux, ms, stage, lane, warp); break;
        case 2: run_task<2>(b, m, ct, R, aux, ms, stage, lane, warp); break;
        default: run_task<3>(b, m, ct, R, aux, ms, stage, lane, warp); break;
    }
  }
}

}  // namespace

size_t viterbi_profile_smem_bytes() { return (size_t)PROF_AUX_BYTES + (size_t)PROF_WARPS * PROF_STAGE_BYTES; }

// Largest useful grid: every resident CTA slot of the device (persistent CTAs pulling from the queues).
int viterbi_profile_max_grid(strique_ctx *ctx, int *warps_per_cta) {
    static int cached[64] = {0};                      // per device: the shared-memory attribute is per device too
    int &per_sm_cached = cached[ctx->device & 63];
    if (warps_per_cta) *warps_per_cta = PROF_WARPS * PROF_SEQS;   // sequences per CTA task
    if (per_sm_cached == 0) {
        const size_t smem = viterbi_profile_smem_bytes();
        if (cudaFuncSetAttribute(viterbi_profile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 0;
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, viterbi_profile_kernel, PROF_WARPS * 32, smem) != cudaSuccess) return 0;
        per_sm_cached = per_sm < 1 ? 1 : per_sm;
    }
    return ctx->num_sms * per_sm_cached;
}

// grid persistent CTAs pulling CTA tasks from the queue
int viterbi_profile_launch(strique_ctx *ctx, const VitProfBatch &b, int grid) {
    if (grid <= 0) return STRIQUE_OK;
    viterbi_profile_kernel<<<grid, PROF_WARPS * 32, viterbi_profile_smem_bytes(), ctx->stream>>>(b);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return STRIQUE_OK;
}

// Packs the model for the profile kernel when its layout hints describe a linear profile (profile_pack.h);
// otherwise leaves m->profile.tab == nullptr.
int viterbi_profile_pack(strique_ctx *ctx, const strique_hmm_desc *d, HmmModel *m) {
    m->has_profile = false;
    ProfileImage img;
    std::string why;
    if (!profile_pack(d, &img, &why)) return STRIQUE_OK;
    VitProfModelDev &f = m->profile;
    memset(&f, 0, sizeof(f));
    auto upload = [&](const void *src, size_t bytes, const void **dst) -> int {
        void *p = nullptr;
        CUDA_TRY(ctx, cudaMalloc(&p, bytes));
        CUDA_TRY(ctx, cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice));
        ctx->owned.push_back(p);
        *dst = p;
        return STRIQUE_OK;
    };
    TRY(upload(img.tab.data(), img.tab.size() * 8, (const void **)&f.tab));
    TRY(upload(img.em_kind.data(), img.em_kind.size(), (const void **)&f.em_kind));
    TRY(upload(img.em_a.data(), img.em_a.size() * 8, (const void **)&f.em_a));
    TRY(upload(img.em_b.data(), img.em_b.size() * 8, (const void **)&f.em_b));
    TRY(upload(img.em_c.data(), img.em_c.size() * 8, (const void **)&f.em_c));
    TRY(upload(img.flags.data(), img.flags.size(), (const void **)&f.flags));
    TRY(upload(img.state_id.data(), img.state_id.size() * 4, (const void **)&f.state_id));
    f.trace = img.trace;
    f.lo = img.lo;
    f.hi = img.hi;
    f.p_start = img.p_off - 1;
    f.n_end = img.n_end;
    for (int e = 0; e < img.n_end; ++e) { f.end_p[e] = img.end_p[e]; f.end_slot[e] = img.end_slot[e]; f.end_w[e] = img.end_w[e]; }
    m->has_profile = true;
    return viterbi_profile_q_pack(ctx, img, &f);
}

}  // namespace strique
