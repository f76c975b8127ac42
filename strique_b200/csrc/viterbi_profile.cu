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
#include "viterbi_profile_dev.cuh"

namespace strique {

namespace {

using namespace f64;

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
        block<2, XQ, AuxShared>(R, aux, ms, S, bits);                     // column 0: delete chain from START
        c[0].bp[lane] = bits[0];
        c[1].bp[lane] = bits[1];
        forward<2, XQ, AuxShared>(R, aux, ms, m, lane, S, c, xcur, 1, c[1].T);
        double best1, best0;
        int barg1, barg0;
        end_edges(m, S[1], stage, lane, best1, barg1);
        {
            pf::State S0[1] = {S[0]};
            const SeqCtx c0[1] = {c[0]};
            double x0[1] = {xcur[0]};
            forward<1, XQ, AuxShared>(R, aux, ms, m, lane, S0, c0, x0, c[1].T + 1, c[0].T);
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
            block<1, XQ, AuxShared>(R, aux, ms, S0, bits);
            c0[0].bp[lane] = bits[0];
            forward<1, XQ, AuxShared>(R, aux, ms, m, lane, S0, c0, x0, 1, c0[0].T);
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
