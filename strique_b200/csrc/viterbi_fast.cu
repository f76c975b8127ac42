// Team Viterbi: the throughput kernel of boundary #2 (see viterbi.cuh for the model form).
//
// Replaces pomegranate 0.10.0 `HiddenMarkovModel.viterbi` behind flankedRepeatHMM.count_repeats and
// repeatModHMM.mod_repeats (reference scripts/STRique.py:433-441, 492-500).  float64 throughout,
// same candidate order and strict-> maxima as the generic kernel in viterbi.cu, so both decode the
// same path and the same log p.
//
// Mapping ("warp per state block"): a TEAM of WPS warps decodes one sequence; every warp owns a
// block of NH "high" slots (<= 6 in-edges) and NL "low" slots (<= 3 in-edges) of 32 emitting states
// and one group of silent chain states (QC per lane, <= 2 entry edges each).
//   * in-edge weights, source offsets, emission parameters and chain weights live in REGISTERS
//     (loaded once per model), so the only shared-memory traffic of a time step is the gather of the
//     source values and the store of the new ones (double-buffered columns);
//   * the delete chains are a max-plus scan inside a warp (Kogge-Stone over shuffles);
//   * the WPS warps meet at two named barriers per time step;
//   * back-pointers: 4 bits per state, one 32-bit word per lane per step -> (T+1) x WPS x 128 B in HBM;
//   * traceback by the team's first warp over rows staged 32 at a time into shared memory.
// A CTA holds 8/WPS teams and pulls "CTA tasks" (<= 8/WPS sequences of one model, similar lengths,
// longest first) from a global queue.
#include <math.h>

#include <algorithm>
#include <numeric>
#include <type_traits>

#include "viterbi.cuh"

namespace strique {

namespace {

__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xfff0000000000000ll); }

// shared-memory gather from an absolute 32-bit shared address; the column parity is a compile-time
// offset, so the load is one LDS.64 [R + IMM].  A plain (compiler-visible) load: it may be hoisted
// and interleaved inside a phase but never across the barriers.
template <int IMM>
__device__ __forceinline__ double lds_f64(uint32_t addr) {
    return *reinterpret_cast<const double *>(__cvta_shared_to_generic((size_t)(addr + IMM)));
}

template <int WPS>
__device__ __forceinline__ void team_sync(int id) {
    if (WPS > 1)
        asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(WPS * 32) : "memory");
    else
        __syncwarp();
}

template <int WPS, int NH, int NL, int QC>
struct FastShape {
    static constexpr int NSW = NH + NL;                 // emitting slots per warp
    static constexpr int ROWS = NH * 6 + NL * 3;        // in-edge rows per warp
    static constexpr int TEAMS = 8 / WPS;               // sequences in flight per CTA
    static constexpr int CB = WPS * NSW * 32;           // first chain position
    static constexpr int P_START = CB + WPS * QC * 32;
    static constexpr int P_NEG = P_START + 1;
    static constexpr int NV = P_START + 2;
    static constexpr int NVP = (NV + 1) & ~1;           // padded column length (doubles)
    static constexpr int STAGE_ROWS = 32;
    static constexpr int TEAM_DOUBLES = 2 * NVP;
    static constexpr int TEAM_STAGE_WORDS = STAGE_ROWS * WPS * 32;
    // back-pointer word of a lane: one bit per "candidate beat the running best" comparison
    //   high slot k: 5 bits at 5k; low slot k: 2 bits at 5 NH + 2 (k - NH); chain state q: 3 bits at CBIT + 3q
    static constexpr int CBIT = 5 * NH + 2 * NL;
    __host__ __device__ static constexpr int bit0(int k) { return k < NH ? 5 * k : 5 * NH + 2 * (k - NH); }
    __host__ __device__ static constexpr int row0(int k) { return k < NH ? 6 * k : 6 * NH + 3 * (k - NH); }
    __host__ __device__ static constexpr int deg(int k) { return k < NH ? 6 : 3; }
    static size_t smem_bytes(int blob_cap) {     // dynamic part: model blob + staged back-pointer rows
        return (size_t)blob_cap + (size_t)TEAMS * TEAM_STAGE_WORDS * 4;
    }
};

template <int WPS, int NH, int NL, int QC>
__global__ void __launch_bounds__(256, 2) viterbi_team_kernel(VitFastBatch b) {
    typedef FastShape<WPS, NH, NL, QC> SH;
    constexpr int NSW = SH::NSW, ROWS = SH::ROWS, TEAMS = SH::TEAMS, CB = SH::CB, P_START = SH::P_START;
    constexpr int NVP = SH::NVP, RW = WPS * 32;       // RW: back-pointer words per time step
    constexpr int QCA = QC > 0 ? QC : 1;
    extern __shared__ __align__(16) unsigned char smem[];
    // value columns are STATIC shared memory: their address is a link-time constant, so a gather is
    // one LDS [R + const] with R = the per-edge byte offset (team offset included) kept in a register
    __shared__ __align__(16) double vcols[TEAMS * SH::TEAM_DOUBLES];
    __shared__ int s_task;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int team = warp / WPS, w = warp % WPS;
    const double NINF = neg_inf();
    // dynamic shared layout: [model blob] [per team: staged back-pointer rows]
    unsigned char *blob = smem;
    double *vteam = vcols + (size_t)team * SH::TEAM_DOUBLES;
    // absolute shared address of this team's column 0 (static shared: < 64 KB, fits the 16-bit fields)
    const uint32_t team_off = (uint32_t)__cvta_generic_to_shared(vteam);
    uint32_t *stage = reinterpret_cast<uint32_t *>(smem + b.blob_cap) + (size_t)team * SH::TEAM_STAGE_WORDS;

    // per-lane model constants (registers)
    double wreg[ROWS];
    uint32_t srcreg[(ROWS + 1) / 2];      // two 16-bit BYTE offsets into a value column per register
    double em0[NSW], em1[NSW], em2[NSW];
    uint32_t kindmask = 0;
    uint32_t csrc[QCA];                   // two 16-bit byte offsets (entry edge 0, entry edge 1)
    const double *cpw_s = nullptr, *cew_s = nullptr, *cwr_s = nullptr;   // chain weights stay in the shared model image
    int cur_model = -1;
    const int bar_id = 1 + team;

    for (;;) {
        __syncthreads();
        if (tid == 0) s_task = atomicAdd(b.queue, 1);
        __syncthreads();
        const int task = s_task;
        if (task >= b.n_tasks) break;
        const VitCtaTask ct = b.tasks[task];
        if (ct.model != cur_model) {
            cur_model = ct.model;
            const VitFastModelDev m = b.models[ct.model];
            for (int i = tid * 16; i < m.blob_bytes; i += blockDim.x * 16)
                *reinterpret_cast<uint4 *>(blob + i) = *reinterpret_cast<const uint4 *>(m.blob + i);
            __syncthreads();
            const double *bw = reinterpret_cast<const double *>(blob + m.off_w) + (size_t)w * ROWS * 32 + lane;
            const uint16_t *bs = reinterpret_cast<const uint16_t *>(blob + m.off_src) + (size_t)w * ROWS * 32 + lane;
#pragma unroll
            for (int r = 0; r < ROWS; ++r) wreg[r] = bw[r * 32];
#pragma unroll
            for (int r = 0; r < ROWS; r += 2) {
                const uint32_t lo = (uint32_t)bs[r * 32] * 8u + team_off;
                const uint32_t hi = r + 1 < ROWS ? (uint32_t)bs[(r + 1) * 32] * 8u + team_off : 0u;
                srcreg[r / 2] = lo | (hi << 16);
            }
            const double *be = reinterpret_cast<const double *>(blob + m.off_em);
            const uint8_t *bk = blob + m.off_flags;
            kindmask = 0;
#pragma unroll
            for (int k = 0; k < NSW; ++k) {
                const int li = (w * NSW + k) * 32 + lane;
                em0[k] = be[li];
                em1[k] = be[WPS * NSW * 32 + li];
                em2[k] = be[2 * WPS * NSW * 32 + li];
                kindmask |= (uint32_t)((bk[li] >> 7) & 1) << k;
            }
            if (QC > 0) {
                const double *bpw = reinterpret_cast<const double *>(blob + m.off_predw);
                const double *bc = reinterpret_cast<const double *>(blob + m.off_cw);
                const uint16_t *bcs = reinterpret_cast<const uint16_t *>(blob + m.off_csrc);
                cpw_s = bpw + (size_t)w * QC * 32 + lane;
                cew_s = bc + (size_t)w * QC * 2 * 32 + lane;
                cwr_s = reinterpret_cast<const double *>(blob + m.off_wr) + (size_t)w * 5 * 32 + lane;
#pragma unroll
                for (int q = 0; q < QC; ++q) {
                    csrc[q] = ((uint32_t)bcs[((w * QC + q) * 2 + 0) * 32 + lane] * 8u + team_off) |
                              (((uint32_t)bcs[((w * QC + q) * 2 + 1) * 32 + lane] * 8u + team_off) << 16);
                }
            }
        }
        if (team >= ct.count) continue;          // uniform per team; rejoins the CTA at __syncthreads

        const int seq = b.order[ct.first + team];
        const int64_t xo = b.x_off[seq];
        const int T = (int)(b.x_off[seq + 1] - xo);
        const double *x = b.x + xo;
        uint32_t *bp = b.bp + b.bp_off[seq];

        // silent chain states of column PAR: v[c] = max(entry edges from v, v[c-1] + w); returns nibbles
        // E1 of one slot for the upcoming step (see the forward pass): reads column PS
        double part[NSW];                                     // best E1 candidate of the upcoming step
        double vprev[NSW];                                    // this lane's emitting values of the last column
        uint32_t pbits = 0u;                                  // comparison bits of E1
        auto e1_slot = [&](auto par_src, auto slot) {
            constexpr int PS = decltype(par_src)::value;
            constexpr int k = decltype(slot)::value;
            // row 0 of a slot is the state's self loop: its value is still in this lane's register
            double best = vprev[k] + wreg[SH::row0(k)];
#pragma unroll
            for (int d = 1; d < SH::deg(k) - 1; ++d) {
                const int r = SH::row0(k) + d;
                const uint32_t addr = (r & 1) ? (srcreg[r / 2] >> 16) : (srcreg[r / 2] & 0xffffu);
                const double cand = lds_f64<PS * NVP * 8>(addr) + wreg[r];
                if (cand > best) {                            // bit d-1: candidate d beat the running best
                    best = cand;
                    pbits |= 1u << (SH::bit0(k) + d - 1);
                }
            }
            part[k] = best;
        };
        // Silent chain states of column PAR: v[c] = max(entry edges from v, v[c-1] + w), a max-plus scan
        // (Kogge-Stone over the lanes; the summed hop weights of every round are model constants, cwr,
        // -inf for the lanes a round does not reach).  When WITH_E1, the E1 work of the next step is
        // issued between the scan rounds: both only read the emitting values of column PAR, and the
        // independent compare chains fill the shuffle / fp64 latencies of the scan.
        auto chain_phase = [&](auto par, auto with_e1) -> uint32_t {
            constexpr int PAR = decltype(par)::value;
            constexpr bool WITH_E1 = decltype(with_e1)::value;
            typedef std::integral_constant<int, PAR> P;
            if (WITH_E1) pbits = 0u;
            if (QC == 0) {
                if (WITH_E1) {
                    e1_slot(P(), std::integral_constant<int, 0>());
                    if (NSW > 1) e1_slot(P(), std::integral_constant<int, (NSW > 1 ? 1 : 0)>());
                    if (NSW > 2) e1_slot(P(), std::integral_constant<int, (NSW > 2 ? 2 : 0)>());
                    if (NSW > 3) e1_slot(P(), std::integral_constant<int, (NSW > 3 ? 3 : 0)>());
                }
                return 0u;
            }
            double a[QCA], cpw[QCA];
            double A = NINF;
            uint32_t bits = 0u;
            double *v = vteam + PAR * NVP;
#pragma unroll
            for (int q = 0; q < QC; ++q) {
                cpw[q] = cpw_s[q * 32];
                const double c0 = lds_f64<PAR * NVP * 8>(csrc[q] & 0xffffu) + cew_s[(2 * q) * 32];
                const double c1 = lds_f64<PAR * NVP * 8>(csrc[q] >> 16) + cew_s[(2 * q + 1) * 32];
                double best = c0;
                if (c0 > NINF) bits |= 1u << (SH::CBIT + 3 * q);
                if (c1 > best) { best = c1; bits |= 2u << (SH::CBIT + 3 * q); }
                a[q] = best;
                const double t0 = A + cpw[q];
                A = best >= t0 ? best : t0;
            }
            auto round = [&](const int r) {
                const double Al = __shfl_up_sync(0xffffffffu, A, 1 << r);
                const double t0 = Al + cwr_s[r * 32];
                A = A >= t0 ? A : t0;
            };
            round(0);
            if (WITH_E1) e1_slot(P(), std::integral_constant<int, 0>());
            round(1);
            if (WITH_E1 && NSW > 1) e1_slot(P(), std::integral_constant<int, (NSW > 1 ? 1 : 0)>());
            round(2);
            if (WITH_E1 && NSW > 2) e1_slot(P(), std::integral_constant<int, (NSW > 2 ? 2 : 0)>());
            round(3);
            if (WITH_E1 && NSW > 3) e1_slot(P(), std::integral_constant<int, (NSW > 3 ? 3 : 0)>());
            round(4);
            double D = __shfl_up_sync(0xffffffffu, A, 1);
            if (lane == 0) D = NINF;
#pragma unroll
            for (int q = 0; q < QC; ++q) {
                const double t0 = D + cpw[q];
                if (a[q] >= t0) { D = a[q]; bits |= 4u << (SH::CBIT + 3 * q); } else { D = t0; }
                v[CB + (w * 32 + lane) * QC + q] = D;
            }
            return bits;
        };
        typedef std::integral_constant<int, 0> Par0;
        typedef std::integral_constant<int, 1> Par1;

        // ---- initial column ------------------------------------------------------------------------
        double *buf0 = vteam, *buf1 = vteam + NVP;
        for (int i = w * 32 + lane; i < 2 * NVP; i += RW) vteam[i] = NINF;
        team_sync<WPS>(bar_id);
        if (w == 0 && lane == 0) buf0[P_START] = 0.0;
        team_sync<WPS>(bar_id);
        bp[w * 32 + lane] = chain_phase(Par0(), std::false_type());
        team_sync<WPS>(bar_id);

        // ---- forward pass ----------------------------------------------------------------------------
        // Software pipeline inside the warp: the in-edges of an emitting state are split into those whose
        // source is an emitting state / START (rows 0..deg-2 of the slot, "E1") and the one whose source
        // is a silent chain state (last row, "E2").  E1 of step t+1 only needs the emitting values of
        // column t, so it runs in the same basic block as the delete-chain scan of column t and fills the
        // shuffle / fp64 latencies of that scan; E2 + emission follow the barrier that publishes the chain.
        // Two steps per loop iteration make the column parity a compile-time constant.
        double xr = lane < T ? x[lane] : 0.0, xn = 0.0;      // samples of the current / next 32-step block
#pragma unroll
        for (int k = 0; k < NSW; ++k) { part[k] = NINF; vprev[k] = NINF; }
        auto step = [&](const int t, auto par) {
            constexpr int PAR = decltype(par)::value;         // parity of the NEW column: t odd -> 1
            const int tl = (t - 1) & 31;
            if (tl == 0) {
                const int idx = t + 31 + lane;
                xn = idx < T ? x[idx] : 0.0;
            }
            const double xt = __shfl_sync(0xffffffffu, xr, tl);
            const bool xnan = xt != xt;
            double *vnew = vteam + PAR * NVP;
            uint32_t word = pbits;
            // E2: the chain-source candidate (last row of the slot) against the best E1 candidate
            double cc[NSW];
#pragma unroll
            for (int k = 0; k < NSW; ++k) {
                const int r = SH::row0(k) + SH::deg(k) - 1;
                const uint32_t addr = (r & 1) ? (srcreg[r / 2] >> 16) : (srcreg[r / 2] & 0xffffu);
                cc[k] = lds_f64<(1 - PAR) * NVP * 8>(addr) + wreg[r];
            }
#pragma unroll
            for (int k = 0; k < NSW; ++k) {
                double best = part[k];
                if (cc[k] > best) {
                    best = cc[k];
                    word |= 1u << (SH::bit0(k) + SH::deg(k) - 2);
                }
                double e;
                if ((kindmask >> k) & 1u) {
                    e = (xt >= em0[k] && xt <= em1[k]) ? em2[k] : NINF;     // Uniform: -log(hi - lo) inside [lo, hi]
                } else {
                    const double dx = xt - em0[k];            // Normal: c0 - (x - mu)^2 * 1/(2 sigma^2)
                    e = em1[k] - (dx * dx) * em2[k];
                }
                if (xnan) e = 0.0;
                vprev[k] = best + e;
            }
#pragma unroll
            for (int k = 0; k < NSW; ++k) vnew[(w * NSW + k) * 32 + lane] = vprev[k];
            team_sync<WPS>(bar_id);
            if (t == 1 && w == 0 && lane == 0) buf0[P_START] = NINF;    // START exists before the first sample only
            word |= chain_phase(par, std::true_type());       // delete chains of column t + E1 of step t+1
            bp[(size_t)t * RW + w * 32 + lane] = word;
            if (tl == 31) xr = xn;
            team_sync<WPS>(bar_id);
        };
        // E1 of step 1 from the initial column (its START entry is still 0)
        pbits = 0u;
        e1_slot(Par0(), std::integral_constant<int, 0>());
        if (NSW > 1) e1_slot(Par0(), std::integral_constant<int, (NSW > 1 ? 1 : 0)>());
        if (NSW > 2) e1_slot(Par0(), std::integral_constant<int, (NSW > 2 ? 2 : 0)>());
        if (NSW > 3) e1_slot(Par0(), std::integral_constant<int, (NSW > 3 ? 3 : 0)>());
        {
            int t = 1;
            for (; t + 1 <= T; t += 2) {
                step(t, Par1());
                step(t + 1, Par0());
            }
            if (t <= T) step(t, Par1());
        }
        if (w != 0) continue;                     // the team's first warp finishes the sequence

        // ---- END edges: log p = max(v[T][src] + w) ---------------------------------------------------
        const VitFastModelDev m = b.models[cur_model];
        const double *vlast = (T & 1) ? buf1 : buf0;
        const uint16_t *end_src = reinterpret_cast<const uint16_t *>(blob + m.off_end_src);
        const double *end_w = reinterpret_cast<const double *>(blob + m.off_end_w);
        double best = NINF;
        int barg = -1;
        for (int e = lane; e < m.n_end; e += 32) {
            const double cand = vlast[end_src[e]] + end_w[e];
            if (cand > best) { best = cand; barg = e; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ob = __shfl_down_sync(0xffffffffu, best, off);
            const int oa = __shfl_down_sync(0xffffffffu, barg, off);
            if (ob > best || (ob == best && oa >= 0 && (barg < 0 || oa < barg))) { best = ob; barg = oa; }
        }
        best = __shfl_sync(0xffffffffu, best, 0);
        barg = __shfl_sync(0xffffffffu, barg, 0);

        // ---- traceback (all lanes walk in lock step; lane 0 writes) ----------------------------------
        VitResult r;
        r.logp = best; r.n_count = 0; r.t_first = -1; r.t_last = -1; r.pattern_len = 0; r.status = 0; r.reserved = 0;
        if (!(best > NINF) || barg < 0) {
            r.status = 1;
        } else {
            const uint16_t *src_tab = reinterpret_cast<const uint16_t *>(blob + m.off_src);
            const uint16_t *csrc_tab = reinterpret_cast<const uint16_t *>(blob + m.off_csrc);
            const uint8_t *em_flags = blob + m.off_flags;
            int s = end_src[barg], t = T;
            uint8_t *pat = b.pattern ? b.pattern + xo : nullptr;
            uint16_t *path = b.path ? b.path + xo : nullptr;
            bool in_group = false;
            uint8_t last_mod = '0';
            int plen = 0;
            int stage_lo = T + 1;                 // rows [stage_lo, stage_lo + 32) are staged
            long long guard = (long long)(T + 2) * (m.C + 2);
            while (s != P_START) {
                if (--guard < 0 || s > P_START || t < 0) { r.status = 2; break; }
                if (t < stage_lo) {
                    // stage the next STAGE_ROWS back-pointer rows: 16-byte async copies, all in flight at once
                    __syncwarp();
                    stage_lo = t - (SH::STAGE_ROWS - 1) > 0 ? t - (SH::STAGE_ROWS - 1) : 0;
                    const uint4 *src = reinterpret_cast<const uint4 *>(bp + (size_t)stage_lo * RW);
                    const int nvec = (t - stage_lo + 1) * (RW / 4);
                    constexpr int VPL = SH::STAGE_ROWS * (RW / 4) / 32;      // vectors per lane for a full stage
                    const uint32_t sdst = (uint32_t)__cvta_generic_to_shared(stage);
#pragma unroll
                    for (int i = 0; i < VPL; ++i)
                        if (lane + 32 * i < nvec)      // cp.async: global -> shared without staging registers
                            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst + (lane + 32 * i) * 16),
                                         "l"(src + lane + 32 * i)
                                         : "memory");
                    asm volatile("cp.async.wait_all;" ::: "memory");
                    __syncwarp();
                }
                const uint32_t *row = stage + (size_t)(t - stage_lo) * RW;
                if (s < CB) {
                    if (t < 1) { r.status = 2; break; }
                    const unsigned fl = em_flags[s];
                    if (fl & HMM_FLAG_COUNT) ++r.n_count;
                    if (fl & HMM_FLAG_REPEAT) { if (r.t_last < 0) r.t_last = t - 1; r.t_first = t - 1; }
                    if (fl & HMM_FLAG_SEP) {
                        if (in_group) { if (pat && lane == 0) pat[T - 1 - plen] = last_mod; ++plen; in_group = false; }
                    } else {
                        in_group = true;
                        last_mod = (fl & HMM_FLAG_MOD) ? '1' : '0';
                    }
                    if (path && lane == 0) path[t - 1] = (uint16_t)m.perm[s];
                    const int ls = s & 31, slot = s >> 5, sw = slot / NSW, k = slot - sw * NSW;
                    const uint32_t wd = row[sw * 32 + ls];
                    const uint32_t mask = k < NH ? (wd >> (5 * k)) & 31u : (wd >> (5 * NH + 2 * (k - NH))) & 3u;
                    const int arg = mask ? 32 - __clz(mask) : 0;          // the last candidate that beat the running best
                    const int r0 = k < NH ? 6 * k : 6 * NH + 3 * (k - NH);
                    s = src_tab[((size_t)sw * ROWS + r0 + arg) * 32 + ls];
                    --t;
                } else {
                    const int c = s - CB, cw_ = c / (32 * QCA), lc = (c / QCA) & 31, q = c % QCA;
                    const uint32_t wd = (row[cw_ * 32 + lc] >> (SH::CBIT + 3 * q)) & 7u;
                    const int arg = (wd & 4u) ? ((wd & 2u) ? 2 : (wd & 1u)) : 0;   // entry edge 1/2, or 0: previous chain state
                    if (arg == 0) s = s - 1; else s = csrc_tab[((size_t)(cw_ * QCA + q) * 2 + arg - 1) * 32 + lc];
                }
            }
            if (in_group) { if (pat && lane == 0) pat[T - 1 - plen] = last_mod; ++plen; }
            if (r.status == 0 && t != 0) r.status = 2;
            r.pattern_len = plen;
        }
        if (lane == 0) b.res[seq] = r;
    }
}

// instantiated shapes: (WPS, NH, NL, QC)
#define STRIQUE_VIT_SHAPES(X) X(1, 1, 0, 0) X(1, 2, 0, 0) X(2, 2, 2, 2)

template <int WPS, int NH, int NL, int QC>
int launch_team_t(strique_ctx *ctx, const VitFastBatch &b) {
    typedef FastShape<WPS, NH, NL, QC> SH;
    const size_t smem = SH::smem_bytes(b.blob_cap);
    if (smem > 110 * 1024) FAIL(ctx, STRIQUE_EUNSUPPORTED, "team Viterbi: model image too large");
    auto kern = viterbi_team_kernel<WPS, NH, NL, QC>;
    CUDA_TRY(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int grid = ctx->num_sms * 2;
    if (grid > b.n_tasks) grid = b.n_tasks;
    kern<<<grid, 256, smem, ctx->stream>>>(b);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return STRIQUE_OK;
}

struct ShapeInfo {
    int wps, nh, nl, qc;
};
#define X(a, b_, c, d) {a, b_, c, d},
const ShapeInfo kShapes[] = {STRIQUE_VIT_SHAPES(X)};
#undef X

}  // namespace

int viterbi_fast_launch(strique_ctx *ctx, const VitFastShape &shape, const VitFastBatch &b) {
    if (b.n_tasks == 0) return STRIQUE_OK;
#define X(a, b_, c, d) \
    if (shape.wps == a && shape.nh == b_ && shape.nl == c && shape.qc == d) return launch_team_t<a, b_, c, d>(ctx, b);
    STRIQUE_VIT_SHAPES(X)
#undef X
    FAIL(ctx, STRIQUE_EUNSUPPORTED, "team Viterbi: shape not instantiated");
}

int viterbi_fast_teams(const VitFastShape &shape) { return 8 / shape.wps; }

// Packs a compiled HMM for the team kernel if one of the instantiated shapes fits; otherwise leaves
// m->shape.wps == 0 and the generic kernel (viterbi.cu) serves the model.
int viterbi_fast_pack(strique_ctx *ctx, const strique_hmm_desc *d, HmmModel *m) {
    static const double SQRT_2_PI = 2.50662827463;   // pomegranate's truncated constant (distributions.pyx)
    const int E = d->n_emit, C = d->n_chain, START = E + C;
    m->shape = VitFastShape();
    auto degree = [&](int l) { return d->in_ptr[l + 1] - d->in_ptr[l]; };
    int n_hi = 0, max_deg = 0;
    for (int l = 0; l < E; ++l) {
        max_deg = std::max(max_deg, degree(l));
        int nc = 0;
        for (int e = d->in_ptr[l]; e < d->in_ptr[l + 1]; ++e) nc += (d->in_src[e] >= E && d->in_src[e] < E + C) ? 1 : 0;
        int nself = 0;
        for (int e = d->in_ptr[l]; e < d->in_ptr[l + 1]; ++e) nself += d->in_src[e] == l ? 1 : 0;
        if (degree(l) - nc - nself > 1) ++n_hi;                // needs a high slot (self + 4 gathers + chain row)
    }
    if (max_deg > 6 || d->n_end > 64) return STRIQUE_OK;
    // chains: maximal runs of chain states linked by finite predecessor weights
    std::vector<std::pair<int, int>> chains;   // (first, length)
    for (int c = 0; c < C; ++c) {
        if (c == 0 || !(d->chain_pred_logw[c] > -INFINITY)) chains.push_back({c, 1});
        else chains.back().second++;
        if (d->chain_in_ptr[c + 1] - d->chain_in_ptr[c] > 2) return STRIQUE_OK;
    }
    int max_chain = 0;
    for (auto &ch : chains) max_chain = std::max(max_chain, ch.second);
    const ShapeInfo *pick = nullptr;
    for (const ShapeInfo &s : kShapes) {
        if (n_hi > s.wps * s.nh * 32 || E > s.wps * (s.nh + s.nl) * 32) continue;
        if ((int)chains.size() > s.wps || max_chain > s.qc * 32) continue;
        if (C > 0 && s.qc == 0) continue;
        pick = &s;
        break;
    }
    if (!pick) return STRIQUE_OK;
    const int WPS = pick->wps, NH = pick->nh, NL = pick->nl, QC = pick->qc, NSW = NH + NL, ROWS = NH * 6 + NL * 3;
    const int QCA = QC > 0 ? QC : 1;
    const int CB = WPS * NSW * 32, P_START = CB + WPS * QC * 32, P_NEG = P_START + 1;
    auto row0 = [&](int k) { return k < NH ? 6 * k : 6 * NH + 3 * (k - NH); };
    // ---- placement of emitting states: high-degree first into the high slots, spread over the warps ----
    std::vector<int32_t> byDeg(E);
    std::iota(byDeg.begin(), byDeg.end(), 0);
    std::stable_sort(byDeg.begin(), byDeg.end(), [&](int a, int b) { return degree(a) > degree(b); });
    std::vector<int> pos_of(E, -1);
    std::vector<int> state_at(CB, -1);
    // slot visiting order: high slots of all warps (round robin over warps), then low slots
    std::vector<std::pair<int, int>> hi_slots, lo_slots;
    for (int k = 0; k < NH; ++k) for (int w = 0; w < WPS; ++w) hi_slots.push_back({w, k});
    for (int k = NH; k < NSW; ++k) for (int w = 0; w < WPS; ++w) lo_slots.push_back({w, k});
    // A high slot has 5 rows for emitting / START sources + 1 row for a chain source, a low slot 2 + 1.
    auto n_chain_src = [&](int l) {
        int n = 0;
        for (int e = d->in_ptr[l]; e < d->in_ptr[l + 1]; ++e) n += (d->in_src[e] >= E && d->in_src[e] < E + C) ? 1 : 0;
        return n;
    };
    // gather rows needed besides the self loop (row 0) and the chain source (last row): <= 4 high, <= 1 low
    auto n_other = [&](int l) {
        int n = 0;
        for (int e = d->in_ptr[l]; e < d->in_ptr[l + 1]; ++e)
            n += (d->in_src[e] != l && !(d->in_src[e] >= E && d->in_src[e] < E + C)) ? 1 : 0;
        return n;
    };
    std::vector<int32_t> hi_list, lo_list;
    for (int l : byDeg) {
        if (n_chain_src(l) > 1 || n_other(l) > 4) return STRIQUE_OK;   // does not fit the row template: generic kernel
        if (n_other(l) > 1) hi_list.push_back(l);
    }
    for (int l : byDeg)
        if (n_other(l) <= 1) {
            if ((int)hi_list.size() < WPS * NH * 32 && E - (int)hi_list.size() > WPS * NL * 32)
                hi_list.push_back(l);                          // only when the low slots would overflow
            else
                lo_list.push_back(l);
        }
    if ((int)hi_list.size() > WPS * NH * 32 || (int)lo_list.size() > WPS * NL * 32) return STRIQUE_OK;
    auto place = [&](const std::vector<std::pair<int, int>> &slots, const std::vector<int32_t> &list) {
        size_t cursor = 0;
        for (auto &sl : slots)
            for (int lane = 0; lane < 32 && cursor < list.size(); ++lane) {
                const int p = (sl.first * NSW + sl.second) * 32 + lane;
                state_at[p] = list[cursor];
                pos_of[list[cursor]] = p;
                ++cursor;
            }
    };
    place(hi_slots, hi_list);
    place(lo_slots, lo_list);
    // ---- chain placement: chain i -> warp i -----------------------------------------------------------
    std::vector<int> cpos(C, -1);
    for (size_t i = 0; i < chains.size(); ++i)
        for (int j = 0; j < chains[i].second; ++j)
            cpos[chains[i].first + j] = CB + ((int)i * 32 + j / QCA) * QCA + j % QCA;
    auto vpos = [&](int src) -> int {
        if (src >= 0 && src < E) return pos_of[src];
        if (src >= E && src < E + C) return cpos[src - E];
        if (src == START) return P_START;
        return -1;
    };
    // ---- image ------------------------------------------------------------------------------------------
    VitFastModelDev &f = m->fast;
    memset(&f, 0, sizeof(f));
    size_t off = 0;
    auto section = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 16); return (int)o; };
    f.off_w = section((size_t)WPS * ROWS * 32 * 8);
    f.off_em = section((size_t)3 * CB * 8);
    f.off_predw = section((size_t)WPS * QCA * 32 * 8);
    f.off_cw = section((size_t)WPS * QCA * 2 * 32 * 8);
    f.off_wr = section((size_t)WPS * 5 * 32 * 8);
    f.off_end_w = section((size_t)d->n_end * 8);
    f.off_src = section((size_t)WPS * ROWS * 32 * 2);
    f.off_csrc = section((size_t)WPS * QCA * 2 * 32 * 2);
    f.off_end_src = section((size_t)d->n_end * 2);
    f.off_flags = section((size_t)CB);
    f.blob_bytes = (int)off;
    f.n_end = d->n_end;
    f.C = C;
    std::vector<unsigned char> blob(off, 0);
    double *bw = (double *)(blob.data() + f.off_w);
    uint16_t *bs = (uint16_t *)(blob.data() + f.off_src);
    double *be = (double *)(blob.data() + f.off_em);
    uint8_t *bf = blob.data() + f.off_flags;
    double *bpw = (double *)(blob.data() + f.off_predw);
    double *bcw = (double *)(blob.data() + f.off_cw);
    uint16_t *bcs = (uint16_t *)(blob.data() + f.off_csrc);
    double *bew = (double *)(blob.data() + f.off_end_w);
    uint16_t *bes = (uint16_t *)(blob.data() + f.off_end_src);
    for (int i = 0; i < WPS * ROWS * 32; ++i) { bw[i] = 0.0; bs[i] = (uint16_t)P_NEG; }
    std::vector<int32_t> perm(CB, 0);
    for (int p = 0; p < CB; ++p) {
        const int slot = p / 32, lane = p % 32, w = slot / NSW, k = slot % NSW;
        const int l = state_at[p];
        if (l < 0) {   // padding state: uniform over an empty range, never reachable
            bf[p] = 0x80 | 0x40; be[p] = 1.0; be[CB + p] = 0.0; be[2 * CB + p] = -INFINITY;
            continue;
        }
        perm[p] = l;
        // row 0: the self loop (weight -inf if the state has none; the kernel reads the value from its
        // own register), rows 1..: other emitting / START sources, last row: the chain source
        int kk = 1;
        const int last_row = row0(k) + (k < NH ? 5 : 2);
        bw[((size_t)w * ROWS + row0(k)) * 32 + lane] = -INFINITY;
        bs[((size_t)w * ROWS + row0(k)) * 32 + lane] = (uint16_t)p;
        for (int e = d->in_ptr[l]; e < d->in_ptr[l + 1]; ++e) {
            const int v = vpos(d->in_src[e]);
            if (v < 0) FAIL(ctx, STRIQUE_EINVAL, "hmm: in-edge source out of range");
            const bool from_chain = d->in_src[e] >= E && d->in_src[e] < E + C;
            const int row = from_chain ? last_row : (d->in_src[e] == l ? row0(k) : row0(k) + kk++);
            if (row >= last_row && !from_chain) FAIL(ctx, STRIQUE_EINVAL, "hmm: team-kernel row template overflow");
            bw[((size_t)w * ROWS + row) * 32 + lane] = d->in_logw[e];
            bs[((size_t)w * ROWS + row) * 32 + lane] = (uint16_t)v;
        }
        const uint8_t fl = d->emit_flags ? (d->emit_flags[l] & 0x7f) : 0;
        const double a = d->emit_a[l], b = d->emit_b[l];
        if (d->emit_kind[l] == 0) {
            bf[p] = fl;
            be[p] = a; be[CB + p] = -log(b * SQRT_2_PI); be[2 * CB + p] = b > 0 ? 1.0 / (2.0 * (b * b)) : 0.0;
        } else {
            bf[p] = fl | 0x80;
            be[p] = a; be[CB + p] = b; be[2 * CB + p] = -log(b - a);
        }
    }
    for (int i = 0; i < WPS * QCA * 32; ++i) bpw[i] = -INFINITY;
    for (int i = 0; i < WPS * QCA * 2 * 32; ++i) { bcw[i] = 0.0; bcs[i] = (uint16_t)P_NEG; }
    for (size_t i = 0; i < chains.size(); ++i)
        for (int j = 0; j < chains[i].second; ++j) {
            const int c = chains[i].first + j, lane = j / QCA, q = j % QCA, w = (int)i;
            bpw[(w * QCA + q) * 32 + lane] = j == 0 ? -INFINITY : d->chain_pred_logw[c];
            int kk = 0;
            for (int e = d->chain_in_ptr[c]; e < d->chain_in_ptr[c + 1]; ++e, ++kk) {
                const int src = d->chain_in_src[e];
                if (!((src >= 0 && src < E) || src == START)) FAIL(ctx, STRIQUE_EINVAL, "hmm: chain entry edges must come from emitting states or START");
                bcw[((w * QCA + q) * 2 + kk) * 32 + lane] = d->chain_in_logw[e];
                bcs[((w * QCA + q) * 2 + kk) * 32 + lane] = (uint16_t)vpos(src);
            }
        }
    // summed hop weights seen by every round of the kernel's Kogge-Stone scan (same association as
    // the shuffle recurrence W <- W(lane - off) + W), -inf where the round does not reach
    double *bwr = (double *)(blob.data() + f.off_wr);
    for (int w = 0; w < WPS; ++w) {
        double W[32];
        for (int lane = 0; lane < 32; ++lane) {
            W[lane] = 0.0;
            for (int q = 0; q < QCA; ++q) W[lane] += bpw[(w * QCA + q) * 32 + lane];
        }
        for (int r = 0; r < 5; ++r) {
            const int off_ = 1 << r;
            double Wn[32];
            for (int lane = 0; lane < 32; ++lane) {
                bwr[(w * 5 + r) * 32 + lane] = lane >= off_ ? W[lane] : -INFINITY;
                Wn[lane] = lane >= off_ ? W[lane - off_] + W[lane] : W[lane];
            }
            memcpy(W, Wn, sizeof(W));
        }
    }
    for (int e = 0; e < d->n_end; ++e) {
        const int v = vpos(d->end_src[e]);
        if (v < 0 || v == P_START) FAIL(ctx, STRIQUE_EINVAL, "hmm: END edge source out of range");
        bew[e] = d->end_logw[e];
        bes[e] = (uint16_t)v;
    }
    void *dblob = nullptr, *dperm = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&dblob, blob.size()));
    CUDA_TRY(ctx, cudaMalloc(&dperm, (size_t)CB * 4));
    CUDA_TRY(ctx, cudaMemcpy(dblob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(dperm, perm.data(), (size_t)CB * 4, cudaMemcpyHostToDevice));
    ctx->owned.push_back(dblob);
    ctx->owned.push_back(dperm);
    f.blob = (const unsigned char *)dblob;
    f.perm = (const int32_t *)dperm;
    m->shape.wps = WPS; m->shape.nh = NH; m->shape.nl = NL; m->shape.qc = QC;
    return STRIQUE_OK;
}

}  // namespace strique
