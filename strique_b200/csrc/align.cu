// Alignment kernels (see align.cuh for the design summary).
#include "align.cuh"

#include <math.h>

#include <type_traits>

namespace strique {

namespace {

// SeqAn TraceBitMap_ (seqan/align/dp_trace_segment.h / dp_profile.h:124-133)
constexpr unsigned T_DIAG = 1, T_HOR = 2, T_VER = 4, T_HOPEN = 8, T_VOPEN = 16, T_MAXH = 32, T_MAXV = 64;

__device__ __forceinline__ int clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

// ---------------------------------------------------------------------------------------------
// Score table: lut[code][level] = max(off - (float)pow((double)|v_code - y_level|, 1.2), min)
// (src/score_distance.h:117-122; the subtraction is fp32, pow is fp64, the cast rounds to fp32).
// CUDA's fp64 pow is accurate to 2 ulp; an entry whose fp64 result lies within 16 fp64-ulps of
// an fp32 rounding midpoint is reported in lut_fix so the host can re-evaluate it with libm and
// patch it -- this keeps the table bit-identical to the reference's.
// ---------------------------------------------------------------------------------------------
// Distances beyond `far` (host: ((off - min) * (1 + 1e-5))^(1/1.2)) score exactly dist_min without evaluating pow:
// there pow(d, 1.2) exceeds off - min by far more than any rounding in the chain, so off - fx <= min.  With the
// reference's parameters (off 16, min 0) that is every |v - y| > 10.08, about three quarters of the table.
__global__ void __launch_bounds__(256) align_build_lut_kernel(AlignBatch b, const int32_t *task_K, const int32_t *task_S,
                                                              int n_tasks, const float far, const int tile_codes) {
    // One CTA = tile_codes (32; 16 for the 64-rows-per-lane kernels, whose rows would not fit) codes of one task; a warp takes one flank level at a time with one code per lane.  The code
    // values rise with the code, so the lanes near a level form a contiguous run and most warps see only far
    // pairs and skip pow altogether.  The tile goes through shared memory so that the table rows (code-major)
    // are still written coalesced.
    extern __shared__ float tile[];                  // [tile_codes][row_len + 1]
    const int t = blockIdx.x;             // (tasks along x: no 65535 limit)
    if (t >= n_tasks) return;
    const int K = task_K[t], S = task_S[t];
    const int f = b.task_flank[t], sg = b.task_sig[t];
    const int nlev_in = b.flank_off[f + 1] - b.flank_off[f];
    const int rows = nlev_in * b.samples;
    const int nlev = rows / S;            // levels as seen by the kernel
    const int row_len = 32 * K, pitch = row_len + 1;
    const float *vals = b.code_values + (size_t)sg * b.n_code_values;
    const float *lev = b.flank_levels + b.flank_off[f];
    float *lut = b.lut + (size_t)t * b.lut_task_stride;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    for (int c0 = blockIdx.y * tile_codes; c0 < b.n_code_values; c0 += gridDim.y * tile_codes) {
        const int c = c0 + lane;
        const bool mine = lane < tile_codes && c < b.n_code_values;
        const float h = mine ? vals[c] : 0.f;
        for (int u = warp; u < row_len; u += n_warps) {
            float out = 0.f;
            if (u < nlev && mine) {
                const float v = lev[(u * S) / b.samples];
                const float d = h > v ? h - v : v - h;
                if (d > far) {
                    out = b.p.dist_min;
                } else {
                    const double x = pow((double)d, 1.2);
                    const float fx = (float)x;
                    const unsigned long long bits = (unsigned long long)__double_as_longlong(x);
                    const long long low = (long long)(bits & 0x1FFFFFFFull) - 0x10000000ll;
                    if ((low < 0 ? -low : low) <= 16 && b.lut_fix != nullptr) {
                        unsigned long long k = atomicAdd(b.lut_fix, 1ull);
                        if (k < (unsigned long long)b.lut_fix_cap) {       // where, and the two operands (one read-back)
                            b.lut_fix[1 + 2 * k] = ((unsigned long long)t << 40) | (unsigned long long)(c * row_len + u);
                            b.lut_fix[2 + 2 * k] = ((unsigned long long)__float_as_uint(h) << 32) | __float_as_uint(v);
                        }
                    }
                    const float s = b.p.dist_offset - fx;
                    out = s > b.p.dist_min ? s : b.p.dist_min;
                }
            }
            if (lane < tile_codes) tile[lane * pitch + u] = out;
        }
        __syncthreads();
        const int n_codes = min(tile_codes, b.n_code_values - c0);
        // storage order inside a code row: [k][lane] for level u = lane * K + k, so that the scan's K loads per
        // column are one 128-byte line each
        for (int e = threadIdx.x; e < n_codes * row_len; e += blockDim.x) {
            const int cc = e / row_len, p = e - cc * row_len;
            const int u = (p & 31) * K + (p >> 5);
            lut[(size_t)(c0 + cc) * row_len + p] = tile[cc * pitch + u];
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// One systolic sweep of a warp over signal columns (j0, j1] of one task.
//   Sv/Hv      : this lane's R rows of the S and H matrices at column j0 on entry, j1 on exit
//   diag_next  : S[j0][lane*R] (value above the strip in the previous column) on entry
// TRACE = false: score scan; tracks the best last-row cell and writes checkpoints.
// TRACE = true : additionally packs SeqAn's trace flags, 4 bits per cell, into `trace`.
// ---------------------------------------------------------------------------------------------
template <int K, int S, bool TRACE>
struct Sweep {
    static constexpr int R = K * S;
    static constexpr int WP = (R + 31) / 32;     // trace words per flag plane
    static constexpr int W = 4 * WP;             // trace words per lane and column: planes gap | maxh | hopen | vopen

    __device__ __forceinline__ static void run(
        const uint16_t *__restrict__ codes, const int N, const float *__restrict__ lut, const int j0, const int j1,
        const int lane, const int nl, const strique_align_params &p, float (&Sv)[R], float (&Hv)[R],
        float diag_next,
        // scan
        const int lastlane, const int kL, float &best, int &bestj, float *__restrict__ ckS,
        float *__restrict__ ckH, const int ckpt_rows,
        // trace
        uint32_t *__restrict__ trace, const int capture_j, float &capS, float &capH, float &capV,
        // floats per code in the score table (the table may have been built for more levels per lane than K)
        const int row_len = 32 * K) {
        const float INF = STRIQUE_SEQAN_INF;
        const float geh = p.gap_extension_h, gev = p.gap_extension_v, goh = p.gap_open_h, gov = p.gap_open_v;
        float lutc[K], lutn[K];
        float botS = 0.f, botV = INF;
        const int last_step = (j1 - j0) + nl - 1;
        // software pipeline: scores of the next column are fetched while this one is computed
        // position of level u = lane * K + k inside a code row of a table built for Kt levels per lane: [u % Kt][u / Kt]
        int loff[K];
        {
            const int Kt = row_len >> 5;
#pragma unroll
            for (int k = 0; k < K; ++k) { const int u = lane * K + k; loff[k] = (u % Kt) * 32 + u / Kt; }
        }
        {
            const int c = codes[clampi(j0 + 1 - lane - 1, 0, N - 1)];
            const float *row = lut + (size_t)c * row_len;
#pragma unroll
            for (int k = 0; k < K; ++k) lutc[k] = __ldg(row + loff[k]);
        }
        int code_nx = codes[clampi(j0 + 2 - lane - 1, 0, N - 1)];
        for (int s = 1; s <= last_step; ++s) {
            const int j = j0 + s - lane;
            {
                const float *row = lut + (size_t)code_nx * row_len;
#pragma unroll
                for (int k = 0; k < K; ++k) lutn[k] = __ldg(row + loff[k]);
            }
            const int code_nx2 = codes[clampi(j + 1, 0, N - 1)];   // code of column j + 2
            float inS = __shfl_up_sync(0xffffffffu, botS, 1);
            float inV = __shfl_up_sync(0xffffffffu, botV, 1);
            if (lane == 0) { inS = 0.f; inV = INF; }   // DP row 0: free begin, S = 0, V = "infinity"
            if (j > j0 && j <= j1 && lane < nl) {
                float diag = diag_next;
                diag_next = inS;
                float cS = inS, cV = inV;
                uint32_t tw[W];
                if (TRACE) {
#pragma unroll
                    for (int w = 0; w < W; ++w) tw[w] = 0u;
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const float sc = lutc[r / S];
                    const float pS = Sv[r], pH = Hv[r];
                    const float inter = diag + sc;
                    diag = pS;
                    const float eh = pH + geh, oh = pS + goh;
                    const float ev = cV + gev, ov = cS + gov;
                    if (TRACE) {
                        // exact restatement of the SeqAn cell (dp_formula_affine.h:64-128)
                        const bool hopen = eh < oh;
                        const float h = hopen ? oh : eh;
                        const bool vopen = ev < ov;
                        cV = vopen ? ov : ev;
                        const bool maxh = cV < h;
                        const float g = maxh ? h : cV;
                        const bool gap = inter < g;
                        cS = gap ? g : inter;
                        Hv[r] = h;
                        // one bit per flag and row, in four planes: a predicated OR per flag instead of assembling and
                        // shifting a nibble (the trace pass is bound by the integer pipe: 27 -> 22 instructions per cell)
                        if (gap) tw[r / 32] |= 1u << (r % 32);
                        if (maxh) tw[WP + r / 32] |= 1u << (r % 32);
                        if (hopen) tw[2 * WP + r / 32] |= 1u << (r % 32);
                        if (vopen) tw[3 * WP + r / 32] |= 1u << (r % 32);
                        if ((r + 1) % S == 0 && j == capture_j && lane == lastlane && kL == r / S) {
                            capS = cS; capH = h; capV = cV;
                        }
                    } else {
                        const float h = fmaxf(eh, oh);
                        cV = fmaxf(ev, ov);
                        cS = fmaxf(fmaxf(inter, h), cV);
                        Hv[r] = h;
                    }
                    Sv[r] = cS;
                }
                botS = cS;
                botV = cV;
                if (TRACE) {
                    // [lane][column][W]: the traceback mostly moves along a diagonal, i.e. through consecutive columns of
                    // one lane's strip -- 8 of them share a 128-byte line of the walker's L1
                    uint32_t *dst = trace + ((size_t)lane * ALIGN_CKPT + (j - j0 - 1)) * W;
#pragma unroll
                    for (int w = 0; w < W; ++w) dst[w] = tw[w];
                } else {
                    if (lane == lastlane) {
                        float last = Sv[S - 1];
#pragma unroll
                        for (int k = 1; k < K; ++k) last = (kL == k) ? Sv[(k + 1) * S - 1] : last;
                        if (last > best) { best = last; bestj = j; }   // strict >: first maximum wins (dp_scout.h:175)
                    }
                    if ((j & (ALIGN_CKPT - 1)) == 0) {
                        const size_t o = (size_t)(j / ALIGN_CKPT - 1) * 2 * ckpt_rows + lane * R + 1;
#pragma unroll
                        for (int r = 0; r < R; ++r) { ckS[o + r] = Sv[r]; ckH[o + r] = Hv[r]; }
                    }
                }
            }
#pragma unroll
            for (int k = 0; k < K; ++k) lutc[k] = lutn[k];
            code_nx = code_nx2;
        }
    }
};


// ---------------------------------------------------------------------------------------------
// Score scan for LINEAR gap costs (gap_open == gap_extension in both directions -- the reference's
// own configuration, scripts/STRique.py:507-512 and configs/STRique.json).  Then, bit for bit,
//     H[j][i] = max(H[j-1][i] + g_h, S[j-1][i] + g_h) = S[j-1][i] + g_h        for j >= 2
//     V[j][i] = max(V[j][i-1] + g_v, S[j][i-1] + g_v) = S[j][i-1] + g_v        for j >= 1
// because S >= H and S >= V in every computed cell and fp32 rounding is monotonic
// (max(fl(a+g), fl(b+g)) == fl(max(a,b)+g)).  The exceptions are SeqAn's initial "infinity":
// H of DP column 0 is the positive denormal INF > S, handled in the j == 1 column, and V of DP row
// 0 is INF > S[j][0] = 0, where INF + g_v == 0 + g_v exactly (the host checks |g| >= 1e-20).
// The cell is 3 FADD + one 3-input FMNMX, no H / V state: half the registers of the affine sweep.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float fmax3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

template <int K, int S>
struct LinSweep {
    static constexpr int R = K * S;

    template <bool FIRST>
    __device__ __forceinline__ static float column(float (&Sv)[R], const float (&lutc)[K], float diag, float cS,
                                                   const float gh, const float gv) {
        const float INF = STRIQUE_SEQAN_INF;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float pS = Sv[r];
            const float inter = diag + lutc[r / S];
            diag = pS;
            const float h = (FIRST ? fmaxf(INF, pS) : pS) + gh;
            const float v = cS + gv;
            cS = fmax3(inter, h, v);
            Sv[r] = cS;
        }
        return cS;
    }

    // Loop structure: the first nl steps (where some lane is at DP column 1 and needs the FIRST
    // variant) run a small general loop; the steady state has ONE straight-line column body (no code
    // variants that would have to be merged with register moves), two steps per iteration so the
    // score rows ping-pong between two register sets, checkpoint columns only add stores around the
    // body, and the best-cell tracking is branch free.
    __device__ __forceinline__ static void run(const uint16_t *__restrict__ codes, const int N,
                                               const float *__restrict__ lut, const int lane, const int nl,
                                               const strique_align_params &p, float (&Sv)[R], float diag_next,
                                               const int lastlane, const int kL, float &best, int &bestj,
                                               float *__restrict__ ckS, float *__restrict__ ckH, const int ckpt_rows) {
        const float gh = p.gap_extension_h, gv = p.gap_extension_v;
        const int row_len = 32 * K;
        float lutc[K], lutn[K];
        float botS = 0.f;
        const int last_step = N + nl - 1;
        const float *lut_lane = lut + lane;      // code row stored [k][lane]
        {
            const int c = codes[clampi(-lane, 0, N - 1)];
            const float *row = lut_lane + (size_t)c * row_len;
#pragma unroll
            for (int k = 0; k < K; ++k) lutc[k] = __ldg(row + k * 32);
        }
        int code_nx = codes[clampi(1 - lane, 0, N - 1)];
        // LASTFULL: the flank ends on the last row of its lane (L == nl * R, e.g. 870 = 29 * 30), so the last DP row
        // is the lane's bottom cell and needs no select chain
        auto track_best = [&](const int j, auto lastfull) {
            float last = Sv[R - 1];
            if (!decltype(lastfull)::value) {
                last = Sv[S - 1];
#pragma unroll
                for (int k = 1; k < K; ++k) last = (kL == k) ? Sv[(k + 1) * S - 1] : last;
            }
            // strict >: first maximum wins (dp_scout.h:175).  Every lane tracks its own bottom row; only the last
            // lane's (best, bestj) is read after the scan, so the lane test is not needed here
            const bool upd = last > best;
            best = upd ? last : best;
            bestj = upd ? j : bestj;
        };
        // general step: any column (DP column 1, checkpoint columns, lanes outside [1, N]); used for the
        // first nl steps and the last nl - 1, where the lanes are not all inside the signal
        auto general = [&](const int st) {
            const int j = st - lane;
            {
                const float *row = lut_lane + (size_t)code_nx * row_len;
#pragma unroll
                for (int k = 0; k < K; ++k) lutn[k] = __ldg(row + k * 32);
            }
            const int code_nx2 = codes[clampi(j + 1, 0, N - 1)];
            float inS = __shfl_up_sync(0xffffffffu, botS, 1);
            if (lane == 0) inS = 0.f;            // DP row 0: free begin, S = 0
            if (j > 0 && j <= N && lane < nl) {
                const bool ck = (j & (ALIGN_CKPT - 1)) == 0;
                const size_t o = (size_t)(j / ALIGN_CKPT - 1) * 2 * ckpt_rows + lane * R + 1;
                if (ck) {
#pragma unroll
                    for (int r = 0; r < R; ++r) ckH[o + r] = Sv[r] + gh;
                }
                const float diag = diag_next;
                diag_next = inS;
                if (j == 1) botS = column<true>(Sv, lutc, diag, inS, gh, gv);
                else botS = column<false>(Sv, lutc, diag, inS, gh, gv);
                if (ck) {
#pragma unroll
                    for (int r = 0; r < R; ++r) ckS[o + r] = Sv[r];
                }
                track_best(j, std::false_type{});
            }
#pragma unroll
            for (int k = 0; k < K; ++k) lutc[k] = lutn[k];
            code_nx = code_nx2;
        };
        // steady step: every lane is at a column 2 <= j <= N, so the column body runs unguarded (idle
        // lanes >= nl compute on zeros).  CK: some lane may sit on a checkpoint column at this step (only the
        // first nl steps of every ALIGN_CKPT); the other steps carry no checkpoint code at all.
        const uint16_t *cptr = codes;           // steady loop: &codes[j + 1] of this lane, advanced every step
        auto steady = [&](const int st, const float (&lc)[K], float (&ln)[K], auto with_ck, auto lastfull) {
            constexpr bool CK = decltype(with_ck)::value;
            const int j = st - lane;
            {
                const float *row = lut_lane + (size_t)code_nx * row_len;
#pragma unroll
                for (int k = 0; k < K; ++k) ln[k] = __ldg(row + k * 32);
            }
            // every lane is at 1 <= j <= N here, so j + 1 needs no clamp: it reads at most 2 codes past the signal,
            // inside the (padded) code buffer, and those values are never used
            const int code_nx2 = *cptr++;
            float inS = __shfl_up_sync(0xffffffffu, botS, 1);
            if (lane == 0) inS = 0.f;
            const bool ck = CK && (j & (ALIGN_CKPT - 1)) == 0 && lane < nl;
            const size_t o = CK ? (size_t)(j / ALIGN_CKPT - 1) * 2 * ckpt_rows + lane * R + 1 : 0;
            if (CK && ck) {                       // H of the checkpoint column: S of the previous column + g_h
#pragma unroll
                for (int r = 0; r < R; ++r) ckH[o + r] = Sv[r] + gh;
            }
            const float diag = diag_next;
            diag_next = inS;
            botS = column<false>(Sv, lc, diag, inS, gh, gv);
            if (CK && ck) {
#pragma unroll
                for (int r = 0; r < R; ++r) ckS[o + r] = Sv[r];
            }
            track_best(j, lastfull);
            code_nx = code_nx2;
        };
        auto steady_loop = [&](int &s, auto lastfull) {
            cptr = codes + max(s - lane + 1, 0);     // lanes >= nl are idle (never negative for the others)
            while (s + 1 <= N) {
                const int m = s & (ALIGN_CKPT - 1);
                if (m >= nl && m <= ALIGN_CKPT - 2) {
                    // both steps of every pair stay inside [nl, ALIGN_CKPT - 1] (mod ALIGN_CKPT): no lane checkpoints
                    int pairs = min((ALIGN_CKPT - m) >> 1, (N - s + 1) >> 1);
                    for (; pairs > 0; --pairs, s += 2) {
                        steady(s, lutc, lutn, std::false_type{}, lastfull);
                        steady(s + 1, lutn, lutc, std::false_type{}, lastfull);
                    }
                } else {
                    steady(s, lutc, lutn, std::true_type{}, lastfull);
                    steady(s + 1, lutn, lutc, std::true_type{}, lastfull);
                    s += 2;
                }
            }
        };
        int s = 1;
        for (; s <= nl && s <= last_step; ++s) general(s);
        if (kL == K - 1) steady_loop(s, std::true_type{});
        else steady_loop(s, std::false_type{});
        for (; s <= last_step; ++s) general(s);
    }
};

// ---------------------------------------------------------------------------------------------
// LinSweep2: the linear-gap scan with PACKED adds and a low register footprint (S even).
//
// Same recurrence as LinSweep, bit for bit (add.rn.f32x2 is two independent IEEE fp32 adds).  Of the three adds of a
// cell, inter = S[j-1][i-1] + score and h = S[j-1][i] + g_h only read the previous column, so two rows share one
// instruction each; v = S[j][i-1] + g_v is the serial chain down the column and stays scalar: 2 packed + 2 scalar
// adds + 2 three-input maxima per TWO cells instead of 6 + 2.
// For both packed adds to take the SAME aligned register pair (S[2m], S[2m+1]) of the previous column, the inter
// pair must be the rows (2m+1, 2m+2) -- and for a pair never to straddle two flank levels (different scores) the
// lane's strip starts one row early: lane l owns the DP rows i = l R + r, r = 0..R-1, i.e. its local row 0 is the
// last row of the previous lane's last level (its score comes from that lane's table entry) and local rows
// 1..R-1 are whole levels; lane 0's local row 0 is DP row 0 (free begin: S = 0, forced).  Checkpoints are stored
// by DP row, so the trace pass is unaffected.
// Measured (C2, 8192 reads): 97.4 ms against 99.5 ms for LinSweep at 16 single-warp CTAs per SM -- far less than the
// column body alone promises (tools/micro/cell.cu, profiles/cell_r02.txt: 173 against 199 cycles per column and
// scheduler at 16 warps per SM, 141 at 28), and MORE resident warps make the real kernel slower (18: 102 ms, 24:
// 102 ms with spills, 28: 107 ms): every task streams its own 164 KB score table through L1, and beyond 16 tables
// per SM the hit rate collapses.  What bounds the scan is the serial add -> max chain down every column, issued in
// order (ncu scan2_r02: 39 % of the stall samples are fixed-latency waits on that chain, issue slots 72 % busy).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(unsigned long long v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
    unsigned long long r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

template <int K, int S>
struct LinSweep2 {
    static constexpr int R = K * S;
    static_assert(S % 2 == 0, "packed pairs must not straddle flank levels");

    // sc[k], k < K: score of this lane's level k for the column's code; sc[K]: of the previous lane's last level
    // scalar form: DP column 1 (FIRST: H of column 0 is SeqAn's "infinity") and the ragged first / last steps
    template <bool FIRST>
    __device__ __forceinline__ static float column(float (&Sv)[R], const float (&sc)[K + 1], float diag, float cS,
                                                   const float gh, const float gv, const bool lane0) {
        const float INF = STRIQUE_SEQAN_INF;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float pS = Sv[r];
            const float inter = diag + (r == 0 ? sc[K] : sc[(r - 1) / S]);
            diag = pS;
            const float h = (FIRST ? fmaxf(INF, pS) : pS) + gh;
            cS = fmax3(inter, h, cS + gv);
            if (r == 0) cS = lane0 ? 0.f : cS;               // DP row 0: free begin
            Sv[r] = cS;
        }
        return cS;
    }

    __device__ __forceinline__ static float column2(float (&Sv)[R], const float (&sc)[K + 1], const float diag, float cS,
                                                    const unsigned long long gh2, const float gv, const bool lane0) {
        float inter = diag + sc[K];
#pragma unroll
        for (int m = 0; m < R / 2; ++m) {
            const unsigned long long P = f2_pack(Sv[2 * m], Sv[2 * m + 1]);
            const float sck = sc[m / (S / 2)];               // rows 2m+1 and 2m+2 lie in the same level
            float h0, h1, i1, i2;
            f2_unpack(f2_add(P, gh2), h0, h1);
            f2_unpack(f2_add(P, f2_pack(sck, sck)), i1, i2);
            cS = fmax3(inter, h0, cS + gv);
            if (m == 0) cS = lane0 ? 0.f : cS;
            Sv[2 * m] = cS;
            cS = fmax3(i1, h1, cS + gv);
            Sv[2 * m + 1] = cS;
            inter = i2;
        }
        return cS;
    }

    // rlk: the last DP row L is local row rlk * S of the last lane (L is a multiple of S).  RL0: rlk == 0 (the
    // reference's 870-sample flanks: 870 = 29 * 30), the last row is the lane's first register -- no select chain
    template <bool RL0>
    __device__ __forceinline__ static void run(const uint16_t *__restrict__ codes, const int N,
                                               const float *__restrict__ lut, const int lane, const int nl,
                                               const strique_align_params &p, float (&Sv)[R], float diag_next,
                                               const int rlk, float &best, int &bestj, float *__restrict__ ckS,
                                               float *__restrict__ ckH, const int ckpt_rows, const int L) {
        const float gh = p.gap_extension_h, gv = p.gap_extension_v;
        const unsigned long long gh2 = f2_pack(gh, gh);
        const int row_len = 32 * K;
        const bool lane0 = lane == 0;
        float lutc[K + 1], lutn[K + 1];
        float botS = 0.f;
        const int last_step = N + nl - 1;
        const float *lut_lane = lut + lane;                       // code row stored [k][lane]
        const int prev_off = (K - 1) * 32 - (lane0 ? 0 : 1);     // previous lane's last level (lane 0: unused)
        auto fetch = [&](const int code, float (&dst)[K + 1]) {
            const float *row = lut_lane + (size_t)code * row_len;
#pragma unroll
            for (int k = 0; k < K; ++k) dst[k] = __ldg(row + k * 32);
            dst[K] = __ldg(row + prev_off);
        };
        fetch(codes[clampi(-lane, 0, N - 1)], lutc);
        int code_nx = codes[clampi(1 - lane, 0, N - 1)];
        auto track_best = [&](const int j) {
            float last = Sv[0];
            if (!RL0) {
#pragma unroll
                for (int k = 1; k < K; ++k) last = (rlk == k) ? Sv[k * S] : last;
            }
            // strict >: first maximum wins (dp_scout.h:175); only the last lane's (best, bestj) is read afterwards
            const bool upd = last > best;
            best = upd ? last : best;
            bestj = upd ? j : bestj;
        };
        // checkpoint column j: H = S[j-1] + g_h (before), S (after).  One lane of the warp is at a checkpoint column in
        // a step, so these stores cost the whole warp their issue slots: pairs (R, the lane offsets and the block
        // offsets are even, the buffer holds 32 R + 1 rows) and no row predicates -- DP row 0 and the rows behind L
        // are written but never read.
        auto store_ck = [&](const int j, const bool before) {
            const size_t o = (size_t)(j / ALIGN_CKPT - 1) * 2 * ckpt_rows + lane * R;
            float2 *dst = reinterpret_cast<float2 *>((before ? ckH : ckS) + o);
#pragma unroll
            for (int m = 0; m < R / 2; ++m) {
                if (before) {
                    float h0, h1;
                    f2_unpack(f2_add(f2_pack(Sv[2 * m], Sv[2 * m + 1]), gh2), h0, h1);
                    dst[m] = make_float2(h0, h1);
                } else {
                    dst[m] = make_float2(Sv[2 * m], Sv[2 * m + 1]);
                }
            }
        };
        // general step: any column (DP column 1, checkpoint columns, lanes outside [1, N])
        auto general = [&](const int st) {
            const int j = st - lane;
            fetch(code_nx, lutn);
            const int code_nx2 = codes[clampi(j + 1, 0, N - 1)];
            const float inS = __shfl_up_sync(0xffffffffu, botS, 1);
            if (j > 0 && j <= N && lane < nl) {
                const bool ck = (j & (ALIGN_CKPT - 1)) == 0;
                if (ck) store_ck(j, true);
                const float diag = diag_next;
                diag_next = inS;
                if (j == 1) botS = column<true>(Sv, lutc, diag, inS, gh, gv, lane0);
                else botS = column<false>(Sv, lutc, diag, inS, gh, gv, lane0);
                if (ck) store_ck(j, false);
                track_best(j);
            }
#pragma unroll
            for (int k = 0; k <= K; ++k) lutc[k] = lutn[k];
            code_nx = code_nx2;
        };
        // steady step: every lane is at a column 2 <= j <= N (idle lanes >= nl compute on zeros)
        const uint16_t *cptr = codes;
        auto steady = [&](const int st, const float (&lc)[K + 1], float (&ln)[K + 1], auto with_ck) {
            constexpr bool CK = decltype(with_ck)::value;
            const int j = st - lane;
            fetch(code_nx, ln);
            const int code_nx2 = *cptr++;                           // reads at most 2 codes past the signal (padded buffer)
            const float inS = __shfl_up_sync(0xffffffffu, botS, 1);
            const bool ck = CK && (j & (ALIGN_CKPT - 1)) == 0 && lane < nl;
            if (CK && ck) store_ck(j, true);
            const float diag = diag_next;
            diag_next = inS;
            botS = column2(Sv, lc, diag, inS, gh2, gv, lane0);
            if (CK && ck) store_ck(j, false);
            track_best(j);
            code_nx = code_nx2;
        };
        int s = 1;
        for (; s <= nl && s <= last_step; ++s) general(s);
        cptr = codes + max(s - lane + 1, 0);
        while (s + 1 <= N) {
            const int m = s & (ALIGN_CKPT - 1);
            if (m >= nl && m <= ALIGN_CKPT - 2) {
                // both steps of every pair stay inside [nl, ALIGN_CKPT - 1] (mod ALIGN_CKPT): no lane checkpoints
                int pairs = min((ALIGN_CKPT - m) >> 1, (N - s + 1) >> 1);
                for (; pairs > 0; --pairs, s += 2) {
                    steady(s, lutc, lutn, std::false_type{});
                    steady(s + 1, lutn, lutc, std::false_type{});
                }
            } else {
                steady(s, lutc, lutn, std::true_type{});
                steady(s + 1, lutn, lutc, std::true_type{});
                s += 2;
            }
        }
        for (; s <= last_step; ++s) general(s);
    }
};

// ---------------------------------------------------------------------------------------------
// LinSweepPair: the linear-gap scan of TWO tasks over the same signal (a read's prefix and suffix flank) by one
// warp.  The two tasks share the column's code, the loop, the shuffle and the step bookkeeping, and ALL three adds
// of the recurrence are packed across the two tasks (add.rn.f32x2 = two independent IEEE fp32 adds, so every value
// is bit-identical to LinSweep's): 3 packed adds + 2 three-input maxima per row for two cells, 2.5 instructions per
// cell instead of 3, and the per-column overhead is paid once for two tasks: 88 instructions per column and task
// against 111 for LinSweep2.  Measured (C2, 8192 reads, 16 warps per SM): 87.1 ms against 94.3 ms; ncu: issue slots
// 62 % busy, the packed adds (48 % of the instructions) wait on the FMA pipe (math-pipe throttle 19 % of the stall
// samples) and on the column chain (fixed-latency wait 26 %).
// Sv[2 r + t] is S of local row r of task t; checkpoints of the pair are stored interleaved (AlignBatch::ckpt_step).
// ---------------------------------------------------------------------------------------------
template <int K, int S>
struct LinSweepPair {
    static constexpr int R = K * S;

    template <bool FIRST>
    __device__ __forceinline__ static void column_scalar(float (&Sv)[2 * R], const float (&sc)[2 * K], const float (&diag)[2],
                                                         float (&cS)[2], const float gh, const float gv) {
        const float INF = STRIQUE_SEQAN_INF;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            float d = diag[t], c = cS[t];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float pS = Sv[2 * r + t];
                const float inter = d + sc[2 * (r / S) + t];
                d = pS;
                const float h = (FIRST ? fmaxf(INF, pS) : pS) + gh;
                c = fmax3(inter, h, c + gv);
                Sv[2 * r + t] = c;
            }
            cS[t] = c;
        }
    }

    __device__ __forceinline__ static unsigned long long column_pair(float (&Sv)[2 * R], const float (&sc)[2 * K],
                                                                     unsigned long long diag2, unsigned long long cS2,
                                                                     const unsigned long long gh2, const unsigned long long gv2) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const unsigned long long P = f2_pack(Sv[2 * r], Sv[2 * r + 1]);
            float iA, iB, hA, hB, vA, vB;
            f2_unpack(f2_add(diag2, f2_pack(sc[2 * (r / S)], sc[2 * (r / S) + 1])), iA, iB);
            diag2 = P;
            f2_unpack(f2_add(P, gh2), hA, hB);
            f2_unpack(f2_add(cS2, gv2), vA, vB);
            const float a = fmax3(iA, hA, vA), b = fmax3(iB, hB, vB);
            Sv[2 * r] = a;
            Sv[2 * r + 1] = b;
            cS2 = f2_pack(a, b);
        }
        return cS2;
    }

    // kL[t]: level (inside its last lane) of the last DP row of task t; LASTFULL: both flanks end on the last row of
    // their last lane (870 = 29 * 30), so the tracked cell is the lane's bottom register
    template <bool LASTFULL>
    __device__ __forceinline__ static void run(const uint16_t *__restrict__ codes, const int N, const float *__restrict__ lutA,
                                               const float *__restrict__ lutB, const int lane, const int nl,
                                               const strique_align_params &p, float (&Sv)[2 * R], unsigned long long diag_next2,
                                               const int kLA, const int kLB, float (&best)[2], int (&bestj)[2],
                                               float *__restrict__ ck, const int ckpt_rows) {
        const float gh = p.gap_extension_h, gv = p.gap_extension_v;
        const unsigned long long gh2 = f2_pack(gh, gh), gv2 = f2_pack(gv, gv);
        const int row_len = 32 * K;
        float lutc[2 * K], lutn[2 * K];
        unsigned long long botS2 = 0ull;
        const int last_step = N + nl - 1;
        const float *lutA_lane = lutA + lane, *lutB_lane = lutB + lane;      // code row stored [k][lane]
        auto fetch = [&](const int code, float (&dst)[2 * K]) {
            const float *ra = lutA_lane + (size_t)code * row_len, *rb = lutB_lane + (size_t)code * row_len;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                dst[2 * k] = __ldg(ra + k * 32);
                dst[2 * k + 1] = __ldg(rb + k * 32);
            }
        };
        fetch(codes[clampi(-lane, 0, N - 1)], lutc);
        int code_nx = codes[clampi(1 - lane, 0, N - 1)];
        auto track_best = [&](const int j) {
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                float last = Sv[2 * (R - 1) + t];
                if (!LASTFULL) {
                    const int kL = t ? kLB : kLA;
                    last = Sv[2 * (S - 1) + t];
#pragma unroll
                    for (int k = 1; k < K; ++k) last = (kL == k) ? Sv[2 * ((k + 1) * S - 1) + t] : last;
                }
                // strict >: first maximum wins (dp_scout.h:175); only the last lane's (best, bestj) is read afterwards
                const bool upd = last > best[t];
                best[t] = upd ? last : best[t];
                bestj[t] = upd ? j : bestj[t];
            }
        };
        // checkpoint column j of both tasks, interleaved: [block][S | H][row i][task]; H = S[j-1] + g_h (before), S (after)
        auto store_ck = [&](const int j, const bool before) {
            float2 *dst = reinterpret_cast<float2 *>(ck + (size_t)(j / ALIGN_CKPT - 1) * 4 * ckpt_rows + (before ? 2 * ckpt_rows : 0)) +
                          (lane * R + 1);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (before) {
                    float h0, h1;
                    f2_unpack(f2_add(f2_pack(Sv[2 * r], Sv[2 * r + 1]), gh2), h0, h1);
                    dst[r] = make_float2(h0, h1);
                } else {
                    dst[r] = make_float2(Sv[2 * r], Sv[2 * r + 1]);
                }
            }
        };
        auto general = [&](const int st) {
            const int j = st - lane;
            fetch(code_nx, lutn);
            const int code_nx2 = codes[clampi(j + 1, 0, N - 1)];
            unsigned long long inS2 = __shfl_up_sync(0xffffffffu, botS2, 1);
            if (lane == 0) inS2 = 0ull;              // DP row 0: free begin, S = 0
            if (j > 0 && j <= N && lane < nl) {
                const bool ck_col = (j & (ALIGN_CKPT - 1)) == 0;
                if (ck_col) store_ck(j, true);
                float diag[2], cS[2];
                f2_unpack(diag_next2, diag[0], diag[1]);
                f2_unpack(inS2, cS[0], cS[1]);
                diag_next2 = inS2;
                if (j == 1) column_scalar<true>(Sv, lutc, diag, cS, gh, gv);
                else column_scalar<false>(Sv, lutc, diag, cS, gh, gv);
                botS2 = f2_pack(cS[0], cS[1]);
                if (ck_col) store_ck(j, false);
                track_best(j);
            }
#pragma unroll
            for (int k = 0; k < 2 * K; ++k) lutc[k] = lutn[k];
            code_nx = code_nx2;
        };
        const uint16_t *cptr = codes;
        auto steady = [&](const int st, const float (&lc)[2 * K], float (&ln)[2 * K], auto with_ck) {
            constexpr bool CK = decltype(with_ck)::value;
            const int j = st - lane;
            fetch(code_nx, ln);
            const int code_nx2 = *cptr++;            // reads at most 2 codes past the signal (padded buffer)
            unsigned long long inS2 = __shfl_up_sync(0xffffffffu, botS2, 1);
            if (lane == 0) inS2 = 0ull;
            const bool ck_col = CK && (j & (ALIGN_CKPT - 1)) == 0 && lane < nl;
            if (CK && ck_col) store_ck(j, true);
            const unsigned long long diag2 = diag_next2;
            diag_next2 = inS2;
            botS2 = column_pair(Sv, lc, diag2, inS2, gh2, gv2);
            if (CK && ck_col) store_ck(j, false);
            track_best(j);
            code_nx = code_nx2;
        };
        int s = 1;
        for (; s <= nl && s <= last_step; ++s) general(s);
        cptr = codes + max(s - lane + 1, 0);
        while (s + 1 <= N) {
            const int m = s & (ALIGN_CKPT - 1);
            if (m >= nl && m <= ALIGN_CKPT - 2) {
                int pairs = min((ALIGN_CKPT - m) >> 1, (N - s + 1) >> 1);
#ifdef ALIGN_PAIR_UNROLL
                constexpr int kUnroll = ALIGN_PAIR_UNROLL;
#pragma unroll kUnroll
#endif
                for (; pairs > 0; --pairs, s += 2) {
                    steady(s, lutc, lutn, std::false_type{});
                    steady(s + 1, lutn, lutc, std::false_type{});
                }
            } else {
                steady(s, lutc, lutn, std::true_type{});
                steady(s + 1, lutn, lutc, std::true_type{});
                s += 2;
            }
        }
        for (; s <= last_step; ++s) general(s);
    }
};

struct TaskGeom {
    int t, N, f, L, nl, lastlane, kL;
    const uint16_t *codes;
    const float *lut, *col0;
};

template <int K, int S>
__device__ __forceinline__ TaskGeom task_geom(const AlignBatch &b, int t) {
    constexpr int R = K * S;
    TaskGeom g;
    g.t = t;
    const int sg = b.task_sig[t];
    g.f = b.task_flank[t];
    g.N = (int)(b.sig_off[sg + 1] - b.sig_off[sg]);
    g.codes = b.codes + b.sig_off[sg];
    g.L = (b.flank_off[g.f + 1] - b.flank_off[g.f]) * b.samples;
    g.nl = (g.L + R - 1) / R;
    g.lastlane = (g.L - 1) / R;
    g.kL = ((g.L - 1) % R) / S;
    g.lut = b.lut + (size_t)t * b.lut_task_stride;
    g.col0 = b.col0 + (size_t)g.f * b.col0_stride;
    return g;
}

// ---------------------------------------------------------------------------------------------
// Pass 1: score scan.  Single-warp CTAs pull tasks (longest first) from a global queue.
// ---------------------------------------------------------------------------------------------
template <int K, int S, bool LIN>
__global__ void __launch_bounds__(32) align_scan_kernel(AlignBatch b, AlignGroup grp) {
    constexpr int R = K * S;
    const int lane = threadIdx.x;
    const float INF = STRIQUE_SEQAN_INF;
    for (;;) {
        int q = 0;
        if (lane == 0) q = atomicAdd(b.queue + 0, 1);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= grp.n_tasks) break;
        const TaskGeom g = task_geom<K, S>(b, grp.order[q]);
        float Sv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = lane * R + r + 1;
            Sv[r] = i <= g.L ? g.col0[i] : 0.f;
        }
        const float diag0 = lane * R <= g.L ? g.col0[lane * R] : 0.f;
        float best = INF;
        int bestj = -1;
        if (lane == g.lastlane && g.col0[g.L] > INF) { best = g.col0[g.L]; bestj = 0; }
        float *ck = b.ckpt + b.ckpt_off[g.t];
        if (LIN) {
            LinSweep<K, S>::run(g.codes, g.N, g.lut, lane, g.nl, b.p, Sv, diag0, g.lastlane, g.kL, best, bestj, ck,
                                ck + b.ckpt_rows, b.ckpt_rows);
        } else {
            float Hv[R];
#pragma unroll
            for (int r = 0; r < R; ++r) Hv[r] = INF;
            float d0 = 0.f, d1 = 0.f, d2 = 0.f;
            Sweep<K, S, false>::run(g.codes, g.N, g.lut, 0, g.N, lane, g.nl, b.p, Sv, Hv, diag0, g.lastlane, g.kL,
                                    best, bestj, ck, ck + b.ckpt_rows, b.ckpt_rows, nullptr, -1, d0, d1, d2);
        }
        if (lane == g.lastlane) {
            b.res[g.t].score = best;
            b.res[g.t].best_j = bestj;
        }
    }
}

// Pass 1 with LinSweep2 (linear gap costs, S even).
template <int K, int S>
__global__ void __launch_bounds__(32, ALIGN_WARPS_PER_SM_PACKED) align_scan2_kernel(AlignBatch b, AlignGroup grp) {
    constexpr int R = K * S;
    const int lane = threadIdx.x;
    const float INF = STRIQUE_SEQAN_INF;
    for (;;) {
        int q = 0;
        if (lane == 0) q = atomicAdd(b.queue + 0, 1);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= grp.n_tasks) break;
        const int t = grp.order[q];
        const int sg = b.task_sig[t], f = b.task_flank[t];
        const int N = (int)(b.sig_off[sg + 1] - b.sig_off[sg]);
        const int L = (b.flank_off[f + 1] - b.flank_off[f]) * b.samples;
        const float *col0 = b.col0 + (size_t)f * b.col0_stride;
        const int nl = L / R + 1, lastlane = L / R, rlk = (L % R) / S;      // DP rows 0..L over the lanes
        float Sv[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = lane * R + r;
            Sv[r] = i <= L ? col0[i] : 0.f;
        }
        const float diag0 = (lane > 0 && lane * R - 1 <= L) ? col0[lane * R - 1] : 0.f;
        float best = INF;
        int bestj = -1;
        if (lane == lastlane && col0[L] > INF) { best = col0[L]; bestj = 0; }
        float *ck = b.ckpt + b.ckpt_off[t];
        if (rlk == 0)
            LinSweep2<K, S>::template run<true>(b.codes + b.sig_off[sg], N, b.lut + (size_t)t * b.lut_task_stride, lane, nl,
                                                b.p, Sv, diag0, rlk, best, bestj, ck, ck + b.ckpt_rows, b.ckpt_rows, L);
        else
            LinSweep2<K, S>::template run<false>(b.codes + b.sig_off[sg], N, b.lut + (size_t)t * b.lut_task_stride, lane, nl,
                                                 b.p, Sv, diag0, rlk, best, bestj, ck, ck + b.ckpt_rows, b.ckpt_rows, L);
        if (lane == lastlane) {
            b.res[t].score = best;
            b.res[t].best_j = bestj;
        }
    }
}

template <int K, int S>
__global__ void __launch_bounds__(32, align_pair_warps(K)) align_scan_pair_kernel(AlignBatch b, AlignGroup grp) {
    constexpr int R = K * S;
    const int lane = threadIdx.x;
    const float INF = STRIQUE_SEQAN_INF;
    for (;;) {
        int q = 0;
        if (lane == 0) q = atomicAdd(b.queue + 0, 1);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= grp.n_pairs) break;
        const int tA = grp.pair_order[q];
        const int sg = b.task_sig[tA];
        const int N = (int)(b.sig_off[sg + 1] - b.sig_off[sg]);
        float Sv[2 * R], best[2], diag0[2];
        int bestj[2], lastlane[2], kL[2], nl = 0;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int f = b.task_flank[tA + t];
            const int L = (b.flank_off[f + 1] - b.flank_off[f]) * b.samples;
            const float *col0 = b.col0 + (size_t)f * b.col0_stride;
            nl = max(nl, (L + R - 1) / R);
            lastlane[t] = (L - 1) / R;
            kL[t] = ((L - 1) % R) / S;
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int i = lane * R + r + 1;
                Sv[2 * r + t] = i <= L ? col0[i] : 0.f;
            }
            diag0[t] = lane * R <= L ? col0[lane * R] : 0.f;
            best[t] = INF;
            bestj[t] = -1;
            if (lane == lastlane[t] && col0[L] > INF) { best[t] = col0[L]; bestj[t] = 0; }
        }
        const float *lutA = b.lut + (size_t)tA * b.lut_task_stride, *lutB = lutA + b.lut_task_stride;
        float *ck = b.ckpt + b.ckpt_off[tA];
        if (kL[0] == K - 1 && kL[1] == K - 1)
            LinSweepPair<K, S>::template run<true>(b.codes + b.sig_off[sg], N, lutA, lutB, lane, nl, b.p, Sv,
                                                   f2_pack(diag0[0], diag0[1]), kL[0], kL[1], best, bestj, ck, b.ckpt_rows);
        else
            LinSweepPair<K, S>::template run<false>(b.codes + b.sig_off[sg], N, lutA, lutB, lane, nl, b.p, Sv,
                                                    f2_pack(diag0[0], diag0[1]), kL[0], kL[1], best, bestj, ck, b.ckpt_rows);
#pragma unroll
        for (int t = 0; t < 2; ++t)
            if (lane == lastlane[t]) {
                b.res[tA + t].score = best[t];
                b.res[tA + t].best_j = bestj[t];
            }
    }
}

// ---------------------------------------------------------------------------------------------
// Pass 2: blockwise trace recomputation + SeqAn traceback (dp_traceback_impl.h:377-481,496-547).
// ---------------------------------------------------------------------------------------------
template <int K, int S>
__device__ __forceinline__ unsigned trace_at(const uint32_t *trace, int j0, int j, int i) {
    constexpr int R = K * S;
    constexpr int WP = (R + 31) / 32, W = 4 * WP;
    if (i <= 0 || j <= 0) return 0u;   // DP row 0 carries no trace; column 0 is never followed
    const int li = (i - 1) / R, r = (i - 1) % R;
    const uint32_t *w = trace + ((size_t)li * ALIGN_CKPT + (j - j0 - 1)) * W + r / 32;   // L1-cached (see the writer)
    const int bit = r % 32;
    const bool gap = (w[0] >> bit) & 1u, maxh = (w[WP] >> bit) & 1u, hopen = (w[2 * WP] >> bit) & 1u,
               vopen = (w[3 * WP] >> bit) & 1u;
    return (!gap ? T_DIAG : (maxh ? T_MAXH : T_MAXV)) | (hopen ? T_HOPEN : T_HOR) | (vopen ? T_VOPEN : T_VER);
}

__device__ __forceinline__ int nearest_signal_index(const int32_t *rows, int L, int N, int q) {
    // argmin_p |a_idx[p] - b_idx[q]| with first-minimum ties (numpy argmin, S.py:540-547),
    // expressed on the per-flank-sample records (see align.cuh / DESIGN.md).
    const int rec = rows[q];
    const int j = rec >> 1;
    if (!(rec & 1)) return j - 1;            // aligned to signal sample j-1: exact hit
    int qs = q, qe = q;                      // maximal run of vertical gaps around q
    while (qs > 0 && rows[qs - 1] == rec) --qs;
    while (qe < L - 1 && rows[qe + 1] == rec) ++qe;
    const int dprev = q - qs + 1, dnext = qe - q + 1;
    const bool has_prev = j >= 1, has_next = j < N;
    if (has_prev && (!has_next || dprev <= dnext)) return j - 1;
    return j;
}

struct TraceCursor {
    int tj, ti;          // cell the traceback stands on (meaningful in lane 0)
    unsigned tv;
    int state;
    bool first;
};
enum { ST_MAIN = 0, ST_VRUN = 1, ST_HRUN = 2 };

// One checkpoint block (j0, j1] of one task: recompute SeqAn's trace flags with KP levels per lane, then let lane 0
// follow the traceback through the block.  KP <= the K the score table was built for: the traceback only moves
// towards smaller rows, so a block entered at DP row `need` is swept with the fewest rows per lane that still
// cover rows 1..need (cells below the cursor cannot influence the cells above it).  Returns 1 when the
// traceback ended inside the block.
template <int KP, int S>
__device__ __forceinline__ int trace_block(const AlignBatch &b, const TaskGeom &g, const int lut_row_len, const int blk,
                                           const int j1, const int bj, const int need, uint32_t *trace, int32_t *rows,
                                           TraceCursor &c) {
    constexpr int R = KP * S;
    const int lane = threadIdx.x;
    const float INF = STRIQUE_SEQAN_INF;
    const int j0 = blk * ALIGN_CKPT;
    const int nl = (need + R - 1) / R;
    const int lastlane = (g.L - 1) / R, kL = ((g.L - 1) % R) / S;   // only used by the first block (need == L)
    float Sv[R], Hv[R];
    float diag0;
    if (blk == 0) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = lane * R + r + 1;
            Sv[r] = i <= g.L ? g.col0[i] : 0.f;
            Hv[r] = INF;
        }
        diag0 = lane * R <= g.L ? g.col0[lane * R] : 0.f;
    } else {
        const int step = b.ckpt_step[g.t];          // 2: interleaved with the other flank of the read (LinSweepPair)
        const float *ckS = b.ckpt + b.ckpt_off[g.t] + (size_t)(blk - 1) * 2 * step * b.ckpt_rows;
        const float *ckH = ckS + (size_t)step * b.ckpt_rows;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int i = lane * R + r + 1;
            Sv[r] = i <= g.L ? ckS[(size_t)i * step] : 0.f;
            Hv[r] = i <= g.L ? ckH[(size_t)i * step] : INF;
        }
        diag0 = lane == 0 ? 0.f : (lane * R <= g.L ? ckS[(size_t)lane * R * step] : 0.f);
    }
    float capS = 0.f, capH = 0.f, capV = 0.f, dummy_best = 0.f;
    int dummy_j = 0;
    Sweep<KP, S, true>::run(g.codes, g.N, g.lut, j0, j1, lane, nl, b.p, Sv, Hv, diag0, lastlane, kL, dummy_best, dummy_j,
                            nullptr, nullptr, 0, trace, c.first ? bj : -1, capS, capH, capV, lut_row_len);
    __syncwarp();
    if (c.first) {
        capS = __shfl_sync(0xffffffffu, capS, lastlane);
        capH = __shfl_sync(0xffffffffu, capH, lastlane);
        capV = __shfl_sync(0xffffffffu, capV, lastlane);
    }
    int done = 0;
    if (lane == 0) {
        int tj = c.tj, ti = c.ti, state = c.state;
        unsigned tv = c.tv;
        bool need_fetch = false;
        if (c.first) {
            tv = trace_at<KP, S>(trace, j0, tj, ti);
            // _correctTraceValue (dp_algorithm_impl.h:1168-1185)
            if (capV == capS) tv = (tv & ~T_DIAG) | T_MAXV;
            else if (capH == capS) tv = (tv & ~T_DIAG) | T_MAXH;
            // _retrieveInitialTraceDirection, PreferGapsAtEnd (dp_traceback_impl.h:456-481)
            if (tv & T_MAXV) tv &= (T_VER | T_VOPEN | T_MAXV);
            else if (tv & T_MAXH) tv &= (T_HOR | T_HOPEN | T_MAXH);
        } else {
            need_fetch = true;   // paused on a cell of this (earlier) block
        }
        for (;;) {
            if (need_fetch) {
                if (tj <= j0 && tj > 0 && ti > 0) break;   // cell lies in an earlier block
                tv = trace_at<KP, S>(trace, j0, tj, ti);
                need_fetch = false;
            }
            if (state == ST_MAIN) {
                if (!(tj > 0 && ti > 0 && tv != 0)) { done = 1; break; }
                if (tv & T_DIAG) {
                    rows[ti - 1] = tj << 1; --tj; --ti; need_fetch = true;
                } else if ((tv & T_MAXV) && (tv & T_VER)) {
                    state = ST_VRUN;
                } else if ((tv & T_MAXV) && (tv & T_VOPEN)) {
                    rows[ti - 1] = (tj << 1) | 1; --ti; need_fetch = true;
                } else if ((tv & T_MAXH) && (tv & T_HOR)) {
                    state = ST_HRUN;
                } else if ((tv & T_MAXH) && (tv & T_HOPEN)) {
                    --tj; need_fetch = true;
                } else { done = 1; break; }
            } else if (state == ST_VRUN) {
                const bool cont = (!(tv & T_VOPEN) || (tv & T_VER)) && ti != 1;
                rows[ti - 1] = (tj << 1) | 1; --ti; need_fetch = true;
                if (!cont) state = ST_MAIN;
            } else {
                const bool cont = (!(tv & T_HOPEN) || (tv & T_HOR)) && tj != 1;
                --tj; need_fetch = true;
                if (!cont) state = ST_MAIN;
            }
        }
        c.tj = tj; c.ti = ti; c.state = state; c.tv = tv;
    }
    c.first = false;
    return __shfl_sync(0xffffffffu, done, 0);
}

template <int K, int S>
__global__ void __launch_bounds__(32, align_trace_warps(K * S)) align_trace_kernel(AlignBatch b, AlignGroup grp) {
    constexpr int R = K * S;
    constexpr int W = 4 * ((R + 31) / 32);       // trace words per lane and column (Sweep::W of the widest variant)
    constexpr int K2 = (K + 1) / 2, K4 = (K + 3) / 4, K8 = (K + 7) / 8;   // fewer levels per lane for blocks entered at a small row
    const int lane = threadIdx.x;
    uint32_t *trace = b.trace + (size_t)blockIdx.x * ALIGN_CKPT * 32 * W;
    for (;;) {
        int q = 0;
        if (lane == 0) q = atomicAdd(b.queue + 1, 1);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= grp.n_tasks) break;
        const TaskGeom g = task_geom<K, S>(b, grp.order[q]);
        int32_t *rows = b.rows + (size_t)g.t * b.rows_stride;
        const int bj = b.res[g.t].best_j;
        int n_blocks = 0;
        // traceback cursor (meaningful in lane 0).  bj < 0: no last-row score ever exceeded SeqAn's
        // "infinity" -> traceback starts at (0,0) and every flank sample is a trailing vertical gap
        // behind all N signal samples; bj == 0: best cell in DP column 0 -> all leading gaps.
        TraceCursor c;
        c.tj = bj < 0 ? g.N : bj; c.ti = g.L; c.tv = 0; c.state = ST_MAIN; c.first = true;
        if (bj > 0) {
            int blk = (bj - 1) / ALIGN_CKPT;
            int j1 = bj;
            for (;;) {
                // rows the traceback can still visit in this block (lane 0 owns the cursor)
                const int need = max(1, __shfl_sync(0xffffffffu, c.ti, 0));
                int done;
                if (need <= 32 * K8 * S) done = trace_block<K8, S>(b, g, 32 * K, blk, j1, bj, need, trace, rows, c);
                else if (need <= 32 * K4 * S) done = trace_block<K4, S>(b, g, 32 * K, blk, j1, bj, need, trace, rows, c);
                else if (need <= 32 * K2 * S) done = trace_block<K2, S>(b, g, 32 * K, blk, j1, bj, need, trace, rows, c);
                else done = trace_block<K, S>(b, g, 32 * K, blk, j1, bj, need, trace, rows, c);
                ++n_blocks;
                if (done || blk == 0) break;
                --blk;
                j1 = blk * ALIGN_CKPT + ALIGN_CKPT;
                __syncwarp();
            }
        }
        const int tj = c.tj, ti = c.ti;
        if (lane == 0) {
            // leading gaps (head) or the degenerate all-gap cases: remaining flank rows are vertical gaps
            for (int i = ti; i >= 1; --i) rows[i - 1] = (tj << 1) | 1;
        }
        __syncwarp();
        if (lane == 0) {
            strique_align_result &r = b.res[g.t];
            const int pre = b.task_pre[g.t], post = b.task_post[g.t];
            r.begin0 = nearest_signal_index(rows, g.L, g.N, 0);
            r.end0 = nearest_signal_index(rows, g.L, g.N, g.L - 1);
            r.begin_trim = nearest_signal_index(rows, g.L, g.N, clampi(pre, 0, g.L - 1));
            r.end_trim = nearest_signal_index(rows, g.L, g.N, clampi(g.L - 1 - post, 0, g.L - 1));
            r.n_blocks = n_blocks;
            r.status = 0;
        }
        __syncwarp();
    }
}

template <int K, int S>
int launch_scan_t(strique_ctx *ctx, const AlignBatch &b, const AlignGroup &g) {
    const bool lin = align_params_linear(b.p);
    if constexpr (S == 6) {
        if (g.n_pairs > 0) {                         // (only set for linear gap costs, see align_run_device)
            int grid = ctx->num_sms * align_pair_warps(K);
            if (grid > g.n_pairs) grid = g.n_pairs;
            align_scan_pair_kernel<K, S><<<grid, 32, 0, ctx->stream>>>(b, g);
            ctx->launches++;
            CUDA_TRY(ctx, cudaGetLastError());
            if (g.n_single == 0) return STRIQUE_OK;
            CUDA_TRY(ctx, cudaMemsetAsync(b.queue, 0, 4, ctx->stream));
            AlignGroup rest = g;
            rest.n_pairs = 0;
            rest.n_tasks = g.n_single;
            rest.order = g.single_order;
            return launch_scan_t<K, S>(ctx, b, rest);
        }
    }
    if constexpr (S % 2 == 0) {
        if (lin && !getenv("STRIQUE_NO_PACKED_SCAN")) {
            int grid = ctx->num_sms * ALIGN_WARPS_PER_SM_PACKED;
            if (grid > g.n_tasks) grid = g.n_tasks;
            align_scan2_kernel<K, S><<<grid, 32, 0, ctx->stream>>>(b, g);
            ctx->launches++;
            CUDA_TRY(ctx, cudaGetLastError());
            return STRIQUE_OK;
        }
    }
    int grid = ctx->num_sms * (lin ? (getenv("STRIQUE_SCAN_WARPS") ? atoi(getenv("STRIQUE_SCAN_WARPS")) : ALIGN_WARPS_PER_SM_LINEAR) : ALIGN_WARPS_PER_SM);
    if (grid > g.n_tasks) grid = g.n_tasks;
    if (lin)
        align_scan_kernel<K, S, true><<<grid, 32, 0, ctx->stream>>>(b, g);
    else
        align_scan_kernel<K, S, false><<<grid, 32, 0, ctx->stream>>>(b, g);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return STRIQUE_OK;
}

template <int K, int S>
int launch_trace_t(strique_ctx *ctx, const AlignBatch &b, const AlignGroup &g, int n_warps) {
    int grid = std::min(n_warps, ctx->num_sms * align_trace_warps(K * S));     // (n_warps sized the trace buffer)
    if (grid > g.n_tasks) grid = g.n_tasks;
    align_trace_kernel<K, S><<<grid, 32, 0, ctx->stream>>>(b, g);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return STRIQUE_OK;
}

}  // namespace

// (K levels per lane, S samples per level) instantiations.  S = 6 is the reference's flank
// sampling (scripts/STRique.py:507-513 'samples'); S = 1 serves arbitrary flank vectors.
#define STRIQUE_ALIGN_INSTANCES(X) \
    X(2, 6) X(3, 6) X(4, 6) X(5, 6) X(6, 6) X(7, 6) X(8, 6) X(9, 6) X(10, 6) X(4, 1) X(8, 1) X(16, 1) X(32, 1) X(64, 1)

// linear gap costs: the scan may use LinSweep (see there for the exactness argument)
bool align_params_linear(const strique_align_params &p) {
    if (getenv("STRIQUE_NO_LINEAR_SCAN")) return false;
    return p.gap_open_h == p.gap_extension_h && p.gap_open_v == p.gap_extension_v &&
           fabsf(p.gap_extension_h) >= 1e-20f && fabsf(p.gap_extension_v) >= 1e-20f &&
           fabsf(p.gap_extension_h) < 1e30f && fabsf(p.gap_extension_v) < 1e30f;
}

bool align_pick_kernel(int nlev, int samples, int *K, int *S) {
    if (samples == 6) {
        // strictly more rows than the flank has: the packed scan (LinSweep2) also gives DP row 0 a place in lane 0
        for (int k = 2; k <= 10; ++k)
            if (k * 32 > nlev) { *K = k; *S = 6; return true; }
    }
    const int rows = nlev * samples;
    const int ks[5] = {4, 8, 16, 32, 64};
    for (int i = 0; i < 5; ++i)
        if (ks[i] * 32 >= rows) { *K = ks[i]; *S = 1; return true; }
    return false;
}

int align_launch_scan(strique_ctx *ctx, const AlignBatch &b, const AlignGroup &g) {
#define X(k, s) \
    if (g.K == k && g.S == s) return launch_scan_t<k, s>(ctx, b, g);
    STRIQUE_ALIGN_INSTANCES(X)
#undef X
    FAIL(ctx, STRIQUE_EUNSUPPORTED, "no alignment kernel for this flank shape");
}

int align_launch_trace(strique_ctx *ctx, const AlignBatch &b, const AlignGroup &g, int n_warps) {
#define X(k, s) \
    if (g.K == k && g.S == s) return launch_trace_t<k, s>(ctx, b, g, n_warps);
    STRIQUE_ALIGN_INSTANCES(X)
#undef X
    FAIL(ctx, STRIQUE_EUNSUPPORTED, "no alignment kernel for this flank shape");
}

int align_launch_build_lut(strique_ctx *ctx, const AlignBatch &b, const int32_t *task_K, const int32_t *task_S,
                           int n_tasks) {
    if (n_tasks == 0) return STRIQUE_OK;
    dim3 grid(n_tasks, 8);
    float far = INFINITY;
    const double span = (double)b.p.dist_offset - (double)b.p.dist_min;
    if (span > 0.0 && span < 1e30) far = (float)(pow(span * (1.0 + 1e-5), 1.0 / 1.2) * (1.0 + 1e-6));
    const int tile_codes = b.lut_row > 1024 ? 16 : 32;                    // lut_row = 32 * Kmax of the batch
    const size_t smem = (size_t)tile_codes * (b.lut_row + 1) * sizeof(float);
    if (smem > 48 * 1024)
        CUDA_TRY(ctx, cudaFuncSetAttribute(align_build_lut_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    align_build_lut_kernel<<<grid, 256, smem, ctx->stream>>>(b, task_K, task_S, n_tasks, far, tile_codes);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return STRIQUE_OK;
}

__global__ void align_patch_lut_kernel(float *lut, const unsigned long long *patch, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) lut[patch[2 * i]] = __uint_as_float((unsigned)patch[2 * i + 1]);
}

// patch = n pairs {element index into lut, fp32 bits}
int align_launch_patch_lut(strique_ctx *ctx, float *lut, const unsigned long long *patch, int n) {
    if (n == 0) return STRIQUE_OK;
    align_patch_lut_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(lut, patch, n);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return STRIQUE_OK;
}

}  // namespace strique
