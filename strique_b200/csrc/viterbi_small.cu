// Small-model Viterbi: the kernel of boundary #2 for HMMs with at most 32 emitting states and no silent chain --
// the reference's repeatModHMM (scripts/STRique.py:447-500: base / mCpG lanes of 6 match + 6 insert states between
// the emitting s0 / e0, 26 states), decoded once per read in methylation mode (S.py:605-609).
//
// Replaces pomegranate 0.10.0 `HiddenMarkovModel.viterbi` for that model: float64, value + edge weight then
// + emission, strict-'>' maxima in the generic kernel's candidate order (self loop first), so the decoded path is
// the same as viterbi_kernel's and the oracle's (equal up to exact ties).
//
// Mapping: ONE warp per sequence, ONE state per lane.  The state values never leave registers: an in-edge is a
// shuffle from the source lane plus a weight held in a register (<= SMALL_DEG in-edges per state), no shared
// memory, no barrier.  Samples are fetched 32 at a time (one coalesced load per 32 columns, broadcast by shuffle).
// Back-pointers: 4 bits per state and column, 8 columns per 32-bit word -> 16 B per time step.  Traceback by the
// same warp: lane i looks at column t - i, so a whole run of self loops (a sample dwells ~8 columns in a state) is
// skipped per iteration.
#include <math.h>

#include <algorithm>

#include "viterbi.cuh"

namespace strique {

namespace {

constexpr int SMALL_WARPS = 8;     // warps (sequences in flight) per CTA
constexpr unsigned FULLMASK = 0xffffffffu;

__device__ __forceinline__ double ninf() { return __longlong_as_double(0xfff0000000000000ll); }

template <int DEG>
__global__ void __launch_bounds__(SMALL_WARPS * 32) viterbi_small_kernel(VitBatch b, VitModelDev m) {
    extern __shared__ __align__(16) unsigned char smem[];
    for (int i = threadIdx.x * 16; i < m.blob_bytes; i += blockDim.x * 16)
        *reinterpret_cast<uint4 *>(smem + i) = *reinterpret_cast<const uint4 *>(m.blob + i);
    __syncthreads();
    const double *edge_w = reinterpret_cast<const double *>(smem + m.off_edge_w);
    const uint16_t *edge_src = reinterpret_cast<const uint16_t *>(smem + m.off_edge_src);
    const uint8_t *em_kind = smem + m.off_em_kind;
    const double *em_p = reinterpret_cast<const double *>(smem + m.off_em_p);   // [3][32]
    const uint8_t *em_flags = smem + m.off_flags;
    const uint16_t *end_src = reinterpret_cast<const uint16_t *>(smem + m.off_end_src);
    const double *end_w = reinterpret_cast<const double *>(smem + m.off_end_w);
    const int lane = threadIdx.x & 31;
    const int P_START = 32;
    const double NINF = ninf();

    // this lane's state: in-edges (register resident), emission
    double w_first[DEG], w_rest[DEG];          // column 1 sees START (value 0) and nothing else; later columns never do
    int src[DEG];
#pragma unroll
    for (int d = 0; d < DEG; ++d) {
        const bool have = d < m.deg[0];
        const int s = have ? edge_src[d * 32 + lane] : P_START + 1;
        const double w = have ? edge_w[d * 32 + lane] : 0.0;
        src[d] = s & 31;
        w_first[d] = s == P_START ? w : NINF;
        w_rest[d] = s < 32 ? w : NINF;
    }
    const int kind = em_kind[lane];
    const double p0 = em_p[lane], p1 = em_p[32 + lane], p2 = em_p[64 + lane];
    const bool has_self = m.deg[0] > 0 && edge_src[lane] == lane;   // the self loop, if any, is candidate 0

    for (;;) {
        int qi = 0;
        if (lane == 0) qi = atomicAdd(b.queue, 1);
        qi = __shfl_sync(FULLMASK, qi, 0);
        if (qi >= b.n_seq) break;
        const int seq = b.order[qi];
        const int64_t xo = b.x_off[seq];
        const int T = (int)(b.x_off[seq + 1] - xo);
        const double *x = b.x + xo;
        uint32_t *bp = reinterpret_cast<uint32_t *>(b.bp + b.bp_off[seq]);   // [ceil(T / 8)][32] words

        double v = NINF;
        uint32_t bits = 0u;
        double xblk = lane < T ? __ldg(x + lane) : 0.0;
        for (int t0 = 0; t0 < T; t0 += 32) {
            const double xnext = t0 + 32 + lane < T ? __ldg(x + t0 + 32 + lane) : 0.0;
            const int nk = min(32, T - t0);
#pragma unroll 8
            for (int k = 0; k < nk; ++k) {
                const double xt = __shfl_sync(FULLMASK, xblk, k);
                double best = NINF;
                uint32_t arg = 0u;
                if (t0 + k == 0) {
#pragma unroll
                    for (int d = 0; d < DEG; ++d) {
                        const double cand = w_first[d];                      // START value 0 + w
                        if (cand > best) { best = cand; arg = d; }
                    }
                } else {
#pragma unroll
                    for (int d = 0; d < DEG; ++d) {
                        const double cand = __shfl_sync(FULLMASK, v, src[d]) + w_rest[d];
                        if (cand > best) { best = cand; arg = d; }
                    }
                }
                const double dx = xt - p0;
                const double e_n = p1 - (dx * dx) * p2;                      // Normal: c0 - (x - mu)^2 / (2 sigma^2)
                const double e_u = (xt >= p0 && xt <= p1) ? p2 : NINF;       // Uniform: -log(hi - lo) inside [lo, hi]
                double e = kind == 0 ? e_n : e_u;
                if (xt != xt) e = 0.0;                                       // NaN sample: log 1 (pomegranate)
                v = best + e;
                bits |= arg << (4 * (k & 7));
                if ((k & 7) == 7 || k == nk - 1) {
                    bp[(size_t)((t0 + k) >> 3) * 32 + lane] = bits;
                    bits = 0u;
                }
            }
            xblk = xnext;
        }
        // END edges: log p = max(v[T][src] + w), first maximum
        double best = NINF;
        int barg = -1;
        for (int e0 = 0; e0 < m.n_end; e0 += 32) {
            const int e = e0 + lane;
            const int s = e < m.n_end ? end_src[e] : 0;
            const double sv = __shfl_sync(FULLMASK, v, s & 31);
            if (e < m.n_end && s < 32) {
                const double cand = sv + end_w[e];
                if (cand > best) { best = cand; barg = e; }
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ob = __shfl_down_sync(FULLMASK, best, off);
            const int oa = __shfl_down_sync(FULLMASK, barg, off);
            if (ob > best || (ob == best && oa >= 0 && (barg < 0 || oa < barg))) { best = ob; barg = oa; }
        }
        best = __shfl_sync(FULLMASK, best, 0);
        barg = __shfl_sync(FULLMASK, barg, 0);
        __syncwarp();
        __threadfence_block();

        // traceback: all lanes keep the same cursor (s, t); lane i inspects column t - i
        VitResult r;
        r.logp = best; r.n_count = 0; r.t_first = -1; r.t_last = -1; r.pattern_len = 0; r.status = 0; r.reserved = 0;
        if (T < 1 || !(best > NINF) || barg < 0) {
            r.status = 1;
        } else {
            int s = end_src[barg], t = T;
            uint8_t *pat = b.pattern ? b.pattern + xo : nullptr;
            uint16_t *path = b.path ? b.path + xo : nullptr;
            bool in_group = false;
            uint8_t last_mod = '0';
            int plen = 0;
            long long guard = (long long)T + 2;
            while (s != P_START) {
                if (--guard < 0 || s > P_START || t < 1) { r.status = 2; break; }
                const int ti = t - lane;
                const bool valid = ti >= 1;
                const uint32_t word = valid ? __ldcg(bp + (size_t)((ti - 1) >> 3) * 32 + s) : 0u;
                const uint32_t a = (word >> (4 * ((ti - 1) & 7))) & 15u;
                const bool self_ok = m.deg[0] > 0 && edge_src[s] == s;      // candidate 0 of state s is its self loop
                const bool self = valid && self_ok && a == 0u;
                const unsigned other = ~__ballot_sync(FULLMASK, self);
                const int k = other ? __ffs(other) - 1 : 32;               // columns t .. t-k+1 are self loops
                const bool step = k < 32 && ((__ballot_sync(FULLMASK, valid) >> k) & 1u);
                const int visits = k + (step ? 1 : 0);
                if (visits == 0) { r.status = 2; break; }
                const unsigned fl = em_flags[s];
                if (fl & HMM_FLAG_COUNT) r.n_count += visits;
                if (fl & HMM_FLAG_REPEAT) { if (r.t_last < 0) r.t_last = t - 1; r.t_first = t - visits; }
                if (fl & HMM_FLAG_SEP) {
                    if (in_group) { if (pat && lane == 0) pat[T - 1 - plen] = last_mod; ++plen; in_group = false; }
                } else {
                    in_group = true;
                    last_mod = (fl & HMM_FLAG_MOD) ? '1' : '0';
                }
                if (path && lane < visits) path[t - 1 - lane] = (uint16_t)m.perm[s];
                t -= k;
                if (step) {
                    const uint32_t ak = __shfl_sync(FULLMASK, a, k);
                    s = edge_src[ak * 32 + s];
                    --t;
                }
            }
            if (in_group) { if (pat && lane == 0) pat[T - 1 - plen] = last_mod; ++plen; }
            if (r.status == 0 && t != 0) r.status = 2;
            r.pattern_len = plen;
        }
        if (lane == 0) b.res[seq] = r;
        __syncwarp();
    }
    (void)has_self;
}

}  // namespace

// models the small kernel serves: one slot of emitting states, no chain, few in-edges per state
bool viterbi_small_fits(const VitModelDev &m) { return m.NS == 1 && m.QC == 0 && m.deg[0] >= 1 && m.deg[0] <= 8; }

int viterbi_small_launch(strique_ctx *ctx, const HmmModel &m, const VitBatch &b) {
    if (b.n_seq == 0) return STRIQUE_OK;
    const VitModelDev &dm = m.dev;
    const size_t smem = align_up(dm.blob_bytes, 16);
    auto launch = [&](auto kernel) -> int {
        CUDA_TRY(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
        int per_sm = 0;
        CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, SMALL_WARPS * 32, smem));
        int grid = ctx->num_sms * std::max(per_sm, 1);
        grid = std::min(grid, (b.n_seq + SMALL_WARPS - 1) / SMALL_WARPS);
        kernel<<<grid, SMALL_WARPS * 32, smem, ctx->stream>>>(b, dm);
        ctx->launches++;
        CUDA_TRY(ctx, cudaGetLastError());
        return STRIQUE_OK;
    };
    if (dm.deg[0] <= 4) return launch(viterbi_small_kernel<4>);
    if (dm.deg[0] <= 6) return launch(viterbi_small_kernel<6>);
    return launch(viterbi_small_kernel<8>);
}

}  // namespace strique
