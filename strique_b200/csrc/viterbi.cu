// Viterbi kernels (see viterbi.cuh for the model form and the mapping onto a warp).
#include <math.h>

#include <algorithm>

#include "viterbi.cuh"

namespace strique {

namespace {

#define NEG_INF (-CUDART_INF)

__device__ __forceinline__ double neg_inf() { return __longlong_as_double(0xfff0000000000000ll); }

// Silent chain states of one time step: v[chain c] = max(entry edges from v, v[chain c-1] + w).
// Each lane owns QC consecutive chain states; lanes are combined with a max-plus Kogge-Stone scan.
// Returns the lane's packed 4-bit back-pointers (0 = previous chain state, k = entry edge k-1).
template <int QCMAX>
__device__ __forceinline__ unsigned long long chain_phase(double *v, const int QC, const int chain_base, const int lane,
                                                          const double *__restrict__ predw,
                                                          const double *__restrict__ ew,
                                                          const uint16_t *__restrict__ es) {
    if (QC == 0) return 0ull;
    double a[QCMAX];
    int ka[QCMAX];
    double A = neg_inf(), W = 0.0;
#pragma unroll
    for (int q = 0; q < QCMAX; ++q) {
        if (q < QC) {
            double best = neg_inf();
            int k = 0;
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                const int idx = (q * 3 + e) * 32 + lane;
                const double cand = v[es[idx]] + ew[idx];
                if (cand > best) { best = cand; k = e + 1; }
            }
            a[q] = best;
            ka[q] = k;
            const double pw = predw[q * 32 + lane];
            const double t0 = A + pw;
            A = best >= t0 ? best : t0;
            W += pw;
        }
    }
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const double Al = __shfl_up_sync(0xffffffffu, A, off);
        const double Wl = __shfl_up_sync(0xffffffffu, W, off);
        if (lane >= off) {
            const double t0 = Al + W;
            A = A >= t0 ? A : t0;
            W = Wl + W;
        }
    }
    double D = __shfl_up_sync(0xffffffffu, A, 1);
    if (lane == 0) D = neg_inf();
    unsigned long long nib = 0ull;
#pragma unroll
    for (int q = 0; q < QCMAX; ++q) {
        if (q < QC) {
            const double t0 = D + predw[q * 32 + lane];
            int arg;
            if (a[q] >= t0) { D = a[q]; arg = ka[q]; } else { D = t0; arg = 0; }
            v[chain_base + lane * QC + q] = D;
            nib |= (unsigned long long)arg << (4 * q);
        }
    }
    return nib;
}

__global__ void __launch_bounds__(VIT_WARPS * 32) viterbi_kernel(VitBatch b, VitModelDev m) {
    extern __shared__ __align__(16) unsigned char smem[];
    for (int i = threadIdx.x * 16; i < m.blob_bytes; i += blockDim.x * 16)
        *reinterpret_cast<uint4 *>(smem + i) = *reinterpret_cast<const uint4 *>(m.blob + i);
    __syncthreads();
    const double *edge_w = reinterpret_cast<const double *>(smem + m.off_edge_w);
    const uint16_t *edge_src = reinterpret_cast<const uint16_t *>(smem + m.off_edge_src);
    const uint8_t *em_kind = smem + m.off_em_kind;
    const double *em_p = reinterpret_cast<const double *>(smem + m.off_em_p);   // [3][NS][32]
    const uint8_t *em_flags = smem + m.off_flags;
    const double *ch_predw = reinterpret_cast<const double *>(smem + m.off_chain_predw);
    const double *ch_ew = reinterpret_cast<const double *>(smem + m.off_chain_ew);
    const uint16_t *ch_es = reinterpret_cast<const uint16_t *>(smem + m.off_chain_es);
    const uint16_t *end_src = reinterpret_cast<const uint16_t *>(smem + m.off_end_src);
    const double *end_w = reinterpret_cast<const double *>(smem + m.off_end_w);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int NS = m.NS, QC = m.QC;
    const int chain_base = NS * 32;
    const int P_START = (NS + QC) * 32, NV = P_START + 2;
    double *vbuf = reinterpret_cast<double *>(smem + ((m.blob_bytes + 15) / 16) * 16) + (size_t)warp * 2 * NV;
    const double NINF = neg_inf();

    for (;;) {
        int qi = 0;
        if (lane == 0) qi = atomicAdd(b.queue, 1);
        qi = __shfl_sync(0xffffffffu, qi, 0);
        if (qi >= b.n_seq) break;
        const int seq = b.order[qi];
        const int64_t xo = b.x_off[seq];
        const int T = (int)(b.x_off[seq + 1] - xo);
        const double *x = b.x + xo;
        unsigned long long *bp = b.bp + b.bp_off[seq];
        double *vp = vbuf, *vc = vbuf + NV;
        for (int i = lane; i < NV; i += 32) { vp[i] = NINF; vc[i] = NINF; }
        __syncwarp();
        if (lane == 0) vp[P_START] = 0.0;
        __syncwarp();
        {
            const unsigned long long nib = chain_phase<4>(vp, QC, chain_base, lane, ch_predw, ch_ew, ch_es);
            bp[lane] = nib << (4 * NS);
        }
        __syncwarp();
        for (int t = 1; t <= T; ++t) {
            const double xt = x[t - 1];
            unsigned long long word = 0ull;
            for (int s = 0; s < NS; ++s) {
                double best = NINF;
                int arg = 0;
                const int base = m.row_base[s], dg = m.deg[s];
                for (int d = 0; d < dg; ++d) {
                    const int idx = (base + d) * 32 + lane;
                    const double cand = vp[edge_src[idx]] + edge_w[idx];
                    if (cand > best) { best = cand; arg = d; }
                }
                const int li = s * 32 + lane;
                const double p0 = em_p[li], p1 = em_p[NS * 32 + li], p2 = em_p[2 * NS * 32 + li];
                double e;
                if (em_kind[li] == 0) {
                    const double dx = xt - p0;          // Normal: c0 - (x - mu)^2 * 1/(2 sigma^2)
                    e = p1 - (dx * dx) * p2;
                } else {
                    e = (xt >= p0 && xt <= p1) ? p2 : NINF;   // Uniform: -log(hi - lo) inside [lo, hi]
                }
                if (xt != xt) e = 0.0;
                vc[li] = best + e;
                word |= (unsigned long long)arg << (4 * s);
            }
            __syncwarp();
            if (t == 1 && lane == 0) vp[P_START] = NINF;   // START exists before the first sample only
            word |= chain_phase<4>(vc, QC, chain_base, lane, ch_predw, ch_ew, ch_es) << (4 * NS);
            bp[(size_t)t * 32 + lane] = word;
            __syncwarp();
            double *tmp = vp; vp = vc; vc = tmp;
        }
        // END edges: log p = max(v[T][src] + w)
        double best = NINF;
        int barg = -1;
        for (int e = lane; e < m.n_end; e += 32) {
            const double cand = vp[end_src[e]] + end_w[e];
            if (cand > best) { best = cand; barg = e; }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const double ob = __shfl_down_sync(0xffffffffu, best, off);
            const int oa = __shfl_down_sync(0xffffffffu, barg, off);
            if (ob > best || (ob == best && oa >= 0 && (barg < 0 || oa < barg))) { best = ob; barg = oa; }
        }
        __syncwarp();
        if (lane == 0) {
            VitResult r;
            r.logp = best; r.n_count = 0; r.t_first = -1; r.t_last = -1; r.pattern_len = 0; r.status = 0; r.reserved = 0;
            if (!(best > NINF) || barg < 0) {
                r.status = 1;
            } else {
                int s = end_src[barg], t = T;
                uint8_t *pat = b.pattern ? b.pattern + xo : nullptr;
                uint16_t *path = b.path ? b.path + xo : nullptr;
                bool in_group = false;
                uint8_t last_mod = '0';
                int plen = 0;
                long long guard = (long long)(T + 2) * (m.C + 2);
                while (s != P_START) {
                    if (--guard < 0 || s > P_START) { r.status = 2; break; }
                    if (s < chain_base) {
                        if (t < 1) { r.status = 2; break; }
                        const unsigned fl = em_flags[s];
                        if (fl & HMM_FLAG_COUNT) ++r.n_count;
                        if (fl & HMM_FLAG_REPEAT) { if (r.t_last < 0) r.t_last = t - 1; r.t_first = t - 1; }
                        if (fl & HMM_FLAG_SEP) {
                            if (in_group) { if (pat) pat[T - 1 - plen] = last_mod; ++plen; in_group = false; }
                        } else {
                            in_group = true;
                            last_mod = (fl & HMM_FLAG_MOD) ? '1' : '0';
                        }
                        if (path) path[t - 1] = (uint16_t)m.perm[s];
                        const int ls = s & 31, slot = s >> 5;
                        const unsigned long long w = __ldcg(bp + (size_t)t * 32 + ls);
                        const int arg = (int)((w >> (4 * slot)) & 15ull);
                        s = edge_src[(m.row_base[slot] + arg) * 32 + ls];
                        --t;
                    } else {
                        const int c = s - chain_base, lc = c / QC, q = c - lc * QC;
                        const unsigned long long w = __ldcg(bp + (size_t)t * 32 + lc);
                        const int arg = (int)((w >> (4 * (NS + q))) & 15ull);
                        if (arg == 0) s = s - 1; else s = ch_es[(q * 3 + arg - 1) * 32 + lc];
                    }
                }
                if (in_group) { if (pat) pat[T - 1 - plen] = last_mod; ++plen; }
                if (r.status == 0 && t != 0) r.status = 2;
                r.pattern_len = plen;
            }
            b.res[seq] = r;
        }
        __syncwarp();
    }
}

template <typename T>
__global__ void prepare_x_kernel(const T *__restrict__ src, const PrepSeg *__restrict__ segs, int n_segs,
                                 const double *__restrict__ stats, int stat_c1, double c3, double c4, double lo, double hi,
                                 double lo2, double hi2, double *__restrict__ x) {
    for (int sgi = blockIdx.x; sgi < n_segs; sgi += gridDim.x) {
        const PrepSeg sg = segs[sgi];
        const double c1 = stats[(size_t)sg.read * 12 + stat_c1], c2 = stats[(size_t)sg.read * 12 + stat_c1 + 1];
        for (int i = threadIdx.x; i < sg.len; i += blockDim.x) {
            double y = (((double)src[sg.src_off + i] - c1) / c2) * c3 + c4;   // numpy operation order, S.py:157-158
            y = y < lo ? lo : (y > hi ? hi : y);                              // S.py:178-179
            y = y < lo2 ? lo2 : (y > hi2 ? hi2 : y);                          // S.py:493 (methylation HMM only)
            x[sg.dst_off + i] = y;
        }
    }
}

}  // namespace

int viterbi_prepare_x(strique_ctx *ctx, int raw_kind, const void *src, const PrepSeg *segs_dev, int n_segs,
                      const double *stats, int stat_c1, double c3, double c4, double lo, double hi, double lo2,
                      double hi2, double *x_out) {
    if (n_segs == 0) return STRIQUE_OK;
    const int grid = std::min(n_segs, ctx->num_sms * 8);
    if (raw_kind == 0)
        prepare_x_kernel<int16_t><<<grid, 256, 0, ctx->stream>>>((const int16_t *)src, segs_dev, n_segs, stats, stat_c1,
                                                                  c3, c4, lo, hi, lo2, hi2, x_out);
    else
        prepare_x_kernel<double><<<grid, 256, 0, ctx->stream>>>((const double *)src, segs_dev, n_segs, stats, stat_c1, c3,
                                                                 c4, lo, hi, lo2, hi2, x_out);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return STRIQUE_OK;
}

int viterbi_launch(strique_ctx *ctx, const HmmModel &m, const VitBatch &b) {
    if (b.n_seq == 0) return STRIQUE_OK;
    const VitModelDev &dm = m.dev;
    const int NV = (dm.NS + dm.QC) * 32 + 2;
    const size_t smem = align_up(dm.blob_bytes, 16) + (size_t)VIT_WARPS * 2 * NV * sizeof(double);
    if (smem > 200 * 1024) FAIL(ctx, STRIQUE_EUNSUPPORTED, "HMM too large for the shared-memory Viterbi kernel");
    CUDA_TRY(ctx, cudaFuncSetAttribute(viterbi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int per_sm = (int)std::min<size_t>(4, (220 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    int grid = ctx->num_sms * per_sm;
    const int need = (b.n_seq + VIT_WARPS - 1) / VIT_WARPS;
    if (grid > need) grid = need;
    viterbi_kernel<<<grid, VIT_WARPS * 32, smem, ctx->stream>>>(b, dm);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return STRIQUE_OK;
}

}  // namespace strique
