// Read-conditioning kernel (see condition.cuh).  One CTA per read; phases separated by
// __syncthreads(); intermediates live in global scratch (L2 resident: a read is 40 KB - 2 MB).
#include "condition.cuh"

namespace strique {
namespace {

constexpr int CT = 512;     // threads per CTA
// resident CTAs per SM: the kernel waits on L2 (12 dependent passes over a read), so threads in flight matter more than
// registers per thread -- 2 CTAs (63 registers) 10.1 ms per 8192 reads, 3 (40 registers) 7.2 ms, 4 (32, spills) 8.1 ms
#ifndef COND_MIN_CTAS
#define COND_MIN_CTAS 3
#endif
constexpr int MAXR = 6;     // ranks resolved per radix-select round

template <typename T> struct KeyTraits;
template <> struct KeyTraits<int16_t> {
    typedef uint32_t Key;
    static constexpr int BYTES = 2;
    __device__ static Key to_key(int16_t x) { return (uint32_t)((uint16_t)x ^ 0x8000u); }
    __device__ static double from_key(Key k) { return (double)(int16_t)(uint16_t)(k ^ 0x8000u); }
};
template <> struct KeyTraits<double> {
    typedef unsigned long long Key;
    static constexpr int BYTES = 8;
    __device__ static Key to_key(double x) {
        const unsigned long long u = (unsigned long long)__double_as_longlong(x);
        return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
    }
    __device__ static double from_key(Key k) {
        const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
        return __longlong_as_double((long long)u);
    }
};

struct Smem {
    unsigned hist[MAXR][256];
    unsigned long long prefix[MAXR];
    int krem[MAXR];
    int ranks[MAXR];
    double sel[MAXR];
    double red[CT / 32];
    unsigned chist[256];
    double bc[8];
    int ibc[8];
    // ranks that share a key prefix share one histogram (pass 0: all of them; later passes: usually the two
    // neighbours of a median / percentile)
    unsigned long long gpre[MAXR];
    int grp[MAXR];
    int ngrp;
};

__device__ __forceinline__ double block_sum(double v, Smem &sm) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sm.red[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        double w = threadIdx.x < CT / 32 ? sm.red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) w += __shfl_down_sync(0xffffffffu, w, o);
        if (threadIdx.x == 0) sm.red[0] = w;
    }
    __syncthreads();
    const double r = sm.red[0];
    __syncthreads();
    return r;
}

// Exact order statistics: values at sorted positions sm.ranks[0..nr) -> sm.sel[0..nr).
// MSB-first radix select, 8 bits per pass, all ranks resolved in the same passes.
template <typename T>
__device__ void multi_select(const T *__restrict__ data, const int n, const int nr, Smem &sm) {
    typedef KeyTraits<T> KT;
    typedef typename KT::Key Key;
    const int tid = threadIdx.x;
    if (tid < MAXR) { sm.prefix[tid] = 0; sm.krem[tid] = tid < nr ? sm.ranks[tid] : 0; }
    __syncthreads();
    for (int pass = 0; pass < KT::BYTES; ++pass) {
        const int shift = (KT::BYTES - 1 - pass) * 8;
        for (int i = tid; i < MAXR * 256; i += CT) (&sm.hist[0][0])[i] = 0;
        if (tid == 0) {
            int ng = 0;
            for (int r = 0; r < nr; ++r) {
                int g = 0;
                while (g < ng && sm.gpre[g] != sm.prefix[r]) ++g;
                if (g == ng) sm.gpre[ng++] = sm.prefix[r];
                sm.grp[r] = g;
            }
            sm.ngrp = ng;
        }
        __syncthreads();
        const int ng = sm.ngrp;
        Key pre[MAXR];
        unsigned lastbin[MAXR], cnt[MAXR];
#pragma unroll
        for (int g = 0; g < MAXR; ++g) { pre[g] = (Key)sm.gpre[g < ng ? g : 0]; lastbin[g] = 0; cnt[g] = 0; }
        for (int i = tid; i < n; i += CT) {
            const Key k = KT::to_key(data[i]);
            const unsigned bin = (unsigned)(k >> shift) & 255u;
            const Key hi = pass == 0 ? (Key)0 : (Key)(k >> (shift + 8 < (int)sizeof(Key) * 8 ? shift + 8 : 0));
#pragma unroll
            for (int g = 0; g < MAXR; ++g) {
                if (g < ng && hi == pre[g]) {
                    if (cnt[g] && lastbin[g] == bin) {
                        ++cnt[g];
                    } else {
                        if (cnt[g]) atomicAdd(&sm.hist[g][lastbin[g]], cnt[g]);
                        lastbin[g] = bin;
                        cnt[g] = 1;
                    }
                }
            }
        }
#pragma unroll
        for (int g = 0; g < MAXR; ++g)
            if (cnt[g]) atomicAdd(&sm.hist[g][lastbin[g]], cnt[g]);
        __syncthreads();
        if (tid < nr) {
            const unsigned *h = sm.hist[sm.grp[tid]];
            int k = sm.krem[tid];
            unsigned d = 0;
            for (; d < 255; ++d) {
                const int c = (int)h[d];
                if (k < c) break;
                k -= c;
            }
            sm.krem[tid] = k;
            sm.prefix[tid] = (sm.prefix[tid] << 8) | d;
        }
        __syncthreads();
    }
    if (tid < nr) sm.sel[tid] = KT::from_key((Key)sm.prefix[tid]);
    __syncthreads();
}

// numpy's linear-interpolation percentile pieces (numpy/lib/_function_base_impl.py: _quantile,
// _get_indexes, _get_gamma, _lerp with method='linear'): virtual index (n-1)*q.
__device__ __forceinline__ void percentile_index(int n, double q, int *lo, int *hi, double *gamma) {
    const double vi = (double)(n - 1) * q;
    if (vi >= (double)(n - 1)) { *lo = n - 1; *hi = n - 1; *gamma = 0.0; return; }
    const double f = floor(vi);
    *lo = (int)f;
    *hi = (int)f + 1;
    *gamma = vi - f;
}
__device__ __forceinline__ double np_lerp(double a, double b, double t) {
    const double d = b - a;
    return t >= 0.5 ? b - d * (1.0 - t) : a + d * t;
}

__device__ __forceinline__ int reflect(int k, int n) {   // scipy.ndimage mode='reflect': d c b a | a b c d | d c b a
    while (k < 0 || k >= n) k = k < 0 ? -k - 1 : 2 * n - k - 1;
    return k;
}

template <bool IS_MAX>
__device__ void window_pass(const uint8_t *__restrict__ in, uint8_t *__restrict__ out8, uint16_t *__restrict__ out16,
                            int n, int lo, int hi, unsigned *chist) {
    // flat 8-wide window [i + lo, i + hi]; a thread produces 4 consecutive outputs from 11 inputs (the 5 inputs all
    // four windows share are reduced once)
    auto op = [](int a, int b) { return IS_MAX ? max(a, b) : min(a, b); };
    for (int i = threadIdx.x * 4; i < n; i += CT * 4) {
        int v[4];
        if (i + lo >= 0 && i + 3 + hi < n) {
            int x[11];
#pragma unroll
            for (int k = 0; k < 11; ++k) x[k] = in[i + lo + k];
            const int c = op(op(op(x[3], x[4]), op(x[5], x[6])), x[7]);
            v[0] = op(op(x[0], x[1]), op(x[2], c));
            v[1] = op(op(x[1], x[2]), op(c, x[8]));
            v[2] = op(op(x[2], c), op(x[8], x[9]));
            v[3] = op(op(c, x[8]), op(x[9], x[10]));
        } else {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int w = IS_MAX ? 0 : 255;
                if (i + j < n)
                    for (int k = lo; k <= hi; ++k) w = op(w, (int)in[reflect(i + j + k, n)]);
                v[j] = w;
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (i + j < n) {
                if (out8) out8[i + j] = (uint8_t)v[j];
                if (out16) out16[i + j] = (uint16_t)v[j];
            }
        }
        if (out16) {                    // code histogram: neighbours are mostly equal after the morphology
            int run = 0, val = -1;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (i + j >= n) break;
                if (v[j] == val) { ++run; continue; }
                if (run) atomicAdd(&chist[val], (unsigned)run);
                val = v[j];
                run = 1;
            }
            if (run) atomicAdd(&chist[val], (unsigned)run);
        }
    }
    __syncthreads();
}

// minmax statistics (tail medians) of a signal once its 1st/99th percentile values are in
// sm.bc[0], sm.bc[1]; result c1 = m5 + (m95-m5)/2, c2 = (m95-m5)/2 in sm.bc[2], sm.bc[3];
// sm.ibc[0] = 1 if a tail is empty (numpy would produce NaN).
template <typename T>
__device__ void tail_medians(const T *__restrict__ data, int n, Smem &sm) {
    const double q_lo = sm.bc[0], q_hi = sm.bc[1];
    double below = 0.0, above = 0.0;
    for (int i = threadIdx.x; i < n; i += CT) {
        const double x = (double)data[i];
        below += x < q_lo ? 1.0 : 0.0;
        above += x > q_hi ? 1.0 : 0.0;
    }
    const int m_lo = (int)block_sum(below, sm);
    const int m_hi = (int)block_sum(above, sm);
    if (m_lo == 0 || m_hi == 0) {
        if (threadIdx.x == 0) { sm.ibc[0] = 1; sm.bc[2] = 0.0; sm.bc[3] = 1.0; }
        __syncthreads();
        return;
    }
    if (threadIdx.x == 0) {
        sm.ranks[0] = (m_lo - 1) / 2; sm.ranks[1] = m_lo / 2;
        sm.ranks[2] = n - m_hi + (m_hi - 1) / 2; sm.ranks[3] = n - m_hi + m_hi / 2;
    }
    __syncthreads();
    multi_select<T>(data, n, 4, sm);
    if (threadIdx.x == 0) {
        const double m5 = (sm.sel[0] + sm.sel[1]) / 2, m95 = (sm.sel[2] + sm.sel[3]) / 2;
        sm.bc[2] = m5 + (m95 - m5) / 2;
        sm.bc[3] = (m95 - m5) / 2;
        sm.ibc[0] = 0;
    }
    __syncthreads();
}

template <typename T>
__device__ void minmax_stats(const T *__restrict__ data, int n, Smem &sm) {
    if (threadIdx.x == 0) {
        double g;
        percentile_index(n, 0.01, &sm.ranks[0], &sm.ranks[1], &g); sm.bc[4] = g;
        percentile_index(n, 0.99, &sm.ranks[2], &sm.ranks[3], &g); sm.bc[5] = g;
    }
    __syncthreads();
    multi_select<T>(data, n, 4, sm);
    if (threadIdx.x == 0) {
        sm.bc[0] = np_lerp(sm.sel[0], sm.sel[1], sm.bc[4]);
        sm.bc[1] = np_lerp(sm.sel[2], sm.sel[3], sm.bc[5]);
    }
    __syncthreads();
    tail_medians<T>(data, n, sm);
}

__device__ __forceinline__ int hist_at(const unsigned *h, int rank) {   // value at sorted position `rank`
    int c = 0;
    for (int v = 0; v < 256; ++v) { c += (int)h[v]; if (rank < c) return v; }
    return 255;
}

template <typename T>
__global__ void __launch_bounds__(CT, COND_MIN_CTAS) condition_kernel(const T *__restrict__ raw_all, const int64_t *__restrict__ off,
                                                       int n_reads, CondModel model, int want_raw, T *__restrict__ flt_all,
                                                       uint16_t *__restrict__ codes_all, uint8_t *__restrict__ tmpA_all,
                                                       uint8_t *__restrict__ tmpB_all, float *__restrict__ vals_all,
                                                       double *__restrict__ stats_all) {
    __shared__ Smem sm;
    const int tid = threadIdx.x;
    double c3, c4, vlo, vhi;
    minmax_model_constants(model, &c3, &c4, &vlo, &vhi);
    for (int read = blockIdx.x; read < n_reads; read += gridDim.x) {
        const int64_t o = off[read];
        const int n = (int)(off[read + 1] - o);
        const T *raw = raw_all + o;
        T *flt = flt_all + o;
        uint16_t *codes = codes_all + o;
        uint8_t *tmpA = tmpA_all + o, *tmpB = tmpB_all + o;
        double *stats = stats_all + (size_t)read * CS_STRIDE;
        int status = 0;
        // 1. scipy.signal.medfilt(raw, 3): zero padded, dtype preserved (S.py:590)
        for (int i = tid; i < n; i += CT) {
            const T a = i > 0 ? raw[i - 1] : (T)0, b = raw[i], c = i + 1 < n ? raw[i + 1] : (T)0;
            const T lo = a < b ? a : b, hi = a < b ? b : a;
            const T m = hi < c ? hi : c;
            flt[i] = lo < m ? m : lo;
        }
        if (tid < 256) sm.chist[tid] = 0;
        __syncthreads();
        // 2. median and 1st / 99th percentile of the filtered signal in one select round
        if (tid == 0) {
            double g;
            sm.ranks[0] = (n - 1) / 2; sm.ranks[1] = n / 2;
            percentile_index(n, 0.01, &sm.ranks[2], &sm.ranks[3], &g); sm.bc[4] = g;
            percentile_index(n, 0.99, &sm.ranks[4], &sm.ranks[5], &g); sm.bc[5] = g;
        }
        __syncthreads();
        multi_select<T>(flt, n, 6, sm);
        const double med = (sm.sel[0] + sm.sel[1]) / 2;
        const double fq_lo = np_lerp(sm.sel[2], sm.sel[3], sm.bc[4]), fq_hi = np_lerp(sm.sel[4], sm.sel[5], sm.bc[5]);
        __syncthreads();
        // 3. MAD = mean |x - median| (S.py:142-143)
        double acc = 0.0;
        for (int i = tid; i < n; i += CT) acc += fabs((double)flt[i] - med);
        const double mad = block_sum(acc, sm) / (double)n;
        if (!(mad > 0.0)) status = 1;
        // 4. z*24+127 clipped and truncated to uint8 (S.py:591-592)
        for (int i = tid; i < n; i += CT) {
            double y = (((double)flt[i] - med) / mad) * 24 + 127;
            y = y < 0.0 ? 0.0 : (y > 255.0 ? 255.0 : y);
            tmpA[i] = status ? (uint8_t)0 : (uint8_t)y;
        }
        __syncthreads();
        // 5. closing(opening(u8, rectangle(1,8))) with scikit-image<0.15 windows (S.py:593-595)
        window_pass<false>(tmpA, tmpB, nullptr, n, -3, 4, nullptr);
        window_pass<true>(tmpB, tmpA, nullptr, n, -4, 3, nullptr);
        window_pass<true>(tmpA, tmpB, nullptr, n, -3, 4, nullptr);
        window_pass<false>(tmpB, nullptr, codes, n, -4, 3, sm.chist);
        // 6. minmax normalisation of the uint8 signal -> 256-entry value table (S.py:596)
        if (tid == 0) {
            int lo, hi; double g;
            percentile_index(n, 0.01, &lo, &hi, &g);
            const double q_lo = np_lerp((double)hist_at(sm.chist, lo), (double)hist_at(sm.chist, hi), g);
            percentile_index(n, 0.99, &lo, &hi, &g);
            const double q_hi = np_lerp((double)hist_at(sm.chist, lo), (double)hist_at(sm.chist, hi), g);
            int m_lo = 0, m_hi = 0;
            for (int v = 0; v < 256; ++v) {
                if ((double)v < q_lo) m_lo += (int)sm.chist[v];
                if ((double)v > q_hi) m_hi += (int)sm.chist[v];
            }
            if (m_lo == 0 || m_hi == 0) {
                sm.ibc[1] = 1; sm.bc[6] = 0.0; sm.bc[7] = 1.0;
            } else {
                const double m5 = ((double)hist_at(sm.chist, (m_lo - 1) / 2) + (double)hist_at(sm.chist, m_lo / 2)) / 2;
                const double m95 = ((double)hist_at(sm.chist, n - m_hi + (m_hi - 1) / 2) +
                                    (double)hist_at(sm.chist, n - m_hi + m_hi / 2)) / 2;
                sm.ibc[1] = 0; sm.bc[6] = m5 + (m95 - m5) / 2; sm.bc[7] = (m95 - m5) / 2;
            }
        }
        __syncthreads();
        if (sm.ibc[1]) status = 1;
        const double u8c1 = sm.bc[6], u8c2 = sm.bc[7];
        if (tid < 256) {
            double y = (((double)tid - u8c1) / u8c2) * c3 + c4;
            y = y < vlo ? vlo : (y > vhi ? vhi : y);
            vals_all[(size_t)read * 256 + tid] = (float)y;
        }
        __syncthreads();
        // 7. minmax statistics of the filtered signal (count HMM input, S.py:597)
        if (tid == 0) { sm.bc[0] = fq_lo; sm.bc[1] = fq_hi; }
        __syncthreads();
        tail_medians<T>(flt, n, sm);
        const double fc1 = sm.bc[2], fc2 = sm.bc[3];
        if (sm.ibc[0]) status = 1;
        __syncthreads();
        // 8. minmax statistics of the raw signal (methylation HMM input, S.py:607)
        double rc1 = 0.0, rc2 = 1.0;
        if (want_raw) {
            minmax_stats<T>(raw, n, sm);
            rc1 = sm.bc[2]; rc2 = sm.bc[3];
            if (sm.ibc[0]) status = 1;
            __syncthreads();
        }
        if (tid == 0) {
            stats[CS_FLT_MEDIAN] = med; stats[CS_FLT_MAD] = mad;
            stats[CS_FLT_C1] = fc1; stats[CS_FLT_C2] = fc2;
            stats[CS_RAW_C1] = rc1; stats[CS_RAW_C2] = rc2;
            stats[CS_U8_C1] = u8c1; stats[CS_U8_C2] = u8c2;
            stats[CS_STATUS] = (double)status;
        }
        __syncthreads();
    }
}

}  // namespace

int condition_run_device(strique_ctx *ctx, int raw_kind, const void *raw, const int64_t *sig_off_dev,
                         const int64_t *sig_off_host, int n_reads, const CondModel &model, bool want_raw_stats,
                         void *flt_out, uint16_t *codes_out, float *code_values_out, double *stats_out) {
    if (n_reads == 0) return STRIQUE_OK;
    const int64_t total = sig_off_host[n_reads];
    for (int r = 0; r < n_reads; ++r)
        if (sig_off_host[r + 1] - sig_off_host[r] <= 0 || sig_off_host[r + 1] - sig_off_host[r] >= (1ll << 30))
            FAIL(ctx, STRIQUE_EINVAL, "read with no samples (or more than 2^30)");
    DevBuf &tmpA = ctx->buf("cond.tmpA"), &tmpB = ctx->buf("cond.tmpB");
    TRY(tmpA.ensure(ctx, total));
    TRY(tmpB.ensure(ctx, total));
    const int grid = std::min(n_reads, ctx->num_sms * 3);
    if (raw_kind == 0)
        condition_kernel<int16_t><<<grid, CT, 0, ctx->stream>>>((const int16_t *)raw, sig_off_dev, n_reads, model,
                                                                 want_raw_stats ? 1 : 0, (int16_t *)flt_out, codes_out,
                                                                 tmpA.as<uint8_t>(), tmpB.as<uint8_t>(), code_values_out,
                                                                 stats_out);
    else if (raw_kind == 1)
        condition_kernel<double><<<grid, CT, 0, ctx->stream>>>((const double *)raw, sig_off_dev, n_reads, model,
                                                                want_raw_stats ? 1 : 0, (double *)flt_out, codes_out,
                                                                tmpA.as<uint8_t>(), tmpB.as<uint8_t>(), code_values_out,
                                                                stats_out);
    else
        FAIL(ctx, STRIQUE_EINVAL, "raw_kind must be 0 (int16) or 1 (float64)");
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return STRIQUE_OK;
}

}  // namespace strique
