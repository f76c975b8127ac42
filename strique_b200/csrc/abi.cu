// C ABI of libstrique_b200 (declared in include/strique_b200.h): context management and the
// batched alignment entry point.  Host-side batch planning lives here; kernels in align.cu.
#include <math.h>

#include <algorithm>
#include <numeric>

#include "align.cuh"
#include "pipeline.cuh"

std::string g_strique_create_error;

strique_ctx::~strique_ctx() {
    for (auto &kv : bufs)
        if (kv.second.p) cudaFree(kv.second.p);
    for (void *p : owned) cudaFree(p);
    for (auto *m : models) delete m;
    for (auto *t : targets) delete t;
    for (auto &e : stage_ev)
        if (e) cudaEventDestroy(e);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (stream) cudaStreamDestroy(stream);
    if (copy_stream) cudaStreamDestroy(copy_stream);
    for (cudaEvent_t e : copy_ev) if (e) cudaEventDestroy(e);
}

extern "C" int strique_version(void) { return 200; }

// 1 when a flank of n_levels k-mer levels, `samples` samples each, fits an alignment kernel (callers check when a
// target is defined, not when the first batch fails)
extern "C" int strique_align_supported(int n_levels, int samples) {
    int K = 0, S = 0;
    return n_levels > 0 && samples > 0 && strique::align_pick_kernel(n_levels, samples, &K, &S) ? 1 : 0;
}

// Page-locked host memory for the batch staging buffers of the caller (strique_detect_batch then uploads at PCIe
// speed instead of through the driver's bounce buffer).
extern "C" void *strique_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
extern "C" void strique_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

extern "C" int strique_ctx_create(int device, strique_ctx **out) {
    if (!out) return STRIQUE_EINVAL;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_strique_create_error = std::string("no CUDA device available: ") + cudaGetErrorString(e);
        cudaGetLastError();
        return STRIQUE_ECUDA;
    }
    if (device < 0 || device >= count) {
        g_strique_create_error = "device index out of range";
        return STRIQUE_EINVAL;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) {
        g_strique_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return STRIQUE_ECUDA;
    }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
        g_strique_create_error = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e);
        return STRIQUE_ECUDA;
    }
    if (prop.major != 10) {
        g_strique_create_error = "libstrique_b200 carries sm_100a code only; device is sm_" +
                                 std::to_string(prop.major) + std::to_string(prop.minor);
        return STRIQUE_EUNSUPPORTED;
    }
    strique_ctx *ctx = new strique_ctx();
    ctx->device = device;
    ctx->num_sms = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaEventCreate(&ctx->ev0)) != cudaSuccess || (e = cudaEventCreate(&ctx->ev1)) != cudaSuccess) {
        g_strique_create_error = std::string("stream/event creation: ") + cudaGetErrorString(e);
        delete ctx;
        return STRIQUE_ECUDA;
    }
    *out = ctx;
    return STRIQUE_OK;
}

extern "C" void strique_ctx_destroy(strique_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    delete ctx;
}

extern "C" const char *strique_last_error(const strique_ctx *ctx) {
    return ctx ? ctx->error.c_str() : g_strique_create_error.c_str();
}

extern "C" int64_t strique_launch_count(const strique_ctx *ctx) { return ctx ? ctx->launches : 0; }
extern "C" void *strique_ctx_stream(const strique_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
extern "C" int64_t strique_last_align_cells(const strique_ctx *ctx) { return ctx ? ctx->last_align_cells : 0; }
extern "C" float strique_last_scan_ms(const strique_ctx *ctx) { return ctx ? ctx->last_scan_ms : 0.f; }

extern "C" int strique_ctx_synchronize(strique_ctx *ctx) {
    if (!ctx) return STRIQUE_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return STRIQUE_OK;
}

namespace strique {

__global__ void widen_u8_kernel(const uint8_t *__restrict__ in, uint16_t *__restrict__ out, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = in[i];
}

// DP column 0 of the S matrix for a flank of L rows (first column is NOT free: vertical gaps only;
// seqan/align/dp_formula_affine.h:289-327 via dp_meta_info.h:186-194).  fp32, same operation order.
static void column0(const strique_align_params &p, int L, float *out /* [L+1] */) {
    const volatile float INF = STRIQUE_SEQAN_INF;
    float cS = 0.f, cV = INF;
    out[0] = 0.f;
    for (int i = 1; i <= L; ++i) {
        volatile float e = cV + p.gap_extension_v, o = cS + p.gap_open_v;
        cV = e < o ? o : e;
        out[i] = cV;
        cS = cV;
    }
}

// Runs the alignment stage on DEVICE-resident inputs.  Small descriptor arrays are host pointers.
int align_run_device(strique_ctx *ctx, const strique_align_params &params, const AlignDeviceInputs &in,
                     int n_tasks, const int32_t *task_signal, const int32_t *task_flank, const int32_t *task_pre,
                     const int32_t *task_post, strique_align_result *results_host, int32_t *rows_out_host,
                     int64_t rows_out_stride, strique_align_result *results_dev_out) {
    ctx->last_align_cells = 0;
    ctx->last_scan_ms = 0.f;
    if (n_tasks == 0) return STRIQUE_OK;
    // ---- plan ---------------------------------------------------------------------------------
    std::vector<int32_t> tK(n_tasks), tS(n_tasks);
    int maxRows = 0, maxK = 0, maxW = 0;
    for (int t = 0; t < n_tasks; ++t) {
        const int sg = task_signal[t], f = task_flank[t];
        if (sg < 0 || sg >= in.n_signals || f < 0 || f >= in.n_flanks) FAIL(ctx, STRIQUE_EINVAL, "task index out of range");
        const int nlev = in.flank_off_host[f + 1] - in.flank_off_host[f];
        const int64_t N = in.sig_off_host[sg + 1] - in.sig_off_host[sg];
        if (nlev <= 0 || N <= 0) FAIL(ctx, STRIQUE_EINVAL, "empty signal or flank (the reference returns FLT_MIN; handle it in the caller)");
        if (N >= (1ll << 30)) FAIL(ctx, STRIQUE_EUNSUPPORTED, "signal longer than 2^30 samples");
        int K, S;
        if (!align_pick_kernel(nlev, in.samples, &K, &S)) FAIL(ctx, STRIQUE_EUNSUPPORTED, "flank too long for the alignment kernels (at most 2048 flank samples = levels x samples per level)");
        tK[t] = K;
        tS[t] = S;
        maxRows = std::max(maxRows, 32 * K * S);
        maxK = std::max(maxK, K);
        maxW = std::max(maxW, 4 * ((K * S + 31) / 32));       // trace words per lane and column: four flag planes
    }
    const int ckpt_rows = (int)align_up(maxRows + 1, 32);
    const int64_t lut_task_stride = (int64_t)in.n_code_values * 32 * maxK;
    int maxL = 0;
    for (int f = 0; f < in.n_flanks; ++f) maxL = std::max(maxL, (in.flank_off_host[f + 1] - in.flank_off_host[f]) * in.samples);
    const int col0_stride = (int)align_up(std::max(maxL, maxRows) + 1, 32);
    std::vector<float> col0((size_t)in.n_flanks * col0_stride, 0.f);
    for (int f = 0; f < in.n_flanks; ++f)
        column0(params, (in.flank_off_host[f + 1] - in.flank_off_host[f]) * in.samples, col0.data() + (size_t)f * col0_stride);
    const int64_t rows_stride = maxL;

    DevBuf &d_col0 = ctx->buf("al.col0"), &d_tsig = ctx->buf("al.tsig"), &d_tflank = ctx->buf("al.tflank"),
           &d_tpre = ctx->buf("al.tpre"), &d_tpost = ctx->buf("al.tpost"), &d_tK = ctx->buf("al.tK"),
           &d_tS = ctx->buf("al.tS"), &d_order = ctx->buf("al.order"), &d_ckoff = ctx->buf("al.ckoff"),
           &d_lut = ctx->buf("al.lut"), &d_ckpt = ctx->buf("al.ckpt"), &d_trace = ctx->buf("al.trace"),
           &d_rows = ctx->buf("al.rows"), &d_res = ctx->buf("al.res"), &d_queue = ctx->buf("al.queue"),
           &d_fix = ctx->buf("al.fix"), &d_patch = ctx->buf("al.patch"), &d_ckstep = ctx->buf("al.ckstep"),
           &d_pair_order = ctx->buf("al.pair_order"), &d_single_order = ctx->buf("al.single_order");
    // chunk tasks so LUT + checkpoints stay within a device-memory budget
    size_t free_b = 0, total_b = 0;
    CUDA_TRY(ctx, ctx_mem_info(ctx, &free_b, &total_b));
    const size_t budget = std::max<size_t>((size_t)1 << 30, (size_t)(free_b * 0.6));
    const int trace_warps = ctx->num_sms * (getenv("STRIQUE_TRACE_WARPS") ? atoi(getenv("STRIQUE_TRACE_WARPS")) : ALIGN_TRACE_WARPS_MAX);   // single-warp CTAs (see align_trace_warps)
    const int fix_cap = 1 << 16;

    TRY(d_col0.ensure(ctx, col0.size() * sizeof(float)));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_col0.as<float>(), col0.data(), col0.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    TRY(d_trace.ensure(ctx, (size_t)trace_warps * ALIGN_CKPT * 32 * maxW * sizeof(uint32_t)));
    TRY(d_queue.ensure(ctx, 64));
    TRY(d_fix.ensure(ctx, (size_t)(2 * fix_cap + 1) * sizeof(unsigned long long)));
    TRY(d_patch.ensure(ctx, (size_t)2 * fix_cap * sizeof(unsigned long long)));

    std::vector<strique_align_result> res_chunk;
    std::vector<unsigned long long> fix_host(2 * fix_cap + 1), patch_host;
    const unsigned long long fix_head = 255;        // entries read back together with the count
    int64_t cells_total = 0;
    float scan_ms_total = 0.f;
    int t0 = 0;
    while (t0 < n_tasks) {
        // ---- chunk extent ---------------------------------------------------------------------
        size_t used = 0;
        int t1 = t0;
        std::vector<int64_t> ckoff;
        int64_t ck_floats = 0;
        // at most 60000 tasks per chunk, the remaining tasks in equal (even: pairs) shares -- a short last chunk would
        // run the kernels at a fraction of their occupancy; the memory budget below may still cut a chunk short.
        // (Nothing forces the 60000: one chunk of 65536 tasks measured 407 ms of scan against 378 ms for two of 32768.)
        const int max_chunk = 60000;
        const int remaining = n_tasks - t0, n_chunks_left = (remaining + max_chunk - 1) / max_chunk;
        const int chunk_cap = (((remaining + n_chunks_left - 1) / n_chunks_left) + 1) & ~1;
        while (t1 < n_tasks && t1 - t0 < chunk_cap) {
            const int64_t N = in.sig_off_host[task_signal[t1] + 1] - in.sig_off_host[task_signal[t1]];
            const int64_t nck = N / ALIGN_CKPT;
            const size_t need = (size_t)lut_task_stride * 4 + (size_t)nck * 2 * ckpt_rows * 4 + (size_t)rows_stride * 4;
            if (t1 > t0 && used + need > budget) break;
            ckoff.push_back(ck_floats);
            ck_floats += nck * 2 * ckpt_rows;
            used += need;
            ++t1;
        }
        const int n = t1 - t0;
        // ---- pairs: two flanks over the same signal with the same kernel shape are scanned by one warp
        // (LinSweepPair, linear gap costs only); their checkpoint areas are adjacent and become one interleaved area
        std::vector<int32_t> ckstep(n, 1), pair_first(n, 0);
        if (align_params_linear(params) && !getenv("STRIQUE_NO_PAIR_SCAN")) {
            for (int t = 0; t + 1 < n;) {
                if (task_signal[t0 + t] == task_signal[t0 + t + 1] && tS[t0 + t] == 6 && tS[t0 + t + 1] == 6 &&
                    tK[t0 + t] == tK[t0 + t + 1]) {
                    ckstep[t] = ckstep[t + 1] = 2;
                    pair_first[t] = 1;
                    ckoff[t + 1] = ckoff[t] + 1;
                    t += 2;
                } else {
                    ++t;
                }
            }
        }
        TRY(d_ckstep.ensure(ctx, n * 4)); TRY(d_pair_order.ensure(ctx, n * 4)); TRY(d_single_order.ensure(ctx, n * 4));
        TRY(d_tsig.ensure(ctx, n * 4)); TRY(d_tflank.ensure(ctx, n * 4)); TRY(d_tpre.ensure(ctx, n * 4));
        TRY(d_tpost.ensure(ctx, n * 4)); TRY(d_tK.ensure(ctx, n * 4)); TRY(d_tS.ensure(ctx, n * 4));
        TRY(d_order.ensure(ctx, n * 4)); TRY(d_ckoff.ensure(ctx, n * 8));
        TRY(d_lut.ensure(ctx, (size_t)n * lut_task_stride * 4));
        TRY(d_ckpt.ensure(ctx, std::max<size_t>(16, (size_t)ck_floats * 4)));
        TRY(d_rows.ensure(ctx, (size_t)n * rows_stride * 4));
        TRY(d_res.ensure(ctx, (size_t)n * sizeof(strique_align_result)));
        auto up = [&](DevBuf &d, const void *src, size_t bytes) {
            return cudaMemcpyAsync(d.p, src, bytes, cudaMemcpyHostToDevice, ctx->stream);
        };
        CUDA_TRY(ctx, up(d_tsig, task_signal + t0, n * 4));
        CUDA_TRY(ctx, up(d_tflank, task_flank + t0, n * 4));
        CUDA_TRY(ctx, up(d_tpre, task_pre + t0, n * 4));
        CUDA_TRY(ctx, up(d_tpost, task_post + t0, n * 4));
        CUDA_TRY(ctx, up(d_tK, tK.data() + t0, n * 4));
        CUDA_TRY(ctx, up(d_tS, tS.data() + t0, n * 4));
        CUDA_TRY(ctx, up(d_ckoff, ckoff.data(), n * 8));
        CUDA_TRY(ctx, up(d_ckstep, ckstep.data(), n * 4));
        CUDA_TRY(ctx, cudaMemsetAsync(d_res.p, 0, (size_t)n * sizeof(strique_align_result), ctx->stream));
        CUDA_TRY(ctx, cudaMemsetAsync(d_fix.p, 0, sizeof(unsigned long long), ctx->stream));

        AlignBatch b;
        b.p = params;
        b.codes = in.codes; b.sig_off = in.sig_off; b.code_values = in.code_values; b.n_code_values = in.n_code_values;
        b.flank_levels = in.flank_levels; b.flank_off = in.flank_off;
        b.col0 = d_col0.as<float>(); b.col0_stride = col0_stride; b.samples = in.samples;
        b.task_sig = d_tsig.as<int32_t>(); b.task_flank = d_tflank.as<int32_t>();
        b.task_pre = d_tpre.as<int32_t>(); b.task_post = d_tpost.as<int32_t>();
        b.lut = d_lut.as<float>(); b.lut_task_stride = lut_task_stride; b.lut_row = 32 * maxK;
        b.ckpt = d_ckpt.as<float>(); b.ckpt_off = d_ckoff.as<int64_t>(); b.ckpt_rows = ckpt_rows;
        b.ckpt_step = d_ckstep.as<int32_t>();
        b.trace = d_trace.as<uint32_t>(); b.rows = d_rows.as<int32_t>(); b.rows_stride = rows_stride;
        b.res = d_res.as<strique_align_result>(); b.queue = d_queue.as<int>();
        b.lut_fix = d_fix.as<unsigned long long>(); b.lut_fix_cap = fix_cap;

        stage_mark(ctx, 2 * STRIQUE_STAGE_ALIGN_TABLE);
        TRY(align_launch_build_lut(ctx, b, d_tK.as<int32_t>(), d_tS.as<int32_t>(), n));
        stage_mark(ctx, 2 * STRIQUE_STAGE_ALIGN_TABLE + 1);
        // ---- groups by kernel instantiation, longest signal first (host work while the table kernel runs) ----
        std::vector<int32_t> order(n);
        std::iota(order.begin(), order.end(), 0);
        auto len = [&](int t) { return in.sig_off_host[task_signal[t0 + t] + 1] - in.sig_off_host[task_signal[t0 + t]]; };
        std::stable_sort(order.begin(), order.end(), [&](int a, int c) {
            const int ka = tK[t0 + a] * 64 + tS[t0 + a], kc = tK[t0 + c] * 64 + tS[t0 + c];
            if (ka != kc) return ka < kc;
            return len(a) > len(c);
        });
        CUDA_TRY(ctx, up(d_order, order.data(), n * 4));
        std::vector<AlignGroup> groups;
        std::vector<int32_t> pair_order, single_order;      // per group: segments at the group's own offset
        pair_order.reserve(n); single_order.reserve(n);
        for (int i = 0; i < n;) {
            int k = i;
            while (k < n && tK[t0 + order[k]] == tK[t0 + order[i]] && tS[t0 + order[k]] == tS[t0 + order[i]]) ++k;
            AlignGroup g;
            g.K = tK[t0 + order[i]]; g.S = tS[t0 + order[i]]; g.n_tasks = k - i; g.order = d_order.as<int32_t>() + i;
            const size_t p0 = pair_order.size(), s0 = single_order.size();
            for (int q = i; q < k; ++q) {
                const int t = order[q];
                if (pair_first[t]) pair_order.push_back(t);
                else if (ckstep[t] == 1) single_order.push_back(t);
            }
            g.n_pairs = (int)(pair_order.size() - p0); g.pair_order = d_pair_order.as<int32_t>() + p0;
            g.n_single = (int)(single_order.size() - s0); g.single_order = d_single_order.as<int32_t>() + s0;
            groups.push_back(g);
            i = k;
        }
        if (!pair_order.empty()) CUDA_TRY(ctx, up(d_pair_order, pair_order.data(), pair_order.size() * 4));
        if (!single_order.empty()) CUDA_TRY(ctx, up(d_single_order, single_order.data(), single_order.size() * 4));
        for (int t = 0; t < n; ++t) {
            const int f = task_flank[t0 + t];
            cells_total += len(t) * (int64_t)((in.flank_off_host[f + 1] - in.flank_off_host[f]) * in.samples);
        }
        // ---- patch table entries too close to an fp32 rounding midpoint with libm's pow -------
        CUDA_TRY(ctx, cudaMemcpyAsync(fix_host.data(), d_fix.p, (1 + 2 * fix_head) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        const unsigned long long nfix = fix_host[0];
        if (nfix > (unsigned long long)fix_cap) FAIL(ctx, STRIQUE_EUNSUPPORTED, "too many borderline score-table entries");
        if (nfix > fix_head)
            CUDA_TRY(ctx, cudaMemcpy(fix_host.data() + 1 + 2 * fix_head, (char *)d_fix.p + (1 + 2 * fix_head) * 8, (nfix - fix_head) * 16, cudaMemcpyDeviceToHost));
        if (nfix) {
            patch_host.resize(2 * nfix);
            for (unsigned long long k = 0; k < nfix; ++k) {
                const int t = (int)(fix_host[1 + 2 * k] >> 40);
                const int64_t e = (int64_t)(fix_host[1 + 2 * k] & ((1ull << 40) - 1));
                const int Kt = tK[t0 + t], row_len = 32 * Kt;
                const int c = (int)(e / row_len), u = (int)(e % row_len);
                float h, v;
                const uint32_t hb = (uint32_t)(fix_host[2 + 2 * k] >> 32), vb = (uint32_t)fix_host[2 + 2 * k];
                memcpy(&h, &hb, 4);
                memcpy(&v, &vb, 4);
                const float d = h > v ? h - v : v - h;
                volatile float fx = (float)pow((double)d, 1.2);
                volatile float sc = params.dist_offset - fx;
                const float out = sc > params.dist_min ? sc : params.dist_min;
                uint32_t ob;
                memcpy(&ob, &out, 4);
                const int64_t pos = (int64_t)c * row_len + (u % Kt) * 32 + u / Kt;       // stored [u % K][u / K] inside the code row
                patch_host[2 * k] = (unsigned long long)((int64_t)t * lut_task_stride + pos);
                patch_host[2 * k + 1] = ob;
            }
            // (pageable source: the copy returns once the host buffer has been staged, so patch_host may be reused)
            CUDA_TRY(ctx, cudaMemcpyAsync(d_patch.p, patch_host.data(), nfix * 16, cudaMemcpyHostToDevice, ctx->stream));
            TRY(align_launch_patch_lut(ctx, b.lut, d_patch.as<unsigned long long>(), (int)nfix));
        }
        // ---- pass 1: scan (timed) -------------------------------------------------------------
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev0, ctx->stream));
        for (const AlignGroup &g : groups) {
            CUDA_TRY(ctx, cudaMemsetAsync(d_queue.p, 0, 64, ctx->stream));
            TRY(align_launch_scan(ctx, b, g));
        }
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev1, ctx->stream));
        // ---- pass 2: trace blocks + traceback -------------------------------------------------
        stage_mark(ctx, 2 * STRIQUE_STAGE_ALIGN_TRACE);
        for (const AlignGroup &g : groups) {
            CUDA_TRY(ctx, cudaMemsetAsync(d_queue.p, 0, 64, ctx->stream));
            TRY(align_launch_trace(ctx, b, g, trace_warps));
        }
        stage_mark(ctx, 2 * STRIQUE_STAGE_ALIGN_TRACE + 1);
        if (results_host)
            CUDA_TRY(ctx, cudaMemcpyAsync(results_host + t0, d_res.p, (size_t)n * sizeof(strique_align_result), cudaMemcpyDeviceToHost, ctx->stream));
        if (results_dev_out)
            CUDA_TRY(ctx, cudaMemcpyAsync(results_dev_out + t0, d_res.p, (size_t)n * sizeof(strique_align_result), cudaMemcpyDeviceToDevice, ctx->stream));
        if (rows_out_host)
            CUDA_TRY(ctx, cudaMemcpy2DAsync(rows_out_host + (size_t)t0 * rows_out_stride, rows_out_stride * 4, d_rows.p,
                                            rows_stride * 4, std::min<int64_t>(rows_stride, rows_out_stride) * 4, n,
                                            cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        float ms = 0.f;
        CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        scan_ms_total += ms;
        ctx->stage_ms[STRIQUE_STAGE_ALIGN_SCAN] += ms;
        stage_collect(ctx, STRIQUE_STAGE_ALIGN_TABLE);
        stage_collect(ctx, STRIQUE_STAGE_ALIGN_TRACE);
        t0 = t1;
    }
    ctx->last_align_cells = cells_total;
    ctx->last_scan_ms = scan_ms_total;
    return STRIQUE_OK;
}

}  // namespace strique

using namespace strique;

extern "C" int strique_align_batch(strique_ctx *ctx, const strique_align_params *params, int n_signals,
                                   const void *codes, int code_bytes, const int64_t *sig_offsets,
                                   const float *code_values, int n_code_values, int n_flanks,
                                   const float *flank_levels, const int32_t *flank_offsets, int samples, int n_tasks,
                                   const int32_t *task_signal, const int32_t *task_flank,
                                   const int32_t *task_pre_trim, const int32_t *task_post_trim, int memspace,
                                   strique_align_result *results, int32_t *rows_out, int64_t rows_stride) {
    if (!ctx) return STRIQUE_EINVAL;
    if (!params || n_signals < 0 || n_flanks < 0 || n_tasks < 0 || samples <= 0 || (code_bytes != 1 && code_bytes != 2) ||
        n_code_values <= 0 || n_code_values > 65536 || (code_bytes == 1 && n_code_values > 256))
        FAIL(ctx, STRIQUE_EINVAL, "strique_align_batch: bad argument");
    if (n_tasks == 0) return STRIQUE_OK;
    if (!codes || !sig_offsets || !code_values || !flank_levels || !flank_offsets || !task_signal || !task_flank ||
        !task_pre_trim || !task_post_trim || !results)
        FAIL(ctx, STRIQUE_EINVAL, "strique_align_batch: null pointer");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    stage_reset(ctx);
    const int64_t total = sig_offsets[n_signals];
    const int total_lev = flank_offsets[n_flanks];
    DevBuf &d_in8 = ctx->buf("ab.in8"), &d_codes = ctx->buf("ab.codes"), &d_sigoff = ctx->buf("ab.sigoff"),
           &d_vals = ctx->buf("ab.vals"), &d_lev = ctx->buf("ab.lev"), &d_flankoff = ctx->buf("ab.flankoff");
    TRY(d_codes.ensure(ctx, std::max<int64_t>(1, total) * 2 + 16));   // + 16: the scan prefetches up to 2 codes past a signal
    TRY(d_sigoff.ensure(ctx, (size_t)(n_signals + 1) * 8));
    TRY(d_flankoff.ensure(ctx, (size_t)(n_flanks + 1) * 4));
    const cudaMemcpyKind kind = memspace == STRIQUE_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const void *codes8 = codes;
    if (code_bytes == 1) {
        if (memspace != STRIQUE_DEVICE) {
            TRY(d_in8.ensure(ctx, std::max<int64_t>(1, total)));
            CUDA_TRY(ctx, cudaMemcpyAsync(d_in8.p, codes, total, kind, ctx->stream));
            codes8 = d_in8.p;
        }
        widen_u8_kernel<<<ctx->num_sms * 4, 256, 0, ctx->stream>>>((const uint8_t *)codes8, d_codes.as<uint16_t>(), total);
        ctx->launches++;
        CUDA_TRY(ctx, cudaGetLastError());
    } else {
        CUDA_TRY(ctx, cudaMemcpyAsync(d_codes.p, codes, total * 2, kind, ctx->stream));
    }
    const float *vals = code_values, *lev = flank_levels;
    if (memspace != STRIQUE_DEVICE) {
        TRY(d_vals.ensure(ctx, (size_t)n_signals * n_code_values * 4));
        TRY(d_lev.ensure(ctx, std::max(1, total_lev) * 4));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_vals.p, code_values, (size_t)n_signals * n_code_values * 4, kind, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_lev.p, flank_levels, (size_t)total_lev * 4, kind, ctx->stream));
        vals = d_vals.as<float>();
        lev = d_lev.as<float>();
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(d_sigoff.p, sig_offsets, (size_t)(n_signals + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_flankoff.p, flank_offsets, (size_t)(n_flanks + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
    AlignDeviceInputs in;
    in.n_signals = n_signals; in.codes = d_codes.as<uint16_t>(); in.sig_off = d_sigoff.as<int64_t>();
    in.sig_off_host = sig_offsets; in.code_values = vals; in.n_code_values = n_code_values;
    in.n_flanks = n_flanks; in.flank_levels = lev; in.flank_off = d_flankoff.as<int32_t>();
    in.flank_off_host = flank_offsets; in.samples = samples;
    return align_run_device(ctx, *params, in, n_tasks, task_signal, task_flank, task_pre_trim, task_post_trim, results,
                            rows_out, rows_stride, nullptr);
}
