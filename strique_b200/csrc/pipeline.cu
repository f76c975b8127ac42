// strique_detect_batch: the batched per-read path (reference repeatCounter.detect,
// scripts/STRique.py:581-618) and strique_condition_batch / strique_target_create.
#include <math.h>

#include <algorithm>
#include <numeric>
#include <string>
#include <vector>

#include "pipeline.cuh"

namespace strique {

__global__ void gather_patterns_kernel(const uint8_t *__restrict__ pat, const int64_t *__restrict__ src_end,
                                       const int64_t *__restrict__ dst_off, const int32_t *__restrict__ len, int n,
                                       uint8_t *__restrict__ out) {
    for (int s = blockIdx.x; s < n; s += gridDim.x) {
        const int l = len[s];
        const uint8_t *src = pat + src_end[s] - l;   // patterns are right-aligned in their slot
        for (int i = threadIdx.x; i < l; i += blockDim.x) out[dst_off[s] + i] = src[i];
    }
}

static CondModel to_cond_model(const strique_pore_constants &p) {
    CondModel m;
    m.m5_mod = p.m5_mod; m.m95_mod = p.m95_mod; m.model_min = p.model_min; m.model_max = p.model_max;
    return m;
}

// uploads raw + offsets (when host resident) and runs the conditioning kernel
static int stage_condition(strique_ctx *ctx, const strique_pore_constants &pore, int n_reads, const void *raw,
                           int raw_kind, const int64_t *raw_offsets, int memspace, bool want_raw, const void **raw_dev_out) {
    const int64_t total = raw_offsets[n_reads];
    const size_t esz = raw_kind == 0 ? 2 : 8;
    DevBuf &d_raw = ctx->buf("pl.raw"), &d_off = ctx->buf("pl.off"), &d_flt = ctx->buf("pl.flt"),
           &d_codes = ctx->buf("pl.codes"), &d_vals = ctx->buf("pl.vals"), &d_stats = ctx->buf("pl.stats");
    TRY(d_off.ensure(ctx, (size_t)(n_reads + 1) * 8));
    TRY(d_flt.ensure(ctx, (size_t)total * esz));
    TRY(d_codes.ensure(ctx, (size_t)total * 2 + 16));       // + 16: the scan prefetches up to 2 codes past a read
    TRY(d_vals.ensure(ctx, (size_t)n_reads * 256 * 4));
    TRY(d_stats.ensure(ctx, (size_t)n_reads * CS_STRIDE * 8));
    const void *raw_dev = raw;
    const CondModel cm = to_cond_model(pore);
    stage_mark(ctx, 2 * STRIQUE_STAGE_H2D);
    CUDA_TRY(ctx, cudaMemcpyAsync(d_off.p, raw_offsets, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    stage_mark(ctx, 2 * STRIQUE_STAGE_H2D + 1);
    stage_mark(ctx, 2 * STRIQUE_STAGE_CONDITION);
    if (memspace != STRIQUE_DEVICE) {
        // Host-resident reads: upload in up to 8 slices of whole reads on a second stream; the conditioning of
        // slice k (one CTA per read) runs while slice k + 1 crosses PCIe.
        TRY(d_raw.ensure(ctx, (size_t)total * esz));
        TRY(ctx->buf("cond.tmpA").ensure(ctx, total));           // sized once: the slices must not re-allocate them
        TRY(ctx->buf("cond.tmpB").ensure(ctx, total));
        raw_dev = d_raw.p;
        if (!ctx->copy_stream) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        const int n_slices = (int)std::max<int64_t>(1, std::min<int64_t>(8, std::min<int64_t>(n_reads, total * (int64_t)esz >> 22)));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));       // earlier work may still read pl.raw (and the buffers may have moved)
        int r0 = 0;
        for (int k = 0; k < n_slices; ++k) {
            // slice boundaries by samples, so slices carry equal bytes
            int r1 = r0;
            const int64_t want = total * (k + 1) / n_slices;
            while (r1 < n_reads && (raw_offsets[r1] < want || r1 == r0)) ++r1;
            if (k == n_slices - 1) r1 = n_reads;
            if (r1 == r0) continue;
            const size_t b0 = (size_t)raw_offsets[r0] * esz, b1 = (size_t)raw_offsets[r1] * esz;
            CUDA_TRY(ctx, cudaMemcpyAsync((char *)d_raw.p + b0, (const char *)raw + b0, b1 - b0, cudaMemcpyHostToDevice, ctx->copy_stream));
            if (!ctx->copy_ev[k]) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->copy_ev[k], cudaEventDisableTiming));
            CUDA_TRY(ctx, cudaEventRecord(ctx->copy_ev[k], ctx->copy_stream));
            CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->copy_ev[k], 0));
            TRY(condition_run_device(ctx, raw_kind, raw_dev, d_off.as<int64_t>() + r0, raw_offsets + r0, r1 - r0, cm, want_raw,
                                     d_flt.p, d_codes.as<uint16_t>(), d_vals.as<float>() + (size_t)r0 * 256,
                                     d_stats.as<double>() + (size_t)r0 * CS_STRIDE));
            r0 = r1;
        }
    } else {
        TRY(condition_run_device(ctx, raw_kind, raw_dev, d_off.as<int64_t>(), raw_offsets, n_reads, cm, want_raw, d_flt.p,
                                 d_codes.as<uint16_t>(), d_vals.as<float>(), d_stats.as<double>()));
    }
    stage_mark(ctx, 2 * STRIQUE_STAGE_CONDITION + 1);
    *raw_dev_out = raw_dev;
    return STRIQUE_OK;
}

}  // namespace strique

using namespace strique;

extern "C" int64_t strique_last_viterbi_edges(const strique_ctx *ctx) { return ctx ? ctx->last_viterbi_edges : 0; }
extern "C" int64_t strique_last_mod_bytes(const strique_ctx *ctx) { return ctx ? ctx->last_mod_bytes : 0; }
extern "C" int64_t strique_last_viterbi_fixed(const strique_ctx *ctx) { return ctx ? ctx->last_viterbi_fixed : 0; }
extern "C" int64_t strique_last_viterbi_declined(const strique_ctx *ctx) { return ctx ? ctx->last_viterbi_declined : 0; }
extern "C" int strique_set_viterbi_exact(strique_ctx *ctx, int exact) {
    if (!ctx) return STRIQUE_EINVAL;
    ctx->viterbi_exact = exact != 0;
    return STRIQUE_OK;
}
extern "C" float strique_last_stage_ms(const strique_ctx *ctx, int stage) {
    return (ctx && stage >= 0 && stage < STRIQUE_N_STAGES) ? ctx->stage_ms[stage] : 0.f;
}

extern "C" int strique_condition_batch(strique_ctx *ctx, const strique_pore_constants *pore, int n_reads,
                                       const void *raw, int raw_kind, const int64_t *raw_offsets, int want_raw_stats,
                                       void *flt_out, uint16_t *codes_out, float *values_out,
                                       strique_condition_stats *stats_out) {
    if (!ctx) return STRIQUE_EINVAL;
    if (!pore || n_reads < 0 || (raw_kind != 0 && raw_kind != 1)) FAIL(ctx, STRIQUE_EINVAL, "strique_condition_batch: bad argument");
    if (n_reads == 0) return STRIQUE_OK;
    if (!raw || !raw_offsets) FAIL(ctx, STRIQUE_EINVAL, "strique_condition_batch: null pointer");
    static_assert(sizeof(strique_condition_stats) == CS_STRIDE * 8, "stats layout");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    stage_reset(ctx);
    const void *raw_dev = nullptr;
    TRY(stage_condition(ctx, *pore, n_reads, raw, raw_kind, raw_offsets, STRIQUE_HOST, want_raw_stats != 0, &raw_dev));
    const int64_t total = raw_offsets[n_reads];
    const size_t esz = raw_kind == 0 ? 2 : 8;
    if (flt_out) CUDA_TRY(ctx, cudaMemcpyAsync(flt_out, ctx->buf("pl.flt").p, (size_t)total * esz, cudaMemcpyDeviceToHost, ctx->stream));
    if (codes_out) CUDA_TRY(ctx, cudaMemcpyAsync(codes_out, ctx->buf("pl.codes").p, (size_t)total * 2, cudaMemcpyDeviceToHost, ctx->stream));
    if (values_out) CUDA_TRY(ctx, cudaMemcpyAsync(values_out, ctx->buf("pl.vals").p, (size_t)n_reads * 256 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    if (stats_out) CUDA_TRY(ctx, cudaMemcpyAsync(stats_out, ctx->buf("pl.stats").p, (size_t)n_reads * CS_STRIDE * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    stage_collect(ctx, STRIQUE_STAGE_CONDITION);
    return STRIQUE_OK;
}

extern "C" int strique_target_create(strique_ctx *ctx, const strique_target_desc *d, int32_t *target_id) {
    if (!ctx || !d || !target_id) return STRIQUE_EINVAL;
    if (d->n_prefix_levels <= 0 || d->n_suffix_levels <= 0 || !d->prefix_levels || !d->suffix_levels)
        FAIL(ctx, STRIQUE_EINVAL, "target: empty flank");
    if (d->count_model < 0 || d->count_model >= (int)ctx->models.size() || d->mod_model >= (int)ctx->models.size())
        FAIL(ctx, STRIQUE_EINVAL, "target: unknown HMM id");
    Target *t = new Target();
    t->prefix_levels.assign(d->prefix_levels, d->prefix_levels + d->n_prefix_levels);
    t->suffix_levels.assign(d->suffix_levels, d->suffix_levels + d->n_suffix_levels);
    t->pre_trim = d->pre_trim; t->post_trim = d->post_trim;
    t->count_model = d->count_model; t->mod_model = d->mod_model; t->count_offset = d->count_offset;
    ctx->targets.push_back(t);
    *target_id = (int32_t)ctx->targets.size() - 1;
    return STRIQUE_OK;
}

extern "C" int strique_detect_batch(strique_ctx *ctx, const strique_detect_config *cfg, int n_reads, const void *raw,
                                    int raw_kind, const int64_t *raw_offsets, const int32_t *read_target, int memspace,
                                    strique_detect_result *results, uint8_t *mod_out, int64_t mod_cap) {
    if (!ctx) return STRIQUE_EINVAL;
    if (!cfg || n_reads < 0 || (raw_kind != 0 && raw_kind != 1) || cfg->samples <= 0)
        FAIL(ctx, STRIQUE_EINVAL, "strique_detect_batch: bad argument");
    if (n_reads == 0) return STRIQUE_OK;
    if (!raw || !raw_offsets || !read_target || !results) FAIL(ctx, STRIQUE_EINVAL, "strique_detect_batch: null pointer");
    for (int r = 0; r < n_reads; ++r)
        if (read_target[r] < 0 || read_target[r] >= (int)ctx->targets.size()) FAIL(ctx, STRIQUE_EINVAL, "unknown target id");
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    stage_reset(ctx);
    ctx->last_viterbi_edges = 0;
    ctx->last_viterbi_fixed = ctx->last_viterbi_declined = 0;
    ctx->last_mod_bytes = 0;
    const bool use_mod = cfg->use_mod != 0;
    // ---- 1. conditioning ------------------------------------------------------------------------
    const void *raw_dev = nullptr;
    TRY(stage_condition(ctx, cfg->pore, n_reads, raw, raw_kind, raw_offsets, memspace, use_mod, &raw_dev));
    // ---- 2. two flank alignments per read -------------------------------------------------------
    const int n_targets = (int)ctx->targets.size();
    std::vector<float> levels;
    std::vector<int32_t> flank_off(1, 0);
    for (int g = 0; g < n_targets; ++g) {
        const Target &t = *ctx->targets[g];
        levels.insert(levels.end(), t.prefix_levels.begin(), t.prefix_levels.end());
        flank_off.push_back((int32_t)levels.size());
        levels.insert(levels.end(), t.suffix_levels.begin(), t.suffix_levels.end());
        flank_off.push_back((int32_t)levels.size());
    }
    DevBuf &d_lev = ctx->buf("pl.levels"), &d_foff = ctx->buf("pl.flankoff");
    TRY(d_lev.ensure(ctx, levels.size() * 4));
    TRY(d_foff.ensure(ctx, flank_off.size() * 4));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_lev.p, levels.data(), levels.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_foff.p, flank_off.data(), flank_off.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<int32_t> tsig(2 * n_reads), tflank(2 * n_reads), tpre(2 * n_reads), tpost(2 * n_reads);
    for (int r = 0; r < n_reads; ++r) {
        const Target &t = *ctx->targets[read_target[r]];
        tsig[2 * r] = r; tflank[2 * r] = 2 * read_target[r]; tpre[2 * r] = t.pre_trim; tpost[2 * r] = 0;
        tsig[2 * r + 1] = r; tflank[2 * r + 1] = 2 * read_target[r] + 1; tpre[2 * r + 1] = 0; tpost[2 * r + 1] = t.post_trim;
    }
    AlignDeviceInputs in;
    in.n_signals = n_reads; in.codes = ctx->buf("pl.codes").as<uint16_t>(); in.sig_off = ctx->buf("pl.off").as<int64_t>();
    in.sig_off_host = raw_offsets; in.code_values = ctx->buf("pl.vals").as<float>(); in.n_code_values = 256;
    in.n_flanks = 2 * n_targets; in.flank_levels = d_lev.as<float>(); in.flank_off = d_foff.as<int32_t>();
    in.flank_off_host = flank_off.data(); in.samples = cfg->samples;
    std::vector<strique_align_result> ares(2 * n_reads);
    std::vector<double> stats((size_t)n_reads * CS_STRIDE);
    TRY(align_run_device(ctx, cfg->align, in, 2 * n_reads, tsig.data(), tflank.data(), tpre.data(), tpost.data(),
                         ares.data(), nullptr, 0, nullptr));
    CUDA_TRY(ctx, cudaMemcpyAsync(stats.data(), ctx->buf("pl.stats").p, stats.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    stage_collect(ctx, STRIQUE_STAGE_H2D);
    stage_collect(ctx, STRIQUE_STAGE_CONDITION);
    // ---- 3. gate + count HMM --------------------------------------------------------------------
    std::vector<int> hmm_reads;
    for (int r = 0; r < n_reads; ++r) {
        strique_detect_result &o = results[r];
        memset(&o, 0, sizeof(o));
        const strique_align_result &p = ares[2 * r], &s = ares[2 * r + 1];
        o.status = stats[(size_t)r * CS_STRIDE + CS_STATUS] != 0.0 ? 1 : 0;
        o.score_prefix = p.end0 > p.begin0 ? (double)p.score / (double)(p.end0 - p.begin0) : 0.0;   // S.py:542-545
        o.score_suffix = s.end0 > s.begin0 ? (double)s.score / (double)(s.end0 - s.begin0) : 0.0;
        o.prefix_begin = p.begin_trim; o.prefix_end = p.end_trim;
        o.suffix_begin = s.begin_trim; o.suffix_end = s.end_trim;
        o.offset = o.prefix_end;
        o.ticks = std::max(o.suffix_begin - o.prefix_end, 0);
        o.mod_len = -1;
        if (o.status) { o.score_prefix = 0.0; o.score_suffix = 0.0; continue; }
        if (o.prefix_begin < o.suffix_end && o.score_prefix > 0.0 && o.score_suffix > 0.0) hmm_reads.push_back(r);   // S.py:603
    }
    double c3, c4, lo, hi;
    minmax_model_constants(to_cond_model(cfg->pore), &c3, &c4, &lo, &hi);
    const void *flt_dev = ctx->buf("pl.flt").p;
    DevBuf &d_segs = ctx->buf("pl.segs"), &d_x = ctx->buf("pl.x");
    std::vector<int> mod_reads;
    // sequences grouped by model so each Viterbi launch sees one model image
    auto run_hmm_stage = [&](const std::vector<int> &reads, bool mod_stage, std::vector<strique_viterbi_result> &vres,
                             std::vector<int64_t> &xoff_all, std::vector<int> &seq_read) -> int {
        vres.clear(); xoff_all.assign(1, 0); seq_read.clear();
        if (reads.empty()) return STRIQUE_OK;
        HostTimer ht_stage(mod_stage ? "hmm stage (mod)" : "hmm stage (count)");
        std::vector<int> ord(reads);
        auto model_of = [&](int r) { const Target &t = *ctx->targets[read_target[r]]; return mod_stage ? t.mod_model : t.count_model; };
        std::stable_sort(ord.begin(), ord.end(), [&](int a, int b) { return model_of(a) < model_of(b); });
        std::vector<PrepSeg> segs;
        for (int r : ord) {
            const strique_detect_result &o = results[r];
            PrepSeg sg;
            if (!mod_stage) {
                sg.src_off = raw_offsets[r] + o.prefix_begin; sg.len = o.suffix_end - o.prefix_begin;
            } else {
                sg.src_off = raw_offsets[r] + o.prefix_begin + o.mod_off /* t_first, stashed */; sg.len = o.mod_len /* stashed length */;
            }
            sg.dst_off = xoff_all.back(); sg.read = r;
            segs.push_back(sg);
            seq_read.push_back(r);
            xoff_all.push_back(xoff_all.back() + sg.len);
        }
        const int n = (int)segs.size();
        HostTimer ht_prep("hmm stage: after seg build");
        TRY(d_segs.ensure(ctx, (size_t)n * sizeof(PrepSeg)));
        TRY(d_x.ensure(ctx, std::max<int64_t>(1, xoff_all.back()) * 8));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_segs.p, segs.data(), (size_t)n * sizeof(PrepSeg), cudaMemcpyHostToDevice, ctx->stream));
        const int stage = mod_stage ? STRIQUE_STAGE_VITERBI_MOD : STRIQUE_STAGE_VITERBI_COUNT;
        stage_mark(ctx, 2 * stage);
        if (!mod_stage)
            TRY(viterbi_prepare_x(ctx, raw_kind, flt_dev, d_segs.as<PrepSeg>(), n, ctx->buf("pl.stats").as<double>(), CS_FLT_C1,
                                  c3, c4, lo, hi, -INFINITY, INFINITY, d_x.as<double>()));
        else
            TRY(viterbi_prepare_x(ctx, raw_kind, raw_dev, d_segs.as<PrepSeg>(), n, ctx->buf("pl.stats").as<double>(), CS_RAW_C1,
                                  c3, c4, lo, hi, cfg->mod_clip_lo, cfg->mod_clip_hi, d_x.as<double>()));
        vres.resize(n);
        // all groups write their patterns at absolute x offsets: size the shared buffer once
        TRY(ctx->buf("vit.pattern").ensure(ctx, std::max<int64_t>(16, xoff_all.back())));
        std::vector<int32_t> seq_model(n);
        for (int i = 0; i < n; ++i) {
            seq_model[i] = model_of(ord[i]);
            if (seq_model[i] < 0 || seq_model[i] >= (int)ctx->models.size()) FAIL(ctx, STRIQUE_EINVAL, "target without the requested HMM");
        }
        TRY(viterbi_run_device_multi(ctx, seq_model.data(), d_x.as<double>(), xoff_all.data(), n, vres.data(), nullptr, nullptr));
        stage_mark(ctx, 2 * stage + 1);
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        stage_collect(ctx, stage);
        return STRIQUE_OK;
    };
    std::vector<strique_viterbi_result> vres;
    std::vector<int64_t> xoff;
    std::vector<int> seq_read;
    TRY(run_hmm_stage(hmm_reads, false, vres, xoff, seq_read));
    for (size_t i = 0; i < seq_read.size(); ++i) {
        const int r = seq_read[i];
        strique_detect_result &o = results[r];
        const strique_viterbi_result &v = vres[i];
        if (v.status == 0) {
            o.hmm_ran = 1;
            o.count = v.n_count + ctx->targets[read_target[r]]->count_offset;   // S.py:437
            o.log_p = v.logp;
            if (use_mod) {
                if (v.t_first >= 0) {
                    o.mod_off = v.t_first;                  // stash: first repeat sample, relative to prefix_begin
                    o.mod_len = v.t_last - v.t_first + 1;   // stash: number of repeat samples
                    mod_reads.push_back(r);
                } else {
                    o.mod_len = -1;                         // empty repeat signal: pomegranate finds no path -> '-'
                }
            }
        } else if (v.status == 2) {
            FAIL(ctx, STRIQUE_ECUDA, "viterbi traceback inconsistency");
        }
    }
    // ---- 4. methylation HMM over the repeat samples ----------------------------------------------
    int64_t mod_used = 0;
    if (use_mod) {
        std::vector<strique_viterbi_result> mres;
        std::vector<int64_t> mxoff;
        std::vector<int> mseq;
        TRY(run_hmm_stage(mod_reads, true, mres, mxoff, mseq));
        const int n = (int)mseq.size();
        std::vector<int64_t> src_end(n), dst_off(n);
        std::vector<int32_t> plen(n);
        for (int i = 0; i < n; ++i) {
            strique_detect_result &o = results[mseq[i]];
            if (mres[i].status == 0) { plen[i] = mres[i].pattern_len; o.mod_len = plen[i]; }
            else { plen[i] = 0; o.mod_len = -1; if (mres[i].status == 2) FAIL(ctx, STRIQUE_ECUDA, "viterbi traceback inconsistency"); }
            o.mod_off = mod_used;
            src_end[i] = mxoff[i + 1];
            dst_off[i] = mod_used;
            mod_used += plen[i];
        }
        ctx->last_mod_bytes = mod_used;
        // the caller retries with a buffer of strique_last_mod_bytes() bytes
        if (mod_used > mod_cap || (mod_used > 0 && !mod_out)) FAIL(ctx, STRIQUE_ENOSPC, "mod_out buffer too small");
        if (mod_used > 0) {
            DevBuf &d_se = ctx->buf("pl.pat_srcend"), &d_do = ctx->buf("pl.pat_dstoff"), &d_pl = ctx->buf("pl.pat_len"),
                   &d_out = ctx->buf("pl.pat_out");
            TRY(d_se.ensure(ctx, (size_t)n * 8)); TRY(d_do.ensure(ctx, (size_t)n * 8)); TRY(d_pl.ensure(ctx, (size_t)n * 4));
            TRY(d_out.ensure(ctx, (size_t)mod_used));
            CUDA_TRY(ctx, cudaMemcpyAsync(d_se.p, src_end.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(ctx, cudaMemcpyAsync(d_do.p, dst_off.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(ctx, cudaMemcpyAsync(d_pl.p, plen.data(), (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
            gather_patterns_kernel<<<std::min(n, ctx->num_sms * 4), 128, 0, ctx->stream>>>(
                ctx->buf("vit.pattern").as<uint8_t>(), d_se.as<int64_t>(), d_do.as<int64_t>(), d_pl.as<int32_t>(), n,
                d_out.as<uint8_t>());
            ctx->launches++;
            CUDA_TRY(ctx, cudaGetLastError());
            CUDA_TRY(ctx, cudaMemcpyAsync(mod_out, d_out.p, (size_t)mod_used, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        }
    }
    for (int r = 0; r < n_reads; ++r)
        if (results[r].mod_len < 0) results[r].mod_off = 0;
    return STRIQUE_OK;
}
