// Host side of boundary #2: packing a compiled HMM into its device image, batch planning of the
// Viterbi stage and the C ABI entry points strique_hmm_create / strique_viterbi_batch.
#include <math.h>
#include <stdlib.h>

#include <algorithm>
#include <numeric>

#include "pipeline.cuh"
#include "viterbi.cuh"

namespace strique {

// pomegranate's NormalDistribution uses this truncated constant (distributions.pyx, SQRT_2_PI)
static const double SQRT_2_PI = 2.50662827463;

int hmm_create(strique_ctx *ctx, const strique_hmm_desc *d, HmmModel *out) {
    const int E = d->n_emit, C = d->n_chain;
    if (E <= 0 || C < 0 || d->n_end <= 0) FAIL(ctx, STRIQUE_EINVAL, "hmm: empty model");
    const int NS = (E + 31) / 32, QC = (C + 31) / 32;
    if (NS + QC > 16 || NS > VIT_MAX_SLOTS || QC > 4)
        FAIL(ctx, STRIQUE_EUNSUPPORTED, "hmm: more than 384 emitting or 128 chain states");
    const int START = E + C;
    // emitting states sorted by in-degree (descending) so every lane slot has a uniform edge count
    std::vector<int32_t> perm(E), inv(E);
    std::iota(perm.begin(), perm.end(), 0);
    auto degree = [&](int l) { return d->in_ptr[l + 1] - d->in_ptr[l]; };
    std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return degree(a) > degree(b); });
    for (int p = 0; p < E; ++p) inv[perm[p]] = p;
    VitModelDev &m = out->dev;
    memset(&m, 0, sizeof(m));
    m.E = E; m.C = C; m.NS = NS; m.QC = QC; m.n_end = d->n_end;
    int rows = 0;
    for (int s = 0; s < NS; ++s) {
        int dg = 0;
        for (int p = s * 32; p < std::min(E, s * 32 + 32); ++p) dg = std::max(dg, degree(perm[p]));
        if (dg > 15) FAIL(ctx, STRIQUE_EUNSUPPORTED, "hmm: emitting state with more than 15 in-edges");
        m.deg[s] = dg;
        m.row_base[s] = rows;
        rows += dg;
    }
    m.rows = rows;
    const int P_START = (NS + QC) * 32, P_NEG = P_START + 1;
    auto vpos = [&](int src) -> int {
        if (src >= 0 && src < E) return inv[src];
        if (src >= E && src < E + C) return NS * 32 + (src - E);
        if (src == START) return P_START;
        return -1;
    };
    // ---- image layout --------------------------------------------------------------------------
    size_t off = 0;
    auto section = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 16); return (int)o; };
    m.off_edge_w = section((size_t)rows * 32 * 8);
    m.off_em_p = section((size_t)3 * NS * 32 * 8);
    m.off_chain_predw = section((size_t)std::max(QC, 1) * 32 * 8);
    m.off_chain_ew = section((size_t)std::max(QC, 1) * 3 * 32 * 8);
    m.off_end_w = section((size_t)d->n_end * 8);
    m.off_edge_src = section((size_t)rows * 32 * 2);
    m.off_chain_es = section((size_t)std::max(QC, 1) * 3 * 32 * 2);
    m.off_end_src = section((size_t)d->n_end * 2);
    m.off_em_kind = section((size_t)NS * 32);
    m.off_flags = section((size_t)NS * 32);
    m.blob_bytes = (int)off;
    std::vector<unsigned char> blob(off, 0);
    double *edge_w = (double *)(blob.data() + m.off_edge_w);
    uint16_t *edge_src = (uint16_t *)(blob.data() + m.off_edge_src);
    double *em_p = (double *)(blob.data() + m.off_em_p);
    uint8_t *em_kind = blob.data() + m.off_em_kind, *flags = blob.data() + m.off_flags;
    double *predw = (double *)(blob.data() + m.off_chain_predw);
    double *ch_ew = (double *)(blob.data() + m.off_chain_ew);
    uint16_t *ch_es = (uint16_t *)(blob.data() + m.off_chain_es);
    double *end_w = (double *)(blob.data() + m.off_end_w);
    uint16_t *end_src = (uint16_t *)(blob.data() + m.off_end_src);
    for (int i = 0; i < rows * 32; ++i) { edge_w[i] = 0.0; edge_src[i] = (uint16_t)P_NEG; }
    int64_t n_edges = 0;
    for (int p = 0; p < NS * 32; ++p) {
        const int s = p / 32, lane = p % 32;
        if (p >= E) {   // padding state: uniform over an empty range, never reachable
            em_kind[p] = 1; em_p[p] = 1.0; em_p[NS * 32 + p] = 0.0; em_p[2 * NS * 32 + p] = -INFINITY;
            continue;
        }
        const int l = perm[p];
        // candidate order: the self loop, then the other edges from emitting states / START, then the edges from
        // chain states (stable within each class)
        int k = 0;
        for (int pass = 0; pass < 3; ++pass)
            for (int e = d->in_ptr[l]; e < d->in_ptr[l + 1]; ++e) {
                const bool from_chain = d->in_src[e] >= E && d->in_src[e] < E + C;
                const int cls = from_chain ? 2 : (d->in_src[e] == l ? 0 : 1);
                if (cls != pass) continue;
                const int v = vpos(d->in_src[e]);
                if (v < 0) FAIL(ctx, STRIQUE_EINVAL, "hmm: in-edge source out of range");
                edge_w[(m.row_base[s] + k) * 32 + lane] = d->in_logw[e];
                edge_src[(m.row_base[s] + k) * 32 + lane] = (uint16_t)v;
                ++n_edges;
                ++k;
            }
        flags[p] = d->emit_flags ? d->emit_flags[l] : 0;
        const double a = d->emit_a[l], b = d->emit_b[l];
        if (d->emit_kind[l] == 0) {
            em_kind[p] = 0;
            em_p[p] = a;
            em_p[NS * 32 + p] = -log(b * SQRT_2_PI);
            em_p[2 * NS * 32 + p] = b > 0 ? 1.0 / (2.0 * (b * b)) : 0.0;
        } else {
            em_kind[p] = 1;
            em_p[p] = a;
            em_p[NS * 32 + p] = b;
            em_p[2 * NS * 32 + p] = -log(b - a);
        }
    }
    for (int i = 0; i < std::max(QC, 1) * 32; ++i) predw[i] = -INFINITY;
    for (int i = 0; i < std::max(QC, 1) * 3 * 32; ++i) { ch_ew[i] = 0.0; ch_es[i] = (uint16_t)P_NEG; }
    for (int c = 0; c < C; ++c) {
        const int lane = c / QC, q = c % QC;
        // a chain may not continue across the lane-0 boundary implicitly: pred weight of c = 0 is ignored
        predw[q * 32 + lane] = c == 0 ? -INFINITY : d->chain_pred_logw[c];
        if (c > 0 && d->chain_pred_logw[c] > -INFINITY) ++n_edges;
        const int n_in = d->chain_in_ptr[c + 1] - d->chain_in_ptr[c];
        if (n_in > 3) FAIL(ctx, STRIQUE_EUNSUPPORTED, "hmm: chain state with more than 3 entry edges");
        for (int k = 0; k < n_in; ++k) {
            const int e = d->chain_in_ptr[c] + k;
            const int src = d->chain_in_src[e];
            if (!((src >= 0 && src < E) || src == START)) FAIL(ctx, STRIQUE_EINVAL, "hmm: chain entry edges must come from emitting states or START");
            ch_ew[(q * 3 + k) * 32 + lane] = d->chain_in_logw[e];
            ch_es[(q * 3 + k) * 32 + lane] = (uint16_t)vpos(src);
            ++n_edges;
        }
    }
    for (int e = 0; e < d->n_end; ++e) {
        const int v = vpos(d->end_src[e]);
        if (v < 0 || v == P_START) FAIL(ctx, STRIQUE_EINVAL, "hmm: END edge source out of range");
        end_w[e] = d->end_logw[e];
        end_src[e] = (uint16_t)v;
    }
    out->n_edges = n_edges;
    void *dblob = nullptr, *dperm = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&dblob, blob.size()));
    CUDA_TRY(ctx, cudaMalloc(&dperm, (size_t)NS * 32 * 4));
    std::vector<int32_t> permpad(NS * 32, 0);
    std::copy(perm.begin(), perm.end(), permpad.begin());
    CUDA_TRY(ctx, cudaMemcpy(dblob, blob.data(), blob.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(dperm, permpad.data(), permpad.size() * 4, cudaMemcpyHostToDevice));
    ctx->owned.push_back(dblob);
    ctx->owned.push_back(dperm);
    m.blob = (const unsigned char *)dblob;
    m.perm = (const int32_t *)dperm;
    return viterbi_profile_pack(ctx, d, out);
}

// Decodes n_seq device-resident sequences, sequence s with model ctx->models[seq_model[s]].
// Sequences of linear profile models (every count model of the reference) are decoded together, one launch for
// both strands / all loci; the rest go model by model through the small-model kernel or the generic kernel.
int viterbi_run_device_multi(strique_ctx *ctx, const int32_t *seq_model, const double *x_dev, const int64_t *x_off_host,
                             int n_seq, strique_viterbi_result *results_host, uint8_t *pattern_host,
                             uint16_t *path_host) {
    if (n_seq == 0) return STRIQUE_OK;
    HostTimer ht_all("viterbi_run_device_multi");
    const int64_t total = x_off_host[n_seq];
    const int n_models = (int)ctx->models.size();
    for (int s = 0; s < n_seq; ++s) {
        if (x_off_host[s + 1] - x_off_host[s] >= (1ll << 30) || x_off_host[s + 1] < x_off_host[s])
            FAIL(ctx, STRIQUE_EINVAL, "viterbi: bad sequence offsets");
        if (seq_model[s] < 0 || seq_model[s] >= n_models) FAIL(ctx, STRIQUE_EINVAL, "unknown HMM id");
    }
    DevBuf &d_xoff = ctx->buf("vit.xoff"), &d_order = ctx->buf("vit.order"), &d_bpoff = ctx->buf("vit.bpoff"),
           &d_bp = ctx->buf("vit.bp"), &d_res = ctx->buf("vit.res"), &d_pat = ctx->buf("vit.pattern"),
           &d_path = ctx->buf("vit.path"), &d_queue = ctx->buf("vit.queue"), &d_tasks = ctx->buf("vit.tasks");
    TRY(d_xoff.ensure(ctx, (size_t)(n_seq + 1) * 8));
    TRY(d_res.ensure(ctx, (size_t)n_seq * sizeof(VitResult)));
    TRY(d_pat.ensure(ctx, std::max<int64_t>(total, 16)));
    if (path_host) TRY(d_path.ensure(ctx, std::max<int64_t>(total, 16) * 2));
    TRY(d_queue.ensure(ctx, 64));
    TRY(d_bpoff.ensure(ctx, (size_t)n_seq * 8));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_xoff.p, x_off_host, (size_t)(n_seq + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
    auto len = [&](int s) { return x_off_host[s + 1] - x_off_host[s]; };
    // STRIQUE_VITERBI_GENERIC forces the generic kernel (A/B parity tests)
    const bool force_generic = getenv("STRIQUE_VITERBI_GENERIC") != nullptr;
    // STRIQUE_VITERBI_EXACT: float64 kernel only (the fixed-point kernel is the default for models inside its bounds)
    const bool exact_only = getenv("STRIQUE_VITERBI_EXACT") != nullptr || ctx->viterbi_exact;
    // ---- groups: all profile-kernel models together, one group per model otherwise
    struct Group { int model; bool profile; std::vector<int32_t> ids; };
    std::vector<Group> groups;
    for (int s = 0; s < n_seq; ++s) {
        const HmmModel &m = *ctx->models[seq_model[s]];
        const bool profile = m.has_profile && !force_generic;
        Group *g = nullptr;
        for (Group &c : groups)
            if (c.profile == profile && (profile || c.model == seq_model[s])) { g = &c; break; }
        if (!g) { groups.push_back(Group{seq_model[s], profile, {}}); g = &groups.back(); }
        g->ids.push_back(s);
        ctx->last_viterbi_edges += len(s) * m.n_edges;
    }
    size_t free_b = 0, total_b = 0;
    { HostTimer ht("vit cudaMemGetInfo"); CUDA_TRY(ctx, ctx_mem_info(ctx, &free_b, &total_b)); }
    const int64_t budget_bytes = (int64_t)std::max<size_t>((size_t)2 << 30, (size_t)((free_b + d_bp.cap) * 0.7));
    std::vector<int64_t> bpoff(n_seq, 0);
    for (Group &g : groups) {
        std::stable_sort(g.ids.begin(), g.ids.end(), [&](int a, int b) { return len(a) > len(b); });
        const int64_t unit = g.profile ? 4 : 8;                            // bytes per back-pointer word
        const int64_t words_per_step = 32;
        size_t i0 = 0;
        while (i0 < g.ids.size()) {
            size_t i1 = i0;
            int64_t words = 0;
            while (i1 < g.ids.size()) {
                const int64_t wds = (len(g.ids[i1]) + 1) * words_per_step;
                if (i1 > i0 && (words + wds) * unit > budget_bytes) break;
                bpoff[g.ids[i1]] = words;
                words += wds;
                ++i1;
            }
            const int n = (int)(i1 - i0);
            TRY(d_bp.ensure(ctx, (size_t)words * unit));
            TRY(d_order.ensure(ctx, (size_t)n * 4));
            CUDA_TRY(ctx, cudaMemcpyAsync(d_bpoff.p, bpoff.data(), (size_t)n_seq * 8, cudaMemcpyHostToDevice, ctx->stream));
            CUDA_TRY(ctx, cudaMemsetAsync(d_queue.p, 0, 64, ctx->stream));
            if (g.profile) {
                // one warp per sequence.  CTA task = up to warps_per_cta sequences of ONE model that are neighbours
                // in its length order (they finish together, so the CTA barrier between tasks costs nothing); one
                // queue over the tasks of all models (loci / strands), longest first: every hand-out is a whole CTA
                // and no warp ever waits for a sequence of another warp's model.
                // Pass 1: the fixed-point kernel for every model that has a fixed-point image; pass 2: the float64
                // kernel for the other models and for the sequences pass 1 declined (status 3).
                DevBuf &d_pmodels = ctx->buf("vit.prof_models");
                std::vector<VitProfModelDev> pm(n_models);
                for (int i = 0; i < n_models; ++i)
                    if (ctx->models[i]->has_profile) pm[i] = ctx->models[i]->profile; else memset(&pm[i], 0, sizeof(pm[i]));
                TRY(d_pmodels.ensure(ctx, (size_t)n_models * sizeof(VitProfModelDev)));
                CUDA_TRY(ctx, cudaMemcpyAsync(d_pmodels.p, pm.data(), pm.size() * sizeof(VitProfModelDev), cudaMemcpyHostToDevice, ctx->stream));
                DevBuf &d_seqmodel = ctx->buf("vit.seq_model");
                TRY(d_seqmodel.ensure(ctx, (size_t)n_seq * 4));
                CUDA_TRY(ctx, cudaMemcpyAsync(d_seqmodel.p, seq_model, (size_t)n_seq * 4, cudaMemcpyHostToDevice, ctx->stream));
                std::vector<int32_t> order;
                std::vector<VitCtaTask> ctas;
                auto run_pass = [&](const std::vector<int32_t> &ids, bool fixed_point) -> int {   // ids in length order
                    if (ids.empty()) return STRIQUE_OK;
                    VitProfBatch b;
                    b.x = x_dev; b.x_off = d_xoff.as<int64_t>(); b.order = d_order.as<int32_t>();
                    b.n_models = n_models; b.models = d_pmodels.as<VitProfModelDev>();
                    b.counters = d_queue.as<int>(); b.seq_model = d_seqmodel.as<int32_t>();
                    b.bp = d_bp.as<uint32_t>(); b.bp_off = d_bpoff.as<int64_t>(); b.res = d_res.as<VitResult>();
                    b.pattern = d_pat.as<uint8_t>(); b.path = path_host ? d_path.as<uint16_t>() : nullptr;
                    CUDA_TRY(ctx, cudaMemsetAsync(d_queue.p, 0, 64, ctx->stream));
                    if (fixed_point) {
                        // one warp per sequence, one queue over all models, longest first
                        int warps_per_cta = 1;
                        int grid = viterbi_profile_q_max_grid(ctx, &warps_per_cta);
                        if (grid <= 0) FAIL(ctx, STRIQUE_ECUDA, "viterbi_profile_q_kernel: occupancy query failed");
                        grid = std::min<int>(grid, ((int)ids.size() + warps_per_cta - 1) / warps_per_cta);
                        CUDA_TRY(ctx, cudaMemcpyAsync(d_order.p, ids.data(), ids.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
                        b.tasks = nullptr; b.n_tasks = (int)ids.size();
                        return viterbi_profile_q_launch(ctx, b, grid);
                    }
                    std::vector<std::vector<int32_t>> per_model(n_models);
                    for (int32_t s : ids) per_model[seq_model[s]].push_back(s);
                    int warps_per_cta = 1;
                    int grid = viterbi_profile_max_grid(ctx, &warps_per_cta);
                    if (grid <= 0) FAIL(ctx, STRIQUE_ECUDA, "viterbi_profile_kernel: occupancy query failed");
                    struct HostTask { int model; size_t first; int count; int64_t maxlen; };
                    std::vector<HostTask> tasks;
                    for (int mi = 0; mi < n_models; ++mi)
                        for (size_t k = 0; k < per_model[mi].size(); k += warps_per_cta)
                            tasks.push_back(HostTask{mi, k, (int)std::min<size_t>(warps_per_cta, per_model[mi].size() - k),
                                                     len(per_model[mi][k])});
                    std::stable_sort(tasks.begin(), tasks.end(), [](const HostTask &a, const HostTask &b) { return a.maxlen > b.maxlen; });
                    order.clear();
                    ctas.clear();
                    for (const HostTask &t : tasks) {
                        ctas.push_back(VitCtaTask{t.model, (int32_t)order.size(), t.count});
                        order.insert(order.end(), per_model[t.model].begin() + t.first, per_model[t.model].begin() + t.first + t.count);
                    }
                    grid = std::min<int>(grid, (int)ctas.size());
                    TRY(d_tasks.ensure(ctx, ctas.size() * sizeof(VitCtaTask)));
                    CUDA_TRY(ctx, cudaMemcpyAsync(d_tasks.p, ctas.data(), ctas.size() * sizeof(VitCtaTask), cudaMemcpyHostToDevice, ctx->stream));
                    CUDA_TRY(ctx, cudaMemcpyAsync(d_order.p, order.data(), order.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
                    b.tasks = d_tasks.as<VitCtaTask>(); b.n_tasks = (int)ctas.size();
                    return viterbi_profile_launch(ctx, b, grid);
                };
                std::vector<int32_t> fixed_ids, exact_ids;
                for (size_t i = i0; i < i1; ++i) {
                    const int32_t s = g.ids[i];
                    (ctx->models[seq_model[s]]->profile.qgrp && !exact_only ? fixed_ids : exact_ids).push_back(s);
                }
                if (!fixed_ids.empty()) {
                    TRY(run_pass(fixed_ids, true));
                    // the declined ones: read the status words back (40 bytes per sequence)
                    std::vector<VitResult> &tmp = ctx->vit_res_scratch;
                    tmp.resize(n_seq);
                    CUDA_TRY(ctx, cudaMemcpyAsync(tmp.data(), d_res.p, (size_t)n_seq * sizeof(VitResult), cudaMemcpyDeviceToHost, ctx->stream));
                    { HostTimer ht("vit fixed-point kernel sync"); CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); }
                    // sequences the fixed-point pass could not vouch for were decoded in float64 inside the same launch
                    // (reserved = 1); status 3 would be a sequence left undecided -- the float64 kernel takes it
                    size_t declined = 0, undecided = 0;
                    for (int32_t s : fixed_ids) {
                        if (tmp[s].status == 3) { exact_ids.push_back(s); ++undecided; }
                        else if (tmp[s].reserved == 1) ++declined;
                    }
                    ctx->last_viterbi_fixed += (int64_t)(fixed_ids.size() - declined - undecided);
                    ctx->last_viterbi_declined += (int64_t)(declined + undecided);
                    if (undecided) std::stable_sort(exact_ids.begin(), exact_ids.end(), [&](int a, int b) { return len(a) > len(b); });
                }
                if (!exact_ids.empty()) {
                    TRY(run_pass(exact_ids, false));
                    { HostTimer ht("vit profile kernel sync"); CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream)); }   // host vectors are read by the async copies
                }
            } else {
                CUDA_TRY(ctx, cudaMemcpyAsync(d_order.p, g.ids.data() + i0, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
                VitBatch b;
                b.x = x_dev; b.x_off = d_xoff.as<int64_t>(); b.n_seq = n; b.order = d_order.as<int32_t>();
                b.bp = d_bp.as<unsigned long long>(); b.bp_off = d_bpoff.as<int64_t>(); b.res = d_res.as<VitResult>();
                b.pattern = d_pat.as<uint8_t>(); b.path = path_host ? d_path.as<uint16_t>() : nullptr;
                b.queue = d_queue.as<int>();
                if (!force_generic && viterbi_small_fits(ctx->models[g.model]->dev))
                    TRY(viterbi_small_launch(ctx, *ctx->models[g.model], b));
                else
                    TRY(viterbi_launch(ctx, *ctx->models[g.model], b));
            }
            // host vectors above are read by the async copies: drain before they go out of scope
            CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
            i0 = i1;
        }
    }
    HostTimer ht_res("vit results d2h + sync");
    CUDA_TRY(ctx, cudaMemcpyAsync(results_host, d_res.p, (size_t)n_seq * sizeof(VitResult), cudaMemcpyDeviceToHost, ctx->stream));
    if (pattern_host) CUDA_TRY(ctx, cudaMemcpyAsync(pattern_host, d_pat.p, total, cudaMemcpyDeviceToHost, ctx->stream));
    if (path_host) CUDA_TRY(ctx, cudaMemcpyAsync(path_host, d_path.p, total * 2, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return STRIQUE_OK;
}

}  // namespace strique

using namespace strique;

extern "C" int strique_hmm_create(strique_ctx *ctx, const strique_hmm_desc *desc, int32_t *model_id) {
    if (!ctx || !desc || !model_id) return STRIQUE_EINVAL;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    HmmModel *m = new HmmModel();
    const int rc = hmm_create(ctx, desc, m);
    if (rc != STRIQUE_OK) { delete m; return rc; }
    ctx->models.push_back(m);
    *model_id = (int32_t)ctx->models.size() - 1;
    return STRIQUE_OK;
}

extern "C" int strique_hmm_kernel_shape(const strique_ctx *ctx, int32_t model_id) {
    if (!ctx || model_id < 0 || model_id >= (int)ctx->models.size()) return -1;
    if (ctx->models[model_id]->has_profile) return 4000;
    if (viterbi_small_fits(ctx->models[model_id]->dev)) return 32;
    return 0;
}

extern "C" int strique_viterbi_batch(strique_ctx *ctx, int32_t model_id, int n_seq, const double *x,
                                     const int64_t *x_offsets, int memspace, strique_viterbi_result *results,
                                     uint8_t *pattern_out, uint16_t *path_out) {
    if (!ctx) return STRIQUE_EINVAL;
    if (model_id < 0 || model_id >= (int)ctx->models.size()) FAIL(ctx, STRIQUE_EINVAL, "unknown HMM id");
    if (n_seq < 0 || (n_seq > 0 && (!x || !x_offsets || !results))) FAIL(ctx, STRIQUE_EINVAL, "strique_viterbi_batch: bad argument");
    if (n_seq == 0) return STRIQUE_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const double *xd = x;
    if (memspace != STRIQUE_DEVICE) {
        DevBuf &d_x = ctx->buf("vb.x");
        TRY(d_x.ensure(ctx, std::max<int64_t>(1, x_offsets[n_seq]) * 8));
        CUDA_TRY(ctx, cudaMemcpyAsync(d_x.p, x, (size_t)x_offsets[n_seq] * 8, cudaMemcpyHostToDevice, ctx->stream));
        xd = d_x.as<double>();
    }
    ctx->last_viterbi_edges = 0;
    ctx->last_viterbi_fixed = ctx->last_viterbi_declined = 0;
    std::vector<int32_t> seq_model(n_seq, model_id);
    return viterbi_run_device_multi(ctx, seq_model.data(), xd, x_offsets, n_seq, results, pattern_out, path_out);
}
