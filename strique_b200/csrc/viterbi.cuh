// Batched Viterbi decoding of STRique's profile / repeat HMMs for sm_100a.
//
// Replaces pomegranate 0.10.0 `HiddenMarkovModel.viterbi` as used by
// flankedRepeatHMM.count_repeats and repeatModHMM.mod_repeats (reference
// scripts/STRique.py:433-441, 492-500, 374-378).  float64 like the reference.
//
// Model form ("compiled HMM", built on the host by strique_b200/hmm.py):
//   * E emitting states with sparse in-edges from the previous time step; sources are emitting
//     states, silent *chain* states or the START pseudo state;
//   * C silent chain states (the delete states D[i] of a profile HMM): value at time t is
//     max(entry edges from emitting states at t, previous chain state + w) -- a max-plus scan;
//   * every other silent state of the reference graph (s1/s2/e1/e2 glue) has been composed away
//     on the host, which does not change any path score (their edges carry log 1 = 0);
//   * END edges evaluated once after the last sample.
// One warp decodes one sequence: state l lives on lane l%32, values of the current and previous
// column in shared memory, back-pointers packed 4 bits per state per time step in HBM
// (one 8/16-byte word per lane per step, coalesced), traceback by the same warp.
#pragma once
#include "common.cuh"
#include "profile_core.h"

namespace strique {

constexpr int VIT_WARPS = 8;          // warps (= sequences in flight) per CTA
constexpr int VIT_MAX_SLOTS = 12;     // emitting slots + chain slots per lane (4 bits each, <= 48 bits)

typedef strique_viterbi_result VitResult;

enum { HMM_FLAG_COUNT = 1, HMM_FLAG_REPEAT = 2, HMM_FLAG_SEP = 4, HMM_FLAG_MOD = 8 };

// Device view of a compiled model: a packed image (copied to shared memory by every CTA) plus
// the offsets of its sections.  Emitting state at device position p lives on lane p%32, slot p/32;
// value-array positions: [0, NS*32) emitting, [NS*32, (NS+QC)*32) chain, then START, then NEG.
struct VitModelDev {
    int E, C, NS, QC, n_end, rows;
    int deg[VIT_MAX_SLOTS];        // in-edges evaluated per emitting slot
    int row_base[VIT_MAX_SLOTS];   // first edge row of the slot
    const unsigned char *blob;
    int blob_bytes;
    int off_edge_w, off_edge_src, off_em_kind, off_em_p, off_flags, off_chain_predw, off_chain_ew, off_chain_es,
        off_end_src, off_end_w;
    const int32_t *perm;           // device position -> caller's emitting state id
};

// Device view of a model packed for the profile kernel (viterbi_profile.cu, profile_pack.h): the per-lane
// constant table, per (position, slot) emission / flag tables for the slow emission path and the
// traceback, the long-range (repeat loop) sources and the END edges.
struct VitProfModelDev {
    const double *tab;             // [pf::K_TOTAL][32]
    const uint8_t *em_kind;        // [pf::NPOS * 2], index position * 2 + slot
    const double *em_a, *em_b, *em_c;
    const uint8_t *flags;
    const int32_t *state_id;       // caller's emitting state id
    pf::TraceCfg trace;
    double lo, hi;                 // samples inside [lo, hi] take the fast emission path
    int p_start;                   // START = match slot of this position, value 0 in column 0
    int n_end;
    int32_t end_p[16], end_slot[16];
    double end_w[16];
    // fixed-point image (profile_q.h) when the model fits its bounds, else nullptr: the model then stays on the
    // float64 kernel
    const int32_t *qgrp;           // [pq::G_TOTAL][32][4]
    const double *qem;             // [pq::E_TOTAL][32][2]
    // its traceback re-scores the path in float64 and re-adds it in fixed point: one record per (position, slot in
    // {M, I}), index position * 2 + slot: {a, b, c, self-loop weight} with emission = b - (x - a)^2 c (Uniform / unused
    // slot: a = 0, c = 0), then the constants of the forward pass's own emission {A, c, C, 0} (profile_q.h);
    // tmeta: flags | caller's state id << 16; tq: {quantised self-loop weight, quantised I-slot emission}; qtab: the
    // quantised weights without their tags, indexed like `tab` (INT32_MIN: the model has no such edge)
    const double *trec;            // [pf::NPOS * 2][8]
    const uint32_t *tmeta;         // [pf::NPOS * 2]
    const int32_t *tq;             // [pf::NPOS * 2][2]
    const int32_t *qtab;           // [pf::K_TOTAL][32]
};

struct HmmModel {                 // host-side handle; device arrays owned by the context
    VitModelDev dev;
    int64_t n_edges = 0;          // in-edges of emitting + chain states (work unit of the Viterbi stage)
    bool has_profile = false;
    VitProfModelDev profile;
};

struct VitCtaTask {                // a few sequences of one model, consecutive in `order`
    int32_t model, first, count;
};

struct VitProfBatch {
    const double *x;
    const int64_t *x_off;
    const int32_t *order;          // sequence ids, the sequences of a CTA task are consecutive
    int n_models;
    const VitProfModelDev *models; // device array indexed by model
    const VitCtaTask *tasks;       // [n_tasks] <= 4 sequences of one model with similar lengths; longest task first
    int n_tasks;                   // (fixed-point kernel: one task = one sequence of `order`, its model in seq_model)
    const int32_t *seq_model;      // [all sequences] model of every sequence
    int *counters;                 // [0]: CTA tasks handed out so far (zeroed by the caller)
    uint32_t *bp;
    const int64_t *bp_off;         // [all sequences] offset in 32-bit words (per sequence: (T+1) * 32 words)
    VitResult *res;
    uint8_t *pattern;
    uint16_t *path;
};

struct VitBatch {
    const double *x;            // normalised samples, all sequences concatenated
    const int64_t *x_off;       // [n_seq + 1]
    int n_seq;
    const int32_t *order;       // [n_seq] longest first
    unsigned long long *bp;     // back-pointer area
    const int64_t *bp_off;      // [n_seq] offset in 64-bit words (per sequence: (T+1) * 32 words)
    VitResult *res;             // [n_seq]
    uint8_t *pattern;           // [sum T] pattern chars, written right-aligned per sequence (x_off)
    uint16_t *path;             // optional [sum T] emitting state per sample (caller's ids)
    int *queue;
};

int viterbi_launch(strique_ctx *ctx, const HmmModel &m, const VitBatch &b);
// small-model kernel (viterbi_small.cu): <= 32 emitting states, no chain, <= 8 in-edges per state
bool viterbi_small_fits(const VitModelDev &m);
int viterbi_small_launch(strique_ctx *ctx, const HmmModel &m, const VitBatch &b);
// profile kernel: packs the model if its layout hints describe a linear profile (sets m->has_profile), launch
int viterbi_profile_pack(strique_ctx *ctx, const strique_hmm_desc *d, HmmModel *m);
int viterbi_profile_max_grid(strique_ctx *ctx, int *warps_per_cta);
int viterbi_profile_launch(strique_ctx *ctx, const VitProfBatch &b, int grid);
// fixed-point profile kernel (viterbi_profile_q.cu): same batch layout; sequences it declines come back with status 3
struct ProfileImage;
int viterbi_profile_q_pack(strique_ctx *ctx, const ProfileImage &img, VitProfModelDev *f);
int viterbi_profile_q_max_grid(strique_ctx *ctx, int *warps_per_cta);
int viterbi_profile_q_launch(strique_ctx *ctx, const VitProfBatch &b, int grid);
// decodes sequences of several models in one pass: seq_model[s] indexes ctx->models
int viterbi_run_device_multi(strique_ctx *ctx, const int32_t *seq_model, const double *x_dev, const int64_t *x_off_host,
                             int n_seq, strique_viterbi_result *results_host, uint8_t *pattern_host,
                             uint16_t *path_host);
int hmm_create(strique_ctx *ctx, const strique_hmm_desc *d, HmmModel *out);

// x[t] = clip(clip(((src[t] - c1) / c2) * c3 + c4, lo, hi), lo2, hi2) for each segment
struct PrepSeg {
    int64_t src_off;   // element offset into the source signal array
    int64_t dst_off;   // element offset into x
    int32_t len;
    int32_t read;      // index into the stats array
};
int viterbi_prepare_x(strique_ctx *ctx, int raw_kind, const void *src, const PrepSeg *segs_dev, int n_segs,
                      const double *stats, int stat_c1, double c3, double c4, double lo, double hi, double lo2,
                      double hi2, double *x_out);

}  // namespace strique
