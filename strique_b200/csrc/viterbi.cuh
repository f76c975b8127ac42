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

namespace strique {

constexpr int VIT_WARPS = 8;          // warps (= sequences in flight) per CTA
constexpr int VIT_MAX_SLOTS = 12;     // emitting slots + chain slots per lane (4 bits each, <= 48 bits)

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

struct HmmModel {                 // host-side handle; device arrays owned by the context
    VitModelDev dev;
    int64_t n_edges = 0;          // in-edges of emitting + chain states (work unit of the Viterbi stage)
};

typedef strique_viterbi_result VitResult;

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
int hmm_create(strique_ctx *ctx, const strique_hmm_desc *d, HmmModel *out);
// decodes n_seq sequences whose samples are device resident (x_dev) -> host results
int viterbi_run_device(strique_ctx *ctx, const HmmModel &m, const double *x_dev, const int64_t *x_off_host, int n_seq,
                       strique_viterbi_result *results_host, uint8_t *pattern_host, uint16_t *path_host);

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
