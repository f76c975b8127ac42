// Batched semi-global flank alignment for sm_100a.
//
// Replaces `align_raw<float,float>::semiglobal` (reference src/align_raw.h:106-158: SeqAn 2.4
// globalAlignment, AlignConfig<true,false,false,true>, AffineGaps, Score<float,Distance> from
// src/score_distance.h:115-122) and the index reduction of repeatCounter.__detect_range__
// (scripts/STRique.py:538-548).  Results are bit-identical to the reference: same fp32 recursion
// (adds / max / compares only), SeqAn's denormal "infinity", its tie rules and its traceback.
//
// Design (see DESIGN.md "Alignment"):
//   * one warp per alignment task; lane l owns R = K*S consecutive flank rows (K levels of S
//     identical samples) and keeps their S and H column values in registers;
//   * columns flow through the warp as a systolic wavefront: at step s lane l processes signal
//     column j = s - l and hands (S, V) of its bottom row to lane l+1 with two warp shuffles;
//   * the substitution score max(off - |h - v|^1.2, min) is an exact per-task table
//     lut[code][level], built once per task in fp64 (align_build_lut_kernel);
//   * pass 1 (align_scan_kernel) computes scores only -- 5 FADD + 4 FMNMX per cell, no trace --
//     records the best last-row cell and checkpoints the DP column every CKPT columns;
//   * pass 2 (align_trace_kernel) recomputes only the CKPT-column blocks the optimal path runs
//     through, now with SeqAn's trace flags packed 4 bits/cell, and walks SeqAn's traceback.
#pragma once
#include "common.cuh"

namespace strique {

#ifndef STRIQUE_ALIGN_CKPT
#define STRIQUE_ALIGN_CKPT 512
#endif
constexpr int ALIGN_CKPT = STRIQUE_ALIGN_CKPT;   // columns between DP-column checkpoints (power of two)
constexpr int ALIGN_WARPS_PER_SM = 8;    // resident single-warp CTAs per SM for the scan
constexpr int ALIGN_WARPS_PER_SM_LINEAR = 16;   // ... for the linear-gap scan (half the registers)
// ... for the trace pass (rows = K * S score and H registers each): the recomputation is issue bound with plenty of
// fixed-latency waits, so resident warps beat registers -- K = 5, 8192 C2 reads: 12 warps per SM (168 registers)
// 18.5 ms, 14 (128 registers, 80 B of spills) 16.4, 16: 15.7, 20 (96 registers): 18.3.
#ifdef ALIGN_TRACE_WARPS
constexpr int align_trace_warps(int) { return ALIGN_TRACE_WARPS; }
#else
constexpr int align_trace_warps(int rows) { return rows <= 32 ? 16 : (rows <= 42 ? 12 : 8); }
#endif
constexpr int ALIGN_TRACE_WARPS_MAX = 16;
// ... for the two-flanks-per-warp scan (LinSweepPair): as many as the registers allow (12 K score registers + 4 K table
// values + ~30 per thread).  Measured for K = 5 on 8192 C2 reads: 8 warps 101.0 ms, 10: 94.8, 12: 88.5, 14: 90.2,
// 16: 87.1, 18 (96 registers, spills): 104.7.
#ifdef ALIGN_PAIR_WARPS
constexpr int align_pair_warps(int) { return ALIGN_PAIR_WARPS; }
#else
constexpr int align_pair_warps(int K) { return K <= 5 ? 16 : (K == 6 ? 12 : (K <= 8 ? 10 : 8)); }
#endif
#ifndef ALIGN_PACKED_WARPS
#define ALIGN_PACKED_WARPS 16
#endif
constexpr int ALIGN_WARPS_PER_SM_PACKED = ALIGN_PACKED_WARPS;   // ... for the packed linear-gap scan (LinSweep2)

struct AlignGroup {        // tasks sharing one (K, S) kernel instantiation
    int K, S;
    int n_tasks;
    const int32_t *order;  // [n_tasks] task ids, longest signal first (device)
    int n_pairs;           // scan only (linear gap costs): the first n_pairs entries of pair_order are tasks t such that
    const int32_t *pair_order;   // t and t + 1 are two flanks over the SAME signal: one warp scans both (LinSweepPair);
    int n_single;                // the remaining tasks follow in single_order
    const int32_t *single_order;
};

// Device-side batch description (all pointers are device pointers).
struct AlignBatch {
    strique_align_params p;
    const uint16_t *codes;        // all signals, concatenated
    const int64_t *sig_off;       // [n_signals + 1]
    const float *code_values;     // [n_signals * n_code_values]
    int n_code_values;
    const float *flank_levels;    // all flanks' levels, concatenated
    const int32_t *flank_off;     // [n_flanks + 1] (levels)
    const float *col0;            // [n_flanks * col0_stride] S of DP column 0, index = row i
    int col0_stride;
    int samples;                  // flank samples per level as given by the caller
    const int32_t *task_sig, *task_flank, *task_pre, *task_post;   // [n_tasks]
    float *lut;                   // [n_tasks][n_code_values][lut_row]; inside a row level u = lane * K + k sits at [k][lane]
    int64_t lut_task_stride;      // floats
    int lut_row;                  // floats per code row (32 * Kmax of the batch)
    float *ckpt;                  // checkpoints, see ckpt_off
    const int64_t *ckpt_off;      // [n_tasks] float offset of the task's checkpoint area
    const int32_t *ckpt_step;     // [n_tasks] 1: the area is the task's own; 2: tasks t, t + 1 of a scanned pair share their
                                  //   two areas, values interleaved (row i of the second task at [2 i + 1]; its offset is
                                  //   the first task's + 1)
    int ckpt_rows;                // padded rows per checkpoint column (per S or H plane)
    uint32_t *trace;              // [n_warps_trace][32][ALIGN_CKPT][W] flag bits: planes gap | maxh | hopen | vopen
    int32_t *rows;                // [n_tasks * rows_stride]
    int64_t rows_stride;
    strique_align_result *res;    // [n_tasks]
    int *queue;                   // work-queue counters: [0] scan, [1] trace
    unsigned long long *lut_fix;  // [0] = count, then (index) entries flagged for host re-evaluation
    int lut_fix_cap;
};

// Device-resident inputs of the alignment stage (+ host copies of the small offset arrays).
struct AlignDeviceInputs {
    int n_signals;
    const uint16_t *codes;
    const int64_t *sig_off;        // device
    const int64_t *sig_off_host;   // host
    const float *code_values;
    int n_code_values;
    int n_flanks;
    const float *flank_levels;
    const int32_t *flank_off;      // device
    const int32_t *flank_off_host; // host
    int samples;
};

int align_run_device(strique_ctx *ctx, const strique_align_params &params, const AlignDeviceInputs &in, int n_tasks,
                     const int32_t *task_signal, const int32_t *task_flank, const int32_t *task_pre,
                     const int32_t *task_post, strique_align_result *results_host, int32_t *rows_out_host,
                     int64_t rows_out_stride, strique_align_result *results_dev_out);

int align_launch_patch_lut(strique_ctx *ctx, float *lut, const unsigned long long *patch, int n);
int align_launch_build_lut(strique_ctx *ctx, const AlignBatch &b, const int32_t *task_K, const int32_t *task_S,
                           int n_tasks);
int align_launch_scan(strique_ctx *ctx, const AlignBatch &b, const AlignGroup &g);
int align_launch_trace(strique_ctx *ctx, const AlignBatch &b, const AlignGroup &g, int n_warps);
bool align_pick_kernel(int nlev, int samples, int *K, int *S);
bool align_params_linear(const strique_align_params &p);

}  // namespace strique
