/*
 * strique_b200 -- C ABI of the B200-native replacement for the native hot path of
 * giesselmann/STRique's per-read repeat detection (`STRique.py count` -> repeatCounter.detect,
 * scripts/STRique.py:581-618).
 *
 * Every entry point below is `extern "C"`, takes plain pointers and sizes (no torch / pybind
 * types) and returns 0 on success or a negative STRIQUE_E* code; `strique_last_error()` holds
 * the message.  "memspace" arguments say whether the data pointers are HOST (0) or DEVICE (1)
 * pointers; small descriptor arrays (offsets, task lists) are always host pointers.
 *
 * Reference interfaces replaced (reference file:line):
 *   strique_align_batch      <- pyseqan.align_raw.align_overlap  (src/pyalign.cpp:47-62,
 *                               src/align_raw.h:106-158, src/score_distance.h:115-122) plus the
 *                               nearest-view-position reduction of repeatCounter.__detect_range__
 *                               (scripts/STRique.py:538-548)
 *   strique_condition_batch  <- the conditioning lines of repeatCounter.detect
 *                               (scripts/STRique.py:590-597) and pore_model.normalize2model
 *                               'minmax' (scripts/STRique.py:151-160,178-179), MAD (142-143)
 *   strique_viterbi_batch    <- pomegranate 0.10.0 HiddenMarkovModel.viterbi as used by
 *                               flankedRepeatHMM.count_repeats / repeatModHMM.mod_repeats
 *                               (scripts/STRique.py:433-441, 492-500, 374-378)
 *   strique_detect_batch     <- repeatCounter.detect (scripts/STRique.py:581-618), batched
 */
#ifndef STRIQUE_B200_H
#define STRIQUE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define STRIQUE_OK 0
#define STRIQUE_EINVAL (-1)   /* bad argument */
#define STRIQUE_ECUDA (-2)    /* CUDA runtime error (no device, launch failure, ...) */
#define STRIQUE_ENOMEM (-3)   /* device or host allocation failed */
#define STRIQUE_EUNSUPPORTED (-4)

#define STRIQUE_HOST 0
#define STRIQUE_DEVICE 1

typedef struct strique_ctx strique_ctx;

/* ---- context -------------------------------------------------------------------------------- */
int strique_ctx_create(int device, strique_ctx **out);
void strique_ctx_destroy(strique_ctx *ctx);
const char *strique_last_error(const strique_ctx *ctx); /* ctx may be NULL: last create error */
int strique_version(void);
/* number of kernel launches issued through this context since creation */
int64_t strique_launch_count(const strique_ctx *ctx);
/* the CUDA stream (cudaStream_t) all work of this context is enqueued on */
void *strique_ctx_stream(const strique_ctx *ctx);
int strique_ctx_synchronize(strique_ctx *ctx);

/* ---- boundary #1: semi-global flank alignment ------------------------------------------------ */
/* the six read/write properties of pyseqan.align_raw (src/pyalign.cpp:50-57) */
typedef struct {
    float gap_open_h, gap_open_v, gap_extension_h, gap_extension_v, dist_offset, dist_min;
} strique_align_params;

typedef struct {
    float score;          /* best last-row score (fp32, bit-exact vs the reference)            */
    int32_t best_j;       /* DP column (1-based signal position) where the alignment ends      */
    int32_t begin0;       /* argmin_p |a_idx[p] - b_idx[0]|            (S.py:540)              */
    int32_t end0;         /* argmin_p |a_idx[p] - b_idx[L-1]|          (S.py:541)              */
    int32_t begin_trim;   /* argmin_p |a_idx[p] - b_idx[pre_trim]|     (S.py:546)              */
    int32_t end_trim;     /* argmin_p |a_idx[p] - b_idx[L-1-post_trim]| (S.py:547)             */
    int32_t n_blocks;     /* trace blocks recomputed for the traceback (diagnostic)            */
    int32_t status;       /* 0 ok                                                              */
} strique_align_result;

/*
 * Signals are passed as CODES into a per-signal table of distinct fp32 sample values (the read
 * signal fed to the reference aligner has <= 256 distinct values), flanks as
 * LEVELS each standing for `samples` identical consecutive flank samples
 * (pore_model.generate_signal(seq, samples), scripts/STRique.py:185-186).
 *
 *   codes        : code_bytes (1 or 2) per sample, all signals concatenated
 *   sig_offsets  : [n_signals+1] sample offsets into codes                       (host)
 *   code_values  : [n_signals * n_code_values] fp32 value of each code
 *   flank_levels : all flanks' levels concatenated (fp32)
 *   flank_offsets: [n_flanks+1] level offsets                                    (host)
 *   task_*       : [n_tasks] signal index, flank index, pre/post trim in flank samples (host)
 *   results      : [n_tasks]                                                     (host)
 *   rows_out     : optional [n_tasks * rows_stride] (host): for flank sample q of task t,
 *                  (j << 1) | is_vertical_gap, j = signal samples consumed up to and including
 *                  that flank sample -- enough to rebuild align_overlap's a_idx / b_idx
 */
int strique_align_batch(strique_ctx *ctx, const strique_align_params *params,
                        int n_signals, const void *codes, int code_bytes, const int64_t *sig_offsets,
                        const float *code_values, int n_code_values,
                        int n_flanks, const float *flank_levels, const int32_t *flank_offsets, int samples,
                        int n_tasks, const int32_t *task_signal, const int32_t *task_flank,
                        const int32_t *task_pre_trim, const int32_t *task_post_trim,
                        int memspace, strique_align_result *results,
                        int32_t *rows_out, int64_t rows_stride);

/* DP cells (signal samples x flank samples, summed over tasks) processed by the last
 * strique_align_batch / strique_detect_batch call, and device time of its scan kernel in ms */
int64_t strique_last_align_cells(const strique_ctx *ctx);
float strique_last_scan_ms(const strique_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* STRIQUE_B200_H */
