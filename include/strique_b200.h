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
#define STRIQUE_ENOSPC (-5)   /* an output buffer is too small; the call says which and how many bytes it needs */

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

/* ---- boundary #2: Viterbi decoding of a compiled HMM ----------------------------------------- */
/*
 * A "compiled" HMM (strique_b200/hmm.py builds it from the reference's topology,
 * scripts/STRique.py:201-500): n_emit emitting states, n_chain silent chain states (profile-HMM
 * delete states), START = index n_emit + n_chain.  All other silent states have been composed away.
 *   in_*        : CSR in-edges of the emitting states; sources index [emitting | chain | START] and
 *                 are read from the PREVIOUS time step
 *   emit_kind   : 0 Normal(a = mean, b = std), 1 Uniform(a = lo, b = hi)
 *   emit_flags  : STRIQUE_HMM_COUNT | _REPEAT | _SEP | _MOD
 *   chain_pred_logw[c]: log weight of chain state c-1 -> c (-inf if c starts a chain)
 *   chain_in_*  : CSR entry edges of the chain states (<= 3 each); sources are emitting states or
 *                 START, read from the SAME time step
 *   end_*       : edges into the END state, evaluated after the last sample
 */
#define STRIQUE_HMM_COUNT 1   /* visits are counted (repeatHMM d1/d2, S.py:374-378)              */
#define STRIQUE_HMM_REPEAT 2  /* state name contains 'repeat' (S.py:608)                           */
#define STRIQUE_HMM_SEP 4     /* s0 / e0 of repeatModHMM: separates repeat passes (S.py:496)       */
#define STRIQUE_HMM_MOD 8     /* state name contains 'mod' (S.py:497)                              */

typedef struct {
    int32_t n_emit, n_chain;
    const int32_t *in_ptr;
    const int32_t *in_src;
    const double *in_logw;
    const int32_t *emit_kind;
    const double *emit_a;
    const double *emit_b;
    const uint8_t *emit_flags;
    const double *chain_pred_logw;
    const int32_t *chain_in_ptr;
    const int32_t *chain_in_src;
    const double *chain_in_logw;
    int32_t n_end;
    const int32_t *end_src;
    const double *end_logw;
    /* Optional layout hints (all three NULL: none).  When the model is a linear profile -- every state
     * belongs to a position of a chain and every edge stays within two positions, except one loop-back
     * edge pair (repeatHMM d1 / d2, S.py:339-346) -- the profile kernel (one warp per sequence, 4 positions
     * per lane, neighbours in registers) serves it; otherwise the hints are ignored.
     *   emit_pos[l]  : position of emitting state l;  emit_slot[l]: 0 match-like, 1 insert-like
     *   chain_pos[c] : position of chain state c */
    const int32_t *emit_pos;
    const uint8_t *emit_slot;
    const int32_t *chain_pos;
} strique_hmm_desc;

typedef struct {
    double logp;          /* Viterbi log probability (float64 like pomegranate)                    */
    int32_t n_count;      /* visits of COUNT states on the best path                               */
    int32_t t_first;      /* first / last sample decoded by a REPEAT state (-1: none)              */
    int32_t t_last;
    int32_t pattern_len;  /* number of repeat passes found between SEP states                      */
    int32_t status;       /* 0 ok, 1 impossible sequence (log p = -inf), 2 internal error
                           * (3 is internal: declined by the fixed-point kernel, never returned)   */
    int32_t reserved;
} strique_viterbi_result;

int strique_hmm_create(strique_ctx *ctx, const strique_hmm_desc *desc, int32_t *model_id);
/* which Viterbi kernel serves the model: 0 = generic warp-per-sequence kernel, 4000 = profile kernels
 * (4 positions per lane; fixed point with float64 fallback), 32 = small-model kernel (one state per lane, values in
 * registers) (diagnostic) */
int strique_hmm_kernel_shape(const strique_ctx *ctx, int32_t model_id);
/*
 *   x, x_offsets : float64 samples of all sequences concatenated; [n_seq+1] offsets (host)
 *   pattern_out  : optional, same layout as x: per sequence the '0'/'1' pattern is written
 *                  RIGHT-aligned in its slot (last pattern_len bytes)
 *   path_out     : optional, same layout as x: emitting state id decoding each sample
 */
int strique_viterbi_batch(strique_ctx *ctx, int32_t model_id, int n_seq, const double *x, const int64_t *x_offsets,
                          int memspace, strique_viterbi_result *results, uint8_t *pattern_out, uint16_t *path_out);

/* work of the last viterbi / detect call: (time steps) x (in-edges of the model), summed */
int64_t strique_last_viterbi_edges(const strique_ctx *ctx);
/*
 * Linear profile models (every count model of the reference) are decoded by a fixed-point kernel: tagged int32
 * scores at 2^-16 nat, log p = float64 re-score of the decoded path (csrc/profile_q.h).  Sequences it cannot
 * vouch for are decoded by the float64 kernel (pomegranate's arithmetic, scripts/STRique.py:433-441) in the same
 * call.  strique_set_viterbi_exact(ctx, 1) sends everything to the float64 kernel.  The two counters report the
 * last viterbi / detect call: sequences decoded in fixed point, and sequences handed on to float64.
 */
int strique_set_viterbi_exact(strique_ctx *ctx, int exact);
int64_t strique_last_viterbi_fixed(const strique_ctx *ctx);
int64_t strique_last_viterbi_declined(const strique_ctx *ctx);

/* ---- conditioning ---------------------------------------------------------------------------- */
typedef struct {
    double m5_mod, m95_mod;       /* medians of the model k-mer means below the 1st / above the 99th percentile */
    double model_min, model_max;  /* pore_model.model_min / model_max (scripts/STRique.py:126-127)             */
} strique_pore_constants;

typedef struct {
    double flt_median, flt_mad;   /* median and mean-absolute-deviation of the median-filtered read */
    double flt_c1, flt_c2;        /* 'minmax' centre and half-range of the filtered read            */
    double raw_c1, raw_c2;        /* same for the raw read (0,1 unless want_raw_stats)              */
    double u8_c1, u8_c2;          /* same for the uint8 morphology signal                           */
    double status;                /* 0 ok, 1 degenerate read (zero MAD or empty percentile tail)    */
    double reserved[3];
} strique_condition_stats;

/*
 * raw_kind 0: int16 samples (fast5 DAC values), 1: float64 samples.  Outputs (host, all optional):
 *   flt_out    : median-filtered signal, same type and layout as raw
 *   codes_out  : uint16 per sample, value 0..255: closing(opening(uint8 z-score))
 *   values_out : [n_reads*256] fp32 value of each code after 'minmax' normalisation to the model
 *   stats_out  : [n_reads]
 */
int strique_condition_batch(strique_ctx *ctx, const strique_pore_constants *pore, int n_reads, const void *raw,
                            int raw_kind, const int64_t *raw_offsets, int want_raw_stats, void *flt_out,
                            uint16_t *codes_out, float *values_out, strique_condition_stats *stats_out);

/* ---- the whole per-read path: repeatCounter.detect, batched ---------------------------------- */
typedef struct {
    const float *prefix_levels;   /* k-mer means of prefix_ext (generate_signal / samples), fp32 */
    int32_t n_prefix_levels;
    const float *suffix_levels;
    int32_t n_suffix_levels;
    int32_t pre_trim;             /* len(prefix_ext) - len(prefix) in samples (S.py:598) */
    int32_t post_trim;            /* len(suffix_ext) - len(suffix) in samples (S.py:599) */
    int32_t count_model;          /* id from strique_hmm_create: flankedRepeatHMM          */
    int32_t mod_model;            /* id of the repeatModHMM or -1                          */
    int32_t count_offset;         /* flanking_count - repeat_offset (S.py:378, 437)        */
} strique_target_desc;

int strique_target_create(strique_ctx *ctx, const strique_target_desc *desc, int32_t *target_id);

typedef struct {
    strique_align_params align;
    int32_t samples;              /* flank samples per k-mer (align config 'samples', S.py:513) */
    int32_t use_mod;              /* run the methylation HMM (reference: --mod_model given)      */
    strique_pore_constants pore;  /* base pore model                                             */
    double mod_clip_lo, mod_clip_hi;  /* min / max of both models' model_min / model_max (S.py:467-468) */
} strique_detect_config;

typedef struct {
    double score_prefix, score_suffix;   /* fp32 alignment score / (end - begin), or 0.0 (S.py:542-545) */
    double log_p;                        /* 0 when the HMM stage did not run                             */
    int32_t count;                       /* repeat count n                                               */
    int32_t offset, ticks;               /* prefix_end, max(suffix_begin - prefix_end, 0)                */
    int32_t prefix_begin, prefix_end, suffix_begin, suffix_end;
    int32_t hmm_ran;                     /* 1 if the count HMM decoded this read                         */
    int32_t mod_len;                     /* length of the methylation pattern; -1 means '-'              */
    int32_t status;                      /* 0 ok, 1 degenerate read                                      */
    int64_t mod_off;                     /* offset of the pattern in mod_out                             */
} strique_detect_result;

/*
 * raw / raw_offsets as in strique_condition_batch; read_target[r] = id from strique_target_create.
 * mod_out (host, capacity mod_cap bytes) receives the '0'/'1' patterns back to back.  When it is too small the call
 * returns STRIQUE_ENOSPC and strique_last_mod_bytes() the size that is needed (one pattern character per repeat pass).
 */
int strique_detect_batch(strique_ctx *ctx, const strique_detect_config *cfg, int n_reads, const void *raw,
                         int raw_kind, const int64_t *raw_offsets, const int32_t *read_target, int memspace,
                         strique_detect_result *results, uint8_t *mod_out, int64_t mod_cap);

int64_t strique_last_mod_bytes(const strique_ctx *ctx);

/* 1 when a flank of n_levels k-mer levels with `samples` samples per level fits the alignment kernels (at most
 * 2048 flank samples = levels x samples; up to 319 levels at samples = 6, the reference's default, run the fast kernels).  The reference aligns
 * flanks of any length (src/align_raw.h:117-158); longer ones are refused when the target is defined. */
int strique_align_supported(int n_levels, int samples);

/* Page-locked host memory for the caller's batch staging buffers (raw of strique_detect_batch with STRIQUE_HOST):
 * the read-batching driver writes decoded fast5 signals straight into it (the reference hands numpy arrays from
 * h5py to its workers, STRique_lib/fast5Index.py:76-84).  NULL when the allocation fails. */
void *strique_host_alloc(size_t bytes);
void strique_host_free(void *p);

/* ---- fast5 Signal chunks -> raw samples on the device -------------------------------------------------------
 * Replaces h5py's read of a deflate-filtered Signal dataset (STRique_lib/fast5Index.py:76-84: `fp[...][()]`, zlib
 * inflate of every chunk on the CPU).  HDF5 stores each chunk as an independent zlib stream; the caller locates the
 * chunks of a batch of reads (B-tree walk, no decompression), packs the stored bytes into `comp` (host or device)
 * and describes where each chunk's samples belong in the batch's raw-sample buffer.  One call inflates all of them
 * into a device buffer owned by the context (valid until the next call), which is then handed to
 * strique_detect_batch / strique_condition_batch as `raw` with STRIQUE_DEVICE.
 *   status_host[i] (written for every chunk): 0 ok, 1 not a zlib stream, 2 bad block header, 3 bad Huffman code,
 *   4 more output than `full`, 5 match before the start of the chunk, 6 stream longer than src_len, 7 Adler-32
 *   mismatch, 8 fewer than `keep` bytes.  The call itself succeeds when some chunks fail; the samples of such a
 *   chunk are undefined and the caller drops the read. */
typedef struct strique_inflate_chunk {
    int64_t src_off;      /* byte offset of the chunk's zlib stream in comp                                       */
    int64_t dst_off;      /* byte offset in the output of the chunk's first sample                                */
    int32_t src_len;      /* stored size of the chunk                                                             */
    int32_t keep;         /* bytes of the chunk that belong to the dataset (HDF5 pads the last chunk)             */
    int32_t full;         /* chunk size in bytes: the most the stream may produce                                 */
    int32_t reserved;
} strique_inflate_chunk;
int strique_inflate_batch(strique_ctx *ctx, const void *comp, int64_t comp_bytes, int comp_memspace,
                          const strique_inflate_chunk *chunks /* host */, int n_chunks, int64_t out_bytes,
                          int32_t *status_host, void **out_dev);

/* device time of the stages of the last strique_detect_batch call (ms) */
#define STRIQUE_STAGE_CONDITION 0
#define STRIQUE_STAGE_ALIGN_TABLE 1
#define STRIQUE_STAGE_ALIGN_SCAN 2
#define STRIQUE_STAGE_ALIGN_TRACE 3
#define STRIQUE_STAGE_VITERBI_COUNT 4
#define STRIQUE_STAGE_VITERBI_MOD 5
#define STRIQUE_STAGE_H2D 6
#define STRIQUE_N_STAGES 8
float strique_last_stage_ms(const strique_ctx *ctx, int stage);

#ifdef __cplusplus
}
#endif
#endif /* STRIQUE_B200_H */
