// Read conditioning on the device (one CTA per read).
//
// Replaces the numpy/scipy/scikit-image lines of repeatCounter.detect
// (reference scripts/STRique.py:590-597): 3-tap median filter, (x - median)/MAD z-score,
// *24+127 -> uint8, grey opening + closing with an 8-wide flat element (scikit-image < 0.15
// window conventions), and pore_model.normalize2model(mode='minmax') (S.py:151-160,178-179) for
// the three consumers (aligner: morphology signal; count HMM: median-filtered signal; methylation
// HMM: raw signal).  Order statistics (median, 1st/99th percentile with numpy's linear
// interpolation, tail medians) are exact: multi-rank radix select over order-preserving keys.
#pragma once
#include "common.cuh"

namespace strique {

// constants of the pore model needed by 'minmax' (computed on the host from the 4096 k-mer means)
struct CondModel {
    double m5_mod, m95_mod;       // medians of the model means below the 1st / above the 99th percentile
    double model_min, model_max;  // S.py:126-127
};

// per-read results (device, doubles)
enum {
    CS_FLT_MEDIAN = 0, CS_FLT_MAD, CS_FLT_C1, CS_FLT_C2, CS_RAW_C1, CS_RAW_C2, CS_U8_C1, CS_U8_C2, CS_STATUS,
    CS_STRIDE = 12
};

// raw_kind: 0 = int16 samples, 1 = float64 samples
int condition_run_device(strique_ctx *ctx, int raw_kind, const void *raw, const int64_t *sig_off_dev,
                         const int64_t *sig_off_host, int n_reads, const CondModel &model, bool want_raw_stats,
                         void *flt_out, uint16_t *codes_out, float *code_values_out /* [n_reads*256] */,
                         double *stats_out /* [n_reads*CS_STRIDE] */);

// x -> clip(((x - c1) / c2) * c3 + c4, lo, hi) with numpy's operation order (S.py:157-158,178-179)
__host__ __device__ inline void minmax_model_constants(const CondModel &m, double *c3, double *c4, double *lo, double *hi) {
    *c3 = (m.m95_mod - m.m5_mod) / 2;
    *c4 = m.m5_mod + (m.m95_mod - m.m5_mod) / 2;
    *lo = m.model_min + .5;
    *hi = m.model_max - .5;
}

}  // namespace strique
