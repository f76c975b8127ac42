/*
 * ORACLE (test infrastructure, NOT product code): plain-C restatement of the reference's
 * semi-global signal alignment, giesselmann/STRique `align_raw<float,float>::semiglobal`
 * (src/align_raw.h:106-158) with `Score<float,Distance>` (src/score_distance.h:115-122) running on
 * SeqAn 2.4's affine-gap DP (seqan/align/dp_formula_affine.h:64-128, dp_formula.h:151-164,270-289,
 * dp_meta_info.h:178-221, dp_scout.h:167-180, dp_algorithm_impl.h:1168-1185,
 * dp_traceback_impl.h:377-481,496-547, dp_cell.h:137-145, dp_traceback_adaptor.h:58-117).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this.
 * Pinned against the compiled reference itself (oracle/_ref/pyseqan, built by oracle/build_ref.sh):
 * tests/test_oracle_align.py + tests/golden/align_*.npz.
 *
 * H = read signal a[0..N) (columns j = 1..N), V = flank b[0..L) (rows i = 1..L).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define T_DIAG 1
#define T_HOR 2
#define T_VER 4
#define T_HOPEN 8
#define T_VOPEN 16
#define T_MAXH 32
#define T_MAXV 64

typedef struct {
    float gap_ext_h, gap_ext_v, gap_open_h, gap_open_v, dist_offset, dist_min;
} oracle_align_params;

/* src/score_distance.h:117-122: subtraction in float, pow in double, cast back to float */
static inline float dist_score(const oracle_align_params *p, float h, float v) {
    float d = h > v ? h - v : v - h;
    float s = p->dist_offset - (float)pow((double)d, 1.2);
    return s > p->dist_min ? s : p->dist_min;
}

/*
 * Returns 0 on success, -1 on allocation failure.
 *   score     : best last-row score (SeqAn scout, strict >, starts at "infinity" = FLT_MIN/2)
 *   a_idx[N]  : view position of every source position of a   (align_raw.h:141-143)
 *   b_idx[L]  : view position of every source position of b   (align_raw.h:144-146)
 *   best_pos  : optional [2] = (j, i) of the traceback start
 */
int strique_oracle_align(const float *a, int64_t N, const float *b, int64_t L,
                         const oracle_align_params *p, float *score_out,
                         uint64_t *a_idx, uint64_t *b_idx, int64_t *best_pos) {
    const float INF = FLT_MIN / 2.0f; /* dp_cell.h:137-145: "minimum" of float / 2 */
    if (N == 0 || L == 0) {
        /* globalAlignment returns MinValue<float> (= FLT_MIN) without running the DP on an empty
         * sequence; all gaps (verified against the compiled reference) */
        for (int64_t k = 0; k < N; ++k) a_idx[k] = (uint64_t)k;
        for (int64_t k = 0; k < L; ++k) b_idx[k] = (uint64_t)k;
        *score_out = FLT_MIN;
        if (best_pos) { best_pos[0] = 0; best_pos[1] = 0; }
        return 0;
    }
    const float geh = p->gap_ext_h, gev = p->gap_ext_v, goh = p->gap_open_h, gov = p->gap_open_v;
    float *S = (float *)malloc(sizeof(float) * (size_t)(L + 1));
    float *Hm = (float *)malloc(sizeof(float) * (size_t)(L + 1));
    uint8_t *T = (uint8_t *)malloc((size_t)(N + 1) * (size_t)(L + 1));
    char *ops = (char *)malloc((size_t)(N + L + 2));
    if (!S || !Hm || !T || !ops) { free(S); free(Hm); free(T); free(ops); return -1; }
#define TR(j, i) T[(size_t)(j) * (size_t)(L + 1) + (size_t)(i)]
    float best = INF, bS = INF, bH = INF, bV = INF;
    int64_t bj = 0, bi = 0;
    /* column 0: row 0 zero, rows >= 1 vertical only (first column is not free) */
    S[0] = 0.0f; Hm[0] = INF; TR(0, 0) = 0;
    {
        float cS = 0.0f, cV = INF;
        for (int64_t i = 1; i <= L; ++i) {
            float e = cV + gev, o = cS + gov;
            uint8_t t;
            if (e < o) { cV = o; t = T_VOPEN; } else { cV = e; t = T_VER; }
            Hm[i] = INF; S[i] = cV; cS = cV;
            TR(0, i) = t | T_MAXV;
        }
        if (S[L] > best) { best = S[L]; bj = 0; bi = L; bS = S[L]; bH = Hm[L]; bV = cV; }
    }
    for (int64_t j = 1; j <= N; ++j) {
        float diag = S[0]; /* = 0 */
        S[0] = 0.0f; TR(j, 0) = 0;
        float cS = 0.0f, cV = INF;
        const float aj = a[j - 1];
        for (int64_t i = 1; i <= L; ++i) {
            float prevS = S[i], prevH = Hm[i];
            float inter = diag + dist_score(p, aj, b[i - 1]);
            diag = prevS;
            float e = prevH + geh, o = prevS + goh, h;
            uint8_t t;
            if (e < o) { h = o; t = T_HOPEN; } else { h = e; t = T_HOR; }
            e = cV + gev; o = cS + gov;
            if (e < o) { cV = o; t |= T_VOPEN; } else { cV = e; t |= T_VER; }
            float g; uint8_t t2;
            if (cV < h) { g = h; t2 = T_MAXH; } else { g = cV; t2 = T_MAXV; }
            float sc; uint8_t tr;
            if (inter < g) { sc = g; tr = t2 | t; } else { sc = inter; tr = T_DIAG | t; }
            Hm[i] = h; S[i] = sc; cS = sc; TR(j, i) = tr;
        }
        if (S[L] > best) { best = S[L]; bj = j; bi = L; bS = S[L]; bH = Hm[L]; bV = cV; }
    }
    /* _correctTraceValue, dp_algorithm_impl.h:1168-1185 */
    int64_t j = bj, i = bi;
    if (bV == bS) TR(j, i) = (uint8_t)((TR(j, i) & ~T_DIAG) | T_MAXV);
    else if (bH == bS) TR(j, i) = (uint8_t)((TR(j, i) & ~T_DIAG) | T_MAXH);
    /* _retrieveInitialTraceDirection (PreferGapsAtEnd), dp_traceback_impl.h:456-481 */
    uint8_t tv = TR(j, i);
    if (tv & T_MAXV) tv &= (T_VER | T_VOPEN | T_MAXV);
    else if (tv & T_MAXH) tv &= (T_HOR | T_HOPEN | T_MAXH);
    /* ops are collected back to front */
    size_t nops = 0;
    /* tail gaps: recorded V first then H => in the final alignment the H block precedes the V block */
    for (int64_t k = 0; k < L - i; ++k) ops[nops++] = 'V';
    for (int64_t k = 0; k < N - j; ++k) ops[nops++] = 'H';
    if (best_pos) { best_pos[0] = j; best_pos[1] = i; }
    while (j > 0 && i > 0 && tv != 0) { /* GapsLeft, dp_traceback_impl.h:377-417 */
        if (tv & T_DIAG) {
            ops[nops++] = 'D'; --j; --i; tv = TR(j, i);
        } else if ((tv & T_MAXV) && (tv & T_VER)) {
            while ((!(tv & T_VOPEN) || (tv & T_VER)) && i != 1) { --i; tv = TR(j, i); ops[nops++] = 'V'; }
            --i; tv = TR(j, i); ops[nops++] = 'V';
        } else if ((tv & T_MAXV) && (tv & T_VOPEN)) {
            --i; tv = TR(j, i); ops[nops++] = 'V';
        } else if ((tv & T_MAXH) && (tv & T_HOR)) {
            while ((!(tv & T_HOPEN) || (tv & T_HOR)) && j != 1) { --j; tv = TR(j, i); ops[nops++] = 'H'; }
            --j; tv = TR(j, i); ops[nops++] = 'H';
        } else if ((tv & T_MAXH) && (tv & T_HOPEN)) {
            --j; tv = TR(j, i); ops[nops++] = 'H';
        } else {
            break; /* undefined trace value: SeqAn asserts (compiled out with NDEBUG) */
        }
    }
    /* head gaps: V segment recorded first, then H (dp_traceback_impl.h:538-545); segments are
     * replayed in reverse recording order, so the H block ends up leftmost */
    for (int64_t k = 0; k < i; ++k) ops[nops++] = 'V';
    for (int64_t k = 0; k < j; ++k) ops[nops++] = 'H';
    /* view positions */
    {
        uint64_t col = 0; int64_t pa = 0, pb = 0;
        for (size_t k = nops; k-- > 0;) {
            char c = ops[k];
            if (c == 'D') { a_idx[pa++] = col; b_idx[pb++] = col; }
            else if (c == 'H') { a_idx[pa++] = col; }
            else { b_idx[pb++] = col; }
            ++col;
        }
    }
    *score_out = best;
    free(S); free(Hm); free(T); free(ops);
    return 0;
#undef TR
}

/* Score-only variant (no trace matrix): used by the CPU-baseline timing of long reads and to
 * cross-check the device scan kernel at sizes where (N+1)(L+1) bytes would not fit. */
int strique_oracle_align_score(const float *a, int64_t N, const float *b, int64_t L,
                               const oracle_align_params *p, float *score_out, int64_t *best_j) {
    const float INF = FLT_MIN / 2.0f;
    const float geh = p->gap_ext_h, gev = p->gap_ext_v, goh = p->gap_open_h, gov = p->gap_open_v;
    float *S = (float *)malloc(sizeof(float) * (size_t)(L + 1));
    float *Hm = (float *)malloc(sizeof(float) * (size_t)(L + 1));
    if (!S || !Hm) { free(S); free(Hm); return -1; }
    float best = INF; int64_t bj = 0;
    S[0] = 0.0f; Hm[0] = INF;
    {
        float cS = 0.0f, cV = INF;
        for (int64_t i = 1; i <= L; ++i) {
            float e = cV + gev, o = cS + gov;
            cV = e < o ? o : e; Hm[i] = INF; S[i] = cV; cS = cV;
        }
        if (S[L] > best) { best = S[L]; bj = 0; }
    }
    for (int64_t j = 1; j <= N; ++j) {
        float diag = 0.0f, cS = 0.0f, cV = INF;
        const float aj = a[j - 1];
        for (int64_t i = 1; i <= L; ++i) {
            float prevS = S[i], prevH = Hm[i];
            float inter = diag + dist_score(p, aj, b[i - 1]);
            diag = prevS;
            float e = prevH + geh, o = prevS + goh;
            float h = e < o ? o : e;
            e = cV + gev; o = cS + gov;
            cV = e < o ? o : e;
            float g = cV < h ? h : cV;
            float sc = inter < g ? g : inter;
            Hm[i] = h; S[i] = sc; cS = sc;
        }
        if (S[L] > best) { best = S[L]; bj = j; }
    }
    *score_out = best; *best_j = (L > 0 || best > INF) ? bj : 0;
    free(S); free(Hm);
    return 0;
}
