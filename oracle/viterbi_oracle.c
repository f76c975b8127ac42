/*
 * ORACLE (test infrastructure, NOT product code): plain-C float64 restatement of the sparse
 * Viterbi decoder of pomegranate 0.10.0 (`HiddenMarkovModel._viterbi`, hmm.pyx; pinned by
 * giesselmann/STRique requirements.txt:10-11, called from scripts/STRique.py:434 and :493).
 * pomegranate is NOT vendored under /root/reference and not installable here, so this follows
 * the published algorithm of that version as recalled in SURVEY.md App. C:
 *   - states ordered [emitting | silent (topological)], `silent_start` = #emitting
 *   - v[0][start] = 0, silent relaxation in index order (only from silent k < l)
 *   - per sample: emitting from v[t] (+edge +emission), then silent from emitting at t+1,
 *     then silent from earlier silent at t+1; every max is a strict '>' over in-edges in order
 *   - log p = v[T][end]; traceback over (time, state) pointers
 * Emissions (distributions.pyx of that version): Normal  c0 - (x-mu)^2 * 1/(2 sigma^2) with
 * c0 = -log(sigma*sqrt(2 pi)); Uniform -log(hi-lo) inside [lo,hi] else -inf; NaN sample -> 0.
 * PARITY UNPINNED for the float outputs: no reference test pins log_p; the integer
 * repeat counts are pinned by the reference's unit-test assertions (tests/test_oracle_pipeline.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may call this.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#define SQRT_2_PI 2.50662827463

typedef struct {
    int32_t n_states;      /* m */
    int32_t silent_start;  /* number of emitting states */
    int32_t start_index, end_index;
    const int32_t *in_ptr;   /* [m+1] CSR over in-edges */
    const int32_t *in_src;   /* [nnz] source state */
    const double *in_logp;   /* [nnz] */
    const int32_t *dist_kind; /* [silent_start] 0 = normal, 1 = uniform */
    const double *dist_a;     /* mu | lo */
    const double *dist_b;     /* sigma | hi */
} oracle_hmm;

static inline double emission(int kind, double a, double b, double c0, double c1, double x) {
    if (isnan(x)) return 0.0;
    if (kind == 0) { double d = x - a; return c0 - (d * d) * c1; }
    return (x >= a && x <= b) ? c0 : -INFINITY;
}

/*
 * path_out: state index per path element (forward order, start first, end last), capacity path_cap.
 * Returns the path length (>0), 0 if the sequence is impossible (log p = -inf), -1 on alloc failure,
 * -2 if path_cap is too small.
 */
/*
 * margin_out (optional): the gap between the best and the second-best complete path = the smallest
 * difference, over the decisions ON the best path, between the winning in-edge and the best other
 * in-edge of that state (any second-best path leaves the best one for the last time at such a decision).
 * Reads with a gap below the arithmetic resolution of a faster decoder are its legitimate exceptions.
 */
int64_t strique_oracle_viterbi_margin(const oracle_hmm *h, const double *x, int64_t T, double *logp_out,
                                      int32_t *path_out, int64_t path_cap, double *margin_out) {
    const int m = h->n_states, p = h->silent_start;
    double *v = (double *)malloc(sizeof(double) * (size_t)(T + 1) * (size_t)m);
    int32_t *tbx = (int32_t *)malloc(sizeof(int32_t) * (size_t)(T + 1) * (size_t)m);
    int32_t *tby = (int32_t *)malloc(sizeof(int32_t) * (size_t)(T + 1) * (size_t)m);
    double *c0 = (double *)malloc(sizeof(double) * (size_t)(p > 0 ? p : 1));
    double *c1 = (double *)malloc(sizeof(double) * (size_t)(p > 0 ? p : 1));
    if (!v || !tbx || !tby || !c0 || !c1) { free(v); free(tbx); free(tby); free(c0); free(c1); return -1; }
    for (int l = 0; l < p; ++l) {
        if (h->dist_kind[l] == 0) {
            c0[l] = -log(h->dist_b[l] * SQRT_2_PI);
            c1[l] = h->dist_b[l] > 0 ? 1.0 / (2.0 * (h->dist_b[l] * h->dist_b[l])) : 0.0;
        } else {
            c0[l] = -log(h->dist_b[l] - h->dist_a[l]);
            c1[l] = 0.0;
        }
    }
    for (size_t k = 0; k < (size_t)(T + 1) * (size_t)m; ++k) { v[k] = -INFINITY; tbx[k] = -1; tby[k] = -1; }
    v[h->start_index] = 0.0;
    for (int l = p; l < m; ++l) {
        if (l == h->start_index) continue;
        for (int k = h->in_ptr[l]; k < h->in_ptr[l + 1]; ++k) {
            int ki = h->in_src[k];
            if (ki < p || ki >= l) continue;
            double s = v[ki] + h->in_logp[k];
            if (s > v[l]) { v[l] = s; tbx[l] = 0; tby[l] = ki; }
        }
    }
    for (int64_t i = 0; i < T; ++i) {
        double *vi = v + (size_t)i * m, *vn = v + (size_t)(i + 1) * m;
        int32_t *bx = tbx + (size_t)(i + 1) * m, *by = tby + (size_t)(i + 1) * m;
        for (int l = 0; l < p; ++l) {
            double e = emission(h->dist_kind[l], h->dist_a[l], h->dist_b[l], c0[l], c1[l], x[i]);
            for (int k = h->in_ptr[l]; k < h->in_ptr[l + 1]; ++k) {
                int ki = h->in_src[k];
                double s = vi[ki] + h->in_logp[k] + e;
                if (s > vn[l]) { vn[l] = s; bx[l] = (int32_t)i; by[l] = ki; }
            }
        }
        for (int l = p; l < m; ++l)
            for (int k = h->in_ptr[l]; k < h->in_ptr[l + 1]; ++k) {
                int ki = h->in_src[k];
                if (ki >= p) continue;
                double s = vn[ki] + h->in_logp[k];
                if (s > vn[l]) { vn[l] = s; bx[l] = (int32_t)(i + 1); by[l] = ki; }
            }
        for (int l = p; l < m; ++l)
            for (int k = h->in_ptr[l]; k < h->in_ptr[l + 1]; ++k) {
                int ki = h->in_src[k];
                if (ki < p || ki >= l) continue;
                double s = vn[ki] + h->in_logp[k];
                if (s > vn[l]) { vn[l] = s; bx[l] = (int32_t)(i + 1); by[l] = ki; }
            }
    }
    double lp = v[(size_t)T * m + h->end_index];
    *logp_out = lp;
    int64_t n = 0;
    if (lp == -INFINITY) { n = 0; goto done; }
    {
        /* walk back, then reverse in place */
        int64_t px = T; int32_t py = h->end_index;
        double margin = INFINITY;
        while (!(px == 0 && py == h->start_index)) {
            if (n >= path_cap) { n = -2; goto done; }
            path_out[n++] = py;
            size_t at = (size_t)px * m + (size_t)py;
            int32_t nx = tbx[at], ny = tby[at];
            if (nx < 0) { n = 0; *logp_out = -INFINITY; goto done; }
            if (margin_out) {
                /* best candidate of (px, py) that does not come from (nx, ny) */
                double second = -INFINITY;
                for (int k = h->in_ptr[py]; k < h->in_ptr[py + 1]; ++k) {
                    int ki = h->in_src[k];
                    double s;
                    if (py < p) {
                        if (px < 1) continue;
                        double e = emission(h->dist_kind[py], h->dist_a[py], h->dist_b[py], c0[py], c1[py], x[px - 1]);
                        if (ki == ny) continue;
                        s = v[(size_t)(px - 1) * m + ki] + h->in_logp[k] + e;
                    } else {
                        if (ki == ny) continue;
                        if (ki >= p && ki >= py) continue;           /* silent sources must precede */
                        if (ki < p && px == 0) continue;             /* column 0 has no emitting values */
                        s = v[(size_t)px * m + ki] + h->in_logp[k];
                    }
                    if (s > second) second = s;
                }
                if (v[at] - second < margin) margin = v[at] - second;
            }
            px = nx; py = ny;
        }
        if (n >= path_cap) { n = -2; goto done; }
        path_out[n++] = h->start_index;
        if (margin_out) *margin_out = margin;
        for (int64_t a = 0, b = n - 1; a < b; ++a, --b) { int32_t t = path_out[a]; path_out[a] = path_out[b]; path_out[b] = t; }
    }
done:
    free(v); free(tbx); free(tby); free(c0); free(c1);
    return n;
}

int64_t strique_oracle_viterbi(const oracle_hmm *h, const double *x, int64_t T, double *logp_out,
                               int32_t *path_out, int64_t path_cap) {
    return strique_oracle_viterbi_margin(h, x, T, logp_out, path_out, path_cap, 0);
}
