/*
 * overlap_oracle.c -- TEST INFRASTRUCTURE ONLY (the parity oracle).
 *
 * A plain-C restatement of the overlap-detection hot path of
 * roiteichman/Genome-Assembly-Using-Overlap-Graphs.  Nothing under the product
 * package may import, link or call this file; only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs use it, as the checker or
 * as the timed CPU arm.
 *
 * Parity pinning: the reference ships no golden vectors for this path
 * (SURVEY.md section 4), so this restatement is pinned against outputs of the live
 * reference (imported unmodified in the build container) -- see
 * tests/golden/make_golden.py and the JSON fixtures beside it.
 *
 * Each function cites the reference lines it restates (paths relative to the
 * reference checkout).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OVO_OK 0
#define OVO_ENOMEM -1

/*
 * Full restatement of aligners.py:27-82 (overlap_alignment).
 *   - dp is int32 storage, traceback int8 storage (aligners.py:28-30)
 *   - candidates are formed in int64 (Numba types the omitted default
 *     indel=-2**31 as int64, so dp+indel never wraps) and truncated to int32 on
 *     store (aligners.py:35-48)
 *   - tie order diag >= up >= left (aligners.py:40-48)
 *   - best = first strict maximum over the last row j=0..m (aligners.py:50-57)
 *   - traceback walk while i>0 and j>0 (aligners.py:59-76)
 * align_s / align_t receive the aligned strings ('-' for gaps), each needs n+m+1
 * bytes of room; *align_len is their common length.  Pass NULL to skip the walk.
 */
int ovo_overlap_alignment(const uint8_t *s, int32_t n, const uint8_t *t, int32_t m,
                          int64_t match, int64_t mismatch, int64_t indel,
                          int32_t *score_out, int32_t *end_out,
                          uint8_t *align_s, uint8_t *align_t, int32_t *align_len)
{
    size_t W = (size_t)m + 1;
    int32_t *dp = (int32_t *)calloc((size_t)(n + 1) * W, sizeof(int32_t));
    int8_t *tb = (int8_t *)calloc((size_t)(n + 1) * W, sizeof(int8_t));
    if (!dp || !tb) { free(dp); free(tb); return OVO_ENOMEM; }

    for (int32_t i = 1; i <= n; ++i) {
        const int32_t *prev = dp + (size_t)(i - 1) * W;
        int32_t *cur = dp + (size_t)i * W;
        int8_t *tbr = tb + (size_t)i * W;
        uint8_t si = s[i - 1];
        for (int32_t j = 1; j <= m; ++j) {
            int64_t diag = (int64_t)prev[j - 1] + (si == t[j - 1] ? match : mismatch);
            int64_t up = (int64_t)prev[j] + indel;
            int64_t left = (int64_t)cur[j - 1] + indel;
            if (diag >= up && diag >= left) { cur[j] = (int32_t)diag; tbr[j] = 0; }
            else if (up >= left)            { cur[j] = (int32_t)up;   tbr[j] = 1; }
            else                            { cur[j] = (int32_t)left; tbr[j] = 2; }
        }
    }

    double max_score = -INFINITY;
    int32_t overlap_len = 0;
    const int32_t *last = dp + (size_t)n * W;
    for (int32_t j = 0; j <= m; ++j) {
        if ((double)last[j] > max_score) { max_score = (double)last[j]; overlap_len = j; }
    }
    *score_out = (int32_t)max_score;
    *end_out = overlap_len;

    if (align_s && align_t && align_len) {
        /* the reference prepends characters; build reversed, then flip */
        int32_t i = n, j = overlap_len, L = 0;
        while (i > 0 && j > 0) {
            int8_t d = tb[(size_t)i * W + j];
            if (d == 0)      { align_s[L] = s[i - 1]; align_t[L] = t[j - 1]; --i; --j; }
            else if (d == 1) { align_s[L] = s[i - 1]; align_t[L] = '-';      --i; }
            else             { align_s[L] = '-';      align_t[L] = t[j - 1]; --j; }
            ++L;
        }
        for (int32_t a = 0, b = L - 1; a < b; ++a, --b) {
            uint8_t x = align_s[a]; align_s[a] = align_s[b]; align_s[b] = x;
            x = align_t[a]; align_t[a] = align_t[b]; align_t[b] = x;
        }
        align_s[L] = 0; align_t[L] = 0;
        *align_len = L;
    }
    free(dp); free(tb);
    return OVO_OK;
}

/*
 * Score/end only, two rolling rows, same arithmetic as above
 * (aligners.py:33-57).  Used where the full matrices would be too slow for a
 * test; cross-checked against ovo_overlap_alignment in tests/test_oracle.py.
 */
int ovo_overlap_score(const uint8_t *s, int32_t n, const uint8_t *t, int32_t m,
                      int64_t match, int64_t mismatch, int64_t indel,
                      int32_t *score_out, int32_t *end_out)
{
    size_t W = (size_t)m + 1;
    int32_t *a = (int32_t *)calloc(2 * W, sizeof(int32_t));
    if (!a) return OVO_ENOMEM;
    int32_t *prev = a, *cur = a + W;
    for (int32_t i = 1; i <= n; ++i) {
        uint8_t si = s[i - 1];
        cur[0] = 0;
        for (int32_t j = 1; j <= m; ++j) {
            int64_t diag = (int64_t)prev[j - 1] + (si == t[j - 1] ? match : mismatch);
            int64_t up = (int64_t)prev[j] + indel;
            int64_t left = (int64_t)cur[j - 1] + indel;
            int64_t v = (diag >= up && diag >= left) ? diag : (up >= left ? up : left);
            cur[j] = (int32_t)v;
        }
        int32_t *x = prev; prev = cur; cur = x;
    }
    int32_t best = prev[0], bj = 0;        /* j = 0 always beats -inf (aligners.py:51-57) */
    for (int32_t j = 1; j <= m; ++j) if (prev[j] > best) { best = prev[j]; bj = j; }
    *score_out = best; *end_out = bj;
    free(a);
    return OVO_OK;
}

/*
 * The DP call site of the graph builder (overlapGraphs.py:53) applied to a list
 * of unique-read index pairs: score[p], end[p] = overlap_alignment(reads[a], reads[b]).
 * `full` != 0 does everything the reference does per call (both matrices and the
 * traceback walk) -- that is the timed CPU baseline; `full` == 0 is the rolling-row
 * variant.  Parallel over pairs with OpenMP (the reference's own parallelism is
 * process-level joblib over experiments, experiments.py:537; pairs are independent).
 */
int ovo_overlap_pairs(const uint8_t *bases, const int64_t *offsets,
                      const int32_t *pair_a, const int32_t *pair_b, int64_t P,
                      int64_t match, int64_t mismatch, int64_t indel,
                      int32_t *score, int32_t *end, int32_t full, int32_t nthreads)
{
    int rc = OVO_OK;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
    {
        uint8_t *as = NULL, *at = NULL;
        size_t cap = 0;
#pragma omp for schedule(dynamic, 16)
        for (int64_t p = 0; p < P; ++p) {
            const uint8_t *s = bases + offsets[pair_a[p]];
            const uint8_t *t = bases + offsets[pair_b[p]];
            int32_t n = (int32_t)(offsets[pair_a[p] + 1] - offsets[pair_a[p]]);
            int32_t m = (int32_t)(offsets[pair_b[p] + 1] - offsets[pair_b[p]]);
            int r;
            if (full) {
                size_t need = (size_t)n + m + 2;
                if (need > cap) {
                    free(as); free(at);
                    as = (uint8_t *)malloc(need); at = (uint8_t *)malloc(need); cap = need;
                }
                int32_t L;
                r = ovo_overlap_alignment(s, n, t, m, match, mismatch, indel,
                                          &score[p], &end[p], as, at, &L);
            } else {
                r = ovo_overlap_score(s, n, t, m, match, mismatch, indel, &score[p], &end[p]);
            }
            if (r != OVO_OK) {
#pragma omp atomic write
                rc = r;
            }
        }
        free(as); free(at);
    }
    return rc;
}

/*
 * Full restatement of aligners.py:106-167 (local_alignment): Smith-Waterman with the reference's
 * branch order (diag if >= up, >= left and >= 0; else up if >= left and >= 0; else left if >= 0;
 * else 0), first strict maximum in row-major order, walk while dp > 0 and traceback != 0.
 * out[0..3] = best_score, start_pos, end_pos, best_i;  aligned strings as in ovo_overlap_alignment
 * (a_q = aligned query, a_r = aligned reference).
 */
int ovo_local_alignment(const uint8_t *q, int32_t n, const uint8_t *r, int32_t m,
                        int64_t match, int64_t mismatch, int64_t indel,
                        int32_t *out, uint8_t *a_q, uint8_t *a_r, int32_t *align_len)
{
    size_t W = (size_t)m + 1;
    int32_t *dp = (int32_t *)calloc((size_t)(n + 1) * W, sizeof(int32_t));
    int8_t *tb = (int8_t *)calloc((size_t)(n + 1) * W, sizeof(int8_t));
    if (!dp || !tb) { free(dp); free(tb); return OVO_ENOMEM; }
    int64_t best = 0;
    int32_t bi = 0, bj = 0;
    for (int32_t i = 1; i <= n; ++i) {
        const int32_t *prev = dp + (size_t)(i - 1) * W;
        int32_t *cur = dp + (size_t)i * W;
        int8_t *tbr = tb + (size_t)i * W;
        for (int32_t j = 1; j <= m; ++j) {
            int64_t diag = (int64_t)prev[j - 1] + (q[i - 1] == r[j - 1] ? match : mismatch);
            int64_t up = (int64_t)prev[j] + indel;
            int64_t left = (int64_t)cur[j - 1] + indel;
            if (diag >= up && diag >= left && diag >= 0) { cur[j] = (int32_t)diag; tbr[j] = 1; }
            else if (up >= left && up >= 0)              { cur[j] = (int32_t)up;   tbr[j] = 2; }
            else if (left >= 0)                          { cur[j] = (int32_t)left; tbr[j] = 3; }
            else                                         { cur[j] = 0; }
            if ((int64_t)cur[j] > best) { best = cur[j]; bi = i; bj = j; }
        }
    }
    int32_t i = bi, j = bj, L = 0;
    while (i > 0 && j > 0 && dp[(size_t)i * W + j] > 0) {
        int8_t d = tb[(size_t)i * W + j];
        if (d == 1)      { a_q[L] = q[i - 1]; a_r[L] = r[j - 1]; --i; --j; }
        else if (d == 2) { a_q[L] = q[i - 1]; a_r[L] = '-';      --i; }
        else if (d == 3) { a_q[L] = '-';      a_r[L] = r[j - 1]; --j; }
        else break;
        ++L;
    }
    for (int32_t a = 0, b = L - 1; a < b; ++a, --b) {
        uint8_t x = a_q[a]; a_q[a] = a_q[b]; a_q[b] = x;
        x = a_r[a]; a_r[a] = a_r[b]; a_r[b] = x;
    }
    a_q[L] = 0; a_r[L] = 0;
    *align_len = L;
    out[0] = (int32_t)best; out[1] = j; out[2] = bj; out[3] = bi;
    free(dp); free(tb);
    return OVO_OK;
}

int ovo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
