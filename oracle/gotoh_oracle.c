/*
 * oracle/gotoh_oracle.c -- CPU restatement of the reference's pairwise-alignment hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under cse305_parallel_sequence_alignment_b200/ may
 * include, link or call this file.  It is used by tests/, by __graft_entry__.smoke() and
 * by bench.py's cpu_baseline / --impl reference legs, as the checker.
 *
 * What it restates (citations relative to /root/reference):
 *   - table borders + recurrence ............ alignment_algorithm/subproblem_alignment.cpp:212-227,
 *                                             :251-292 (borders), :229-249 and :357-400 (recurrence)
 *   - substitution score f / end credit ..... alignment_algorithm/subproblem_alignment.h:83-96
 *   - traceback (first-equality order, truncation at the border, dropped head node)
 *                                             alignment_algorithm/subproblem_alignment.cpp:105-172
 *   - output rows ........................... alignment_algorithm/main_alignment.cpp:32-55
 *   - live start/end types (-1,-1) .......... alignment_algorithm/main_alignment.cpp:396-407,
 *                                             :250-251 (end_type = -pbp[k+1].t)
 *
 * Arithmetic: the reference stores doubles but every value is an exact small integer or
 * -infinity (g, h integral).  Here: int32 with ORC_NEG standing for -infinity and the rule
 * "-inf + c = -inf" (orc_add), so equality tests in the traceback behave like the doubles.
 *
 * Parity status:
 *   GLOBAL mode is pinned against the reference itself: tests/test_oracle_vs_ref.py runs the
 *   (2-line-repaired) reference built by oracle/build_ref.sh on thousands of random pairs and
 *   on the SURVEY section 8c golden vectors (tests/golden/).
 *   LOCAL mode (Smith-Waterman) does not exist in the reference: PARITY UNPINNED by the
 *   reference.  The spec implemented here is normative for this repo:
 *     T1[i][j] = f(i,j) + max(0, T1, T2, T3)[i-1][j-1]; T2, T3 as in global mode; all borders
 *     -inf.  score = max T1 over all cells (0 => empty alignment); end = the maximising cell
 *     with the smallest i, then the smallest j.  Traceback starts at the end cell in state 1;
 *     in state 1 the 0-floor is tested FIRST (T1[i][j] == f(i,j) => this is the first column,
 *     emit it and stop -- the conventional Smith-Waterman stop), otherwise the reference's
 *     predecessor order T1, T2, T3 applies; states 2 and 3 exactly as the reference.
 *
 * Sequences are passed as pointers to base 1 (a[0] is the reference's A[1]) plus lengths.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_NEG (INT32_MIN / 2)

typedef struct {
    int32_t t1, t2, t3;      /* corner values T1/T2/T3[m][n] (global); ORC_NEG = -inf        */
    int32_t score;           /* global: max(t1,t2,t3); local: best T1                         */
    int32_t end_state;       /* state the traceback starts in (1,2,3)                         */
    int64_t end_i, end_j;    /* global: m,n ; local: 1-based end cell (0,0 if score==0)      */
    int64_t start_i, start_j;/* 1-based first emitted cell (0,0 if nothing emitted)          */
    int64_t aln_len;         /* number of emitted columns                                     */
} orc_result;

static inline int32_t orc_add(int32_t x, int32_t c) { return x <= ORC_NEG ? ORC_NEG : x + c; }
static inline int32_t orc_max(int32_t a, int32_t b) { return a > b ? a : b; }
static inline int32_t orc_max3(int32_t a, int32_t b, int32_t c) { return orc_max(orc_max(a, b), c); }

/* Fill the three full (m+1)x(n+1) tables, row-major with row stride n+1.
 * mode 0: global, borders by start_type (subproblem_alignment.cpp:259-292, 212-227).
 * mode 1: local (spec in the header). */
int orc_fill_full(const char* a, const char* b, int64_t m, int64_t n, int g, int h, int start_type,
                  int mode, int32_t* T1, int32_t* T2, int32_t* T3) {
    if (m < 0 || n < 0 || g < 0 || h < 0) return -1;
    const int64_t W = n + 1;
    const int32_t go = g + h;
    /* row 0 */
    T1[0] = ORC_NEG; T2[0] = ORC_NEG; T3[0] = ORC_NEG;
    if (mode == 0) {
        if (start_type == 1 || start_type == -1) T1[0] = 0;
        else if (start_type == -2) T2[0] = 0;
        else if (start_type == -3) T3[0] = 0;
    }
    for (int64_t j = 1; j <= n; ++j) {
        T1[j] = ORC_NEG; T3[j] = ORC_NEG;
        if (mode == 1) T2[j] = ORC_NEG;
        else if (start_type == -2) T2[j] = (int32_t)(-(int64_t)g * j);
        else if (start_type == 1 || start_type == 3) T2[j] = ORC_NEG;
        else T2[j] = (int32_t)(-h - (int64_t)g * j);
    }
    for (int64_t i = 1; i <= m; ++i) {
        int32_t *r1 = T1 + i * W, *r2 = T2 + i * W, *r3 = T3 + i * W;
        const int32_t *p1 = r1 - W, *p2 = r2 - W, *p3 = r3 - W;
        r1[0] = ORC_NEG; r2[0] = ORC_NEG;
        if (mode == 1) r3[0] = ORC_NEG;
        else if (start_type == -3) r3[0] = (int32_t)(-(int64_t)g * i);
        else if (start_type == 1 || start_type == 2) r3[0] = ORC_NEG;
        else r3[0] = (int32_t)(-h - (int64_t)g * i);
        const char ai = a[i - 1];
        for (int64_t j = 1; j <= n; ++j) {
            const int32_t f = (ai == b[j - 1]) ? 1 : 0;
            int32_t d = orc_max3(p1[j - 1], p2[j - 1], p3[j - 1]);
            if (mode == 1) d = orc_max(d, 0);
            r1[j] = orc_add(d, f);
            r3[j] = orc_max3(orc_add(p1[j], -go), orc_add(p2[j], -go), orc_add(p3[j], -g));
            r2[j] = orc_max3(orc_add(r1[j - 1], -go), orc_add(r2[j - 1], -g), orc_add(r3[j - 1], -go));
        }
    }
    return 0;
}

/* Traceback over full tables (subproblem_alignment.cpp:105-172).  Writes the emitted states
 * (1,2,3) in FORWARD order (alignment_begin .. alignment_end) into ops[0..aln_len) -- capacity
 * m+n -- and fills res.  Returns 0, or -2 if no predecessor equality holds (the reference would
 * read an uninitialised node there). */
int orc_traceback_full(const char* a, const char* b, int64_t m, int64_t n, int g, int h, int end_type,
                       int mode, const int32_t* T1, const int32_t* T2, const int32_t* T3,
                       orc_result* res, uint8_t* ops) {
    const int64_t W = n + 1;
    const int32_t go = g + h;
    int64_t i = m, j = n;
    int state;
    memset(res, 0, sizeof(*res));
    if (mode == 0) {
        res->t1 = T1[m * W + n]; res->t2 = T2[m * W + n]; res->t3 = T3[m * W + n];
        res->score = orc_max3(res->t1, res->t2, res->t3);
        if (end_type > 0) state = end_type;
        else {
            /* h_prime (subproblem_alignment.h:91-96): +h credit only if end_type is -2/-3 */
            const int32_t t1 = res->t1;
            const int32_t t2 = orc_add(res->t2, end_type == -2 ? h : 0);
            const int32_t t3 = orc_add(res->t3, end_type == -3 ? h : 0);
            if (t1 >= t2 && t1 >= t3) state = 1;
            else if (t2 >= t1 && t2 >= t3) state = 2;
            else state = 3;
        }
        res->end_i = m; res->end_j = n;
    } else {
        int32_t best = 0; int64_t bi = 0, bj = 0;
        for (int64_t ii = 1; ii <= m; ++ii)
            for (int64_t jj = 1; jj <= n; ++jj)
                if (T1[ii * W + jj] > best) { best = T1[ii * W + jj]; bi = ii; bj = jj; }
        res->score = best; res->end_i = bi; res->end_j = bj;
        res->t1 = best; res->t2 = ORC_NEG; res->t3 = ORC_NEG;
        state = 1; i = bi; j = bj;
    }
    res->end_state = state;
    /* emit in reverse, flip at the end */
    int64_t len = 0;
    while (i > 0 && j > 0) {
        ops[len++] = (uint8_t)state;
        res->start_i = i; res->start_j = j;
        const int32_t f = (a[i - 1] == b[j - 1]) ? 1 : 0;
        if (state == 1) {
            const int32_t v = T1[i * W + j];
            if (mode == 1 && v == f) break;                       /* 0-floor first (local spec) */
            if (v == orc_add(T1[(i - 1) * W + j - 1], f)) state = 1;
            else if (v == orc_add(T2[(i - 1) * W + j - 1], f)) state = 2;
            else if (v == orc_add(T3[(i - 1) * W + j - 1], f)) state = 3;
            else return -2;
            --i; --j;
        } else if (state == 2) {
            const int32_t v = T2[i * W + j];
            if (v == orc_add(T1[i * W + j - 1], -go)) state = 1;
            else if (v == orc_add(T2[i * W + j - 1], -g)) state = 2;
            else if (v == orc_add(T3[i * W + j - 1], -go)) state = 3;
            else return -2;
            --j;
        } else {
            const int32_t v = T3[i * W + j];
            if (v == orc_add(T1[(i - 1) * W + j], -go)) state = 1;
            else if (v == orc_add(T2[(i - 1) * W + j], -go)) state = 2;
            else if (v == orc_add(T3[(i - 1) * W + j], -g)) state = 3;
            else return -2;
            --i;
        }
    }
    for (int64_t k = 0; k < len / 2; ++k) { uint8_t t = ops[k]; ops[k] = ops[len - 1 - k]; ops[len - 1 - k] = t; }
    res->aln_len = len;
    if (len == 0) { res->start_i = 0; res->start_j = 0; }
    return 0;
}

/* Expand forward ops into the two printed rows (main_alignment.cpp:32-55).  rows hold aln_len
 * chars each (no terminator).  (start_i,start_j) is the first emitted cell. */
void orc_render_rows(const char* a, const char* b, const uint8_t* ops, int64_t len, int64_t start_i,
                     int64_t start_j, char* row_a, char* row_b) {
    int64_t i = start_i, j = start_j;
    for (int64_t k = 0; k < len; ++k) {
        const int t = ops[k];
        row_a[k] = (t == 1 || t == 3) ? a[i - 1] : '-';
        row_b[k] = (t == 1 || t == 2) ? b[j - 1] : '-';
        if (k + 1 < len) {            /* next emitted cell: its state says which way we came */
            const int tn = ops[k + 1];
            if (tn == 1) { ++i; ++j; } else if (tn == 2) { ++j; } else { ++i; }
        }
    }
}

/* One call: fill + traceback + rows.  Allocates 3 full tables (12 B/cell).  ops/row_a/row_b
 * need capacity m+n.  t_dump, if non-NULL, receives the three tables (3*(m+1)*(n+1) int32). */
int orc_align_full(const char* a, const char* b, int64_t m, int64_t n, int g, int h, int start_type,
                   int end_type, int mode, orc_result* res, uint8_t* ops, char* row_a, char* row_b,
                   int32_t* t_dump) {
    const size_t cells = (size_t)(m + 1) * (size_t)(n + 1);
    int32_t* T = t_dump ? t_dump : (int32_t*)malloc(3 * cells * sizeof(int32_t));
    if (!T) return -3;
    int rc = orc_fill_full(a, b, m, n, g, h, start_type, mode, T, T + cells, T + 2 * cells);
    if (rc == 0) rc = orc_traceback_full(a, b, m, n, g, h, end_type, mode, T, T + cells, T + 2 * cells, res, ops);
    if (rc == 0 && row_a && row_b) orc_render_rows(a, b, ops, res->aln_len, res->start_i, res->start_j, row_a, row_b);
    if (!t_dump) free(T);
    return rc;
}

/* Linear-space score-only fill (two rows).  Global (start_type -1 only): corner values.
 * Local: score and end cell (smallest i, then smallest j).  Used where the reference cannot
 * allocate O(mn) (C4, C5, full-length records). */
int orc_score_linear(const char* a, const char* b, int64_t m, int64_t n, int g, int h, int mode, orc_result* res) {
    if (m < 0 || n < 0 || g < 0 || h < 0) return -1;
    const int32_t go = g + h;
    int32_t* buf = (int32_t*)malloc(6 * (size_t)(n + 1) * sizeof(int32_t));
    if (!buf) return -3;
    int32_t *p1 = buf, *p2 = p1 + n + 1, *p3 = p2 + n + 1, *c1 = p3 + n + 1, *c2 = c1 + n + 1, *c3 = c2 + n + 1;
    memset(res, 0, sizeof(*res));
    p1[0] = mode == 0 ? 0 : ORC_NEG; p2[0] = ORC_NEG; p3[0] = ORC_NEG;
    for (int64_t j = 1; j <= n; ++j) {
        p1[j] = ORC_NEG; p3[j] = ORC_NEG;
        p2[j] = mode == 0 ? (int32_t)(-h - (int64_t)g * j) : ORC_NEG;
    }
    int32_t best = 0; int64_t bi = 0, bj = 0;
    for (int64_t i = 1; i <= m; ++i) {
        c1[0] = ORC_NEG; c2[0] = ORC_NEG;
        c3[0] = mode == 0 ? (int32_t)(-h - (int64_t)g * i) : ORC_NEG;
        const char ai = a[i - 1];
        for (int64_t j = 1; j <= n; ++j) {
            const int32_t f = (ai == b[j - 1]) ? 1 : 0;
            int32_t d = orc_max3(p1[j - 1], p2[j - 1], p3[j - 1]);
            if (mode == 1) d = orc_max(d, 0);
            c1[j] = orc_add(d, f);
            c3[j] = orc_max3(orc_add(p1[j], -go), orc_add(p2[j], -go), orc_add(p3[j], -g));
            c2[j] = orc_max3(orc_add(c1[j - 1], -go), orc_add(c2[j - 1], -g), orc_add(c3[j - 1], -go));
            if (mode == 1 && c1[j] > best) { best = c1[j]; bi = i; bj = j; }
        }
        int32_t* t;
        t = p1; p1 = c1; c1 = t; t = p2; p2 = c2; c2 = t; t = p3; p3 = c3; c3 = t;
    }
    if (mode == 0) {
        res->t1 = p1[n]; res->t2 = p2[n]; res->t3 = p3[n];
        res->score = orc_max3(res->t1, res->t2, res->t3);
        res->end_i = m; res->end_j = n;
        res->end_state = (res->t1 >= res->t2 && res->t1 >= res->t3) ? 1 : (res->t2 >= res->t1 && res->t2 >= res->t3) ? 2 : 3;
    } else {
        res->score = best; res->t1 = best; res->t2 = ORC_NEG; res->t3 = ORC_NEG;
        res->end_i = bi; res->end_j = bj; res->end_state = 1;
    }
    free(buf);
    return 0;
}

/* Batch helper for the CPU baseline: score-only over many pairs laid out back to back
 * (same layout as psa_batch: offsets + lengths).  Single-threaded; callers thread it. */
int orc_score_batch(const char* a, const int64_t* off_a, const int32_t* len_a, const char* b,
                    const int64_t* off_b, const int32_t* len_b, int64_t n_pairs, int g, int h, int mode,
                    orc_result* out) {
    for (int64_t p = 0; p < n_pairs; ++p) {
        int rc = orc_score_linear(a + off_a[p], b + off_b[p], len_a[p], len_b[p], g, h, mode, out + p);
        if (rc) return rc;
    }
    return 0;
}

int32_t orc_neg_inf(void) { return ORC_NEG; }
