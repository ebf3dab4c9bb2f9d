/* oracle/local_affine_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * CPU restatement of TERefiner's affine-gap local aligner as LocalAlignment::optAlign calls it
 * (/root/reference/TERefiner/algorithms/local_alignment.cpp:1036-1049: aln_stdaln(ref, sgmt, &aln_param_blast,
 * ALN_TYPE_LOCAL, 1)), written for reading rather than speed: whole-column arrays with one field per state, a full
 * traceback table for the global fill, no packing tricks.  It is pinned to the reference itself: tests/test_oracle_affine.py
 * checks it against the golden vectors made by the reference's own code (tests/golden/local_affine.json) and, when
 * oracle/_ref/libla_ref.so is present, against that code live.  Only tests/ may load it; the product never does.
 *
 * Letters: a/A c/C g/G t/T are the four bases, every other byte one class (aln_nt4_table :32-49).
 * Scores (aln_sm_blast :193-199, aln_param_blast :206): equal bases +1, unequal -3, anything against a non-base -2,
 * gap open 5, gap extension 2, band 50.
 * Domain: the reference's 16-bit rescaling (:573-588, :633-646) is not restated; callers keep
 * min(len1, len2) + 7 <= 32000. */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define LA_INF (-1073741823)                 /* MINOR_INF, stdaln.h:84 */
enum { ST_M = 0, ST_I = 1, ST_D = 2 };       /* FROM_M / FROM_I / FROM_D, stdaln.h:74-76 */

typedef struct { int match, mismatch, other, open, ext, band; } la_params;
static const la_params LA_BLAST = {1, -3, -2, 5, 2, 50};

static int la_class(unsigned char c)
{
    switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}
static int la_sub(const la_params *p, unsigned char a, unsigned char b)
{
    const int x = la_class(a), y = la_class(b);
    if (x == 4 || y == 4) return p->other;
    return x == y ? p->match : p->mismatch;
}

/* Pass 1, aln_local_core :565-609.  s1 is walked in the inner loop, s2 in the outer one; 1-based cells.
 * Returns the score; *end_i / *end_j = the first strict maximum in that order. */
static int la_forward(const la_params *p, const char *s1, int len1, const char *s2, int len2, int *end_i, int *end_j)
{
    const int q = p->open, r = p->ext;
    int *H = calloc((size_t)len1 + 1, sizeof(int));      /* H(i, j-1) while column j is computed, then H(i, j) */
    int *E = calloc((size_t)len1 + 1, sizeof(int));      /* horizontal gap state of row i at column j-1 */
    int best = 0;
    *end_i = *end_j = 0;
    for (int j = 1; j <= len2; ++j) {
        int diag = 0;            /* H(i-1, j-1) */
        int up = 0;              /* H(i-1, j) */
        int f = 0;               /* vertical gap state, left as it is while the cell above scores 0 (:590) */
        for (int i = 1; i <= len1; ++i) {
            const int left = H[i];
            int h = diag + la_sub(p, (unsigned char)s1[i - 1], (unsigned char)s2[j - 1]);
            if (h < 0) h = 0;
            if (up > 0) {
                f = (f > up - q) ? f - r : up - q - r;
                if (h < f) h = f;
            }
            int e = 0;
            if (left >= q + r + 1) {                     /* :594: the horizontal state is dropped below this */
                e = (E[i] > left - q) ? E[i] - r : left - q - r;
                if (h < e) h = e;
            }
            E[i] = e;
            diag = left;
            H[i] = h;
            up = h;
            if (best < h) { best = h; *end_i = i; *end_j = j; }
        }
    }
    free(H); free(E);
    return best;
}

/* Pass 2, :617-690: back from the end cell inside a band [lo+1, hi] of rows that is re-cut after every column.
 * Hs[k] / Es[k]: score of row k and horizontal gap state of row k as of the last column that wrote them (the reference
 * keeps both in eh[]: eh[k] >> 16 and eh[k+1] & 0xffff).  Returns 0, or 1 when the band collapses (hi < lo: the
 * reference's loop would leave its array). */
static int la_reverse(const la_params *p, const char *s1, const char *s2, int score_f, int end_i, int end_j,
                      int *start_i, int *start_j, int *score_r)
{
    const int q = p->open, r = p->ext, qr = q + r;
    int *Hs = calloc((size_t)end_i + 2, sizeof(int)), *Es = calloc((size_t)end_i + 2, sizeof(int));
    int sr = la_sub(p, (unsigned char)s1[end_i - 1], (unsigned char)s2[end_j - 1]);
    int si = end_i, sj = end_j, hi = end_i - 1, lo = end_i - 3 > 0 ? end_i - 3 : 0, bad = 0, done = 0;
    Hs[end_i] = qr + sr;                                  /* :627 */
    for (int j = end_j - 1; j >= 1 && !done; --j) {
        if (hi < lo) { bad = 1; break; }
        int above_new = 0;        /* this column's score of the row above (0 before the first cell, :648) */
        int above_old = Hs[hi + 1];
        int f = 0;
        Hs[hi + 1] = 0;           /* what :665 stores with the first cell */
        for (int i = hi; i > lo; --i) {
            const int left = Hs[i];
            int h = above_old + la_sub(p, (unsigned char)s1[i - 1], (unsigned char)s2[j - 1]);
            if (h < 0) h = 0;
            if (above_new > 0) {
                f = (f > above_new - q) ? f - r : above_new - qr;
                if (h < f) h = f;
            }
            int e = (Es[i] > left - q) ? Es[i] - r : left - qr;
            if (e < 0) e = 0;
            if (h < e) h = e;
            Es[i] = e;
            above_old = left;
            Hs[i] = h;
            above_new = h;
            if (sr < h) {
                sr = h; si = i; sj = j;
                if (sr - qr == score_f) { done = 1; break; }     /* :669 */
            }
        }
        if (done) break;
        Es[lo] = 0;                                       /* :674 clears the low half of eh[lo+1] */
        if (Hs[hi] <= qr) --hi;                           /* :676 */
        if (hi < 0) hi = 0;
        lo = si - (sj - j) - (sr + (sj - j) * p->match) / r - 1;      /* :678 */
        if (lo < 0) lo = 0;
    }
    free(Hs); free(Es);
    *start_i = si; *start_j = sj; *score_r = sr - qr;
    return bad;
}

/* Pass 3, aln_global_core :328-508 with gap_end < 0 (:719): banded global alignment of a[1..len1] x b[1..len2] with a full
 * traceback table; returns the score, *pi / *pj = the last element of the traced path (the cell next to the corner it leaves
 * from: (1,1), (0,1) or (1,0)). */
static int la_global(const la_params *p, const char *a, int len1, const char *b, int len2, int band, int *pi, int *pj)
{
    const int q = p->open, r = p->ext;
    int b1, b2;
    if (len1 > len2) { b1 = len1 - len2 + band; b2 = band; } else { b1 = band; b2 = len2 - len1 + band; }
    if (b1 > len1) b1 = len1;
    if (b2 > len2) b2 = len2;
    const size_t w = (size_t)len1 + 1;
    int *M = malloc(2 * w * sizeof(int)), *I = malloc(2 * w * sizeof(int)), *D = malloc(2 * w * sizeof(int));
    unsigned char *tM = calloc(w * ((size_t)len2 + 1), 1), *tI = calloc(w * ((size_t)len2 + 1), 1), *tD = calloc(w * ((size_t)len2 + 1), 1);
    for (size_t k = 0; k < 2 * w; ++k) M[k] = I[k] = D[k] = LA_INF;
    int cur = 0;
#define CM(i) M[cur * w + (i)]
#define CI(i) I[cur * w + (i)]
#define CD(i) D[cur * w + (i)]
#define LM(i) M[(1 - cur) * w + (i)]
#define LI(i) I[(1 - cur) * w + (i)]
#define LD(i) D[(1 - cur) * w + (i)]
#define T(t, i, j) t[(size_t)(j) * w + (i)]
    /* cell (i, j): the three states, each from the better of two or three predecessors; ties as :241-299 */
#define FILL_M(i, j) do { int v_, t_; \
        if (LM((i) - 1) >= LI((i) - 1)) { if (LM((i) - 1) >= LD((i) - 1)) { v_ = LM((i) - 1); t_ = ST_M; } else { v_ = LD((i) - 1); t_ = ST_D; } } \
        else { if (LI((i) - 1) > LD((i) - 1)) { v_ = LI((i) - 1); t_ = ST_I; } else { v_ = LD((i) - 1); t_ = ST_D; } } \
        CM(i) = v_ + la_sub(p, (unsigned char)a[(i) - 1], (unsigned char)b[(j) - 1]); T(tM, i, j) = (unsigned char)t_; } while (0)
#define FILL_I(i, j) do { if (LM(i) - q > LI(i)) { CI(i) = LM(i) - q - r; T(tI, i, j) = ST_M; } else { CI(i) = LI(i) - r; T(tI, i, j) = ST_I; } } while (0)
#define FILL_D(i, j) do { if (CM((i) - 1) - q > CD((i) - 1)) { CD(i) = CM((i) - 1) - q - r; T(tD, i, j) = ST_M; } else { CD(i) = CD((i) - 1) - r; T(tD, i, j) = ST_D; } } while (0)
    /* column 0 (:375-381) */
    CM(0) = 0;
    for (int i = 1; i < b1; ++i) { CM(i) = CI(i) = LA_INF; FILL_D(i, 0); }
    cur = 1 - cur;
    int j = 1;
    const int tmp_end = b2 < len2 ? b2 : len2 - 1;
    for (;;) {                                            /* columns whose band still touches row 0 (:383-423) */
        const int in_loop = j <= tmp_end, extra = !in_loop && j == len2 && b2 != len2 - 1 && j == tmp_end + 1;
        if (!in_loop && !extra) break;
        CM(0) = CD(0) = LA_INF; FILL_I(0, j);
        const int end = (j + b1 <= len1 + 1) ? (j + b1 - 1) : len1;
        for (int i = 1; i < end; ++i) { FILL_M(i, j); FILL_I(i, j); FILL_D(i, j); }
        FILL_M(end, j); FILL_D(end, j);
        if (j + b1 - 1 > len1) FILL_I(end, j); else CI(end) = LA_INF;
        cur = 1 - cur;
        ++j;
        if (extra) break;
    }
    for (; j <= len2 - b2 + 1; ++j) {                     /* both edges inside (:425-439) */
        const int e0 = j - b2, end = j + b1 - 1;
        CM(e0) = CI(e0) = CD(e0) = LA_INF;
        for (int i = e0 + 1; i < end; ++i) { FILL_M(i, j); FILL_I(i, j); FILL_D(i, j); }
        FILL_M(end, j); FILL_D(end, j); CI(end) = LA_INF;
        cur = 1 - cur;
    }
    for (; j <= len2; ++j) {                              /* the band reaches the last row (:441-470) */
        const int e0 = j - b2;
        CM(e0) = CI(e0) = CD(e0) = LA_INF;
        for (int i = e0 + 1; i <= len1; ++i) { FILL_M(i, j); FILL_I(i, j); FILL_D(i, j); }
        cur = 1 - cur;
    }
    /* traceback (:472-498) */
    int i = len1, state = ST_M, best = LM(len1);
    j = len2;
    if (LI(len1) > best) { best = LI(len1); state = ST_I; }
    if (LD(len1) > best) { best = LD(len1); state = ST_D; }
    for (;;) {
        const int ci = i, cj = j;
        const int prev = state == ST_M ? T(tM, i, j) : state == ST_I ? T(tI, i, j) : T(tD, i, j);
        if (state == ST_M) { --i; --j; } else if (state == ST_I) --j; else --i;
        state = prev;
        if (i == 0 && j == 0) { *pi = ci; *pj = cj; break; }
    }
    free(M); free(I); free(D); free(tM); free(tI); free(tD);
    return best;
}

/* out = {score, start1, end1, start2, end2}, 1-based, as aln_stdaln_aux reports them (:817-821).
 * Returns 0; 1 when nothing aligns (the reference reads path[-1]); 2 when the reverse band collapses. */
int lao_local_affine(const char *s1, int len1, const char *s2, int len2, int32_t *out)
{
    const la_params *p = &LA_BLAST;
    memset(out, 0, 5 * sizeof(int32_t));
    if (len1 <= 0 || len2 <= 0) { out[0] = -1; return 1; }             /* :545 */
    int end_i, end_j, start_i, start_j, score_r;
    const int score_f = la_forward(p, s1, len1, s2, len2, &end_i, &end_j);
    out[0] = score_f;
    if (score_f < 1) return 1;                                          /* :611 */
    out[2] = end_i; out[4] = end_j;
    if (la_reverse(p, s1, s2, score_f, end_i, end_j, &start_i, &start_j, &score_r)) return 2;
    const int n1 = end_i - start_i + 1, n2 = end_j - start_j + 1;
    const int widest = (n1 > n2 ? n1 : n2);                             /* :714-715: max extent + 1, extents are n - 1 */
    int score_g = 0, pi = 1, pj = 1;
    for (int band = p->band;; band <<= 1) {                             /* :717-725 */
        score_g = la_global(p, s1 + start_i - 1, n1, s2 + start_j - 1, n2, band, &pi, &pj);
        if (score_g == score_r || score_g == score_f) break;
        if (band > widest) break;
    }
    out[0] = (score_r > score_g && score_f > score_g) ? -1 : score_g;   /* :727-731 */
    pi += start_i - 1; pj += start_j - 1;                               /* :735-738 */
    out[1] = pi ? pi : 1; out[3] = pj ? pj : 1;                         /* :818-821 */
    return 0;
}
