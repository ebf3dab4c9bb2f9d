/* oracle/overlap_oracle.c -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * Plain-C, CPU restatement of the one hot path of GAPPadder's ContigsMerger that this
 * repository re-implements for the B200: the pairwise overlap dynamic programme
 * ContigsCompactor::Evaluate and its integer/double epilogue.  All file:line citations are
 * relative to /root/reference/ContigsCompactor-v0.2.0/ContigsMerger/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this file's shared object.  The product (libgappadder_b200.so, the gp_* host code) never
 * links, loads or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function below against
 * oracle/_ref/libcm_ref.so (the reference's own Evaluate compiled from /root/reference by
 * oracle/build_ref.sh) on seeded random pairs, and tests/golden/ holds vectors generated from
 * that library (tests/golden/make_golden.py) for machines where the reference is absent.
 *
 * Scoring is integer: match = +1 (ContigsCompactor.cpp:1596), mismatch = (int)scoreMismatch
 * (:1640, the assignment to `int matchScoreStep` truncates), indel = scoreIndel (:1654,:1660).
 * GAPPadder always passes -i1 -2.0 -i2 -2.0 (MergeContigs.py:85) so every table value is an
 * integer; a fractional -i2 is outside this oracle's (and the product's) contract.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define GPO_MAX_NEG_SCORE (-1000000000) /* ContigsCompactor.cpp:31 */

/* Result of one Evaluate call before the significance test. */
typedef struct {
    int32_t score;      /* scoreMax   (:1674) */
    int32_t row_end;    /* posRowEnd  (:1675) */
    int32_t col_end;    /* posColEnd  (:1676) */
    int32_t nclip;      /* nclip      (:1677) */
    int32_t tb_row;     /* tbCur.first  after the traceback loop (:1763-1812) */
    int32_t tb_col;     /* tbCur.second after the traceback loop */
    int32_t bcontained; /* :1814,:1834-1837 */
} gpo_dp_result;

/* Best-cell scan, ContigsCompactor.cpp:1679-1709.  `H(i,j)` is supplied through a callback-free
 * macro by both variants below; this helper works on a full table. */
static void scan_full(const int32_t *H, int m, int n, int maxclip, gpo_dp_result *r)
{
    int32_t scoreMax = GPO_MAX_NEG_SCORE;
    int posRowEnd = -1, posColEnd = -1, nclip = -1;
    size_t ld = (size_t)n + 1;
    for (int c = 0; c <= maxclip; c++) {
        for (int i = 0; i <= m; ++i) {              /* :1681-1693 column n-c, top to bottom */
            int icol = n - c;
            if (icol < 0) break;
            if (H[(size_t)i * ld + icol] > scoreMax) {
                scoreMax = H[(size_t)i * ld + icol]; posColEnd = icol; posRowEnd = i; nclip = c;
            }
        }
        for (int j = 0; j <= n; ++j) {              /* :1696-1708 row m-c, left to right */
            int irow = m - c;
            if (irow < 0) break;
            if (H[(size_t)irow * ld + j] > scoreMax) {
                scoreMax = H[(size_t)irow * ld + j]; posColEnd = j; posRowEnd = irow; nclip = c;
            }
        }
    }
    r->score = scoreMax; r->row_end = posRowEnd; r->col_end = posColEnd; r->nclip = nclip;
}

/* Literal restatement: full score table and full predecessor table, as the reference keeps them
 * (ContigsCompactor.cpp:1600-1671), followed by the scan (:1674-1709) and the predecessor walk
 * (:1736-1837).  O(m*n) memory -- for small inputs. Returns 0, or -1 on allocation failure. */
int gpo_evaluate_full(const char *s1, int m, const char *s2, int n,
                      int mismatch, int indel, int maxclip, gpo_dp_result *r)
{
    size_t ld = (size_t)n + 1, cells = ((size_t)m + 1) * ld;
    int32_t *H = (int32_t *)malloc(cells * sizeof(int32_t));
    int32_t *Pi = (int32_t *)malloc(cells * sizeof(int32_t));   /* tblTraceBack[i][j].first  */
    int32_t *Pj = (int32_t *)malloc(cells * sizeof(int32_t));   /* tblTraceBack[i][j].second */
    if (!H || !Pi || !Pj) { free(H); free(Pi); free(Pj); return -1; }
    for (int j = 0; j <= n; ++j) { H[j] = 0; Pi[j] = -1; Pj[j] = -1; }          /* :1611-1617 */
    for (int i = 1; i <= m; ++i) {
        H[(size_t)i * ld] = 0; Pi[(size_t)i * ld] = -1; Pj[(size_t)i * ld] = -1; /* :1628-1630 */
        for (int j = 1; j <= n; ++j) {
            int step = (s1[i - 1] == s2[j - 1]) ? 1 : mismatch;                  /* :1640-1644 */
            int32_t sc = H[(size_t)(i - 1) * ld + (j - 1)] + step;               /* :1651 diag  */
            int32_t pi = i - 1, pj = j - 1;
            if (sc < H[(size_t)(i - 1) * ld + j] + indel) {                      /* :1654 up    */
                sc = H[(size_t)(i - 1) * ld + j] + indel; pi = i - 1; pj = j;
            }
            if (sc < H[(size_t)i * ld + (j - 1)] + indel) {                      /* :1660 left  */
                sc = H[(size_t)i * ld + (j - 1)] + indel; pi = i; pj = j - 1;
            }
            H[(size_t)i * ld + j] = sc; Pi[(size_t)i * ld + j] = pi; Pj[(size_t)i * ld + j] = pj;
        }
    }
    scan_full(H, m, n, maxclip, r);
    /* predecessor walk, :1736,:1763-1812 (the string it builds is discarded by the reference) */
    int ti = r->row_end, tj = r->col_end;
    while (ti > 0 && tj > 0) {
        int pi = Pi[(size_t)ti * ld + tj], pj = Pj[(size_t)ti * ld + tj];
        if (!(tj > pj) && !(ti > pi)) break;                                     /* :1805-1809 */
        ti = pi; tj = pj;
    }
    r->tb_row = ti; r->tb_col = tj;
    int bc = 0;
    if (r->row_end + r->nclip == m && ti == 0) bc = 1;                           /* :1834 */
    if (r->col_end + r->nclip == n && tj == 0) bc = 1;                           /* :1836 */
    r->bcontained = bc;
    free(H); free(Pi); free(Pj);
    return 0;
}

/* Rolling restatement: two score rows plus a 2-bit "origin" per cell
 * (bit0: the predecessor walk from this cell ends in row 0, bit1: it ends in column 0) replace
 * the predecessor table; the scan keeps the last maxclip+1 rows and reads the last maxclip+1
 * columns as they are produced.  Same outputs as gpo_evaluate_full except tb_row/tb_col, which
 * are reported as 0 / >0 indicators only (0 when the walk ends on that border, 1 otherwise).
 * O((maxclip+2)*n) memory. */
int gpo_evaluate(const char *s1, int m, const char *s2, int n,
                 int mismatch, int indel, int maxclip, gpo_dp_result *r)
{
    int keep = maxclip + 1;                  /* rows m-maxclip..m are needed by the row scans */
    size_t ld = (size_t)n + 1;
    /* ring of kept rows for the scan */
    int32_t *ringH = (int32_t *)malloc((size_t)keep * ld * sizeof(int32_t));
    uint8_t *ringO = (uint8_t *)malloc((size_t)keep * ld);
    /* last maxclip+1 columns of every row */
    int32_t *colH = (int32_t *)malloc(((size_t)m + 1) * keep * sizeof(int32_t));
    uint8_t *colO = (uint8_t *)malloc(((size_t)m + 1) * keep);
    int32_t *prevH = (int32_t *)malloc(ld * sizeof(int32_t)), *curH = (int32_t *)malloc(ld * sizeof(int32_t));
    uint8_t *prevO = (uint8_t *)malloc(ld), *curO = (uint8_t *)malloc(ld);
    if (!ringH || !ringO || !colH || !colO || !prevH || !curH || !prevO || !curO) {
        free(ringH); free(ringO); free(colH); free(colO); free(prevH); free(curH); free(prevO); free(curO);
        return -1;
    }
    for (int i = 0; i <= m; ++i) {
        if (i == 0) {
            for (int j = 0; j <= n; ++j) { curH[j] = 0; curO[j] = (uint8_t)(1 | (j == 0 ? 2 : 0)); }
        } else {
            curH[0] = 0; curO[0] = 2;        /* (i,0), i>0: the walk stops at once, tbCur.second==0 */
            char a = s1[i - 1];
            for (int j = 1; j <= n; ++j) {
                int32_t sc = prevH[j - 1] + ((a == s2[j - 1]) ? 1 : mismatch);
                uint8_t o = prevO[j - 1];
                if (sc < prevH[j] + indel) { sc = prevH[j] + indel; o = prevO[j]; }
                if (sc < curH[j - 1] + indel) { sc = curH[j - 1] + indel; o = curO[j - 1]; }
                curH[j] = sc; curO[j] = o;
            }
        }
        if (i >= m - maxclip) {
            memcpy(ringH + (size_t)(i % keep) * ld, curH, ld * sizeof(int32_t));
            memcpy(ringO + (size_t)(i % keep) * ld, curO, ld);
        }
        for (int c = 0; c <= maxclip; ++c) {
            int j = n - c;
            if (j < 0) break;
            colH[(size_t)i * keep + c] = curH[j]; colO[(size_t)i * keep + c] = curO[j];
        }
        int32_t *t = prevH; prevH = curH; curH = t;
        uint8_t *u = prevO; prevO = curO; curO = u;
    }
    int32_t scoreMax = GPO_MAX_NEG_SCORE;
    int posRowEnd = -1, posColEnd = -1, nclip = -1;
    uint8_t org = 0;
    for (int c = 0; c <= maxclip; c++) {
        int icol = n - c;
        if (icol >= 0) {
            for (int i = 0; i <= m; ++i) {
                int32_t v = colH[(size_t)i * keep + c];
                if (v > scoreMax) { scoreMax = v; posColEnd = icol; posRowEnd = i; nclip = c; org = colO[(size_t)i * keep + c]; }
            }
        }
        int irow = m - c;
        if (irow >= 0) {
            const int32_t *rowH = ringH + (size_t)(irow % keep) * ld;
            const uint8_t *rowO = ringO + (size_t)(irow % keep) * ld;
            for (int j = 0; j <= n; ++j) {
                if (rowH[j] > scoreMax) { scoreMax = rowH[j]; posColEnd = j; posRowEnd = irow; nclip = c; org = rowO[j]; }
            }
        }
    }
    r->score = scoreMax; r->row_end = posRowEnd; r->col_end = posColEnd; r->nclip = nclip;
    r->tb_row = (org & 1) ? 0 : 1; r->tb_col = (org & 2) ? 0 : 1;
    int bc = 0;
    if (posRowEnd + nclip == m && (org & 1)) bc = 1;
    if (posColEnd + nclip == n && (org & 2)) bc = 1;
    r->bcontained = bc;
    free(ringH); free(ringO); free(colH); free(colO); free(prevH); free(curH); free(prevO); free(curO);
    return 0;
}

/* Thresholds of IsScoreSignificant, as the doubles main.cpp ends up with (CM/main.cpp:24-42,
 * 87-176: the -s/-c/-x/-y/-z values go through sscanf("%f") into a float and are widened). */
typedef struct {
    double fractionLossScore;          /* -s */
    double fracMinOverlap;             /* -c */
    double minOverlapLen;              /* -x */
    double minOverlapLenWithScaffold;  /* -z */
} gpo_thresholds;

/* ContigsCompactor::IsScoreSignificant, ContigsCompactor.cpp:1876-1976. Returns 0, 1 or 2. */
int gpo_is_score_significant(const gpo_thresholds *t, int scoreMax, int szSeq1, int szSeq2,
                             int rowStart, int colStart, int nclip)
{
    int szOverlap0 = szSeq1 < szSeq2 ? szSeq1 : szSeq2;                 /* :1894 */
    int szOverlap1 = szOverlap0, szOverlap2 = szOverlap0;
    if (rowStart + nclip == szSeq1) szOverlap1 = colStart;              /* :1896 */
    if (colStart + nclip == szSeq2) szOverlap2 = rowStart;              /* :1900 */
    int mn = szOverlap1 < szOverlap2 ? szOverlap1 : szOverlap2;
    int szOverlap = szOverlap0 < mn ? szOverlap0 : mn;                  /* :1904 */
    if (szOverlap < szSeq1 * t->fracMinOverlap && szOverlap < szSeq2 * t->fracMinOverlap)
        return 0;                                                       /* :1911 */
    const int MIN_ASM_EXT_LEN = 5;                                      /* :1916 */
    if (rowStart + nclip == szSeq1) { if (colStart + MIN_ASM_EXT_LEN - 1 >= szSeq2) return 0; }
    if (colStart + nclip == szSeq2) { if (rowStart + MIN_ASM_EXT_LEN - 1 >= szSeq1) return 0; }
    double scoreMinThres = szOverlap * (1 - t->fractionLossScore);      /* :1958 */
    if (scoreMax < scoreMinThres) return 0;                             /* :1960 */
    if (szOverlap < t->minOverlapLenWithScaffold) return 0;             /* :1972 */
    else if (szOverlap >= t->minOverlapLenWithScaffold && szOverlap < t->minOverlapLen) return 1;
    else return 2;
}

/* ContigsCompactorAction::SetMergedStringConcat, ContigsCompactor.cpp:108-153.
 * Writes the merged string (NUL-terminated) into out (capacity >= m+n+1); returns its length. */
int gpo_merged_concat(const char *s1, int m, const char *s2, int n,
                      int posRowEnd, int posColEnd, int nclip, int bcontained, char *out)
{
    int len = 0;
    if (bcontained && (posRowEnd + nclip) == m && m < n) {                       /* :116 */
        memcpy(out, s2, (size_t)n); len = n;
    } else if (bcontained && (posColEnd + nclip) == n && n < m) {                /* :122 */
        memcpy(out, s1, (size_t)m); len = m;
    } else if ((posRowEnd + nclip) == m) {                                       /* :131-139 */
        memcpy(out, s1, (size_t)(m - nclip)); len = m - nclip;
        memcpy(out + len, s2 + posColEnd, (size_t)(n - posColEnd)); len += n - posColEnd;
    } else {                                                                     /* :141-149 */
        memcpy(out, s2, (size_t)(n - nclip)); len = n - nclip;
        memcpy(out + len, s1 + posRowEnd, (size_t)(m - posRowEnd)); len += m - posRowEnd;
    }
    out[len] = 0;
    return len;
}

/* ContigsCompactorAction::IsContainment, ContigsCompactor.cpp:155-159. */
int gpo_is_containment(int m, int n, int posRowEnd, int posColEnd, int nclip, int bcontained)
{
    return bcontained && (((posRowEnd + nclip) == m && m < posColEnd) ||
                          ((posColEnd + nclip) == n && n < posRowEnd));
}

/* ContigsCompactorAction::GetOverlapSize, ContigsCompactor.h:51. */
int gpo_overlap_size(int m, int n, int nclip, int mergedLen) { return m + n - nclip - mergedLen; }

/* GetComplement + FastaSequence::RevsereComplement, GenSeqsUtils.cpp:24-61, fastareader.cpp
 * ("RevsereComplement"): N/n stay, ACGT (any case) complement to upper case, the rest -> 'N'. */
void gpo_revcomp(const char *s, int len, char *out)
{
    for (int i = 0; i < len; ++i) {
        char b = s[len - 1 - i], o;
        if (b == 'N' || b == 'n') o = b;
        else {
            char u = (b >= 'a' && b <= 'z') ? (char)(b - 32) : b;
            o = u == 'A' ? 'T' : u == 'T' ? 'A' : u == 'G' ? 'C' : u == 'C' ? 'G' : 'N';
        }
        out[i] = o;
    }
    out[len] = 0;
}

/* 2-bit code of SetKmerTypeForNtAt, KmerUtils.cpp:22-58: A/other=0, C=1, G=2, T=3. */
static inline uint64_t kcode(char nt)
{
    if (nt == 'c' || nt == 'C') return 1;
    if (nt == 'g' || nt == 'G') return 2;
    if (nt == 't' || nt == 'T') return 3;
    return 0;
}

/* k-mers of GetAllKmersFromSeq, KmerUtils.cpp:90-115.  The reference keeps the k-mer in the TOP
 * 2k bits of a uint64 and never clears what it shifts out of them (FormKmerTypeShortShift,
 * :75-88, only shifts left by 2 and sets position k-1), so bits above the k-mer simply fall off
 * the top: the value is (window code) << (64-2k), the same for equal windows.  Comparing window
 * codes is therefore equivalent. */
static uint64_t window_code(const char *p, int k)
{
    uint64_t v = 0;
    for (int i = 0; i < k; ++i) v = (v << 2) | kcode(p[i]);
    return v;
}

static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

/* Sorted k-mer codes of one node = the key set of mapKmerFreqInRepeat built by
 * QuickCheckerContigsMatch::Init, ContigsCompactor.cpp:2041-2056.  Returns count (len-k+1). */
static int node_kmers(const char *s, int len, int k, uint64_t *codes)
{
    int cnt = len - k + 1;
    uint64_t mask = k >= 32 ? ~(uint64_t)0 : (((uint64_t)1 << (2 * k)) - 1), v = 0;
    for (int i = 0; i < len; ++i) {
        v = ((v << 2) | kcode(s[i])) & mask;
        if (i >= k - 1) codes[i - k + 1] = v;
    }
    qsort(codes, (size_t)cnt, sizeof(uint64_t), cmp_u64);
    return cnt;
}

static int feasible(const uint64_t *codes, int cnt, const char *sj, int lenj, int k)
{
    const int lenContigLen = 30;                                   /* :2024 */
    for (int side = 0; side < 2; ++side) {
        const char *w = side == 0 ? sj : sj + lenj - lenContigLen; /* :2027,:2029 */
        for (int a = 0; a + k <= lenContigLen; ++a) {
            uint64_t q = window_code(w + a, k);
            if (bsearch(&q, codes, (size_t)cnt, sizeof(uint64_t), cmp_u64)) return 1;  /* :2072-2095 */
        }
    }
    return 0;
}

/* QuickCheckerContigsMatch(node i).IsMatchFeasible(node j), ContigsCompactor.cpp:1997-2095:
 * true iff some k-mer of the first 30 or of the last 30 bases of s_j occurs anywhere in s_i.
 * Requires leni >= k, lenj >= 30, k <= 30 (the reference exits / reads out of bounds otherwise). */
int gpo_quickcheck(const char *si, int leni, const char *sj, int lenj, int k)
{
    uint64_t *codes = (uint64_t *)malloc((size_t)(leni - k + 1) * sizeof(uint64_t));
    if (!codes) return -1;
    int cnt = node_kmers(si, leni, k, codes);
    int r = feasible(codes, cnt, sj, lenj, k);
    free(codes);
    return r;
}

/* The candidate list of the pairwise phase in the order the reference produces it with -t 1
 * (threadQuickCheck, ContigsCompactor.cpp:1068-1100: all i <= j INCLUDING j == i, row-major).
 * seqs[x]/lens[x] are the graph nodes [c0, c0_R, c1, c1_R, ...] (:794-799).  pairs receives
 * (i,j) int32 couples, at most cap of them; returns the number of candidates (may exceed cap). */
int64_t gpo_candidate_pairs(const char *const *seqs, const int32_t *lens, int nnodes, int k,
                            int32_t *pairs, int64_t cap)
{
    int64_t np = 0;
    int maxlen = 0;
    for (int i = 0; i < nnodes; ++i) if (lens[i] > maxlen) maxlen = lens[i];
    uint64_t *codes = (uint64_t *)malloc((size_t)(maxlen > 0 ? maxlen : 1) * sizeof(uint64_t));
    if (!codes) return -1;
    for (int i = 0; i < nnodes; ++i) {
        int cnt = node_kmers(seqs[i], lens[i], k, codes);
        for (int j = i; j < nnodes; ++j) {
            if (feasible(codes, cnt, seqs[j], lens[j], k)) {
                if (np < cap) { pairs[2 * np] = i; pairs[2 * np + 1] = j; }
                ++np;
            }
        }
    }
    free(codes);
    return np;
}

/* ---------------------------------------------------------------------------------------------
 * Flank placement (BASELINE.json configs[1]) -- PARITY UNPINNED.
 *
 * GAPPadder places the two flanks of a gap on every contig with `bwa mem -T <s> -a contigs.fa flanks.fa`
 * (/root/reference/pick_contigs.py:83-86) and keeps, per contig and side, strand, clip type, the number of
 * aligned columns and the leftmost contig position of the best record (:99-147).  The arithmetic is BWA's,
 * which is neither vendored nor version-pinned (README.md:28), and the reference holds no test or vector for it:
 * nothing here is checked against GAPPadder's own output.  What follows is the BUILDER-WRITTEN definition that
 * the B200 kernel (gappadder_b200/csrc/flank_place.cuh) is bit-exact against -- the semi-global form the
 * survey prescribes: the flank is aligned end to end inside the contig, with Evaluate's linear scoring
 * (match +1, mismatch, indel; ContigsCompactor.cpp:1596,1640,1654).
 *
 *   rows = flank s1 (m bases, consumed entirely), columns = contig s2 (n bases, free ends)
 *   H(0,j) = 0,  H(i,0) = i*indel,  H(i,j) = max(H(i-1,j-1) + s(i,j), H(i-1,j) + indel, H(i,j-1) + indel)
 *   score     = max_j H(m,j)
 *   col_end   = the smallest j that reaches it            (flank occupies contig[col_start, col_end), 0-based)
 *   col_start = the LARGEST start column over all optimal alignments that end in (m, col_end)
 *               (an order-free definition: start(i,j) = max of start over the predecessors that reach H(i,j);
 *                start(0,j) = j, start(i,0) = 0)
 * m = 0 gives score 0, col_end = 0, col_start = 0. */
typedef struct {
    int32_t score;
    int32_t col_start;
    int32_t col_end;
} gpo_place_result;

int gpo_semiglobal(const char *s1, int m, const char *s2, int n, int mismatch, int indel, gpo_place_result *r)
{
    int32_t *H = (int32_t *)malloc(((size_t)n + 1) * sizeof(int32_t));
    int32_t *S = (int32_t *)malloc(((size_t)n + 1) * sizeof(int32_t));
    if (!H || !S) { free(H); free(S); return -1; }
    for (int j = 0; j <= n; ++j) { H[j] = 0; S[j] = j; }
    for (int i = 1; i <= m; ++i) {
        int32_t dh = H[0], ds = S[0];              /* (i-1, j-1) */
        H[0] = i * indel; S[0] = 0;
        for (int j = 1; j <= n; ++j) {
            const int32_t uh = H[j], us = S[j];    /* (i-1, j) */
            int32_t h = dh + (s1[i - 1] == s2[j - 1] ? 1 : mismatch), s = ds;
            const int32_t hu = uh + indel, hl = H[j - 1] + indel;
            if (hu > h || (hu == h && us > s)) { h = hu; s = us; }
            if (hl > h || (hl == h && S[j - 1] > s)) { h = hl; s = S[j - 1]; }
            dh = uh; ds = us;
            H[j] = h; S[j] = s;
        }
    }
    int best = 0;
    for (int j = 1; j <= n; ++j) if (H[j] > H[best]) best = j;
    r->score = H[best]; r->col_end = best; r->col_start = S[best];
    free(H); free(S);
    return 0;
}
