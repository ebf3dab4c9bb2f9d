// overlap_wf16c.cuh -- packed 16-bit overlap-DP kernel with origin CERTIFICATES instead of tie tags
// ("certificate kernel", sm_100a).  The hot kernel for what GAPPadder produces: A/C/G/T sequences whose
// column sequence has at most 16382 bases.  Same contract as the other overlap kernels
// (ContigsCompactor::Evaluate before the significance test,
// ContigsCompactor-v0.2.0/ContigsMerger/ContigsCompactor.cpp:1596-1709,:1736-1837).
//
// Why.  The table kernel (overlap_wf16t.cuh) spends one of its three ALU-pipe instructions per cell
// pair on the z tag that makes a packed max reproduce the reference's diag > up > left tie rule
// (:1651-1665).  The tie rule never changes a score, only WHICH optimal predecessor walk the
// reference follows, and the walk is used for exactly one thing: whether it ends in row 0 or in
// column 0 (:1834-1837).  This kernel drops the tag (two ALU-pipe instructions per cell pair) and
// proves the walk's end instead:
//
//   V = 2*(P+1) + b,   P = H + (n - j) clamped at -1 (as in overlap_wf16.cuh),   b = one origin bit.
//
//   A packed max over candidates with equal P takes the larger b, so b(cell) = OR of b over ALL
//   optimal predecessor walks into the cell, whatever the tie rule.
//     system U:  b = 1 on row 0 (corner included), 0 on column 0.   b(best) = 0  <=>  every optimal
//                walk into the best cell ends in column 0 below the corner, so the reference's walk
//                does:  tbCur.second == 0, tbCur.first > 0.
//     system L:  b = 1 on column 0 (corner included), 0 on row 0.   b(best) = 0  <=>  every optimal
//                walk ends in row 0 right of the corner: tbCur.first == 0, tbCur.second > 0.
//     system C:  b = 0 on the corner (0,0) only, 1 on the rest of row 0 and column 0.   b(best) = 0  <=>  every
//                optimal walk ends in the corner: tbCur.first == 0 and tbCur.second == 0 (two sequences that
//                start with the same bases, e.g. a merged contig against the node it starts with).
//   Scores, and therefore the best cell (score, posRowEnd, posColEnd, nclip), are exact in every system.
//
// Per pair: a 16-base probe predicts the system (both sequences start with the same 16 bases -> C; the first
// bases of the column sequence occur in the row sequence -> the walk will end in column 0 -> U; the first bases
// of the row sequence occur in the column sequence -> L).  If the certificate fails (b = 1) the other systems
// are run, one after the other, on the sub-table that ends at the best cell (rows 1..posRowEnd, columns
// 1..posColEnd: usually a sliver, because a wrong guess means the best cell sits in the opposite corner) and
// report that one cell's bit.  If all three fail (a genuine tie between walks that end on different borders)
// the pair is appended to a retry list and recomputed by an exact kernel (overlap_wf16t.cuh or
// overlap_wf32.cuh) launched behind this one.  Results are therefore always the reference's; the
// certificates only decide which kernel produces them.
//
// The wavefront machinery is the table kernel's: one warp per pair, strips of 64*K rows, lane l owns 2K
// consecutive rows (K lo rows in the low halves, K hi rows one column behind in the high halves),
// per-warp shared-memory increment table (16 column-symbol combinations x K words per lane), a ring of
// per-column data, lanes 3 columns apart with the shuffle issued one step ahead, 32-step blocks
// instantiated for {edge, filter} flags, candidate filter under the column potential.
//
// Host/device: the per-lane arithmetic is __host__ __device__; tests/emulate_wf16.cu runs it on the CPU
// against the oracle.
#pragma once
#include <type_traits>
#include "overlap_wf16t.cuh"

namespace gp {

constexpr uint32_t WF16C_MAX_N = 16382;           // 2*(n+1)+1 <= 32767
constexpr int WF16C_THREADS = 128;                // 4 warps per CTA
constexpr int WF16C_CTAS_PER_SM = 3;              // 12 warps per SM: 3 x (4 x 17.6 KB) of shared memory
constexpr int WF16C_RING = 128;
constexpr int WF16C_SKEW = 3;
constexpr int WF16C_TAB_WORDS = 16 * 8 * 32;
constexpr int WF16C_ORING = WF16C_RING + 32;        // bottom-row ring: addressed like the input ring, a block never wraps
constexpr int WF16C_WARP_WORDS = WF16C_TAB_WORDS + 2 * WF16C_RING + WF16C_ORING;
constexpr size_t WF16C_SMEM_BYTES = (size_t)(WF16C_THREADS / 32) * WF16C_WARP_WORDS * sizeof(uint32_t);
constexpr int WF16C_PROBE = 16;                   // bases of the orientation probe (two packed words)
// Second value layout ("free moves"): V = 2*Q + b with Q = H + 2*i_rel + 2*j + n + 1 (i_rel: row inside the strip).
// Under it the up and left moves (-2 each) add nothing and the diagonal adds 2*(s+4), so a cell is ONE 3-input
// max after the diagonal add.  Only for the standard scores (-2, -2) and columns short enough for 15 bits:
// Q <= n + 2*512 + 2*n + n + 1.
constexpr uint32_t WF16C_POT2_MAX_N = 3800;         // 2*(4n + 1025) + 1 + 10 <= 32767

inline bool wf16c_pair_ok(uint32_t m, uint32_t n) { return m >= 1 && n >= 1 && n <= WF16C_MAX_N && m <= 0xffffff; }

struct Wf16cParams {
    uint32_t inc_mism;              // 16-bit diagonal increment of a mismatch, 2*(X-1) (a match adds 0)
    uint32_t gup, gleft;            // packed up / left increments under the column potential
    int32_t max_clip;
    int32_t std_scores;             // mismatch == -2 && indel == -2: the kernel with immediate operands
    int32_t pot2;                   // free-moves layout (set per launch: std scores and every n <= WF16C_POT2_MAX_N)
    uint32_t inc2_match, inc2_mism; // its diagonal increments, 2*(s - 2*indel)
};

inline Wf16cParams wf16c_make_params(int mismatch, int indel, int max_clip)
{
    Wf16cParams p;
    p.inc_mism = (uint32_t)((mismatch - 1) * 2) & 0xffffu;
    auto pk = [](int v) { uint32_t h = (uint32_t)(v * 2) & 0xffffu; return h | (h << 16); };
    p.gup = pk(indel);
    p.gleft = pk(indel - 1);
    p.max_clip = max_clip;
    p.std_scores = (mismatch == -2 && indel == -2) ? 1 : 0;
    p.pot2 = 0;
    p.inc2_match = (uint32_t)(2 * (1 - 2 * indel)) & 0xffffu;
    p.inc2_mism = (uint32_t)(2 * (mismatch - 2 * indel)) & 0xffffu;
    return p;
}

// One DP pass: the sub-table rows 1..m, columns 1..n of a pair in one certificate system.
//   scan mode: the reference's best-cell scan (C = max_clip) over the whole table of the pair;
//   cell mode: only cell (m, n) is reported (its score must be s_cell).
constexpr int WF16C_SYS_U = 0, WF16C_SYS_L = 1, WF16C_SYS_C = 2;
struct Wf16cPass {
    int m, n, C;
    int sys;                        // WF16C_SYS_*
    uint32_t brow, bcol, bcorner;   // origin bit on row 0 (j >= 1) / on column 0 (i >= 1) / on the corner (0,0)
    bool cell;
    int s_cell;
    uint32_t gup, gleft;
    bool pot2;                      // free-moves layout (see WF16C_POT2_MAX_N); rows are then strip-relative (irel)
    // Transposed pair: the computed table's rows are the reference's COLUMN sequence and vice versa, cell (i,j) here is
    // the reference's cell (j,i).  Scores are symmetric; what is not -- the scan order (ranks) and which border is
    // "row 0" -- is translated where it is used: cell_rank with swapped arguments, the first scanned cell (0,N) sits at
    // (m,0) here, and certified origins swap ROW0 <-> COL0.  The tie rule (diag > up > left) is not symmetric either, but
    // this kernel never follows it: certificates hold for ALL optimal walks.  The host orients a pair so that the
    // sequence that fills 512-row strips better is the row sequence.
    bool tr;
    // V of a cell with score H and origin bit 0 / H of a cell value
    GP_HD int v_of(int H, int irel, int j) const { return pot2 ? 2 * (H + 2 * irel + 2 * j + n + 1) : 2 * (H + n - j + 1); }
    GP_HD int h_of(int vv, int irel, int j) const { return pot2 ? (vv >> 1) - (2 * irel + 2 * j + n + 1) : (vv >> 1) - 1 - (n - j); }
    GP_HD uint32_t v_col0(int i, int irel) const { return (uint32_t)v_of(0, irel, 0) + (i == 0 ? bcorner : bcol); }
    GP_HD uint32_t v_row0(int j) const { return (uint32_t)v_of(0, 0, j) + (j == 0 ? bcorner : brow); }
};

GP_HD Wf16cPass wf16c_make_pass(int m, int n, const Wf16cParams& P, int sys, bool cell, int s_cell, bool tr = false)
{
    Wf16cPass g;
    g.m = m; g.n = n; g.C = cell ? 0 : P.max_clip;
    g.sys = sys;
    g.brow = sys == WF16C_SYS_L ? 0u : 1u;
    g.bcol = sys == WF16C_SYS_U ? 0u : 1u;
    g.bcorner = sys == WF16C_SYS_C ? 0u : 1u;
    g.cell = cell; g.s_cell = s_cell;
    g.gup = P.gup; g.gleft = P.gleft;
    g.pot2 = P.pot2 != 0;
    g.tr = tr;
    return g;
}

// Boundary-line word of column j (1..n+1): V(0,j) in the low half, the symbol combination
// c_j + 4*c_{j-1} (c_0 = c_{n+1} = 0) above it.
GP_HD uint32_t wf16c_line_word(const Wf16cPass& g, int j, uint32_t cj, uint32_t cjm1)
{
    return g.v_row0(j <= g.n ? j : g.n) | (((cj & 3u) | ((cjm1 & 3u) << 2)) << 16);
}

GP_HD uint32_t wf16c_table_word(uint32_t row_lo, uint32_t row_hi, uint32_t ca, uint32_t cb, const Wf16cParams& P)
{
    const uint32_t mt = P.pot2 ? P.inc2_match : 0u, mm = P.pot2 ? P.inc2_mism : P.inc_mism;
    return (row_lo == ca ? mt : mm) | ((row_hi == cb ? mt : mm) << 16);
}

template <int K>
struct Lane16c {
    uint32_t W[K];       // (row k @ col j | row K+k @ col j-1 << 16)
    uint32_t up0_prev;   // previous step's `up` of W[0] == this step's diagonal of W[0]
};

// irel_top: the lane's first row minus one, relative to the strip (itop - i0).
template <int K>
GP_HD void lane16c_begin(Lane16c<K>& st, const Wf16cPass& g, int itop, int irel_top)
{
#pragma unroll
    for (int k = 0; k < K; ++k)                            // column 0 of rows itop+1+k (lo) and itop+1+K+k (hi)
        st.W[k] = g.v_col0(itop + 1 + k, irel_top + 1 + k) | (g.v_col0(itop + 1 + K + k, irel_top + 1 + K + k) << 16);
    st.up0_prev = g.v_col0(itop, irel_top) | (g.v_col0(itop + K, irel_top + K) << 16);
}

// After a lane's first step the hi group has "computed" column 0: put the boundary back.
template <int K>
GP_HD void lane16c_fix_first(Lane16c<K>& st, const Wf16cPass& g, int itop, int irel_top)
{
#pragma unroll
    for (int k = 0; k < K; ++k) st.W[k] = (st.W[k] & 0xffffu) | (g.v_col0(itop + 1 + K + k, irel_top + 1 + K + k) << 16);
}

// One step: VIADD.16x2 + 2 x VIADDMNMX.S16x2 per register, nothing else.  Two phases so that the kernel can
// reload the increment registers between them (the diagonal adds are the only readers of `inc`):
//   lane16c_diag   d[k] = diagonal + increment, from the values before the step;
//   lane16c_chain  the left and up candidates and the maxima, the `up` chain through the K registers.
template <int K>
GP_HD void lane16c_diag(const Lane16c<K>& st, const uint32_t (&inc)[K], uint32_t (&d)[K])
{
    d[0] = p_add2(st.up0_prev, inc[0]);
#pragma unroll
    for (int k = 1; k < K; ++k) d[k] = p_add2(st.W[k - 1], inc[k]);
}
GP_HD uint32_t p_max3_relu2(uint32_t a, uint32_t b, uint32_t c)           // max(a, b, c, 0) per signed 16-bit half
{
#if defined(__CUDA_ARCH__)
    return __vimax3_s16x2_relu(a, b, c);
#else
    return p_max2(p_max2(p_max2(a, b), c), 0u);
#endif
}
template <int K, bool POT2 = false>
GP_HD void lane16c_chain(Lane16c<K>& st, uint32_t recv, const uint32_t (&d)[K], uint32_t gup, uint32_t gleft)
{
    const uint32_t up0 = p_prmt(recv, st.W[K - 1], 0x5432u);   // (recv.hi16 , old W[K-1].lo16)
    st.up0_prev = up0;
    uint32_t up = up0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        uint32_t w;
        if (POT2) {
            w = p_max3_relu2(st.W[k], up, d[k]);                // left and up moves are free under this layout
        } else {
            const uint32_t t = p_addmax2(st.W[k], gleft, d[k]);
            w = p_addmax2_relu(up, gup, t);
        }
        up = w;
        st.W[k] = w;
    }
}
template <int K>
GP_HD void lane16c_step(Lane16c<K>& st, uint32_t recv, const uint32_t (&inc)[K], uint32_t gup, uint32_t gleft, bool pot2 = false)
{
    uint32_t d[K];
    lane16c_diag<K>(st, inc, d);
    if (pot2) lane16c_chain<K, true>(st, recv, d, gup, gleft);
    else lane16c_chain<K, false>(st, recv, d, gup, gleft);
}

template <int K>
GP_HD uint32_t lane16c_max(const Lane16c<K>& st)
{
    uint32_t a = st.W[0];
    if (K == 2) a = p_max2(a, st.W[1]);
    if (K >= 4) a = p_max3_2(p_max2(a, st.W[1]), st.W[2], st.W[3]);
    if (K == 8) a = p_max3_2(p_max3_2(a, st.W[4], st.W[5]), st.W[6], st.W[7]);
    return a;
}
// Free-moves layout: row k of either group sits 2*k higher than row 0, i.e. 4*k in V.  The offsets go on with independent
// packed adds (the FMA-heavy pipe, which has room) and the maximum is a tree of 3-input maxima: three levels deep instead
// of a chain of K add-max instructions on the ALU pipe.  A filter step hangs off the DP step's own K-deep max chain, so
// the chain doubled the latency of every filter block -- 215 against 88 clocks per step for a warp that has its scheduler
// to itself (the tail of a relax launch; tools/lone_pair_bench.py --trace) -- and cost four more ALU slots per step.
template <int K>
GP_HD uint32_t lane16c_max_pot2(const Lane16c<K>& st)
{
    uint32_t t[K];
    t[0] = st.W[0];
#pragma unroll
    for (int k = 1; k < K; ++k) t[k] = p_add2(st.W[k], (uint32_t)((-4 * k) & 0xffff) * 0x00010001u);
    uint32_t a = t[0];
    if (K == 2) a = p_max2(a, t[1]);
    if (K >= 4) a = p_max2(p_max3_2(a, t[1], t[2]), t[3]);
    if (K == 8) a = p_max3_2(a, p_max3_2(t[4], t[5], t[6]), t[7]);
    return a;
}

// Candidate filter.  nthr = -(the V a cell has when H = 1) for the lane's first lo row at column j (low half)
// and its first hi row at column j-1 (high half): acc = max(V) + nthr = 2*(H-1) + b of the lane's best
// cell, compared with thrS = 2*(max(S,1)-1).
template <int K>
GP_HD uint32_t wf16c_nthr(const Wf16cPass& g, int j, int irel_top)
{
    const uint32_t lo = (uint32_t)(-g.v_of(1, irel_top + 1, j)) & 0xffffu;          // other j wrap mod 2^16
    const uint32_t hi = (uint32_t)(-g.v_of(1, irel_top + 1 + K, j - 1)) & 0xffffu;
    return lo | (hi << 16);
}
constexpr uint32_t WF16C_UNARMED = 0x7fff7fffu;   // threshold no acc reaches (acc <= 2*n + 1 < 32767)
constexpr uint32_t WF16C_NSTEP = 0x00020002u;          // nthr of the next column
constexpr uint32_t WF16C_NSTEP_POT2 = 0xfffcfffcu;
GP_HD uint32_t wf16c_filter_thr(int S) { const uint32_t t = (uint32_t)(2 * ((S > 1 ? S : 1) - 1)); return t | (t << 16); }

// Exact scan of the cells a lane holds after a step (lo column j, hi column j-1).  Scan mode: a cell can
// only change `best` if it scores at least the lane's best score (and at least 1: the initial best is
// cell (0,n) with H = 0 and rank 0).  That test comes first and is PACKED: one signed compare per register
// (two cells) against the threshold of both columns; the rank arithmetic only runs for the cells that pass
// (one or two per call as a rule).  The result is an order-independent maximum over keys, so the order in
// which cells are visited does not matter.  Cell mode: only cell (m, n) counts.  The key's origin field holds b.
GP_HD void p_ge2(uint32_t a, uint32_t b, bool& hi, bool& lo)      // per half: a >= b (signed); one VIMNMX.S16x2 with predicates
{
#if defined(__CUDA_ARCH__)
    (void)__vibmax_s16x2(a, b, &hi, &lo);
#else
    lo = (int16_t)a >= (int16_t)b;
    hi = (int16_t)(a >> 16) >= (int16_t)(b >> 16);
#endif
}
// Layout arithmetic without the pass structure (the cold functions take scalars): pot2 < 0 is the column-potential
// layout, otherwise pot2 = irel_top, the lane's first row minus one relative to its strip (free-moves layout).
GP_HD int wf16c_v_of(int H, int n, int pot2, int rk, int j) { return pot2 >= 0 ? 2 * (H + 2 * (pot2 + rk) + 2 * j + n + 1) : 2 * (H + n - j + 1); }
GP_HD int wf16c_h_of(int vv, int n, int pot2, int rk, int j) { return pot2 >= 0 ? (vv >> 1) - (2 * (pot2 + rk) + 2 * j + n + 1) : (vv >> 1) - 1 - (n - j); }
// V of a cell with H = max(S,1) in register k: lo row (column j) | hi row (column j-1)
GP_HD uint32_t wf16c_scan_thr(int S, int n, int pot2, int K, int k, int j)
{
    const int s1 = S > 1 ? S : 1;
    const int lo = wf16c_v_of(s1, n, pot2, 1 + k, j), hi = wf16c_v_of(s1, n, pot2, 1 + K + k, j - 1);
    return (uint32_t)(lo < 0x7fff ? lo : 0x7fff) | ((uint32_t)(hi < 0x7fff ? hi : 0x7fff) << 16);
}
// The exact scan takes K at run time and loops: it is cold code, and four unrolled copies of it (one per strip
// height) made up a third of the kernel's instructions -- enough to push the kernel out of the instruction cache.
GP_HD long long lane16c_scan_rt(const uint32_t* W, int K, int m, int n, int C, bool cell, int itop, int j, int s_floor, int pot2, long long best, bool tr = false)
{
    if (cell) {
        for (int half = 0; half < 2; ++half) {
            if (j - half != n) continue;
            const int k = m - itop - 1 - half * K;                 // the register that holds row m, if this lane does
            if (k >= 0 && k < K) {
                const int vv = half ? (int)(W[k] >> 16) : (int)(W[k] & 0xffffu);
                const long long key = make_key(wf16c_h_of(vv, n, pot2, 1 + k + half * K, n), 0u, (uint32_t)vv & 1u);
                best = key > best ? key : best;
            }
        }
        return best;
    }
    // s_floor: a score some candidate of the pair is already known to reach (the warp's shared best); cells
    // below it cannot win, cells that equal it can (by rank).
    const bool vlo = (uint32_t)(j - 1) < (uint32_t)n, vhi = (uint32_t)(j - 2) < (uint32_t)n;   // columns inside the table
    int sthr = (int)(best >> 32);
    sthr = sthr > s_floor ? sthr : s_floor;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int k = 0; k < K; ++k) {
        bool hit[2];
        const uint32_t w = W[k];
        p_ge2(w, wf16c_scan_thr(sthr, n, pot2, K, k, j), hit[1], hit[0]);
        hit[0] = hit[0] && vlo; hit[1] = hit[1] && vhi;
        if (!(hit[0] || hit[1])) continue;
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
        for (int half = 0; half < 2; ++half) {
            if (!hit[half]) continue;
            const int jh = j - half;
            const int vv = half ? (int)(w >> 16) : (int)(w & 0xffffu);
            const int i = itop + 1 + k + half * K;
            const uint32_t rk = i <= m ? (tr ? cell_rank(jh, i, n, m, C) : cell_rank(i, jh, m, n, C)) : RANK_MAX + 1u;
            if (rk <= RANK_MAX) {
                const long long key = make_key(wf16c_h_of(vv, n, pot2, 1 + k + half * K, jh), rk, (uint32_t)vv & 1u);
                if (key > best) {
                    best = key;
                    const int sb = (int)(best >> 32);
                    sthr = sb > sthr ? sb : sthr;
                }
            }
        }
    }
    return best;
}
template <int K>
GP_HD long long lane16c_scan_mn(const uint32_t (&W)[K], int m, int n, int C, bool cell, int itop, int j, int s_floor, int pot2, long long best, bool tr = false)
{
    return lane16c_scan_rt(W, K, m, n, C, cell, itop, j, s_floor, pot2, best, tr);
}
template <int K>
GP_HD long long lane16c_scan(const uint32_t (&W)[K], const Wf16cPass& g, int itop, int irel_top, int j, long long best)
{
    return lane16c_scan_mn<K>(W, g.m, g.n, g.C, g.cell, itop, j, 0, g.pot2 ? irel_top : -1, best, g.tr);
}

// Deferred exact scans.  Along an alignment path inside the candidate zone a lane's best cell gains a point at
// every column, so resolving every fired step costs one exact scan per column although only the last one can
// win.  A fired step is instead kept as a snapshot of the lane's registers together with its SCORE: the exact
// maximum of H over the step's candidate cells.  A later step whose score is strictly higher replaces it unseen
// (every cell of the old step loses to the new step's best cell, whatever the ranks); an equal or lower score
// resolves the old snapshot first.  Whatever is pending at the end of the strip is resolved then.  Results are
// the same maxima over the same keys.
//   A step is deferrable when both its columns are inside the table and on the same side of the first candidate
//   column (jswitch): then the candidate cells are whole rows, rows 1..m when both columns are candidate columns,
//   rows m-C..m otherwise, and the score is a maximum over the lane's registers masked to those rows.
//   score form: a = 2*(H-1) + 1 (the filter's accumulator with the origin bit forced).
constexpr int WF16C_NO_SNAP = -(1 << 30);
GP_HD int wf16c_score_of(int a) { return (a >> 1) + 1; }      // H of a step score
GP_HD bool wf16c_deferrable(int n, int C, bool cell, int j)
{
    const int jswitch = n - C > 1 ? n - C : 1;
    return !cell && j >= 2 && j <= n && j != jswitch;
}
// Exact score of a deferrable step; WF16C_NO_SNAP when the lane holds no candidate cell in it.
GP_HD int wf16c_exact_step_score_rt(const uint32_t* W, int K, int m, int n, int C, int itop, int j, int pot2)
{
    const int jswitch = n - C > 1 ? n - C : 1;
    const int rlo = j - 1 >= jswitch ? 1 : m - C;              // candidate rows rlo..m
    int best = WF16C_NO_SNAP;                                   // 2*(H-1) + b of the best candidate cell
    const int v1lo = wf16c_v_of(1, n, pot2, 1, j), v1hi = wf16c_v_of(1, n, pot2, 1 + K, j - 1);
    const int per_row = pot2 >= 0 ? 4 : 0;                      // free-moves layout: row k sits 4*k higher in V
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int k = 0; k < K; ++k) {
        const int il = itop + 1 + k, ih = il + K;
        const int a = (int)(W[k] & 0xffffu), b = (int)(W[k] >> 16);
        if (il >= rlo && il <= m) { const int h = a - v1lo - per_row * k; best = h > best ? h : best; }
        if (ih >= rlo && ih <= m) { const int h = b - v1hi - per_row * k; best = h > best ? h : best; }
    }
    return best == WF16C_NO_SNAP ? best : (best | 1);
}
template <int K>
GP_HD int wf16c_exact_step_score(const uint32_t (&W)[K], int m, int n, int C, int itop, int j, int pot2)
{
    return wf16c_exact_step_score_rt(W, K, m, n, C, itop, j, pot2);
}

// Initial `best` of a pass.  Scan mode: cell (0,n), the first cell the reference scans (rank 0, H = 0);
// its walk ends where it starts, in row 0, so its b is the row-0 bit (1 for the corner).  Cell mode: a
// sentinel just below the expected score.
GP_HD long long wf16c_initial_best(const Wf16cPass& g)
{
    if (g.cell) return make_key(g.s_cell - 1, RANK_MAX, 0u);
    if (g.tr) return make_key(0, 0u, g.m == 0 ? g.bcorner : g.bcol);      // the reference's (0,N) is cell (m,0) here
    return make_key(0, 0u, g.n == 0 ? g.bcorner : g.brow);
}

// What a finished pass proves.  Returns the origin flags (FLAG_ROW0 | FLAG_COL0) when the pass certifies
// them, 0 when it does not.  `key` is the pass's warp-wide best.
GP_HD uint32_t wf16c_certified_origin(const Wf16cPass& g, long long key)
{
    const uint32_t lo = (uint32_t)(key & 0xffffffffll);
    uint32_t f;                                                                 // in the computed table's terms
    if (!g.cell && (int)(key >> 32) == 0 && (RANK_MAX - (lo >> 2)) == 0u)     // the best cell is the first scanned cell itself
        f = g.tr ? (FLAG_COL0 | (g.m == 0 ? FLAG_ROW0 : 0u)) : (FLAG_ROW0 | (g.n == 0 ? FLAG_COL0 : 0u));
    else if (g.cell && (int)(key >> 32) != g.s_cell) return 0u;                 // cannot happen; be safe
    else if (lo & 1u) return 0u;
    // system U proves column 0, system L proves row 0, system C proves the corner
    else f = g.sys == WF16C_SYS_U ? FLAG_COL0 : g.sys == WF16C_SYS_L ? FLAG_ROW0 : (FLAG_ROW0 | FLAG_COL0);
    if (g.tr) f = ((f & FLAG_ROW0) ? FLAG_COL0 : 0u) | ((f & FLAG_COL0) ? FLAG_ROW0 : 0u);   // to the reference's orientation
    return f;
}

// The systems to try after `first` failed, in order.
GP_HD int wf16c_next_system(int first, int attempt)      // attempt = 1, 2
{
    const int order[3][2] = {{WF16C_SYS_L, WF16C_SYS_C}, {WF16C_SYS_U, WF16C_SYS_C}, {WF16C_SYS_U, WF16C_SYS_L}};
    return order[first][attempt - 1];
}

// Strip plan of the certificate kernel: wf16_next_strip's, except that a last strip of up to 64 rows also runs as a
// 128-row strip -- the 64-row variant would be 15 KB more code for 1 % of the work, and the kernel sits just below
// the size at which instruction fetch starts to stall (measured: 105 KB fine, 118 KB 13 % slower).
GP_HD Wf16Strip wf16c_next_strip(int i0, int m, int C)
{
    Wf16Strip s = wf16_next_strip(i0, m, C);
    if (s.rows < 128) s.rows = 128;
    return s;
}

#if defined(__CUDACC__)
// ---- device side ------------------------------------------------------------------------------

// Diagnostic build only (-DGP_WF16C_TRACE, `make trace`; tools/lone_pair_bench.py --trace): lane 0 of every warp stamps the
// phases of a pair -- probe, boundary line, every strip's set-up / first block / end -- with the SM clock.
#if defined(GP_WF16C_TRACE)
__device__ unsigned long long gp_trace_buf[3 * (1 << 16)];
__device__ unsigned int gp_trace_n = 0;
__device__ __forceinline__ void gp_trace(uint32_t tag, uint32_t val)
{
    if ((threadIdx.x & 31) == 0) {
        const unsigned int i = atomicAdd(&gp_trace_n, 1u);
        if (i < (1u << 16)) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
            gp_trace_buf[3 * i] = t;
            gp_trace_buf[3 * i + 1] = ((unsigned long long)blockIdx.x << 32) | ((unsigned long long)(threadIdx.x >> 5) << 16) | tag;
            gp_trace_buf[3 * i + 2] = val;
        }
    }
}
#define GP_TRACE(tag, val) gp_trace(tag, val)
#else
#define GP_TRACE(tag, val) ((void)0)
#endif

struct Wf16cWarp {                      // warp-uniform state of one pass
    const uint32_t* packed;             // the column sequence's table
    const uint32_t* packed_row;         // the row sequence's (the same table, or the relax chain's arena of merged contigs)
    PairDesc pd;
    Wf16cPass g;
    uint32_t* bnd;                      // boundary line in global scratch (see wf16c_line_word)
    uint32_t* smem;                     // this warp's WF16C_WARP_WORDS words of shared memory
    // team mode (the warps of a CTA share one pair, strip s goes to warp s % TEAM and follows strip s-1 through
    // the boundary line): shared-memory address of the TEAM progress words, this warp's index in the team and
    // the index of the strip it is about to run
    uint32_t prog_addr;
    int team_warp;
    int strip_idx;
};

// Progress word of a team warp: (strip index << 15) | columns of that strip's bottom row already in the boundary
// line.  Monotone over a pass (a warp's strips have increasing indices), so "strip s has published column c" is
// one unsigned compare whatever the producer is doing by now.
__device__ __forceinline__ uint32_t lds32_volatile(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void sts32_volatile(uint32_t addr, uint32_t v)
{
    asm volatile("st.volatile.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}

template <int K> struct CVals { uint32_t W[K]; };

// The cold side of the candidate bookkeeping: ONE function for every strip height (K at run time; scalars and
// pointers only).  `bufs` is the lane's scratch in local memory (so that the steady loops carry no extra live
// registers): two buffers of K registers and three words of state -- the pending step's lo column, its score, and
// which buffer holds it.  The caller has put the fired step's registers into the OTHER buffer; deferring a step
// flips the buffers instead of copying.
//   flush == 0: the lane fired at lo column j; acc is the filter's accumulator of that step.  A deferrable step
//               becomes the pending step (the old pending step is scanned first when it ties or beats the new one and
//               can still win); any other step is scanned now.
//   flush != 0: end of the strip, the pending step is scanned.
// Returns the lane's best key.
__device__ __noinline__ long long wf16c_cold(uint32_t* bufs, int K, int flush, uint32_t acc, int m, int n, int C_or_cell,
                                             int itop, int j, int s_floor, int pot2, long long best)
{
    const bool cell = C_or_cell < 0;
    const int C = cell ? 0 : C_or_cell;
    const bool tr = (flush & 2) != 0;                         // bit 1 of `flush`: transposed pair (ranks with swapped coordinates)
    flush &= 1;
    uint32_t* meta = bufs + 2 * K;
    const uint32_t idx = meta[2];
    const uint32_t* pend = bufs + idx * K;
    const uint32_t* cur = bufs + (1u - idx) * K;
    if (flush) return lane16c_scan_rt(pend, K, m, n, C, false, itop, (int)meta[0], s_floor, pot2, best, tr);
    if (!wf16c_deferrable(n, C, cell, j)) return lane16c_scan_rt(cur, K, m, n, C, cell, itop, j, s_floor, pot2, best, tr);
    const int jswitch = n - C > 1 ? n - C : 1;
    const int rlo = j - 1 >= jswitch ? 1 : m - C;              // candidate rows rlo..m
    int a;
    if (itop + 1 >= rlo && itop + 2 * K <= m) {                 // every cell of the lane is a candidate: the filter's maximum is exact
        const int lo = (int)(int16_t)(acc & 0xffffu), hi = (int)(int16_t)(acc >> 16);
        a = (lo > hi ? lo : hi) | 1;
    } else {
        a = wf16c_exact_step_score_rt(cur, K, m, n, C, itop, j, pot2);
    }
    if (a == WF16C_NO_SNAP || wf16c_score_of(a) < (s_floor > 1 ? s_floor : 1)) return best;   // no candidate can matter
    const int old = (int)meta[1];
    if (old != WF16C_NO_SNAP && a <= old && wf16c_score_of(old) >= s_floor)
        best = lane16c_scan_rt(pend, K, m, n, C, false, itop, (int)meta[0], s_floor, pot2, best, tr);
    meta[0] = (uint32_t)j;
    meta[1] = (uint32_t)a;
    meta[2] = 1u - idx;
    return best;
}

// One strip of 64*K rows starting after table row `i0` (see wf16t_strip for the block structure).
//
// TEAM > 1: strip s of a pair runs on warp s % TEAM of the CTA, concurrently with its neighbours.  The strip
// reads column j of the boundary line (the bottom row of strip s-1) one block ahead and replaces it in place
// 94+ steps later, so the only ordering needed is "strip s-1 has published column j before strip s reads it":
// the producer publishes its progress after every block (columns <= tb + cnt - 95), the consumer spins on it
// before each refill.  Boundary reads bypass L1 (ld.global.cg): the line is written by other warps.
template <int K, bool STD, int TEAM, bool POT2>
__device__ __noinline__ long long wf16c_strip(Wf16cWarp& w, const Wf16cParams& P, int i0, bool rowscan, bool store_bottom, long long best)
{
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr uint32_t LANE_BYTES = K >= 4 ? 16u : 4u * K;         // bytes a lane owns per table row
    constexpr uint32_t COMBO_BYTES = K * 128u;                      // table bytes per combination
    constexpr int D = WF16C_SKEW;
    const int lane = threadIdx.x & 31;
    const Wf16cPass g = w.g;
    const int n = g.n, m = g.m;
    const int itop = i0 + lane * 2 * K, irel_top = lane * 2 * K;
    const int pot2 = POT2 ? irel_top : -1;                             // layout argument of the cold functions
    const uint32_t gup = STD ? 0xfffcfffcu : g.gup, gleft = STD ? 0xfffafffau : g.gleft;
    uint32_t* const bnd = w.bnd;
    const uint32_t tab_base = (uint32_t)__cvta_generic_to_shared(w.smem);
    const uint32_t ring_base = tab_base + WF16C_TAB_WORDS * 4;
    const uint32_t oring_base = ring_base + 2 * WF16C_RING * 4;
    constexpr uint32_t OR_OFF = 2 * WF16C_RING * 4;                // bottom-row ring slot of the input-ring slot at the same index
    const uint32_t my_tab = tab_base + lane * LANE_BYTES;
    GP_TRACE(10, (uint32_t)i0);

    // ---- increment table of this lane's rows ---------------------------------------------------
    {
        uint32_t rc[2 * K];
#pragma unroll
        for (int x = 0; x < 2 * K; ++x) rc[x] = (itop + x < m) ? (TEAM > 1 ? load_code_cg(w.packed_row, w.pd.row_off, (uint32_t)(itop + x)) : load_code(w.packed_row, w.pd.row_off, (uint32_t)(itop + x))) : 0u;
#pragma unroll 1
        for (uint32_t combo = 0; combo < 16; ++combo) {
            const uint32_t ca = combo & 3u, cb = combo >> 2;
            uint32_t wd[K];
#pragma unroll
            for (int k = 0; k < K; ++k) wd[k] = wf16c_table_word(rc[k], rc[K + k], ca, cb, P);
            const uint32_t a = my_tab + combo * COMBO_BYTES;
            if constexpr (K >= 4) {
#pragma unroll
                for (int q = 0; q < K / 4; ++q)
                    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(a + q * 512), "r"(wd[4 * q]), "r"(wd[4 * q + 1]),
                                 "r"(wd[4 * q + 2]), "r"(wd[4 * q + 3]) : "memory");
            } else if constexpr (K == 2) {
                asm volatile("st.shared.v2.u32 [%0], {%1,%2};" :: "r"(a), "r"(wd[0]), "r"(wd[1]) : "memory");
            } else {
                sts32(a, wd[0]);
            }
        }
    }
    // ---- ring: every slot valid (offset 0), then columns 1..32 ------------------------------------
    auto ring_word = [&](uint32_t line) {      // boundary-line word -> (table offset << 16 | value)
        return (line & 0xffffu) | (((line >> 16) & 15u) * COMBO_BYTES) << 16;
    };
    auto ring_put = [&](int jj, uint32_t line) {
        const uint32_t v = ring_word(line);
        sts32(ring_base + 4 * (jj & (WF16C_RING - 1)), v);
        sts32(ring_base + 4 * ((jj & (WF16C_RING - 1)) + WF16C_RING), v);
    };
    for (int e = lane; e < 2 * WF16C_RING; e += 32) sts32(ring_base + 4 * e, 0u);
    __syncwarp();
    auto wait_cols = [&](int need) {            // team mode: until strip_idx-1 has published columns <= min(need, n)
        if constexpr (TEAM > 1) {
            if (w.strip_idx > 0) {
                const uint32_t want = ((uint32_t)(w.strip_idx - 1) << 15) | (uint32_t)(need < n ? need : n);
                const uint32_t a = w.prog_addr + 4u * (uint32_t)((w.strip_idx - 1) % TEAM);
                while (lds32_volatile(a) < want) { }
                __syncwarp();
            }
        }
    };
    auto bnd_load = [&](int jj) -> uint32_t {
        if constexpr (TEAM > 1) return __ldcg(bnd + jj);
        else return bnd[jj];
    };
    GP_TRACE(11, (uint32_t)i0);
    wait_cols(32);
    GP_TRACE(12, (uint32_t)i0);
    { const int jj = 1 + lane; ring_put(jj, bnd_load(jj <= n + 1 ? jj : n + 1)); }
    Lane16c<K> st;
    lane16c_begin<K>(st, g, itop, irel_top);

    // ---- candidate filter state --------------------------------------------------------------------
    int S0 = __reduce_max_sync(FULL, (int)(best >> 32));                   // the warp's best score so far
    uint32_t thrS = wf16c_filter_thr(S0);
    const int jswitch = n - g.C > 1 ? n - g.C : 1;                        // first candidate column
    const bool rowlane = rowscan && (itop + 2 * K >= m - g.C) && (itop + 1 <= m);
    const int jarm = rowlane ? 1 : jswitch;
    const int t_end = n + 1 + 31 * D;                                     // lane 31's lo group reaches column n+1
    const bool do_store = store_bottom && lane == 31;
    uint32_t recv_next = 0;                                               // shuffle issued one step ahead

    // deferred exact scan (see wf16c_cold): the pending step's score and buffer index here, the rest in local memory
    uint32_t snap_bufs[2 * K + 3];
    int snapA = WF16C_NO_SNAP;
    uint32_t snap_idx = 0;
    snap_bufs[2 * K + 1] = (uint32_t)WF16C_NO_SNAP;
    snap_bufs[2 * K + 2] = 0u;
    auto fire_path = [&](const uint32_t (&W)[K], int jj, uint32_t acc) { // this lane fired at lo column jj, holding registers W then
        uint32_t* cur = snap_bufs + (1u - snap_idx) * K;
#pragma unroll
        for (int k = 0; k < K; ++k) cur[k] = W[k];
        best = wf16c_cold(snap_bufs, K, g.tr ? 2 : 0, acc, m, n, g.cell ? -1 : g.C, itop, jj, S0, pot2, best);
        snapA = (int)snap_bufs[2 * K + 1];
        snap_idx = snap_bufs[2 * K + 2];
    };
    // After any lane's fire the whole warp learns the new best score at once (one REDUX): without it the
    // lanes below an alignment path, whose cells all gain a point per column, fire at every column of the
    // candidate zone until the next block end although another lane already holds a better cell.
    auto share_floor = [&]() {
        const int mine = (int)(best >> 32), pend = snapA != WF16C_NO_SNAP ? wf16c_score_of(snapA) : mine;
        S0 = __reduce_max_sync(FULL, mine > pend ? mine : pend);
        thrS = wf16c_filter_thr(S0);
    };
    __syncwarp();

    // One register set of increments: a step adds them to the diagonals first, the loads of the NEXT step's
    // increments go out right behind those adds and come back under the max chain.  The ring pointer is the
    // only moving address: the bottom-row ring is addressed relative to it, the loop ends on it.
    auto run_block = [&](auto edge_c, auto filt_c, int tb, int cnt) {
        constexpr bool EDGE = decltype(edge_c)::value, FILT = decltype(filt_c)::value;
        uint32_t p = ring_base + (((uint32_t)(tb - D * lane)) & (WF16C_RING - 1)) * 4u;
        uint32_t p_end = p + 4u * (uint32_t)cnt;
        asm volatile("" : "+r"(p_end));                              // opaque: kept in a register, not recomputed every iteration
        int j = tb - D * lane;                                            // my lo column
        uint32_t nthr = FILT ? wf16c_nthr<K>(g, j, irel_top) : 0u;                     // follows j (mod 2^16 outside 1..n+1)
        uint32_t inc[K];
        uint32_t w0 = lds32(p), w1 = lds32(p + 4);
        lds_inc<K>(inc, my_tab + (w0 >> 16));
        // One step of the DP, nothing else; `in`: this lane's lo column is inside 1..n+1 (always true outside edge blocks).
        auto step = [&](uint32_t word, uint32_t next_word, uint32_t oaddr, int jj, bool& in) {
            // Where the shuffle goes out is a scheduling matter (its result is needed a step from now): first thing
            // under the free-moves layout, whose short max chain hides less latency (ptxas otherwise parks the copy
            // of the loop-carried value right behind it: 7 % of the kernel waiting), behind the loads otherwise.
            uint32_t recv = recv_next;
            if (POT2) recv_next = __shfl_up_sync(FULL, st.W[K - 1], 1);
            uint32_t d[K];
            lane16c_diag<K>(st, inc, d);
            lds_inc<K>(inc, my_tab + (next_word >> 16));
            if (!POT2) recv_next = __shfl_up_sync(FULL, st.W[K - 1], 1);
            if (lane == 0) recv = word << 16;
            if constexpr (!EDGE) {
                in = true;
                lane16c_chain<K, POT2>(st, recv, d, gup, gleft);
                if (do_store) sts32(oaddr, st.W[K - 1]);
            } else {
                // Edge blocks: some lanes are outside columns 1..n+1.  Every lane computes the step and the lanes outside keep
                // their old registers by SELECTS, not by a branch around the max chain: a branch there keeps the compiler from
                // scheduling the step's other work (next increments, ring words, shuffles) under the chain's latency, and a
                // warp that has its scheduler to itself ran such blocks at 184 clocks per step against 88.
                in = (uint32_t)(jj - 1) <= (uint32_t)n;
                Lane16c<K> nx = st;
                lane16c_chain<K, POT2>(nx, recv, d, gup, gleft);
                if (jj == 1) lane16c_fix_first<K>(nx, g, itop, irel_top);
#pragma unroll
                for (int k = 0; k < K; ++k) st.W[k] = in ? nx.W[k] : st.W[k];
                st.up0_prev = in ? nx.up0_prev : st.up0_prev;
                if (do_store && in) sts32(oaddr, st.W[K - 1]);
            }
        };
        if constexpr (!EDGE && !FILT) {
            // the steady loop: full blocks (cnt == 32), four steps per iteration so that the loop-carried copies (ring
            // words, increment registers) are paid once per four steps -- they run on the same pipe as the packed adds
            bool in;
#pragma unroll 1
            do {
                const uint32_t w2 = lds32(p + 8);
                step(w0, w1, p + OR_OFF, j, in);
                const uint32_t w3 = lds32(p + 12);
                step(w1, w2, p + OR_OFF + 4, j + 1, in);
                const uint32_t w4 = lds32(p + 16);
                step(w2, w3, p + OR_OFF + 8, j + 2, in);
                const uint32_t w5 = lds32(p + 20);
                step(w3, w4, p + OR_OFF + 12, j + 3, in);
                w0 = w4; w1 = w5;
                p += 16; j += 4;
            } while (p != p_end);
        } else {
            // Two steps per iteration, then the candidate filter of BOTH steps, then ONE vote.  A filter and a vote after
            // every step make the next step wait for this step's whole max chain (no overlap between steps: 215 against 88
            // clocks per step for a warp alone on its scheduler); written this way the second step's chain follows the first's
            // directly and both filters -- independent of each other -- run behind it.  The first step's registers are kept
            // until the vote; its fire is resolved before the second's, so the cold function sees the steps in order.  The
            // second step's threshold is the one before the first step's fire -- staler, i.e. lower: the fire test is a
            // superset filter, the cold function decides.
#pragma unroll 1
            do {
                const uint32_t w2 = lds32(p + 8);
                bool in0, in1;
                step(w0, w1, p + OR_OFF, j, in0);
                Lane16c<K> first;
                if (FILT) {
#pragma unroll
                    for (int k = 0; k < K; ++k) first.W[k] = st.W[k];
                }
                const uint32_t w3 = lds32(p + 12);
                step(w1, w2, p + OR_OFF + 4, j + 1, in1);
                if (FILT) {
                    const uint32_t nthr1 = p_add2(nthr, POT2 ? WF16C_NSTEP_POT2 : WF16C_NSTEP);
                    const uint32_t a0 = p_add2(POT2 ? lane16c_max_pot2<K>(first) : lane16c_max<K>(first), nthr);
                    const uint32_t a1 = p_add2(POT2 ? lane16c_max_pot2<K>(st) : lane16c_max<K>(st), nthr1);
                    const bool f0 = in0 && filter_fired(a0, j >= jarm ? thrS : WF16C_UNARMED);
                    const bool f1 = in1 && filter_fired(a1, j + 1 >= jarm ? thrS : WF16C_UNARMED);
                    nthr = p_add2(nthr1, POT2 ? WF16C_NSTEP_POT2 : WF16C_NSTEP);
                    if (__any_sync(FULL, f0 || f1)) {                     // warp-uniform: every lane runs every step
                        if (f0) fire_path(first.W, j, a0);
                        if (f1) fire_path(st.W, j + 1, a1);
                        share_floor();
                    }
                }
                w0 = w2; w1 = w3;
                p += 8; j += 2;
            } while (p != p_end);
        }
    };
    using std::true_type;
    using std::false_type;

    for (int tb = 1; tb <= t_end; tb += 32) {
        // the next block's boundary words (L2 latency hidden behind this block)
        uint32_t next_line = 0;
        const bool have_next = tb + 32 <= t_end;
        if (have_next) { wait_cols(tb + 63); const int jj = tb + 32 + lane; next_line = bnd_load(jj <= n + 1 ? jj : n + 1); }
        const bool edge = tb < 31 * D + 1 || tb + 31 > n + 1;             // some lane outside columns 1..n+1
        const bool filt = rowscan || tb + 31 >= jswitch;                  // some lane may hold candidates
        const int cnt = t_end - tb + 1 < 32 ? ((t_end - tb + 2) & ~1) : 32;   // steps past t_end find every lane out of range
        if (!edge) { if (!filt) run_block(false_type(), false_type(), tb, cnt); else run_block(false_type(), true_type(), tb, cnt); }
        else       { if (!filt) run_block(true_type(), false_type(), tb, cnt);  else run_block(true_type(), true_type(), tb, cnt); }
        __syncwarp();
        if (store_bottom) {                                               // bottom row of the columns lane 31 finished
            const int c = tb + lane - (31 * D + 1);
            const uint32_t v = lds32(oring_base + 4 * ((((uint32_t)(tb - 31 * D)) & (WF16C_RING - 1)) + lane));
            // free-moves layout: the bottom row becomes row 0 of the next strip, 64*K rows lower in the row potential
            int vb = (int)(v >> 16);
            if (POT2) { vb -= 4 * 64 * K; vb = vb > 0 ? vb : 0; }
            if (c >= 1 && c <= n && lane < cnt) reinterpret_cast<uint16_t*>(bnd)[2 * c] = (uint16_t)vb;
            if constexpr (TEAM > 1) {                                      // publish: columns <= tb + cnt - 95 are in the line
                __threadfence_block();
                __syncwarp();
                const int done = tb + cnt - 1 - (31 * D + 1);
                if (lane == 0 && done >= 1)
                    sts32_volatile(w.prog_addr + 4u * (uint32_t)w.team_warp, ((uint32_t)w.strip_idx << 15) | (uint32_t)(done < n ? done : n));
            }
        }
        if (have_next) ring_put(tb + 32 + lane, next_line);
        __syncwarp();
        if (tb == 1 || tb == 97 || tb + 32 > t_end || (tb <= n - 160 && tb + 32 > n - 160)) GP_TRACE(13, (uint32_t)tb);
    }
    GP_TRACE(14, (uint32_t)i0);
    if (snapA != WF16C_NO_SNAP && wf16c_score_of(snapA) >= S0)
        best = wf16c_cold(snap_bufs, K, g.tr ? 3 : 1, 0u, m, n, g.C, itop, 0, S0, pot2, best);
    __syncwarp();
    GP_TRACE(15, (uint32_t)i0);
    return best;
}

// One pass (all strips of the sub-table); returns the best key of the pair: warp wide, or CTA wide in team mode
// (every thread of the CTA calls it and gets the same key).
template <bool STD, int TEAM, bool POT2>
__device__ __noinline__ long long wf16c_pass(Wf16cWarp& w, const Wf16cParams& P, long long* team_keys)
{
    const int lane = threadIdx.x & 31;
    const int m = w.g.m, n = w.g.n;
    // boundary line = table row 0 plus the column symbol combinations
    if constexpr (TEAM > 1) {
        __syncthreads();                                                   // the previous pass is over for every warp
        if (lane == 0) sts32_volatile(w.prog_addr + 4u * (uint32_t)w.team_warp, 0u);
    }
    for (int j = 1 + (TEAM > 1 ? (int)threadIdx.x : lane); j <= n + 1; j += 32 * TEAM) {
        const uint32_t cj = (j <= n) ? load_code(w.packed, w.pd.col_off, (uint32_t)(j - 1)) : 0u;
        const uint32_t cp = (j >= 2) ? load_code(w.packed, w.pd.col_off, (uint32_t)(j - 2)) : 0u;
        w.bnd[j] = wf16c_line_word(w.g, j, cj, cp);
    }
    if constexpr (TEAM > 1) { __threadfence_block(); __syncthreads(); } else __syncwarp();
    GP_TRACE(3, (uint32_t)n);
    long long best = wf16c_initial_best(w.g);
    int i0 = 0, idx = 0;
    while (i0 < m) {
        const Wf16Strip s = wf16c_next_strip(i0, m, w.g.C);
        if (TEAM == 1 || idx % TEAM == w.team_warp) {
            const bool sb = !s.last;
            const bool rs = s.rowscan && !w.g.cell;
            w.strip_idx = idx;
            switch (s.rows) {
            case 512: best = wf16c_strip<8, STD, TEAM, POT2>(w, P, i0, rs, sb, best); break;
            case 256: best = wf16c_strip<4, STD, TEAM, POT2>(w, P, i0, rs, sb, best); break;
            default:  best = wf16c_strip<2, STD, TEAM, POT2>(w, P, i0, rs, sb, best); break;
            }
        }
        i0 += s.rows;
        ++idx;
    }
    best = warp_max_key(best);
    GP_TRACE(4, (uint32_t)m);
    if constexpr (TEAM > 1) {
        if (lane == 0) team_keys[w.team_warp] = best;
        __syncthreads();
#pragma unroll
        for (int t = 0; t < TEAM; ++t) { const long long k = team_keys[t]; best = k > best ? k : best; }
        __syncthreads();
    }
    return best;
}

// Does the 16-base word pair (plo, phi) occur in the sequence at word offset `off` with `len` bases?
// Warp-wide; every lane returns the same answer.
// CG: read through L2 (ld.global.cg) instead of the read-only path -- for a row sequence another CTA of the same launch wrote.
template <bool CG> __device__ __forceinline__ uint32_t wf16c_ldw(const uint32_t* p) { return CG ? __ldcg(p) : __ldg(p); }
template <bool CG = false>
__device__ __forceinline__ bool wf16c_probe_hit(const uint32_t* __restrict__ packed, uint32_t off, uint32_t len, uint32_t plo, uint32_t phi)
{
    const int lane = threadIdx.x & 31;
    const uint32_t nw = (len + 7u) >> 3;
    bool hit = false;
    for (uint32_t q0 = 0; q0 < nw; q0 += 32) {
        const uint32_t q = q0 + lane;
        const uint32_t w0 = q < nw ? wf16c_ldw<CG>(packed + off + q) : 0u;
        const uint32_t w1 = q + 1 < nw ? wf16c_ldw<CG>(packed + off + q + 1) : 0u;
        const uint32_t w2 = q + 2 < nw ? wf16c_ldw<CG>(packed + off + q + 2) : 0u;
#pragma unroll
        for (uint32_t s = 0; s < 8; ++s) {
            const uint32_t lo = __funnelshift_r(w0, w1, 4 * s), hi = __funnelshift_r(w1, w2, 4 * s);
            hit |= (lo == plo) && (hi == phi) && (8u * q + s + WF16C_PROBE <= len);
        }
        if (__any_sync(0xffffffffu, hit)) return true;
    }
    return false;
}

// Orientation probe: the system to start with (where the walk is expected to end).
template <bool CG = false>
__device__ __forceinline__ int wf16c_predict_system(const uint32_t* __restrict__ packed_row, const uint32_t* __restrict__ packed, const PairDesc& pd)
{
    if (pd.m < (uint32_t)WF16C_PROBE || pd.n < (uint32_t)WF16C_PROBE) return WF16C_SYS_U;
    const uint32_t c0 = __ldg(packed + pd.col_off), c1 = __ldg(packed + pd.col_off + 1);
    const uint32_t r0 = wf16c_ldw<CG>(packed_row + pd.row_off), r1 = wf16c_ldw<CG>(packed_row + pd.row_off + 1);
    // both sequences start with the same bases: the overlap starts in the corner
    if (c0 == r0 && c1 == r1) return WF16C_SYS_C;
    // the first bases of the column sequence occur in the row sequence: the overlap starts in column 0
    if (wf16c_probe_hit<CG>(packed_row, pd.row_off, pd.m, c0, c1)) return WF16C_SYS_U;
    // the first bases of the row sequence occur in the column sequence: it starts in row 0
    return wf16c_probe_hit(packed, pd.col_off, pd.n, r0, r1) ? WF16C_SYS_L : WF16C_SYS_U;
}

// One pair, start to finish: probe, first pass (scan mode), up to two sub-table passes in the other systems.  Every
// thread of the warp (TEAM = 1) or CTA (TEAM > 1) calls it with w.pd / w.packed / w.packed_row set.  r receives the exact
// score / ends / clip; the return value is the certified origin (FLAG_ROW0 | FLAG_COL0 bits) or 0 when no system
// certifies it (the caller hands the pair to an exact kernel); key_out the first pass's key without the origin bits.
template <bool STD, int TEAM, bool POT2>
__device__ __forceinline__ uint32_t wf16c_solve_pair(Wf16cWarp& w, const Wf16cParams& P, long long* team_keys, uint32_t force_sys,
                                                     bool leader, unsigned int* __restrict__ counters, DevResult& r, long long& key_out, bool tr = false)
{
    const int m = (int)w.pd.m, n = (int)w.pd.n;                           // of the computed table (tr: the reference's n, m)
    GP_TRACE(1, (uint32_t)m);
    const int sys0 = force_sys != 0u ? (int)force_sys - 1 : wf16c_predict_system<(TEAM > 1)>(w.packed_row, w.packed, w.pd);
    w.g = wf16c_make_pass(m, n, P, sys0, false, 0, tr);
    GP_TRACE(2, (uint32_t)sys0);
    const long long key = wf16c_pass<STD, TEAM, POT2>(w, P, team_keys);
    GP_TRACE(5, 0u);
    uint32_t origin = wf16c_certified_origin(w.g, key);
    store_result(&r, key & ~3ll, tr ? n : m, tr ? m : n, FLAG_KERNEL16);  // exact score / ends / clip, the reference's orientation; origin still open
    for (int attempt = 1; attempt <= 2 && origin == 0u; ++attempt) {
        // another system on the sub-table that ends at the best cell, that cell only
        __syncwarp();
        if (leader) atomicAdd(counters, 1u);
        w.g = wf16c_make_pass(tr ? r.col_end : r.row_end, tr ? r.row_end : r.col_end, P, wf16c_next_system(sys0, attempt), true, r.score, tr);
        const long long key2 = wf16c_pass<STD, TEAM, POT2>(w, P, team_keys);
        origin = wf16c_certified_origin(w.g, key2);
    }
    key_out = key & ~3ll;
    return origin;
}

// counters[0] sub-table passes run (second and third passes), counters[1] pairs handed to an exact kernel.
// force_sys: 0 the probe decides, 1 + WF16C_SYS_* starts with that system (tests).
// retry16t / retry32: work lists of the exact kernels behind this one (pairs with n <= WF16T_MAX_N go to
// the table kernel); *retry16t_n / *retry32_n are their fill counts (the host may have pre-filled them).
// TEAM = 1: one warp per pair.  TEAM = warps per CTA: one CTA per pair (few, long pairs: the relax chain).
// TEAM = 1: one warp per pair, 4-warp CTAs.  TEAM = 4 / 8: one CTA of TEAM warps per pair (8 warps: one CTA per SM).
template <bool STD, int TEAM, bool POT2>
__global__ void __launch_bounds__(TEAM > 4 ? 32 * TEAM : WF16C_THREADS, TEAM > 4 ? 1 : WF16C_CTAS_PER_SM)
overlap_wf16c_kernel(const uint32_t* __restrict__ packed, const PairDesc* __restrict__ pairs,
                     const uint32_t* __restrict__ order, uint32_t n_work, unsigned int* __restrict__ queue,
                     Wf16cParams P, uint32_t* __restrict__ scratch, uint32_t scratch_stride,
                     uint32_t* __restrict__ retry16t, unsigned int* __restrict__ retry16t_n,
                     uint32_t* __restrict__ retry32, unsigned int* __restrict__ retry32_n,
                     unsigned int* __restrict__ counters, uint32_t force_sys, DevResult* __restrict__ out)
{
    extern __shared__ uint32_t wf16c_smem[];
    __shared__ uint32_t team_prog[TEAM];
    __shared__ long long team_keys[TEAM];
    __shared__ uint32_t team_qi;
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    Wf16cWarp w;
    w.packed = packed;
    w.packed_row = packed;
    w.team_warp = TEAM > 1 ? (int)(threadIdx.x >> 5) : 0;
    w.strip_idx = 0;
    w.prog_addr = (uint32_t)__cvta_generic_to_shared(team_prog);
    w.bnd = scratch + (size_t)(warp_global - (uint32_t)w.team_warp) * scratch_stride;   // team: the CTA's first line
    w.smem = wf16c_smem + (threadIdx.x >> 5) * WF16C_WARP_WORDS;
    const bool leader = TEAM > 1 ? threadIdx.x == 0 : lane == 0;
    for (;;) {
        uint32_t qi = 0;
        if constexpr (TEAM > 1) {
            if (threadIdx.x == 0) team_qi = atomicAdd(queue, 1u);
            __syncthreads();
            qi = team_qi;
            __syncthreads();
        } else {
            if (lane == 0) qi = atomicAdd(queue, 1u);
            qi = __shfl_sync(0xffffffffu, qi, 0);
        }
        if (qi >= n_work) break;
        const uint32_t ord = order[qi];
        const uint32_t pid = ord & 0x7fffffffu;
        const bool tr = (ord >> 31) != 0u;                                 // the host's orientation choice (Wf16cPass::tr)
        w.pd = pairs[pid];
        const int m = (int)w.pd.m, n = (int)w.pd.n;                        // the reference's orientation
        if (tr) w.pd = PairDesc{w.pd.col_off, w.pd.n, w.pd.row_off, w.pd.m};
        DevResult r;
        long long key;
        const uint32_t origin = wf16c_solve_pair<STD, TEAM, POT2>(w, P, team_keys, force_sys, leader, counters, r, key, tr);
        if (leader) {
            if (origin != 0u) {
                store_result(out + pid, key | (long long)origin, m, n, FLAG_KERNEL16);
            } else {
                out[pid] = r;                                                // overwritten by the exact kernel
                atomicAdd(counters + 1, 1u);
                if ((uint32_t)n <= WF16T_MAX_N) retry16t[atomicAdd(retry16t_n, 1u)] = pid;
                else retry32[atomicAdd(retry32_n, 1u)] = pid;
            }
        }
        __syncwarp();
    }
}

constexpr int WF16C_TEAM = WF16C_THREADS / 32;
constexpr int WF16C_TEAM_BIG = 8;                  // the few, long pairs of a late relax step: 8 warps per pair, one CTA per SM

template <int TEAM> constexpr size_t wf16c_smem_bytes() { return (size_t)(TEAM > 4 ? TEAM : WF16C_THREADS / 32) * WF16C_WARP_WORDS * sizeof(uint32_t); }

template <bool STD, int TEAM, bool POT2>
inline cudaError_t wf16c_configure_one()
{
    cudaError_t e = cudaFuncSetAttribute(overlap_wf16c_kernel<STD, TEAM, POT2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wf16c_smem_bytes<TEAM>());
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(overlap_wf16c_kernel<STD, TEAM, POT2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

inline cudaError_t wf16c_configure()
{
    cudaError_t e;
    if ((e = wf16c_configure_one<true, 1, false>()) != cudaSuccess) return e;
    if ((e = wf16c_configure_one<false, 1, false>()) != cudaSuccess) return e;
    if ((e = wf16c_configure_one<true, WF16C_TEAM, false>()) != cudaSuccess) return e;
    if ((e = wf16c_configure_one<false, WF16C_TEAM, false>()) != cudaSuccess) return e;
    if ((e = wf16c_configure_one<true, 1, true>()) != cudaSuccess) return e;
    if ((e = wf16c_configure_one<true, WF16C_TEAM, true>()) != cudaSuccess) return e;
    if ((e = wf16c_configure_one<true, WF16C_TEAM_BIG, false>()) != cudaSuccess) return e;
    return wf16c_configure_one<true, WF16C_TEAM_BIG, true>();
}

// Launches the kernel on `stream`; grows *scratch (device) as needed.  Returns a cudaError_t as int.
// team: 0 one warp per pair; WF16C_TEAM / WF16C_TEAM_BIG: one CTA of that many warps per pair (the big team exists for
// the standard scores only; other scores take the 4-warp team).
inline int wf16c_launch(cudaStream_t stream, int sm_count, const uint32_t* packed, const PairDesc* pairs,
                        const uint32_t* order, uint32_t n_work, unsigned int* queue, const Wf16cParams& P,
                        uint32_t max_n, void** scratch, size_t* scratch_cap,
                        uint32_t* retry16t, unsigned int* retry16t_n, uint32_t* retry32, unsigned int* retry32_n,
                        unsigned int* counters, uint32_t force_sys, int team, DevResult* out)
{
    if (team == WF16C_TEAM_BIG && !P.std_scores) team = WF16C_TEAM;
    const bool big = team == WF16C_TEAM_BIG;
    const int blocks = big ? sm_count : sm_count * WF16C_CTAS_PER_SM;
    const int threads = big ? 32 * WF16C_TEAM_BIG : WF16C_THREADS;
    const size_t smem = big ? wf16c_smem_bytes<WF16C_TEAM_BIG>() : wf16c_smem_bytes<1>();
    const uint32_t warps = (uint32_t)blocks * (uint32_t)(threads / 32);
    const uint32_t stride = (max_n + 2 + 31 + 32) & ~31u;
    const size_t need = (size_t)warps * stride * sizeof(uint32_t);
    if (need > *scratch_cap) {
        if (*scratch) cudaFree(*scratch);
        *scratch = nullptr; *scratch_cap = 0;
        cudaError_t e = cudaMalloc(scratch, need);
        if (e != cudaSuccess) return (int)e;
        *scratch_cap = need;
    }
#define GP_WF16C_LAUNCH(STD_, TEAM_, POT2_)                                                                            \
    overlap_wf16c_kernel<STD_, TEAM_, POT2_><<<blocks, threads, smem, stream>>>(                                        \
        packed, pairs, order, n_work, queue, P, (uint32_t*)*scratch, stride, retry16t, retry16t_n, retry32, retry32_n, \
        counters, force_sys, out)
    if (P.pot2 && P.std_scores) {
        if (big) GP_WF16C_LAUNCH(true, WF16C_TEAM_BIG, true); else if (team) GP_WF16C_LAUNCH(true, WF16C_TEAM, true); else GP_WF16C_LAUNCH(true, 1, true);
    } else if (P.std_scores) {
        if (big) GP_WF16C_LAUNCH(true, WF16C_TEAM_BIG, false); else if (team) GP_WF16C_LAUNCH(true, WF16C_TEAM, false); else GP_WF16C_LAUNCH(true, 1, false);
    } else {
        if (team) GP_WF16C_LAUNCH(false, WF16C_TEAM, false); else GP_WF16C_LAUNCH(false, 1, false);
    }
#undef GP_WF16C_LAUNCH
    return (int)cudaGetLastError();
}
#endif // __CUDACC__

} // namespace gp
