// overlap_wf16.cuh -- packed 16-bit overlap-DP kernel, one warp per pair (sm_100a).
//
// Same contract as overlap_wf32.cuh (ContigsCompactor::Evaluate before the significance test,
// ContigsCompactor-v0.2.0/ContigsMerger/ContigsCompactor.cpp:1596-1709,:1736-1837), for pairs with
// min(m,n) <= 4094, an alphabet of <= 8 symbols and small penalties -- i.e. everything GAPPadder
// produces.  Two DP cells per 32-bit register, DPX packed instructions (VIADD.16x2,
// VIMNMX.S16x2, VIADDMNMX.S16x2.RELU), a PRMT table lookup for the substitution score.
//
// ---- arithmetic ------------------------------------------------------------------------------
// Potential.  Instead of H the kernel carries P = H + pot, pot(i,j) = m-i when m <= n ("row
// potential") else n-j.  Every DP step then has a non-positive increment
//     diag: match 0, mismatch X-1     up: G-1 (row pot) / G (col pot)     left: G / G-1
// so P never grows along a path.  The answer has H >= 0 (cell (0,n) is scanned first and holds 0),
// hence P >= 0 on every cell of the winning path and on every cell that can tie with it; clamping
// P at -1 therefore changes nothing that matters: by induction the clamped recurrence computes
// exactly max(P,-1) everywhere, so cells with P >= 0 keep their exact value, predecessor choice and
// origin.  0 <= P <= min(m,n) for those cells, i.e. 12 bits for 4094 -- which leaves room for tags.
//
// Cell encoding (16 bits, signed, >= 0):   V = 8*(P+1) + 4*z + code
//   code  origin of the cell's predecessor walk, thermometer coded by where on the border it ends,
//         ordered along the border from bottom-left to top-right:
//         0 = column 0 (row > 0), 1 = the corner (0,0), 3 = row 0 (column > 0).
//   z     set on the diagonal candidate only, cleared after the max.
// Tie rule.  The reference prefers diag, then up, then left on equal scores (:1651-1665).  Walks of
// different cells never cross, so origins are monotone: code(left) <= code(diag) <= code(up).  With
// the code in the low bits a packed max already resolves left-vs-diag and left-vs-up ties the way
// the reference does (the preferred candidate has the larger or equal code, and equal codes are
// interchangeable because only the origin survives the walk, :1834-1837).  The one remaining case,
// diag-vs-up, is fixed by the z bonus.  max is then order independent, which lets the kernel fold
// the `up` candidate (the only one that depends on the row above) last.
//
// ---- layout -----------------------------------------------------------------------------------
// A strip is 32 lanes x 2K rows.  Lane l owns rows [2K*l, 2K*l+2K) of the strip: K "lo" rows in
// the low halves of W[0..K) and K "hi" rows in the high halves, the hi group running one column
// behind the lo group, so W[k] = (row k @ column j , row K+k @ column j-1) and W[k] depends on
// W[k-1] of the same step in BOTH halves.  Only W[0] needs a fix-up (one PRMT).  Lane l runs two
// steps behind lane l-1 and receives, per step, one word from it: the value of its last row and
// the column's base code.  Lane 0 reads the same word from the boundary line the previous strip
// left behind (global scratch, L2 resident, 32 columns per coalesced load).
//
// Host/device: the per-lane arithmetic below is __host__ __device__; tests/emulate_wf16.cu runs
// the very same functions on the CPU, lane by lane, against the oracle.
#pragma once
#include "common.cuh"

#if defined(__CUDACC__)
#define GP_HD __host__ __device__ __forceinline__
#else
#define GP_HD inline
#endif

namespace gp {

// ---- packed primitives (device: one SASS instruction each; host: emulation for the tests) -----
GP_HD uint32_t p_prmt(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
    // prmt.b32 in its default mode: selector nibble bit 3 replicates the sign of the selected byte
    // (__byte_perm() masks that bit away, so it cannot be used here).
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
#else
    uint64_t src = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) {
        uint32_t nib = (sel >> (4 * i)) & 0xf;
        uint32_t byte = (uint32_t)(src >> (8 * (nib & 7))) & 0xff;
        if (nib & 8) byte = (byte & 0x80) ? 0xff : 0x00;
        r |= byte << (8 * i);
    }
    return r;
#endif
}
GP_HD uint32_t p_add2(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __vadd2(a, b);
#else
    return ((a + b) & 0xffffu) | (((a >> 16) + (b >> 16)) << 16);
#endif
}
GP_HD uint32_t p_max2(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __vmaxs2(a, b);
#else
    int16_t al = (int16_t)a, bl = (int16_t)b, ah = (int16_t)(a >> 16), bh = (int16_t)(b >> 16);
    return (uint16_t)(al > bl ? al : bl) | ((uint32_t)(uint16_t)(ah > bh ? ah : bh) << 16);
#endif
}
// max(a + b, c, 0) per signed 16-bit half
GP_HD uint32_t p_addmax2_relu(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
    return __viaddmax_s16x2_relu(a, b, c);
#else
    uint32_t s = p_add2(a, b);
    return p_max2(p_max2(s, c), 0u);
#endif
}

// max(a + b, c) per signed 16-bit half
GP_HD uint32_t p_addmax2(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
    return __viaddmax_s16x2(a, b, c);
#else
    return p_max2(p_add2(a, b), c);
#endif
}

// ---- parameters -------------------------------------------------------------------------------
constexpr uint32_t WF16_MAX_MIN_LEN = 4094;       // 8*(P+1)+7 <= 32767
constexpr uint32_t TAG_Z2 = 0x00040004u;

struct Wf16Params {
    uint32_t tbl_lo, tbl_hi;     // PRMT table of diagonal increments, indexed by the selector nibble (see addsel)
    uint32_t addsel;             // 1: <= 4 symbols, selector = row code + (4 - column code) % 4 (match at 0 and 4),
                                 //    formed with VIADD.16x2 (FMA-side pipe); 0: <= 8 symbols, selector = row ^ column
                                 //    (match at 0), formed with LOP3 (ALU pipe)
    uint32_t gup_row, gleft_row; // packed up/left increments under the row potential
    uint32_t gup_col, gleft_col; // ... under the column potential
    int32_t max_clip;
};

inline bool wf16_params_ok(int mismatch, int indel)
{
    // increments must be <= 0 in the potential domain and fit a signed byte after scaling by 8
    return mismatch <= 1 && mismatch >= -15 && indel <= 0 && indel >= -14;
}

inline bool wf16_pair_ok(uint32_t m, uint32_t n)
{
    return m >= 1 && n >= 1 && (m < n ? m : n) <= WF16_MAX_MIN_LEN && m <= 0xffffff && n <= 0xffffff;
}

inline Wf16Params wf16_make_params(int mismatch, int indel, int max_clip, bool addsel)
{
    Wf16Params p;
    const uint32_t inc_match = (uint32_t)(0 * 8 + 4) & 0xff;                 // +1 - 1, z bonus
    const uint32_t inc_mism = (uint32_t)((mismatch - 1) * 8 + 4) & 0xff;     // X - 1, z bonus
    p.tbl_lo = inc_match | (inc_mism << 8) | (inc_mism << 16) | (inc_mism << 24);
    p.tbl_hi = (addsel ? inc_match : inc_mism) | (inc_mism << 8) | (inc_mism << 16) | (inc_mism << 24);
    p.addsel = addsel ? 1u : 0u;
    auto pk = [](int v) { uint32_t h = (uint32_t)(v * 8) & 0xffffu; return h | (h << 16); };
    p.gup_row = pk(indel - 1); p.gleft_row = pk(indel);
    p.gup_col = pk(indel);     p.gleft_col = pk(indel - 1);
    p.max_clip = max_clip;
    return p;
}

// Geometry of one pair in the potential domain.
struct Wf16Pair {
    int m, n, C;
    bool rowpot;             // pot = m - i, else pot = n - j
    uint32_t gup, gleft;
    GP_HD int pot(int i, int j) const { return rowpot ? m - i : n - j; }
    // boundary values: V(i,0) for i >= 0 and V(0,j) for j >= 1
    GP_HD uint32_t v_col0(int i) const
    {
        int p = pot(i, 0) + 1;
        if (p < 0) p = 0;
        return (uint32_t)(8 * p) + (i == 0 ? 1u : 0u);
    }
    GP_HD uint32_t v_row0(int j) const { return (uint32_t)(8 * (pot(0, j) + 1)) + (j == 0 ? 1u : 3u); }
};

GP_HD Wf16Pair wf16_make_pair(int m, int n, const Wf16Params& P)
{
    Wf16Pair g;
    g.m = m; g.n = n; g.C = P.max_clip;
    g.rowpot = m <= n;
    g.gup = g.rowpot ? P.gup_row : P.gup_col;
    g.gleft = g.rowpot ? P.gleft_row : P.gleft_col;
    return g;
}

// code11 of a column base code: the selector contribution in both nibbles of a byte (PRMT selectors
// for the low and the high byte of one 16-bit increment)
GP_HD uint32_t code11(uint32_t c, uint32_t addsel) { return (addsel ? ((4u - c) & 3u) : (c & 7u)) * 0x11u; }

// ---- per-lane state and arithmetic ------------------------------------------------------------
template <int K>
struct Lane16 {
    uint32_t W[K];       // (row k @ col j | row K+k @ col j-1 << 16)
    uint32_t Rk[K];      // selector constant: row codes in nibbles, bit 3 of the odd nibbles set
    uint32_t cvec;       // byte 0: code11 of column j, byte 1: of column j-1 (older history above)
    uint32_t up0_prev;   // previous step's `up` of W[0] == this step's diagonal of W[0]
    uint32_t negthr[K];  // candidate filter: minus the V a cell needs to matter (0x8001 = never), per half
    uint32_t fstep[K];   // per-step drift of negthr (column potential only)
};

// Start of a strip.  itop = number of table rows above this lane's first row; the lane's lo rows
// are i = itop+1+k, its hi rows i = itop+1+K+k (1-based).  rcode[x] is the 4-bit code of the
// base of row itop+1+x (x = 0..2K-1), or 15 beyond the end of the row sequence.
template <int K>
GP_HD void lane16_begin(Lane16<K>& st, const Wf16Pair& g, int itop, const uint32_t (&rcode)[2 * K])
{
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int ilo = itop + 1 + k, ihi = itop + 1 + K + k;
        st.W[k] = g.v_col0(ilo) | (g.v_col0(ihi) << 16);
        const uint32_t rlo = rcode[k] & 7u, rhi = rcode[K + k] & 7u;     // rows beyond the sequence: any code
        st.Rk[k] = (rlo * 0x11u | 0x80u) | ((rhi * 0x11u | 0x80u) << 8);
    }
    st.cvec = 0;
    st.up0_prev = g.v_col0(itop) | (g.v_col0(itop + K) << 16);
}

// After a lane's first step (lo group at column 1) the hi group has "computed" column 0; put the
// real column-0 boundary back.
template <int K>
GP_HD void lane16_fix_first(Lane16<K>& st, const Wf16Pair& g, int itop)
{
#pragma unroll
    for (int k = 0; k < K; ++k) st.W[k] = (st.W[k] & 0xffffu) | (g.v_col0(itop + 1 + K + k) << 16);
}

// One step: the lo group advances to column j, the hi group to column j-1.
// recv: bits 0-15 V(row above the lane, column j), bits 16-23 code11(column j).
template <int K, bool ADDSEL>
GP_HD void lane16_step(Lane16<K>& st, uint32_t recv, const Wf16Params& P, uint32_t gup, uint32_t gleft)
{
    st.cvec = p_prmt(recv, st.cvec, 0x6542u);                  // (recv.b2, cvec.b0, cvec.b1, cvec.b2)
    const uint32_t up0 = p_prmt(recv, st.W[K - 1], 0x5410u);   // (recv.lo16 , old W[K-1].lo16)
    uint32_t diag = st.up0_prev;
    st.up0_prev = up0;
    uint32_t up = up0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t left = st.W[k];
        const uint32_t inc = p_prmt(P.tbl_lo, P.tbl_hi, ADDSEL ? p_add2(st.Rk[k], st.cvec) : (st.Rk[k] ^ st.cvec));
        const uint32_t d = p_add2(diag, inc);                  // carries the z bonus
        const uint32_t l = p_add2(left, gleft);
        const uint32_t t = p_max2(d, l);
        const uint32_t w = p_addmax2_relu(up, gup, t) & ~TAG_Z2;
        diag = left;
        up = w;
        st.W[k] = w;
    }
}

// Word handed to the next lane after a step: value of this lane's last row (hi half of W[K-1],
// column j-1) and that column's code11.
template <int K>
GP_HD uint32_t lane16_send(const Lane16<K>& st)
{
    return p_prmt(st.W[K - 1], st.cvec, 0x7532u);
}

// Candidate bookkeeping for one 16-bit cell value.
GP_HD long long wf16_cell_key(uint32_t v16, int i, int j, const Wf16Pair& g)
{
    const uint32_t rk = cell_rank(i, j, g.m, g.n, g.C);
    if (rk > RANK_MAX) return (long long)0x8000000000000000ull;
    const int H = (int)(v16 >> 3) - 1 - g.pot(i, j);
    const uint32_t code = v16 & 3u;
    const uint32_t origin = (code & 1u) | ((code & 2u) ? 0u : 2u);   // row0 | col0 << 1
    return make_key(H, rk, origin);
}

// Exact scan of the cells a lane holds after a step (lo column j, hi column j-1); cells scoring
// below the lane's current best are skipped before the rank arithmetic.
template <int K>
GP_HD long long lane16_scan(const Lane16<K>& st, const Wf16Pair& g, int itop, int j, long long best)
{
    const int floor_score = (int)(best >> 32);
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int ilo = itop + 1 + k, ihi = itop + 1 + K + k;
        if (ilo <= g.m && j >= 1 && j <= g.n) {
            const uint32_t v = st.W[k] & 0xffffu;
            if ((int)(v >> 3) - 1 - g.pot(ilo, j) >= floor_score) {
                long long key = wf16_cell_key(v, ilo, j, g);
                best = key > best ? key : best;
            }
        }
        if (ihi <= g.m && j - 1 >= 1 && j - 1 <= g.n) {
            const uint32_t v = st.W[k] >> 16;
            if ((int)(v >> 3) - 1 - g.pot(ihi, j - 1) >= floor_score) {
                long long key = wf16_cell_key(v, ihi, j - 1, g);
                best = key > best ? key : best;
            }
        }
    }
    return best;
}

// ---- candidate filter ---------------------------------------------------------------------------
// Only cells in the last C+1 rows / columns can be the answer, and only if they score at least the
// best score the lane has seen so far (S).  Instead of looking at every such cell the lane keeps,
// per half, negthr = -(the V a cell of that row has when H = 1) and folds acc = max(V + negthr)
// over its registers: one VIADDMNMX per register.  acc = 8*(H-1) + tags for the best candidate
// cell, so "some candidate has H >= max(S,1)" is one packed compare of acc with
// thrS = 8*(max(S,1)-1).  A lane whose compare fires runs the exact scan above on its own cells.
// negthr does not depend on S; under the column potential it drifts by +8 per step (fstep).
enum : int { FILTER_NONE = 0, FILTER_ROWS = 1, FILTER_ALL = 2 };
constexpr uint32_t NEVER16 = 0x8000u;          // V + NEVER16 < 0 for every V <= 32767

// Thresholds for the step whose lo column is jc (hi column jc-1).
template <int K>
GP_HD void lane16_set_filter(Lane16<K>& st, const Wf16Pair& g, int itop, int jc, int mode)
{
#pragma unroll
    for (int k = 0; k < K; ++k) {
        uint32_t nt[2], fs[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = itop + 1 + k + h * K;
            const bool cand = i <= g.m && (mode == FILTER_ALL || (mode == FILTER_ROWS && i >= g.m - g.C));
            const int thr = 8 * (2 + g.pot(i, jc - h));          // <= 8*(2+4094) = 32768
            if (cand) { nt[h] = (uint32_t)(-thr) & 0xffffu; fs[h] = g.rowpot ? 0u : 8u; }
            else { nt[h] = NEVER16; fs[h] = 0u; }
        }
        st.negthr[k] = nt[0] | (nt[1] << 16);
        st.fstep[k] = fs[0] | (fs[1] << 16);
    }
}

// Evaluates the filter on the lane's current cells and advances the thresholds to the next step.
template <int K>
GP_HD uint32_t lane16_filter(Lane16<K>& st)
{
    uint32_t acc = 0x80008000u;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        acc = p_addmax2(st.W[k], st.negthr[k], acc);
        st.negthr[k] = p_add2(st.negthr[k], st.fstep[k]);
    }
    return acc;
}
// thrS for a best score S, both halves
GP_HD uint32_t filter_thr(int S) { const uint32_t t = (uint32_t)(8 * ((S > 1 ? S : 1) - 1)); return t | (t << 16); }
// true iff some half of acc is >= the same half of thr (signed)
GP_HD bool filter_fired(uint32_t acc, uint32_t thr)
{
#if defined(__CUDA_ARCH__)
    bool hi, lo;
    (void)__vibmax_s16x2(acc, thr, &hi, &lo);
    return hi || lo;
#else
    return (int16_t)acc >= (int16_t)thr || (int16_t)(acc >> 16) >= (int16_t)(thr >> 16);
#endif
}

// ---- strip schedule -------------------------------------------------------------------------------
// Full strips of 512 rows (K = 8), a 256-row strip (K = 4) when that keeps the last C+1 rows
// together, and a last strip of the smallest K in {1,2,4,8} that holds the remaining rows.  A strip
// needs the row filter (ROWSCAN) iff it reaches into the last C+1 rows.
struct Wf16Strip { int rows; bool last; bool rowscan; };
GP_HD Wf16Strip wf16_next_strip(int i0, int m, int C)
{
    const int R = m - i0;
    Wf16Strip s;
    if (R > 512) { s.rows = (R - 512 >= C + 1) ? 512 : 256; s.last = false; }
    else { s.rows = R <= 64 ? 64 : R <= 128 ? 128 : R <= 256 ? 256 : 512; s.last = true; }
    s.rowscan = i0 + s.rows >= m - C;
    return s;
}

#if defined(__CUDACC__)
// ---- device side ------------------------------------------------------------------------------
// Code size matters here: 16 warps per SM run different strip variants and phases at the same
// time, and the first version of this kernel (everything inlined, 345 KB of SASS) spent most of
// its cycles waiting for instructions.  The cold paths (exact scan, threshold set-up) are therefore
// out-of-line functions taking their operands by value, and each strip variant has exactly one
// generic loop body and one steady loop body.

constexpr int WF16_THREADS = 256;

struct Wf16Warp {                       // warp-uniform state of one pair
    const uint32_t* packed;
    PairDesc pd;
    Wf16Pair g;
    uint32_t* bnd;                      // boundary line: low half V(i0, j), bits 16-23 code11(column j)
    int S;                              // best score so far (warp wide, refreshed per strip)
};

template <int K> struct WVals { uint32_t W[K]; };
template <int K> struct FVals { uint32_t negthr[K], fstep[K]; };

template <int K>
__device__ __noinline__ long long wf16_scan_cold(WVals<K> v, Wf16Pair g, int itop, int j, long long best)
{
    Lane16<K> st;
#pragma unroll
    for (int k = 0; k < K; ++k) st.W[k] = v.W[k];
    return lane16_scan<K>(st, g, itop, j, best);
}

template <int K>
__device__ __noinline__ FVals<K> wf16_filter_cold(Wf16Pair g, int itop, int jc, int mode)
{
    Lane16<K> st;
    lane16_set_filter<K>(st, g, itop, jc, mode);
    FVals<K> f;
#pragma unroll
    for (int k = 0; k < K; ++k) { f.negthr[k] = st.negthr[k]; f.fstep[k] = st.fstep[k]; }
    return f;
}

// One strip of 64*K rows starting after table row `i0`; with store_bottom the low halves of bnd[]
// are replaced in place by the strip's last row.
template <int K, bool ROWSCAN, bool ADDSEL>
__device__ __noinline__ long long wf16_strip(Wf16Warp& w, const Wf16Params& P, int i0, bool store_bottom, long long best)
{
    const int lane = threadIdx.x & 31;
    const Wf16Pair g = w.g;
    const int itop = i0 + lane * 2 * K;
    Lane16<K> st;
    {
        uint32_t rcode[2 * K];
#pragma unroll
        for (int x = 0; x < 2 * K; ++x) rcode[x] = (itop + x < g.m) ? load_code(w.packed, w.pd.row_off, (uint32_t)(itop + x)) : 0u;
        lane16_begin<K>(st, g, itop, rcode);
    }
    auto set_filter = [&](int jc, int mode) {
        const FVals<K> f = wf16_filter_cold<K>(g, itop, jc, mode);
#pragma unroll
        for (int k = 0; k < K; ++k) { st.negthr[k] = f.negthr[k]; st.fstep[k] = f.fstep[k]; }
    };
    set_filter(1, ROWSCAN ? FILTER_ROWS : FILTER_NONE);
    {   // start from the warp's best score: fewer false alarms than each lane's own
        int s = (int)(best >> 32);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { int other = __shfl_xor_sync(0xffffffffu, s, o); s = other > s ? other : s; }
        w.S = s;
    }
    const int S0 = w.S;
    uint32_t thrS = filter_thr(S0);
    uint32_t send = 0, chunk = 0;
    const int n = g.n;
    const int t_end = n + 1 + 62;                       // lane 31's lo group reaches column n+1
    const int jswitch = n - g.C > 1 ? n - g.C : 1;      // first candidate column
    uint32_t* const bnd = w.bnd;
    uint16_t* const bnd16 = reinterpret_cast<uint16_t*>(bnd);
    const bool do_store = store_bottom && lane == 31;
    bool filtering = ROWSCAN;                           // per lane: any threshold armed

    // exact scan of this lane's cells (divergent: only lanes whose filter fired)
    auto slow_path = [&](int j) {
        WVals<K> v;
#pragma unroll
        for (int k = 0; k < K; ++k) v.W[k] = st.W[k];
        best = wf16_scan_cold<K>(v, g, itop, j, best);
        const int s = (int)(best >> 32);
        thrS = filter_thr(s > S0 ? s : S0);
    };

    // steady blocks: every lane active, no lane in the last C+1 columns yet
    int t_steady1 = (jswitch > 65 ? jswitch : 65);      // first step at which some lane may be in the tail
    t_steady1 = ((t_steady1 - 1) & ~31) + 1;
    for (int tb = 1; tb <= t_end; tb += 32) {
        {
            const int jj = tb + lane;
            if (jj <= n + 1) chunk = bnd[jj];
        }
        if (tb >= 65 && tb < t_steady1) {
#pragma unroll 1
            for (int s = 0; s < 32; ++s) {
                const uint32_t from_line = __shfl_sync(0xffffffffu, chunk, s);
                uint32_t recv = __shfl_up_sync(0xffffffffu, send, 1);
                if (lane == 0) recv = from_line;
                lane16_step<K, ADDSEL>(st, recv, P, g.gup, g.gleft);
                send = lane16_send<K>(st);
                const int j = tb + s - 2 * lane;
                if (do_store) bnd16[2 * (j - 1)] = (uint16_t)(st.W[K - 1] >> 16);
                if (ROWSCAN) {
                    if (filter_fired(lane16_filter<K>(st), thrS)) slow_path(j);
                }
            }
        } else {
            // generic steps: lanes may be idle, first-column fix-up, switch to the column filter
            const int s_end = t_end - tb < 31 ? t_end - tb : 31;
#pragma unroll 1
            for (int s = 0; s <= s_end; ++s) {
                const uint32_t from_line = __shfl_sync(0xffffffffu, chunk, s);
                uint32_t recv = __shfl_up_sync(0xffffffffu, send, 1);
                if (lane == 0) recv = from_line;
                const int j = tb + s - 2 * lane;        // my lo column
                if (j >= 1 && j <= n + 1) {
                    lane16_step<K, ADDSEL>(st, recv, P, g.gup, g.gleft);
                    if (j == 1) lane16_fix_first<K>(st, g, itop);
                    send = lane16_send<K>(st);
                    if (do_store && j >= 2) bnd16[2 * (j - 1)] = (uint16_t)(st.W[K - 1] >> 16);
                    if (j == jswitch) { set_filter(j, FILTER_ALL); filtering = true; }
                    if (filtering) {
                        if (filter_fired(lane16_filter<K>(st), thrS)) slow_path(j);
                    }
                }
            }
        }
    }
    __syncwarp();
    return best;
}

template <bool ADDSEL>
__global__ void __launch_bounds__(WF16_THREADS)
overlap_wf16_kernel(const uint32_t* __restrict__ packed, const PairDesc* __restrict__ pairs,
                    const uint32_t* __restrict__ order, uint32_t n_work, unsigned int* __restrict__ queue,
                    Wf16Params P, uint32_t* __restrict__ scratch, uint32_t scratch_stride,
                    DevResult* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    Wf16Warp w;
    w.packed = packed;
    w.bnd = scratch + (size_t)warp_global * scratch_stride;
    for (;;) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(queue, 1u);
        qi = __shfl_sync(0xffffffffu, qi, 0);
        if (qi >= n_work) break;
        const uint32_t pid = order[qi];
        w.pd = pairs[pid];
        w.g = wf16_make_pair((int)w.pd.m, (int)w.pd.n, P);
        const int m = w.g.m, n = w.g.n;
        // boundary line = table row 0, plus the column codes
        for (int j = 1 + lane; j <= n + 1; j += 32) {
            uint32_t c = (j <= n) ? load_code(packed, w.pd.col_off, (uint32_t)(j - 1)) : 0u;
            w.bnd[j] = w.g.v_row0(j <= n ? j : n) | (code11(c, ADDSEL ? 1u : 0u) << 16);
        }
        __syncwarp();
        long long best = make_key(0, 0u, 1u | (n == 0 ? 2u : 0u));   // cell (0,n): rank 0, H = 0
        w.S = 0;
        int i0 = 0;
        while (i0 < m) {
            const Wf16Strip s = wf16_next_strip(i0, m, w.g.C);
            const bool sb = !s.last;
            switch (s.rows) {
            case 512: best = s.rowscan ? wf16_strip<8, true, ADDSEL>(w, P, i0, sb, best) : wf16_strip<8, false, ADDSEL>(w, P, i0, sb, best); break;
            case 256: best = s.rowscan ? wf16_strip<4, true, ADDSEL>(w, P, i0, sb, best) : wf16_strip<4, false, ADDSEL>(w, P, i0, sb, best); break;
            case 128: best = wf16_strip<2, true, ADDSEL>(w, P, i0, sb, best); break;
            default:  best = wf16_strip<1, true, ADDSEL>(w, P, i0, sb, best); break;
            }
            i0 += s.rows;
        }
        best = warp_max_key(best);
        if (lane == 0) store_result(out + pid, best, m, n, FLAG_KERNEL16);
        __syncwarp();
    }
}

inline cudaError_t wf16_configure() { return cudaSuccess; }

// Launches the kernel on `stream`; grows *scratch (device) as needed.  Returns a cudaError_t as int.
inline int wf16_launch(cudaStream_t stream, int sm_count, const uint32_t* packed, const PairDesc* pairs,
                       const uint32_t* order, uint32_t n_work, unsigned int* queue, const Wf16Params& P,
                       uint32_t max_n, void** scratch, size_t* scratch_cap, DevResult* out)
{
    const int blocks = sm_count * 2;
    const uint32_t warps = (uint32_t)blocks * (WF16_THREADS / 32);
    const uint32_t stride = (max_n + 2 + 31 + 32) & ~31u;
    const size_t need = (size_t)warps * stride * sizeof(uint32_t);
    if (need > *scratch_cap) {
        if (*scratch) cudaFree(*scratch);
        *scratch = nullptr; *scratch_cap = 0;
        cudaError_t e = cudaMalloc(scratch, need);
        if (e != cudaSuccess) return (int)e;
        *scratch_cap = need;
    }
    if (P.addsel)
        overlap_wf16_kernel<true><<<blocks, WF16_THREADS, 0, stream>>>(packed, pairs, order, n_work, queue, P,
                                                                       (uint32_t*)*scratch, stride, out);
    else
        overlap_wf16_kernel<false><<<blocks, WF16_THREADS, 0, stream>>>(packed, pairs, order, n_work, queue, P,
                                                                        (uint32_t*)*scratch, stride, out);
    return (int)cudaGetLastError();
}
#endif // __CUDACC__

} // namespace gp
