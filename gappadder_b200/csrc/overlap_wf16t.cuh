// overlap_wf16t.cuh -- packed 16-bit overlap-DP kernel with a shared-memory increment table
// ("table kernel", sm_100a).  The fast path for what GAPPadder produces: A/C/G/T sequences whose
// column sequence has at most 4094 bases.  Same contract as overlap_wf16.cuh / overlap_wf32.cuh
// (ContigsCompactor::Evaluate before the significance test,
// ContigsCompactor-v0.2.0/ContigsMerger/ContigsCompactor.cpp:1596-1709,:1736-1837).
//
// It keeps overlap_wf16.cuh's arithmetic -- potential-domain values V = 8*(P+1) + 4*z + code clamped
// at P = -1, thermometer origin codes, the z bonus that makes a packed max implement the
// reference's diag > up > left tie rule, the strip/lane layout with the hi group one column behind
// -- and changes what surrounds the two max instructions of a cell, because profiles/ showed the ALU
// pipe (PRMT, LOP3, VIADDMNMX) to be the limiter and the non-steady code to cost as much as the
// steady loop:
//
//  * Column potential only: P = H + (n - j) for every pair (valid whenever n <= 4094).  Column 0 is
//    then one constant, and the value a cell needs to be a scan candidate depends on its column
//    only, so the candidate filter is a max tree over the lane's registers (VIMNMX3) and one packed
//    compare instead of one VIADDMNMX + one VIADD per register.
//  * The diagonal increment comes from a per-warp shared-memory table instead of VIADD + PRMT per
//    register: entry [combo][k] = (inc(row k, c_j), inc(row K+k, c_{j-1})) for the 16 combinations
//    of the two column symbols a lane's lo and hi groups see in one step.  A lane builds its
//    16*K words once per strip and fetches the K words of a step with K/4 LDS.128 (conflict free:
//    a lane's words are 16-byte interleaved across the warp).
//  * Nothing but DP values goes through the shuffle.  The per-column data (top boundary value for
//    lane 0 and the table offset of the column's symbol combination) sits in a 128-entry
//    shared-memory ring, refilled 32 columns at a time from the warp's boundary line one block
//    ahead; the strip's bottom row leaves through a second small ring and one coalesced 16-bit
//    store per 32 steps.  A step's table words are fetched one step ahead.
//  * One loop body per strip height serves the fill, steady and tail phases; two block-uniform
//    flags switch the per-lane range checks and the candidate filter on.
//
// Host/device: the per-lane arithmetic is __host__ __device__; tests/emulate_wf16.cu runs it on
// the CPU against the oracle.
#pragma once
#include <type_traits>
#include "overlap_wf16.cuh"

namespace gp {

constexpr uint32_t WF16T_MAX_N = 4094;            // 8*(n+1)+7 <= 32767
constexpr int WF16T_THREADS = 128;                // 4 warps per CTA
constexpr int WF16T_CTAS_PER_SM = 3;              // 12 warps per SM: 3 x (4 x 17.1 KB) of shared memory
constexpr int WF16T_RING = 128;                   // columns in the ring (stored twice: [0,128) and [128,256))
constexpr int WF16T_SKEW = 3;                     // columns between neighbouring lanes (31*3 + 32 <= WF16T_RING)
constexpr int WF16T_TAB_WORDS = 16 * 8 * 32;      // K = 8: 16 combinations x 8 registers x 32 lanes
constexpr int WF16T_WARP_WORDS = WF16T_TAB_WORDS + 2 * WF16T_RING + 32;
constexpr size_t WF16T_SMEM_BYTES = (size_t)(WF16T_THREADS / 32) * WF16T_WARP_WORDS * sizeof(uint32_t);

inline bool wf16t_pair_ok(uint32_t m, uint32_t n) { return m >= 1 && n >= 1 && n <= WF16T_MAX_N && m <= 0xffffff; }

struct Wf16tParams {
    uint32_t inc_match, inc_mism;   // 16-bit diagonal increments, z bonus included
    uint32_t gup, gleft;            // packed up / left increments under the column potential
    int32_t max_clip;
    int32_t std_scores;             // mismatch == -2 && indel == -2: the kernel with immediate operands
};

inline Wf16tParams wf16t_make_params(int mismatch, int indel, int max_clip)
{
    Wf16tParams p;
    p.inc_match = 4u;
    p.inc_mism = (uint32_t)((mismatch - 1) * 8 + 4) & 0xffffu;
    auto pk = [](int v) { uint32_t h = (uint32_t)(v * 8) & 0xffffu; return h | (h << 16); };
    p.gup = pk(indel);
    p.gleft = pk(indel - 1);
    p.max_clip = max_clip;
    p.std_scores = (mismatch == -2 && indel == -2) ? 1 : 0;
    return p;
}

// Pair geometry: always the column potential.
GP_HD Wf16Pair wf16t_make_pair(int m, int n, const Wf16tParams& P)
{
    Wf16Pair g;
    g.m = m; g.n = n; g.C = P.max_clip;
    g.rowpot = false;
    g.gup = P.gup;
    g.gleft = P.gleft;
    return g;
}

// Boundary-line word of column j (1..n+1): V(0,j) in the low half, the symbol combination
// c_j + 4*c_{j-1} (c_0 = c_{n+1} = 0) above it.
GP_HD uint32_t wf16t_line_word(const Wf16Pair& g, int j, uint32_t cj, uint32_t cjm1)
{
    return g.v_row0(j <= g.n ? j : g.n) | (((cj & 3u) | ((cjm1 & 3u) << 2)) << 16);
}

// Table word of register k for combination (ca = symbol under the lo group, cb = under the hi group).
GP_HD uint32_t wf16t_table_word(uint32_t row_lo, uint32_t row_hi, uint32_t ca, uint32_t cb, const Wf16tParams& P)
{
    return (row_lo == ca ? P.inc_match : P.inc_mism) | ((row_hi == cb ? P.inc_match : P.inc_mism) << 16);
}

template <int K>
struct Lane16t {
    uint32_t W[K];       // (row k @ col j | row K+k @ col j-1 << 16)
    uint32_t up0_prev;   // previous step's `up` of W[0] == this step's diagonal of W[0]
};

template <int K>
GP_HD void lane16t_begin(Lane16t<K>& st, const Wf16Pair& g, int itop)
{
    const uint32_t c0 = g.v_col0(1);                       // the same for every row >= 1
#pragma unroll
    for (int k = 0; k < K; ++k) st.W[k] = c0 | (c0 << 16);
    st.up0_prev = g.v_col0(itop) | (c0 << 16);
}

// After a lane's first step the hi group has "computed" column 0: put the boundary back.
template <int K>
GP_HD void lane16t_fix_first(Lane16t<K>& st, const Wf16Pair& g)
{
    const uint32_t c0 = g.v_col0(1);
#pragma unroll
    for (int k = 0; k < K; ++k) st.W[k] = (st.W[k] & 0xffffu) | (c0 << 16);
}

// One step.  recv: the W[K-1] of the lane above as it was before this step (its high half is the
// value of the row above this lane at column j); inc[k]: this step's table words.
template <int K>
GP_HD void lane16t_step(Lane16t<K>& st, uint32_t recv, const uint32_t (&inc)[K], uint32_t gup, uint32_t gleft)
{
    const uint32_t up0 = p_prmt(recv, st.W[K - 1], 0x5432u);   // (recv.hi16 , old W[K-1].lo16)
    uint32_t diag = st.up0_prev;
    st.up0_prev = up0;
    uint32_t up = up0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const uint32_t left = st.W[k];
        const uint32_t d = p_add2(diag, inc[k]);               // carries the z bonus
        const uint32_t t = p_addmax2(left, gleft, d);
        const uint32_t w = p_addmax2_relu(up, gup, t) & ~TAG_Z2;
        diag = left;
        up = w;
        st.W[k] = w;
    }
}

// max over the lane's registers, per half
GP_HD uint32_t p_max3_2(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
    return __vimax3_s16x2(a, b, c);
#else
    return p_max2(p_max2(a, b), c);
#endif
}
template <int K>
GP_HD uint32_t lane16t_max(const Lane16t<K>& st)
{
    uint32_t a = st.W[0];
    if (K == 2) a = p_max2(a, st.W[1]);
    if (K >= 4) a = p_max3_2(p_max2(a, st.W[1]), st.W[2], st.W[3]);
    if (K == 8) a = p_max3_2(p_max3_2(a, st.W[4], st.W[5]), st.W[6], st.W[7]);
    return a;
}

// Candidate filter under the column potential.  nthr = -(the V a cell has when H = 1) for the lo
// column j (low half) and the hi column j-1 (high half): acc = max(V) + nthr = 8*(H-1) + tags of the
// lane's best cell, compared with thrS = 8*(max(S,1)-1) as in overlap_wf16.cuh.
GP_HD uint32_t wf16t_nthr(const Wf16Pair& g, int j)
{
    const uint32_t lo = (uint32_t)(-8 * (g.n - j + 2)) & 0xffffu;          // 1 <= j: >= -8*(4094+1) = -32760
    const uint32_t hi = (uint32_t)(-8 * (g.n - j + 3)) & 0xffffu;          // column j-1: >= -32768; other j wrap mod 2^16
    return lo | (hi << 16);
}
constexpr uint32_t WF16T_UNARMED = 0x7fff7fffu;   // threshold no acc reaches (acc <= 8*n + 7 < 32767)
constexpr uint32_t WF16T_NSTEP = 0x00080008u;

#if defined(__CUDACC__)
// ---- device side ------------------------------------------------------------------------------

struct Wf16tWarp {                      // warp-uniform state of one pair
    const uint32_t* packed;
    PairDesc pd;
    Wf16Pair g;
    uint32_t* bnd;                      // boundary line in global scratch (see wf16t_line_word)
    uint32_t* smem;                     // this warp's WF16T_WARP_WORDS words of shared memory
    int S;
};

__device__ __forceinline__ uint32_t lds32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(addr), "r"(v) : "memory");
}
template <int K>
__device__ __forceinline__ void lds_inc(uint32_t (&inc)[K], uint32_t addr)
{
    if constexpr (K >= 4) {
#pragma unroll
        for (int q = 0; q < K / 4; ++q)
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(inc[4 * q]), "=r"(inc[4 * q + 1]), "=r"(inc[4 * q + 2]), "=r"(inc[4 * q + 3]) : "r"(addr + q * 512));
    } else if constexpr (K == 2) {
        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(inc[0]), "=r"(inc[1]) : "r"(addr));
    } else {
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(inc[0]) : "r"(addr));
    }
}

// Exact scan of the cells a lane holds after a step (lo column j, hi column j-1), out of line.  A cell can
// only change `best` if it scores at least the lane's best score (and at least 1: the initial best is
// cell (0,n) with H = 0 and rank 0), i.e. V >= 8*(max(S,1) + n - j' + 1) under the column potential; that
// one compare per cell comes first, the rank arithmetic only runs for the cells that pass.
template <int K>
__device__ __noinline__ long long wf16t_scan_cold(WVals<K> v, int m, int n, int C, int itop, int j, long long best)
{
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int jh = j - half;
        if (jh < 1 || jh > n) continue;
        int S = (int)(best >> 32);
        int thrV = 8 * ((S > 1 ? S : 1) + n - jh + 1);
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int vv = half ? (int)(v.W[k] >> 16) : (int)(v.W[k] & 0xffffu);
            if (vv >= thrV) {
                const int i = itop + 1 + k + half * K;
                const uint32_t rk = i <= m ? cell_rank(i, jh, m, n, C) : RANK_MAX + 1u;
                if (rk <= RANK_MAX) {
                    const uint32_t code = (uint32_t)vv & 3u;
                    const long long key = make_key((vv >> 3) - 1 - (n - jh), rk, (code & 1u) | ((code & 2u) ? 0u : 2u));
                    if (key > best) {
                        best = key;
                        S = (int)(best >> 32);
                        thrV = 8 * ((S > 1 ? S : 1) + n - jh + 1);
                    }
                }
            }
        }
    }
    return best;
}

// One strip of 64*K rows starting after table row `i0`.  rowscan: the strip reaches into the last
// C+1 rows.  With store_bottom the low halves of bnd[] are replaced in place by the strip's last row.
//
// Lane l runs WF16T_SKEW = 3 columns behind lane l-1: the value it needs from the lane above at step
// t left that lane at the end of step t-2, so the shuffle is issued a whole step before its result
// is used and two consecutive steps of a lane do not depend on each other through it (with a skew of
// 2 the shuffle sits on the critical path of every step; tools/microbench_step.cu: 9.6 -> 8.6 clocks
// per register-step).  Steps come in blocks of 32.  A block in which every lane is inside columns
// 1..n+1 and no lane can hold a scan candidate runs the branch-free loop; any other block runs the
// checked loop (per-lane range test, first-column fix-up, candidate filter).
template <int K, bool STD>
__device__ __noinline__ long long wf16t_strip(Wf16tWarp& w, const Wf16tParams& P, int i0, bool rowscan, bool store_bottom, long long best)
{
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr uint32_t LANE_BYTES = K >= 4 ? 16u : 4u * K;         // bytes a lane owns per table row
    constexpr uint32_t COMBO_BYTES = K * 128u;                      // table bytes per combination
    constexpr int D = WF16T_SKEW;
    const int lane = threadIdx.x & 31;
    const Wf16Pair g = w.g;
    const int n = g.n, m = g.m;
    const int itop = i0 + lane * 2 * K;
    const uint32_t gup = STD ? 0xfff0fff0u : g.gup, gleft = STD ? 0xffe8ffe8u : g.gleft;
    uint32_t* const bnd = w.bnd;
    const uint32_t tab_base = (uint32_t)__cvta_generic_to_shared(w.smem);
    const uint32_t ring_base = tab_base + WF16T_TAB_WORDS * 4;
    const uint32_t oring_base = ring_base + 2 * WF16T_RING * 4;
    const uint32_t my_tab = tab_base + lane * LANE_BYTES;

    // ---- increment table of this lane's rows ---------------------------------------------------
    {
        uint32_t rc[2 * K];
#pragma unroll
        for (int x = 0; x < 2 * K; ++x) rc[x] = (itop + x < m) ? load_code(w.packed, w.pd.row_off, (uint32_t)(itop + x)) : 0u;
#pragma unroll 1
        for (uint32_t combo = 0; combo < 16; ++combo) {
            const uint32_t ca = combo & 3u, cb = combo >> 2;
            uint32_t wd[K];
#pragma unroll
            for (int k = 0; k < K; ++k) wd[k] = wf16t_table_word(rc[k], rc[K + k], ca, cb, P);
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
        sts32(ring_base + 4 * (jj & (WF16T_RING - 1)), v);
        sts32(ring_base + 4 * ((jj & (WF16T_RING - 1)) + WF16T_RING), v);
    };
    for (int e = lane; e < 2 * WF16T_RING; e += 32) sts32(ring_base + 4 * e, 0u);
    __syncwarp();
    { const int jj = 1 + lane; ring_put(jj, bnd[jj <= n + 1 ? jj : n + 1]); }
    Lane16t<K> st;
    lane16t_begin<K>(st, g, itop);

    // ---- candidate filter state --------------------------------------------------------------------
    int S0;
    {   // start from the warp's best score
        int s = (int)(best >> 32);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { int other = __shfl_xor_sync(FULL, s, o); s = other > s ? other : s; }
        S0 = s;
    }
    uint32_t thrS = filter_thr(S0);
    const int jswitch = n - g.C > 1 ? n - g.C : 1;                        // first candidate column
    const bool rowlane = rowscan && (itop + 2 * K >= m - g.C) && (itop + 1 <= m);
    const int jarm = rowlane ? 1 : jswitch;
    const int t_end = n + 1 + 31 * D;                                     // lane 31's lo group reaches column n+1
    const bool do_store = store_bottom && lane == 31;
    uint32_t recv_next = 0;                                               // shuffle issued one step ahead

    auto slow_path = [&](int j) {                                         // exact scan of this lane's cells
        WVals<K> v;
#pragma unroll
        for (int k = 0; k < K; ++k) v.W[k] = st.W[k];
        best = wf16t_scan_cold<K>(v, m, n, g.C, itop, j, best);
        const int s = (int)(best >> 32);
        thrS = filter_thr(s > S0 ? s : S0);
    };
    __syncwarp();

    // One block of up to 32 steps, two steps per iteration; the table words of a step are fetched one
    // step ahead and the ring word two steps ahead.  EDGE: some lane is outside columns 1..n+1 (range
    // test per lane, first-column fix-up).  FILT: some lane may hold scan candidates (filter).
    auto run_block = [&](auto edge_c, auto filt_c, int tb, int cnt) {
        constexpr bool EDGE = decltype(edge_c)::value, FILT = decltype(filt_c)::value;
        uint32_t p = ring_base + (((uint32_t)(tb - D * lane)) & (WF16T_RING - 1)) * 4u;
        uint32_t optr = oring_base;
        int j = tb - D * lane;                                            // my lo column
        uint32_t nthr = FILT ? wf16t_nthr(g, j) : 0u;                     // follows j (mod 2^16 outside 1..n+1)
        uint32_t incA[K], incB[K];
        uint32_t wordA = lds32(p), wordB = lds32(p + 4);
        lds_inc<K>(incA, my_tab + (wordA >> 16));
        auto step = [&](const uint32_t (&inc)[K], uint32_t word, uint32_t oaddr, int jj) {
            uint32_t recv = recv_next;
            recv_next = __shfl_up_sync(FULL, st.W[K - 1], 1);
            if (lane == 0) recv = word << 16;
            if (!EDGE || (uint32_t)(jj - 1) <= (uint32_t)n) {
                lane16t_step<K>(st, recv, inc, gup, gleft);
                if (EDGE && jj == 1) lane16t_fix_first<K>(st, g);
                if (do_store) sts32(oaddr, st.W[K - 1]);
                if (FILT) {
                    const uint32_t acc = p_add2(lane16t_max<K>(st), nthr);
                    if (filter_fired(acc, jj >= jarm ? thrS : WF16T_UNARMED)) slow_path(jj);
                }
            }
            if (FILT) nthr = p_add2(nthr, WF16T_NSTEP);
        };
#pragma unroll 1
        for (int s = 0; s < cnt; s += 2) {
            lds_inc<K>(incB, my_tab + (wordB >> 16));
            const uint32_t wordA2 = lds32(p + 8);
            step(incA, wordA, optr, j);
            lds_inc<K>(incA, my_tab + (wordA2 >> 16));
            const uint32_t wordB2 = lds32(p + 12);
            step(incB, wordB, optr + 4, j + 1);
            wordA = wordA2; wordB = wordB2;
            p += 8; optr += 8; j += 2;
        }
    };
    using std::true_type;
    using std::false_type;

    for (int tb = 1; tb <= t_end; tb += 32) {
        // the next block's boundary words (L2 latency hidden behind this block)
        uint32_t next_line = 0;
        const bool have_next = tb + 32 <= t_end;
        if (have_next) { const int jj = tb + 32 + lane; next_line = bnd[jj <= n + 1 ? jj : n + 1]; }
        const bool edge = tb < 31 * D + 1 || tb + 31 > n + 1;             // some lane outside columns 1..n+1
        const bool filt = rowscan || tb + 31 >= jswitch;                  // some lane may hold candidates
        const int cnt = t_end - tb + 1 < 32 ? ((t_end - tb + 2) & ~1) : 32;   // steps past t_end find every lane out of range
        if (!edge) { if (!filt) run_block(false_type(), false_type(), tb, cnt); else run_block(false_type(), true_type(), tb, cnt); }
        else       { if (!filt) run_block(true_type(), false_type(), tb, cnt);  else run_block(true_type(), true_type(), tb, cnt); }
        __syncwarp();
        if (store_bottom) {                                               // bottom row of the columns lane 31 finished
            const int c = tb + lane - (31 * D + 1);
            const uint32_t v = lds32(oring_base + 4 * lane);
            if (c >= 1 && c <= n && lane < cnt) reinterpret_cast<uint16_t*>(bnd)[2 * c] = (uint16_t)(v >> 16);
        }
        if (have_next) ring_put(tb + 32 + lane, next_line);
        if (filt) {                                                       // share the best score across the warp
            int s = (int)(best >> 32);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { int other = __shfl_xor_sync(FULL, s, o); s = other > s ? other : s; }
            S0 = s > S0 ? s : S0;
            thrS = filter_thr(S0);
        }
        __syncwarp();
    }
    w.S = S0;
    return best;
}

template <bool STD>
__global__ void __launch_bounds__(WF16T_THREADS, WF16T_CTAS_PER_SM)
overlap_wf16t_kernel(const uint32_t* __restrict__ packed, const PairDesc* __restrict__ pairs,
                     const uint32_t* __restrict__ order, uint32_t n_work, unsigned int* __restrict__ queue,
                     Wf16tParams P, uint32_t* __restrict__ scratch, uint32_t scratch_stride,
                     DevResult* __restrict__ out, const unsigned int* __restrict__ n_work_dev)
{
    extern __shared__ uint32_t wf16t_smem[];
    if (n_work_dev) n_work = *n_work_dev;             // work list filled on the device (overlap_wf16c.cuh's retries)
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    Wf16tWarp w;
    w.packed = packed;
    w.bnd = scratch + (size_t)warp_global * scratch_stride;
    w.smem = wf16t_smem + (threadIdx.x >> 5) * WF16T_WARP_WORDS;
    for (;;) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(queue, 1u);
        qi = __shfl_sync(0xffffffffu, qi, 0);
        if (qi >= n_work) break;
        const uint32_t pid = order[qi];
        w.pd = pairs[pid];
        w.g = wf16t_make_pair((int)w.pd.m, (int)w.pd.n, P);
        const int m = w.g.m, n = w.g.n;
        // boundary line = table row 0 plus the column symbol combinations
        for (int j = 1 + lane; j <= n + 1; j += 32) {
            const uint32_t cj = (j <= n) ? load_code(packed, w.pd.col_off, (uint32_t)(j - 1)) : 0u;
            const uint32_t cp = (j >= 2) ? load_code(packed, w.pd.col_off, (uint32_t)(j - 2)) : 0u;
            w.bnd[j] = wf16t_line_word(w.g, j, cj, cp);
        }
        __syncwarp();
        long long best = make_key(0, 0u, 1u | (n == 0 ? 2u : 0u));   // cell (0,n): rank 0, H = 0
        w.S = 0;
        int i0 = 0;
        while (i0 < m) {
            const Wf16Strip s = wf16_next_strip(i0, m, w.g.C);
            const bool sb = !s.last;
            switch (s.rows) {
            case 512: best = wf16t_strip<8, STD>(w, P, i0, s.rowscan, sb, best); break;
            case 256: best = wf16t_strip<4, STD>(w, P, i0, s.rowscan, sb, best); break;
            case 128: best = wf16t_strip<2, STD>(w, P, i0, s.rowscan, sb, best); break;
            default:  best = wf16t_strip<1, STD>(w, P, i0, s.rowscan, sb, best); break;
            }
            i0 += s.rows;
        }
        best = warp_max_key(best);
        if (lane == 0) store_result(out + pid, best, m, n, FLAG_KERNEL16);
        __syncwarp();
    }
}

inline cudaError_t wf16t_configure()
{
    cudaError_t e = cudaFuncSetAttribute(overlap_wf16t_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WF16T_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(overlap_wf16t_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WF16T_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(overlap_wf16t_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(overlap_wf16t_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

// Launches the kernel on `stream`; grows *scratch (device) as needed.  Returns a cudaError_t as int.
inline int wf16t_launch(cudaStream_t stream, int sm_count, const uint32_t* packed, const PairDesc* pairs,
                        const uint32_t* order, uint32_t n_work, unsigned int* queue, const Wf16tParams& P,
                        uint32_t max_n, void** scratch, size_t* scratch_cap, DevResult* out,
                        const unsigned int* n_work_dev = nullptr)
{
    const int blocks = sm_count * WF16T_CTAS_PER_SM;
    const uint32_t warps = (uint32_t)blocks * (WF16T_THREADS / 32);
    const uint32_t stride = (max_n + 2 + 31 + 32) & ~31u;
    const size_t need = (size_t)warps * stride * sizeof(uint32_t);
    if (need > *scratch_cap) {
        if (*scratch) cudaFree(*scratch);
        *scratch = nullptr; *scratch_cap = 0;
        cudaError_t e = cudaMalloc(scratch, need);
        if (e != cudaSuccess) return (int)e;
        *scratch_cap = need;
    }
    if (P.std_scores)
        overlap_wf16t_kernel<true><<<blocks, WF16T_THREADS, WF16T_SMEM_BYTES, stream>>>(packed, pairs, order, n_work, queue, P,
                                                                                         (uint32_t*)*scratch, stride, out, n_work_dev);
    else
        overlap_wf16t_kernel<false><<<blocks, WF16T_THREADS, WF16T_SMEM_BYTES, stream>>>(packed, pairs, order, n_work, queue, P,
                                                                                          (uint32_t*)*scratch, stride, out, n_work_dev);
    return (int)cudaGetLastError();
}
#endif // __CUDACC__

} // namespace gp
