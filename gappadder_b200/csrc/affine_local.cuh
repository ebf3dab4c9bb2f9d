// affine_local.cuh -- TERefiner's affine-gap local aligner on the device (sm_100a).
//
// What it stands in for: LocalAlignment::optAlign (/root/reference/TERefiner/algorithms/local_alignment.cpp:1036-1049),
// i.e. aln_stdaln(ref, sgmt, &aln_param_blast, ALN_TYPE_LOCAL, 1): match +1, mismatch -3, anything against a non-A/C/G/T
// letter -2 (aln_sm_blast, :193-199), gap open 5, gap extension 2, band 50 (:206).  Unlike BASELINE config 2's BWA call,
// this aligner's source IS in the reference tree, so parity is PINNED: oracle/build_ref.sh compiles that file as it
// lies and the tests compare start/end coordinates and score with it.
//
// The reference works in three passes, and so does this file:
//   1. forward (aln_local_core :565-609): the whole len1 x len2 table, local (scores clamped at 0), affine, keeping the
//      FIRST strict maximum in column-major order (seq2 outer, seq1 inner).  Two details differ from the textbook
//      recurrence and are reproduced: the horizontal gap state of a cell is dropped unless the cell to its left scores
//      more than gap_open + gap_ext (:594-599), and the vertical one is only looked at below a positive cell (:590-593;
//      that one changes no value, see aff_lane_step).  This is the O(len1*len2) part: `affine_forward_kernel`, an anti-diagonal
//      wavefront with one warp per pair, 512-row strips, 16 rows per lane in registers, the lane-to-lane hand-over by
//      shuffle one step ahead, the strip-to-strip hand-over through a boundary line in L2, the substitution scores from
//      a per-warp shared-memory table read with conflict-free LDS.128 (the machinery of flank_place.cuh).
//   2. reverse (:617-690): from the end cell backwards inside a band whose edges follow the running best score, a
//      sequential heuristic (the band of a column depends on what the previous column found, cells outside it keep stale
//      values that are read again when the band widens); it stops at the first cell that scores score + open + ext.
//   3. a banded GLOBAL alignment of the sub-rectangle (:715-739, aln_global_core :328-508), band 50, doubled until its
//      score agrees; the reference walks its traceback matrix only to report where the path starts, which is a
//      two-bit tag carried forward with each state here (no traceback matrix).
//   Passes 2 and 3 exist twice.  `aff_epilogue` follows the reference statement by statement (one cell after the other);
//   `aff_epilogue_warp` computes the same values column by column with the 32 lanes of a warp spread over a column's band,
//   and is what `affine_epilogue_kernel` runs, one warp per pair (the cost of these passes is the square of the ALIGNED
//   length, not of the sequence lengths).  Both are __host__ __device__ text: compiled for the host they are what the CPU
//   tests compare with each other and with the reference (tests/emulate_affine.cu), so the device path carries no
//   arithmetic of its own.
//
// Domain: min(len1, len2) * match + open + ext <= 32000 (below the reference's 16-bit overflow rescaling, :573-588), both
// lengths < 2^20, gap penalties and scores small positive / negative integers (aff_params_ok).  Everything this file does has
// been measured step by step: DESIGN.md section 4 "Affine local aligner", profiles/affine_*.
#pragma once
#include "common.cuh"

namespace gp {

struct AffParams {
    int match, mismatch, nscore;      // substitution scores: equal A/C/G/T, unequal A/C/G/T, anything with another letter
    int q, r;                         // gap open, gap extension (a gap of length g costs q + g*r)
    int band;                         // initial band of the global fill
};

constexpr int AFF_MINOR_INF = -1073741823;        // stdaln.h:84
constexpr int AFF_OVERFLOW = 32000;               // LOCAL_OVERFLOW_THRESHOLD, local_alignment.cpp:235
constexpr uint32_t AFF_MAX_LEN = (1u << 20) - 1u;

constexpr uint32_t AFF_FLAG_NO_MATCH = 1u;        // score 0: the reference reads path[-1] (:611-614, :817); coordinates 0 here
constexpr uint32_t AFF_FLAG_UNDEFINED = 2u;       // the reverse band collapsed (the reference's loop :654 would run out of its array)
constexpr uint32_t AFF_FLAG_POTENTIAL_BUG = 4u;   // the reference prints "Potential bug" and reports score -1 (:727-730)

struct DevLocal {                                 // same layout as gp_local_result
    int32_t score;
    int32_t start1, end1, start2, end2;
    uint32_t flags;
};

inline bool aff_params_ok(const AffParams& P)
{
    return P.match >= 1 && P.match <= 64 && P.mismatch <= 0 && P.mismatch >= -1024 && P.nscore <= 0 && P.nscore >= -1024 &&
           P.q >= 0 && P.q <= 1024 && P.r >= 1 && P.r <= 64 && P.band >= 1 && P.band <= (1 << 20);
}
inline bool aff_pair_ok(uint32_t len1, uint32_t len2, const AffParams& P)
{
    const uint64_t lo = len1 < len2 ? len1 : len2;
    return len1 <= AFF_MAX_LEN && len2 <= AFF_MAX_LEN && lo * (uint64_t)P.match + (uint64_t)(P.q + P.r) <= (uint64_t)AFF_OVERFLOW;
}

// Substitution score of two 4-bit codes (gp_pack_sequences: A C G T = 0..3, everything else >= 4), aln_sm_blast's shape.
__host__ __device__ __forceinline__ int aff_sc(uint32_t a, uint32_t b, const AffParams& P)
{
    return (a > 3u || b > 3u) ? P.nscore : (a == b ? P.match : P.mismatch);
}

// ---- pass 1: forward -------------------------------------------------------------------------------------------------

constexpr int AF_R = 16;                         // rows per lane
constexpr int AF_STRIP = 32 * AF_R;              // 512 rows per strip
constexpr int AF_THREADS = 128;                  // 4 warps per CTA
constexpr int AF_CTAS_PER_SM = 3;
constexpr int AF_SKEW = 2;                       // columns between neighbouring lanes
constexpr int AF_CLASSES = 5;                    // column symbol classes: A C G T other
constexpr int AF_TAB_WORDS = AF_CLASSES * AF_R * 32;   // 10 KB per warp
constexpr int AF_RING = 64;                      // lane 0's inputs: {cell above, symbol class} of the next columns
constexpr int AF_WARP_WORDS = AF_TAB_WORDS + 2 * AF_RING;
constexpr size_t AF_SMEM_BYTES = (size_t)(AF_THREADS / 32) * AF_WARP_WORDS * sizeof(uint32_t);
constexpr int AF_NEG = -(1 << 29);               // substitution score of a virtual row: its diagonal never wins
constexpr int AF_FBIAS = 1 << 15;

// What a row hands to the row below at one column: its score and its vertical gap state, in one word.
__host__ __device__ __forceinline__ uint32_t aff_pack(int h, int f) { return ((uint32_t)h << 16) | (uint32_t)(f + AF_FBIAS); }
__host__ __device__ __forceinline__ int aff_h(uint32_t w) { return (int)(w >> 16); }
__host__ __device__ __forceinline__ int aff_f(uint32_t w) { return (int)(w & 0xffffu) - AF_FBIAS; }

// Best cell: first strict maximum with seq2 (columns, j) outer and seq1 (rows, i) inner (:600-603) == the maximum of
// (score, -j, -i).  i, j < 2^20.
__host__ __device__ __forceinline__ long long aff_key(int score, int i, int j)
{
    return ((long long)score << 40) | ((long long)(AFF_MAX_LEN - (uint32_t)j) << 20) | (long long)(AFF_MAX_LEN - (uint32_t)i);
}

template <int R>
struct AffLane {
    int H[R];          // H(i, j) of my rows at the column of my previous step
    int Hq[R];         // H - (q + r): what a gap opened from that cell starts with (used twice: by the row below, by the next column)
    int E[R];          // the horizontal gap state stored with it (what the reference keeps in the low half of eh[], :597)
    int up_prev;       // H(row above my first, j-1): the diagonal of my first row
    int bscore;        // score of `best`
    long long best;
};

template <int R>
__host__ __device__ __forceinline__ void aff_lane_begin(AffLane<R>& st, int q, int r)
{
#pragma unroll
    for (int x = 0; x < R; ++x) { st.H[x] = 0; st.Hq[x] = -(q + r); st.E[x] = 0; }       // column 0 (:566)
    st.up_prev = 0;
}

__host__ __device__ __forceinline__ int aff_max3_relu(int a, int b, int c)
{
#ifdef __CUDA_ARCH__
    return __vimax3_s32_relu(a, b, c);
#else
    int m = a > b ? a : b; m = m > c ? m : c; return m > 0 ? m : 0;
#endif
}

// One column of my R rows.  recv = {H, F} of the row above my first at this column; inc[x] = substitution score of row x
// against this column's letter (AF_NEG for a virtual row).  Returns {H, F} of my last row.
//
// The reference's cell (:589-603), with q = open, r = ext:
//     h = max(0, H(i-1,j-1) + s)
//     if H(i-1,j) > 0:  f = max(f - r, H(i-1,j) - q - r);  h = max(h, f)          [f is left as it was otherwise]
//     if H(i,j-1) >= q + r + 1:  e = max(E(i,j-1) - r, H(i,j-1) - q - r);  h = max(h, e)      else e = 0
// The `if` around f changes nothing: a stale f is <= 0 when the guard first fails (h >= f held one row up and h = 0
// there) and only decreases until it is replaced, so skipped or not, a non-positive f never shows in h; the plain
// recurrence f = max(f - r, H(i-1,j) - q - r) has the same positive part.
// The guard around e does change values and is kept, in this form: with t = H(i,j-1) - q - r the guard is t >= 1, and
//     e = min(max(E(i,j-1) - r, t), t * 2^15)
// is the reference's e when the guard holds (t * 2^15 >= 2^15 exceeds every score) and some value <= 0 when it does not.
// A stored state <= 0 is as good as the reference's 0: it never shows in h (clamped at 0), and one column on it is either
// discarded again or loses against t' >= 1 (e' = max(E - r, t') = t' for every E <= r + 1).  One min on the ALU pipe and one
// multiply on the other pipe instead of a compare and a select, both on the ALU pipe, which is this kernel's bound.
template <int R>
__host__ __device__ __forceinline__ uint32_t aff_lane_step(AffLane<R>& st, uint32_t recv, const int* inc, int q, int r, int j, int itop)
{
    const int qr = q + r;
    int d[R];
    d[0] = st.up_prev + inc[0];
#pragma unroll
    for (int x = 1; x < R; ++x) d[x] = st.H[x - 1] + inc[x];
    int hu = aff_h(recv), f = aff_f(recv);
    st.up_prev = hu;
    int huq = hu - qr;
    int colmax = 0;
#pragma unroll
    for (int x = 0; x < R; ++x) {
        const int a = f - r;
        f = a > huq ? a : huq;
        const int t = st.Hq[x];
        const int e1 = st.E[x] - r;
        int e = e1 > t ? e1 : t;
        const int cap = t * 32768;
        e = e < cap ? e : cap;
        const int h = aff_max3_relu(d[x], e, f);
        st.E[x] = e;
        st.H[x] = h;
        huq = h - qr;
        st.Hq[x] = huq;
        colmax = h > colmax ? h : colmax;
    }
    if (colmax >= st.bscore && colmax > 0) {
        int x0 = R - 1;
#pragma unroll
        for (int x = R - 2; x >= 0; --x) x0 = st.H[x] == colmax ? x : x0;      // the smallest row that reaches it
        const long long k = aff_key(colmax, itop + x0 + 1, j);
        if (k > st.best) { st.best = k; st.bscore = colmax; }
    }
    return aff_pack(st.H[R - 1], f);
}

#ifdef __CUDACC__
__global__ void __launch_bounds__(AF_THREADS, AF_CTAS_PER_SM)
affine_forward_kernel(const uint32_t* __restrict__ packed, const PairDesc* __restrict__ pairs, const uint32_t* __restrict__ order,
                      uint32_t n_work, unsigned int* __restrict__ queue, AffParams P,
                      uint32_t* __restrict__ scratch, uint32_t scratch_stride, DevLocal* __restrict__ out)
{
    extern __shared__ uint32_t af_smem[];
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr int R = AF_R, D = AF_SKEW;
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t* const bnd = scratch + (size_t)warp_global * scratch_stride;
    uint32_t* const tab = af_smem + (threadIdx.x >> 5) * AF_WARP_WORDS;
    uint32_t* const ring = tab + AF_TAB_WORDS;                      // [slot] = cell above, [AF_RING + slot] = symbol class
    const uint32_t tab_addr = (uint32_t)__cvta_generic_to_shared(tab) + (uint32_t)lane * 16u;
    const uint32_t row0 = aff_pack(0, -(P.q + P.r));                // row 0: H = 0, no vertical gap (:566, f = 0 at :569)

    for (;;) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(queue, 1u);
        qi = __shfl_sync(FULL, qi, 0);
        if (qi >= n_work) break;
        const uint32_t pid = order[qi];
        const PairDesc pd = pairs[pid];
        const int m = (int)pd.m, n = (int)pd.n;                     // m = len1 (rows, the reference's inner loop), n = len2
        const int n_strips = (m + AF_STRIP - 1) / AF_STRIP;
        const int pad = n_strips * AF_STRIP - m;                     // virtual rows above row 1: they stay 0, like row 0

        for (int j = lane; j <= n; j += 32) bnd[j] = row0;
        __syncwarp();

        AffLane<R> st;
        st.best = aff_key(0, 0, 0);
        st.bscore = 0;

        for (int s = 0; s < n_strips; ++s) {
            const int itop = s * AF_STRIP + lane * R - pad;          // 0-based index of my first row; < 0: virtual
            const bool last = s == n_strips - 1;
            {
                uint32_t code[R];
#pragma unroll
                for (int x = 0; x < R; ++x) code[x] = itop + x >= 0 ? load_code(packed, pd.row_off, (uint32_t)(itop + x)) : 0xffu;
                __syncwarp();
#pragma unroll 1
                for (uint32_t c = 0; c < (uint32_t)AF_CLASSES; ++c) {
#pragma unroll
                    for (int g = 0; g < R / 4; ++g) {
                        uint32_t w[4];
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            const uint32_t rcx = code[4 * g + x];
                            w[x] = (uint32_t)(rcx == 0xffu ? AF_NEG : aff_sc(rcx, c, P));
                        }
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(tab_addr + c * (R * 128u) + g * 512u),
                                     "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
                    }
                }
            }
            aff_lane_begin<R>(st, P.q, P.r);
            uint32_t bottom = row0;                                  // my last row at the column of my previous step
            uint32_t mysym = 0;                                      // that column's symbol class
            uint32_t recv_next = row0;
            uint32_t sym_next = 0;
            uint32_t nextV, nextC;                                   // lane 0's inputs, one chunk of 32 columns ahead
            {
                const int j = 1 + lane;
                nextV = j <= n ? __ldcg(bnd + j) : row0;
                nextC = j <= n ? min(load_code(packed, pd.col_off, (uint32_t)(j - 1)), 4u) : 0u;
            }
            const int steps = n + D * 31;
#pragma unroll 1
            for (int t = 1; t <= steps; ++t) {
                if (((t - 1) & 31) == 0) {
                    __syncwarp();
                    const int slot = (t - 1 + lane) & (AF_RING - 1);
                    ring[slot] = nextV;
                    ring[AF_RING + slot] = nextC;
                    const int j = t + 32 + lane;
                    nextV = j <= n ? __ldcg(bnd + j) : row0;
                    nextC = j <= n ? min(load_code(packed, pd.col_off, (uint32_t)(j - 1)), 4u) : 0u;
                    __syncwarp();
                }
                uint32_t recv = recv_next;
                uint32_t csym = sym_next;
                recv_next = __shfl_up_sync(FULL, bottom, 1);           // used one step from now
                sym_next = __shfl_up_sync(FULL, mysym, 1);
                if (lane == 0) {
                    recv = ring[(t - 1) & (AF_RING - 1)];
                    csym = ring[AF_RING + ((t - 1) & (AF_RING - 1))];
                }
                const int j = t - D * lane;
                if ((uint32_t)(j - 1) < (uint32_t)n) {
                    int inc[R];
                    const uint32_t a = tab_addr + csym * (R * 128u);
#pragma unroll
                    for (int g = 0; g < R / 4; ++g)
                        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(inc[4 * g]), "=r"(inc[4 * g + 1]), "=r"(inc[4 * g + 2]), "=r"(inc[4 * g + 3])
                                     : "r"(a + g * 512u));
                    bottom = aff_lane_step<R>(st, recv, inc, P.q, P.r, j, itop);
                    mysym = csym;
                    if (lane == 31 && !last) bnd[j] = bottom;          // the next strip's row above
                }
            }
            __syncwarp();
        }
        long long k = warp_max_key(st.best);
        if (lane == 0) {
            DevLocal res;
            res.score = (int32_t)(k >> 40);
            res.end1 = (int32_t)(AFF_MAX_LEN - (uint32_t)(k & 0xfffffll));
            res.end2 = (int32_t)(AFF_MAX_LEN - (uint32_t)((k >> 20) & 0xfffffll));
            res.start1 = 0; res.start2 = 0;
            res.flags = res.score > 0 ? 0u : AFF_FLAG_NO_MATCH;
            if (res.score <= 0) { res.end1 = 0; res.end2 = 0; }
            out[pid] = res;
        }
        __syncwarp();
    }
}
#endif // __CUDACC__

// ---- passes 2 and 3: where the alignment starts ------------------------------------------------------------------------

// A sequence of 4-bit codes in the packed table (device: global memory; host tests: the same words on the host).
struct AffSeq {
    const uint32_t* words;
    __host__ __device__ __forceinline__ uint32_t at1(int pos1) const      // 1-based, as the reference indexes after --seq (:563)
    {
        const uint32_t p = (uint32_t)(pos1 - 1);
        return (words[p >> 3] >> ((p & 7u) * 4u)) & 15u;
    }
};

// Ints of scratch aff_epilogue needs for a pair whose forward pass ended in row end1.
__host__ __device__ __forceinline__ size_t aff_epilogue_words(int end1) { return 8u * ((size_t)end1 + 2u); }

// One state column of the global fill: scores of the three states and, for each, which of the three cells next to the
// corner its path leaves from (2 bits each in `tag`: bits 0-1 M, 2-3 I, 4-5 D).
struct AffCol { int* M; int* I; int* D; int* tag; };
constexpr int AFF_FROM_M = 0, AFF_FROM_I = 1, AFF_FROM_D = 2;            // stdaln.h:74-76; as tags: path leaves (1,1) / (0,1) / (1,0)

// set_M / set_I / set_D of the reference (:241-299, gap_end < 0 as aln_local_core passes it, :719), followed by the tag
// of the chosen predecessor state instead of a traceback cell.
__host__ __device__ __forceinline__ void aff_set_M(const AffCol& cur, int i, const AffCol& p, int pi, int sc, bool corner)
{
    int v, t;
    if (p.M[pi] >= p.I[pi]) {
        if (p.M[pi] >= p.D[pi]) { v = p.M[pi]; t = AFF_FROM_M; } else { v = p.D[pi]; t = AFF_FROM_D; }
    } else {
        if (p.I[pi] > p.D[pi]) { v = p.I[pi]; t = AFF_FROM_I; } else { v = p.D[pi]; t = AFF_FROM_D; }
    }
    cur.M[i] = v + sc;
    const int tg = corner ? AFF_FROM_M : (p.tag[pi] >> (2 * t)) & 3;
    cur.tag[i] = (cur.tag[i] & ~3) | tg;
}
__host__ __device__ __forceinline__ void aff_set_I(const AffCol& cur, int i, const AffCol& p, int pi, int q, int r, bool corner)
{
    int t;
    if (p.M[pi] - q > p.I[pi]) { t = AFF_FROM_M; cur.I[i] = p.M[pi] - q - r; } else { t = AFF_FROM_I; cur.I[i] = p.I[pi] - r; }
    const int tg = corner ? AFF_FROM_I : (p.tag[pi] >> (2 * t)) & 3;
    cur.tag[i] = (cur.tag[i] & ~12) | (tg << 2);
}
__host__ __device__ __forceinline__ void aff_set_D(const AffCol& cur, int i, const AffCol& p, int pi, int q, int r, bool corner)
{
    int t;
    if (p.M[pi] - q > p.D[pi]) { t = AFF_FROM_M; cur.D[i] = p.M[pi] - q - r; } else { t = AFF_FROM_D; cur.D[i] = p.D[pi] - r; }
    const int tg = corner ? AFF_FROM_D : (p.tag[pi] >> (2 * t)) & 3;
    cur.tag[i] = (cur.tag[i] & ~48) | (tg << 4);
}
__host__ __device__ __forceinline__ void aff_set_inf(const AffCol& c, int i) { c.M[i] = c.I[i] = c.D[i] = AFF_MINOR_INF; }

// aln_global_core (:328-508) on s1[o1+1 .. o1+len1] x s2[o2+1 .. o2+len2] with band b: returns the score, *first = where
// the traced path leaves the corner (the reference's last path element, :486-498).  work: 8*(len1+1) ints.
__host__ __device__ inline int aff_global(const AffSeq& s1, int o1, int len1, const AffSeq& s2, int o2, int len2, const AffParams& P,
                                          int b, int* work, int* first)
{
    const int q = P.q, r = P.r;
    int b1, b2;
    if (len1 > len2) { b1 = len1 - len2 + b; b2 = b; } else { b1 = b; b2 = len2 - len1 + b; }      // :356-362
    if (b1 > len1) b1 = len1;
    if (b2 > len2) b2 = len2;
    const int w = len1 + 1;
    AffCol curr{work, work + w, work + 2 * w, work + 3 * w}, last{work + 4 * w, work + 5 * w, work + 6 * w, work + 7 * w};
    for (int i = 0; i < 8 * w; ++i) work[i] = 0;
    for (int i = 0; i < w; ++i) { aff_set_inf(curr, i); aff_set_inf(last, i); }
    auto swap = [&]() { const AffCol t = curr; curr = last; last = t; };
    auto sc = [&](int i, int j) { return aff_sc(s1.at1(o1 + i), s2.at1(o2 + j), P); };

    // first row (:375-381)
    curr.M[0] = 0;
    for (int i = 1; i < b1; ++i) { aff_set_inf(curr, i); aff_set_D(curr, i, curr, i - 1, q, r, i == 1); }
    swap();

    int j;
    // part 1 (:383-402) and its last row (:403-423): the band still touches row 0
    const int tmp_end = b2 < len2 ? b2 : len2 - 1;
    auto part1 = [&](int jj) {
        aff_set_inf(curr, 0);
        aff_set_I(curr, 0, last, 0, q, r, jj == 1);
        const int end = (jj + b1 <= len1 + 1) ? (jj + b1 - 1) : len1;
        int i;
        for (i = 1; i != end; ++i) {
            aff_set_M(curr, i, last, i - 1, sc(i, jj), i == 1 && jj == 1);
            aff_set_I(curr, i, last, i, q, r, false);
            aff_set_D(curr, i, curr, i - 1, q, r, false);
        }
        aff_set_M(curr, i, last, i - 1, sc(i, jj), i == 1 && jj == 1);
        aff_set_D(curr, i, curr, i - 1, q, r, false);
        if (jj + b1 - 1 > len1) aff_set_I(curr, i, last, i, q, r, false); else curr.I[i] = AFF_MINOR_INF;
        swap();
    };
    for (j = 1; j <= tmp_end; ++j) part1(j);
    if (j == len2 && b2 != len2 - 1) { part1(j); ++j; }
    // part 2 (:425-439): both band edges inside the table
    for (; j <= len2 - b2 + 1; ++j) {
        aff_set_inf(curr, j - b2);
        const int end = j + b1 - 1;
        int i;
        for (i = j - b2 + 1; i != end; ++i) {
            aff_set_M(curr, i, last, i - 1, sc(i, j), false);
            aff_set_I(curr, i, last, i, q, r, false);
            aff_set_D(curr, i, curr, i - 1, q, r, false);
        }
        aff_set_M(curr, i, last, i - 1, sc(i, j), false);
        aff_set_D(curr, i, curr, i - 1, q, r, false);
        curr.I[i] = AFF_MINOR_INF;
        swap();
    }
    // part 3 (:441-455) and the last row (:456-470): the band reaches the last row of seq1
    auto part3 = [&](int jj) {
        aff_set_inf(curr, jj - b2);
        int i;
        for (i = jj - b2 + 1; i < len1; ++i) {
            aff_set_M(curr, i, last, i - 1, sc(i, jj), false);
            aff_set_I(curr, i, last, i, q, r, false);
            aff_set_D(curr, i, curr, i - 1, q, r, false);
        }
        aff_set_M(curr, i, last, len1 - 1, sc(i, jj), false);
        aff_set_I(curr, i, last, i, q, r, false);
        aff_set_D(curr, i, curr, i - 1, q, r, false);
        swap();
    };
    for (; j < len2; ++j) part3(j);
    if (j == len2) part3(j);

    // where the traceback starts (:472-478)
    int mx = last.M[len1], t = last.tag[len1] & 3;
    if (last.I[len1] > mx) { mx = last.I[len1]; t = (last.tag[len1] >> 2) & 3; }
    if (last.D[len1] > mx) { mx = last.D[len1]; t = (last.tag[len1] >> 4) & 3; }
    *first = t;
    return mx;
}

// Passes 2 and 3 of aln_local_core for one pair, given pass 1's result (score > 0, end cell).  s1/s2: the pair's
// sequences; work: aff_epilogue_words(end1) ints.  Fills start1/start2 (1-based, aln_stdaln_aux :817-821), the final score
// and flags.
__host__ __device__ inline void aff_epilogue(const AffSeq& s1, const AffSeq& s2, const AffParams& P, int score_f, int end_i, int end_j,
                                             int* work, DevLocal* res)
{
    const int q = P.q, r = P.r, qr = q + r;
    res->end1 = end_i; res->end2 = end_j; res->flags = 0;
    // reverse pass (:617-690).  eh[k] = H << 16 | E, as there.
    int* eh = work;
    for (int i = 0; i <= end_i; ++i) eh[i] = 0;
    int score_r = aff_sc(s1.at1(end_i), s2.at1(end_j), P);
    int start_i = end_i, start_j = end_j;
    eh[end_i] = (qr + score_r) << 16;
    int hi = end_i - 1, lo = end_i - 3;
    if (lo <= 0) lo = 0;
    bool undefined = false;
    for (int j = end_j - 1; j != 0; --j) {
        if (hi < lo) { undefined = true; break; }
        int last_h = 0, f = 0, i;
        const uint32_t c2 = s2.at1(j);
        bool stop = false;
        for (i = hi; i != lo; --i) {
            const int cell = eh[i + 1];
            int h = (cell >> 16) + aff_sc(s1.at1(i), c2, P);
            if (h < 0) h = 0;
            if (last_h > 0) {
                f = (f > last_h - q) ? f - r : last_h - qr;
                if (h < f) h = f;
            }
            const int hl = eh[i] >> 16;
            int e = ((cell & 0xffff) > hl - q) ? (cell & 0xffff) - r : hl - qr;
            if (e < 0) e = 0;
            if (h < e) h = e;
            eh[i + 1] = (last_h << 16) | e;
            last_h = h;
            if (score_r < h) {
                score_r = h; start_i = i; start_j = j;
                if (score_r - qr == score_f) { stop = true; break; }
            }
        }
        if (stop) break;
        eh[i + 1] = last_h << 16;
        if ((eh[hi] >> 16) <= qr) --hi;
        if (hi <= 0) hi = 0;
        lo = start_i - (start_j - j) - (score_r + (start_j - j) * P.match) / r - 1;
        if (lo <= 0) lo = 0;
    }
    if (undefined) { res->flags |= AFF_FLAG_UNDEFINED; res->start1 = 0; res->start2 = 0; res->score = score_f; return; }
    score_r -= qr;                                                    // :705-706

    // global fill of the rectangle, band doubled until the score agrees (:715-725)
    const int len1 = end_i - start_i + 1, len2 = end_j - start_j + 1;
    int jmax = (end_i - start_i > end_j - start_j) ? end_i - start_i : end_j - start_j;
    ++jmax;
    int score_g = 0, first = AFF_FROM_M;
    for (int b = P.band;; b <<= 1) {
        score_g = aff_global(s1, start_i - 1, len1, s2, start_j - 1, len2, P, b, work, &first);
        if (score_g == score_r || score_f == score_g) break;
        if (b > jmax) break;
    }
    if (score_r > score_g && score_f > score_g) { res->score = -1; res->flags |= AFF_FLAG_POTENTIAL_BUG; }      // :727-730
    else res->score = score_g;
    // last path element (:735-738, :817-820): local (1,1), (0,1) or (1,0), shifted by start - 1; a 0 reads as 1
    const int pi = (first == AFF_FROM_I ? 0 : 1) + start_i - 1, pj = (first == AFF_FROM_D ? 0 : 1) + start_j - 1;
    res->start1 = pi ? pi : 1;
    res->start2 = pj ? pj : 1;
}

// ---- passes 2 and 3, one WARP per pair ----------------------------------------------------------------------------------
//
// aff_epilogue above is the reference's order of evaluation, one cell after the other; with one thread per pair a batch
// waits for its longest pair's half a million dependent cells.  aff_epilogue_warp computes the very same values column
// by column with the 32 lanes spread over a column's band, lane l on the cells l, 32 + l, 64 + l, ... of the column (so
// that a warp's loads and stores are whole cache lines):
//   * a cell's score without its vertical gap state depends on the previous column only: all cells at once;
//   * the vertical state is a running maximum down the column, F(i) = max over cells k already passed of
//     (h'(k) - q - |k - i| r), h' = the score without F (an F-derived score never opens a better gap than the score it
//     came from, q >= 0): a prefix maximum of h'(k) - q -/+ k r, AFF_G chunks of 32 cells per meeting of the lanes.
//     The reference's ties go to the EXTENSION (:262, :285), i.e. to the earliest origin, which the prefix keeps;
//   * the column's effect on the running best (first strict maximum while walking the column, stop at the first record
//     equal to score + q + r, :667-672) is a reduction: the column maximum at its first position, and the first cell
//     that reaches the stop value.
// The band edges of the next column are set by lane 0 exactly as the reference sets them, and exactly the array entries the
// reference writes are written (entries outside the band keep their stale values for the time the band widens again).
// Lanes talk through memory only (AffWarp, the work arrays) with a warp barrier between phases, so the same text runs on
// the host with a loop over the lanes in every phase: that is what the CPU tests compare with aff_epilogue and with
// the reference.  The two places where the lanes' partial results meet are functions with a plain-loop definition (host)
// and the same function of the same inputs by warp shuffles (device).
#ifdef __CUDA_ARCH__
#define AFF_LANES(l) for (int l = (int)(threadIdx.x & 31u), aff_once_ = 1; aff_once_; aff_once_ = 0)
#define AFF_SYNC() __syncwarp()
#else
#define AFF_LANES(l) for (int l = 0; l < 32; ++l)
#define AFF_SYNC() ((void)0)
#endif

constexpr int AFF_G = 4;                          // chunks of 32 cells per meeting
constexpr int AFF_NEG_BIG = -(1 << 29);
constexpr int AFF_NEVER = AFF_MINOR_INF * 2 + 2;  // below anything a state can hold

struct AffWarp {                                  // one per warp, shared memory on the device
    int gA[AFF_G][32], gTag[AFF_G][32];           // in: candidates in column order (chunk, lane), with the tag that rides along
    int gExA[AFF_G][32], gExTag[AFF_G][32];       // out: the best candidate before each position
    int gH[AFF_G][32];                            // reverse pass: the group's scores before F
    int carryH[2];                                // reverse pass: old score of the previous group's last row
    int run0, tag0;                               // the best candidate before this meeting / after it
    int segM[32], segP[32], segT[32];             // per lane: best score of the column so far, its row, first row that reaches T
    int colM, colP, colT;
    int score_r, start_i, start_j, hi, lo, stop, undefined;
    int gl_score, gl_first;
};

// gEx[g][l] = the best of (run0, tag0), gA[0][0..31], gA[1][0..31], ... up to but not including gA[g][l], where a later
// candidate replaces an earlier one only if it is strictly larger; (run0, tag0) = the best after the last position.
__host__ __device__ __forceinline__ void aff_meet_prefix_best(AffWarp* w)
{
#ifdef __CUDA_ARCH__
    const int l = (int)(threadIdx.x & 31u);
    // one key per candidate: value, then "earlier wins" (bits 2-7: 32 for what came before, 31 - lane inside a chunk), then the tag
    long long carry = (long long)w->run0 * 256 + (32 << 2) + w->tag0;
    long long ex[AFF_G];
#pragma unroll
    for (int g = 0; g < AFF_G; ++g) {
        long long x = (long long)w->gA[g][l] * 256 + ((31 - l) << 2) + w->gTag[g][l];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (l >= o && y > x) x = y;
        }
        const long long before = __shfl_up_sync(0xffffffffu, x, 1);
        const long long total = __shfl_sync(0xffffffffu, x, 31);
        ex[g] = (l == 0 || carry > before) ? carry : before;
        if (total > carry) carry = (total & ~0xfcll) | (32 << 2);
    }
#pragma unroll
    for (int g = 0; g < AFF_G; ++g) { w->gExA[g][l] = (int)(ex[g] >> 8); w->gExTag[g][l] = (int)(ex[g] & 3); }
    __syncwarp();
    if (l == 0) { w->run0 = (int)(carry >> 8); w->tag0 = (int)(carry & 3); }
#else
    int run = w->run0, rtag = w->tag0;
    for (int g = 0; g < AFF_G; ++g)
        for (int k = 0; k < 32; ++k) {
            w->gExA[g][k] = run; w->gExTag[g][k] = rtag;
            if (w->gA[g][k] > run) { run = w->gA[g][k]; rtag = w->gTag[g][k]; }
        }
    w->run0 = run; w->tag0 = rtag;
#endif
}

// The same for plain maxima (no tag, ties are immaterial): gExA[g][l] = max(run0, every gA before position (g, l)).
__host__ __device__ __forceinline__ void aff_meet_prefix_max(AffWarp* w)
{
#ifdef __CUDA_ARCH__
    const int l = (int)(threadIdx.x & 31u);
    int carry = w->run0;
    int ex[AFF_G];
#pragma unroll
    for (int g = 0; g < AFF_G; ++g) {
        int x = w->gA[g][l];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int y = __shfl_up_sync(0xffffffffu, x, o);
            if (l >= o && y > x) x = y;
        }
        const int before = __shfl_up_sync(0xffffffffu, x, 1);
        const int total = __shfl_sync(0xffffffffu, x, 31);
        ex[g] = (l == 0 || carry > before) ? carry : before;
        carry = total > carry ? total : carry;
    }
#pragma unroll
    for (int g = 0; g < AFF_G; ++g) w->gExA[g][l] = ex[g];
    __syncwarp();
    if (l == 0) w->run0 = carry;
#else
    int run = w->run0;
    for (int g = 0; g < AFF_G; ++g)
        for (int k = 0; k < 32; ++k) {
            w->gExA[g][k] = run;
            if (w->gA[g][k] > run) run = w->gA[g][k];
        }
    w->run0 = run;
#endif
}

// colM = the largest segM, colP = the largest segP among the lanes that hold it, colT = the largest segT.
// segM >= -1, 0 <= segP, segT < 2^20.
__host__ __device__ __forceinline__ void aff_meet_column_best(AffWarp* w)
{
#ifdef __CUDA_ARCH__
    const int l = (int)(threadIdx.x & 31u);
    long long key = ((long long)(w->segM[l] + 1) << 20) | (long long)w->segP[l];
    int t = w->segT[l];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const long long y = __shfl_xor_sync(0xffffffffu, key, o);
        const int z = __shfl_xor_sync(0xffffffffu, t, o);
        key = y > key ? y : key;
        t = z > t ? z : t;
    }
    if (l == 0) { w->colM = (int)(key >> 20) - 1; w->colP = (int)(key & 0xfffff); w->colT = t; }
#else
    int M = -1, Mp = 0, tpos = 0;
    for (int k = 0; k < 32; ++k) {
        if (w->segM[k] > M || (w->segM[k] == M && w->segP[k] > Mp)) { M = w->segM[k]; Mp = w->segP[k]; }
        if (w->segT[k] > tpos) tpos = w->segT[k];
    }
    w->colM = M; w->colP = Mp; w->colT = tpos;
#endif
}

// Reverse pass (:617-690).  Hh[k] / Ee[k]: the score of row k and the horizontal gap state of row k at the last column
// that touched them (the reference's eh[k] >> 16 and eh[k+1] & 0xffff).
__host__ __device__ inline void aff_reverse_warp(const AffSeq& s1, const AffSeq& s2, const AffParams& P, int score_f, int end_i, int end_j,
                                                 int* __restrict__ Hh, int* __restrict__ Ee, AffWarp* w)
{
    const int q = P.q, r = P.r, qr = q + r;
    AFF_LANES(l) { for (int k = l; k <= end_i + 1; k += 32) { Hh[k] = 0; Ee[k] = 0; } }
    AFF_SYNC();
    AFF_LANES(l) {
        if (l == 0) {
            const int sr = aff_sc(s1.at1(end_i), s2.at1(end_j), P);
            w->score_r = sr; w->start_i = end_i; w->start_j = end_j;
            Hh[end_i] = qr + sr;                                                  // :627
            w->hi = end_i - 1;
            w->lo = end_i - 3 > 0 ? end_i - 3 : 0;
            w->stop = 0; w->undefined = 0;
        }
    }
    AFF_SYNC();
    const int T = score_f + qr;
    for (int j = end_j - 1; j != 0; --j) {
        const int hi = w->hi, lo = w->lo;
        if (hi < lo) {
            AFF_SYNC();
            AFF_LANES(l) { if (l == 0) w->undefined = 1; }
            AFF_SYNC();
            break;
        }
        const int W = hi - lo;                                                    // the column's cells: rows hi - t, t = 0 .. W-1
        const uint32_t c2 = s2.at1(j);
        AFF_LANES(l) {
            w->segM[l] = -1; w->segP[l] = 0; w->segT[l] = 0;
            if (l == 0) w->run0 = AFF_NEG_BIG;
        }
        // AFF_G chunks of 32 cells at a time: scores without F from the previous column's values (all reads of a group come
        // before its writes; the one value a group needs from a row the group before has already overwritten, the old
        // score of that group's last row, is handed over in carryH), F by a prefix maximum, final scores and records.
        for (int t0 = 0, k = 0; t0 < W; t0 += 32 * AFF_G, ++k) {
            AFF_LANES(l) {
                for (int g = 0; g < AFF_G; ++g) {
                    const int t = t0 + 32 * g + l, i = hi - t;
                    int h = 0, a = AFF_NEG_BIG;
                    if (t < W) {
                        const int hl = Hh[i], eo = Ee[i];
                        const int hd = (t == t0 && t0 > 0) ? w->carryH[(k + 1) & 1] : Hh[i + 1];
                        h = hd + aff_sc(s1.at1(i), c2, P);
                        if (h < 0) h = 0;
                        int e = (eo > hl - q) ? eo - r : hl - qr;
                        if (e < 0) e = 0;
                        if (h < e) h = e;
                        Ee[i] = e;                                                // read by this cell only: in place
                        a = h - q - i * r;                                        // as F of a row i' < i: this + i' r
                        if (t == t0 + 32 * AFF_G - 1) w->carryH[k & 1] = hl;
                    }
                    w->gH[g][l] = h;
                    w->gA[g][l] = a;
                }
            }
            AFF_SYNC();
            aff_meet_prefix_max(w);
            AFF_SYNC();
            AFF_LANES(l) {
                int bm = w->segM[l], bp = w->segP[l], tp = w->segT[l];
                for (int g = 0; g < AFF_G; ++g) {
                    const int t = t0 + 32 * g + l, i = hi - t;
                    if (t < W) {
                        const int hp = w->gH[g][l];
                        const int f = w->gExA[g][l] + i * r;                      // max over k > i of h'(k) - q - (k - i) r
                        const int h = hp > f ? hp : f;
                        Hh[i] = h;
                        if (h > bm) { bm = h; bp = i; }
                        if (tp == 0 && h >= T) tp = i;
                    }
                }
                w->segM[l] = bm; w->segP[l] = bp; w->segT[l] = tp;
            }
            AFF_SYNC();
        }
        AFF_LANES(l) { if (l == 0) { Hh[hi + 1] = 0; Ee[lo] = 0; } }              // :665 of the first cell, :674
        aff_meet_column_best(w);
        AFF_SYNC();
        AFF_LANES(l) {
            if (l == 0) {
                const int M = w->colM, Mp = w->colP, tpos = w->colT;
                bool stop = false;
                if (w->score_r < T && tpos != 0 && Hh[tpos] == T) {                // the first record that reaches T is T itself
                    w->score_r = T; w->start_i = tpos; w->start_j = j; stop = true;
                }
                if (stop) w->stop = 1;
                else {
                    if (M > w->score_r) { w->score_r = M; w->start_i = Mp; w->start_j = j; }
                    int nh = hi;
                    if (Hh[nh] <= qr) --nh;                                        // :676-677
                    if (nh <= 0) nh = 0;
                    int nl = w->start_i - (w->start_j - j) - (w->score_r + (w->start_j - j) * P.match) / r - 1;      // :678
                    if (nl <= 0) nl = 0;
                    w->hi = nh; w->lo = nl;
                }
            }
        }
        AFF_SYNC();
        if (w->stop) break;
    }
}

// aln_global_core (:328-508), as aff_global: returns through w->gl_score / w->gl_first.  work: 8*(len1+1) ints.
__host__ __device__ inline void aff_global_warp(const AffSeq& s1, int o1, int len1, const AffSeq& s2, int o2, int len2, const AffParams& P,
                                                int b, int* work, AffWarp* w)
{
    const int q = P.q, r = P.r;
    int b1, b2;
    if (len1 > len2) { b1 = len1 - len2 + b; b2 = b; } else { b1 = b; b2 = len2 - len1 + b; }
    if (b1 > len1) b1 = len1;
    if (b2 > len2) b2 = len2;
    const int wd = len1 + 1;
    AffCol curr{work, work + wd, work + 2 * wd, work + 3 * wd}, last{work + 4 * wd, work + 5 * wd, work + 6 * wd, work + 7 * wd};
    AFF_SYNC();
    AFF_LANES(l) {
        for (int i = l; i < wd; i += 32) {
            curr.M[i] = curr.I[i] = curr.D[i] = AFF_MINOR_INF; curr.tag[i] = 0;
            last.M[i] = last.I[i] = last.D[i] = AFF_MINOR_INF; last.tag[i] = 0;
        }
    }
    AFF_SYNC();
    // first row (:375-381): D(i) = -q - i r, its path leaves the corner downwards
    AFF_LANES(l) {
        if (l == 0) curr.M[0] = 0;
        for (int i = 1 + l; i < b1; i += 32) { curr.D[i] = -q - i * r; curr.tag[i] = AFF_FROM_D << 4; }
    }
    AFF_SYNC();
    { const AffCol t = curr; curr = last; last = t; }

    // One column.  e0: the band's lower edge cell (row 0 while the band touches it, INF otherwise), end: its top cell;
    // the cells in between are rows e0 + 1 + t, t = 0 .. n-1.
    auto column = [&](int j, int e0, int end, bool top_I_inf) {
        const int n = end - e0;
        const uint32_t c2 = s2.at1(o2 + j);
        AFF_LANES(l) {
            if (l == 0) {                                                         // the edge cell
                curr.M[e0] = AFF_MINOR_INF; curr.D[e0] = AFF_MINOR_INF;
                if (e0 == 0) aff_set_I(curr, 0, last, 0, q, r, j == 1); else curr.I[e0] = AFF_MINOR_INF;
                // the D chain starts from it: D(e0) extended, or M(e0) opened (both INF-like here; the reference's order, :280)
                if (curr.M[e0] - q > curr.D[e0]) { w->run0 = curr.M[e0] - q + e0 * r; w->tag0 = curr.tag[e0] & 3; }
                else { w->run0 = curr.D[e0] + e0 * r; w->tag0 = (curr.tag[e0] >> 4) & 3; }
            }
            for (int t = l; t < n; t += 32) {
                const int i = e0 + 1 + t;
                aff_set_M(curr, i, last, i - 1, aff_sc(s1.at1(o1 + i), c2, P), i == 1 && j == 1);
                if (i == end && top_I_inf) curr.I[i] = AFF_MINOR_INF; else aff_set_I(curr, i, last, i, q, r, false);
            }
        }
        AFF_SYNC();
        for (int t0 = 0; t0 < n; t0 += 32 * AFF_G) {                              // D: a prefix over the M states of the rows below
            AFF_LANES(l) {
                for (int g = 0; g < AFF_G; ++g) {
                    const int t = t0 + 32 * g + l, i = e0 + 1 + t;
                    w->gA[g][l] = t < n ? curr.M[i] - q + i * r : AFF_NEVER;      // as D of a row i' > i: this - i' r
                    w->gTag[g][l] = t < n ? curr.tag[i] & 3 : 0;
                }
            }
            AFF_SYNC();
            aff_meet_prefix_best(w);                                              // ties: the earlier origin (:285)
            AFF_SYNC();
            AFF_LANES(l) {
                for (int g = 0; g < AFF_G; ++g) {
                    const int t = t0 + 32 * g + l, i = e0 + 1 + t;
                    if (t < n) {
                        curr.D[i] = w->gExA[g][l] - i * r;
                        curr.tag[i] = (curr.tag[i] & ~48) | (w->gExTag[g][l] << 4);
                    }
                }
            }
            AFF_SYNC();
        }
        { const AffCol t = curr; curr = last; last = t; }
    };

    int j;
    const int tmp_end = b2 < len2 ? b2 : len2 - 1;
    for (j = 1; j <= tmp_end; ++j) column(j, 0, (j + b1 <= len1 + 1) ? (j + b1 - 1) : len1, !(j + b1 - 1 > len1));      // part 1
    if (j == len2 && b2 != len2 - 1) { column(j, 0, (j + b1 <= len1 + 1) ? (j + b1 - 1) : len1, !(j + b1 - 1 > len1)); ++j; }
    for (; j <= len2 - b2 + 1; ++j) column(j, j - b2, j + b1 - 1, true);                                                  // part 2
    for (; j < len2; ++j) column(j, j - b2, len1, false);                                                                 // part 3
    if (j == len2) column(j, j - b2, len1, false);
    AFF_LANES(l) {
        if (l == 0) {
            int mx = last.M[len1], t = last.tag[len1] & 3;
            if (last.I[len1] > mx) { mx = last.I[len1]; t = (last.tag[len1] >> 2) & 3; }
            if (last.D[len1] > mx) { mx = last.D[len1]; t = (last.tag[len1] >> 4) & 3; }
            w->gl_score = mx; w->gl_first = t;
        }
    }
    AFF_SYNC();
}

// Passes 2 and 3 for one pair by one warp (host: by a loop over the lanes); same contract as aff_epilogue.  Lane 0's *res
// is the result.
__host__ __device__ inline void aff_epilogue_warp(const AffSeq& s1, const AffSeq& s2, const AffParams& P, int score_f, int end_i, int end_j,
                                                  int* work, AffWarp* w, DevLocal* res)
{
    const int qr = P.q + P.r;
    const int stride = end_i + 2;
    aff_reverse_warp(s1, s2, P, score_f, end_i, end_j, work, work + stride, w);
    res->end1 = end_i; res->end2 = end_j; res->flags = 0;
    if (w->undefined) { res->flags |= AFF_FLAG_UNDEFINED; res->start1 = 0; res->start2 = 0; res->score = score_f; return; }
    const int score_r = w->score_r - qr, start_i = w->start_i, start_j = w->start_j;
    const int len1 = end_i - start_i + 1, len2 = end_j - start_j + 1;
    int jmax = (end_i - start_i > end_j - start_j) ? end_i - start_i : end_j - start_j;
    ++jmax;
    int score_g = 0, first = AFF_FROM_M;
    for (int b = P.band;; b <<= 1) {
        aff_global_warp(s1, start_i - 1, len1, s2, start_j - 1, len2, P, b, work, w);
        score_g = w->gl_score; first = w->gl_first;
        if (score_g == score_r || score_f == score_g) break;
        if (b > jmax) break;
    }
    if (score_r > score_g && score_f > score_g) { res->score = -1; res->flags |= AFF_FLAG_POTENTIAL_BUG; }
    else res->score = score_g;
    const int pi = (first == AFF_FROM_I ? 0 : 1) + start_i - 1, pj = (first == AFF_FROM_D ? 0 : 1) + start_j - 1;
    res->start1 = pi ? pi : 1;
    res->start2 = pj ? pj : 1;
}

#ifdef __CUDACC__
constexpr int AE_THREADS = 128;                  // 4 warps per CTA, one pair per warp at a time
constexpr int AE_CTAS_PER_SM = 8;                // 64 registers: the kernel waits on memory and on its meetings, more warps hide more

// One warp per pair (pulled from a queue in the forward kernel's order), scratch_stride ints of scratch per warp.
__global__ void __launch_bounds__(AE_THREADS, AE_CTAS_PER_SM)
affine_epilogue_kernel(const uint32_t* __restrict__ packed, const PairDesc* __restrict__ pairs, const uint32_t* __restrict__ order,
                       uint32_t n_work, unsigned int* __restrict__ queue, AffParams P, int* __restrict__ scratch, size_t scratch_stride,
                       DevLocal* __restrict__ out)
{
    __shared__ AffWarp warp_area[AE_THREADS / 32];
    AffWarp* const w = &warp_area[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    int* const work = scratch + (size_t)((blockIdx.x * blockDim.x + threadIdx.x) >> 5) * scratch_stride;
    for (;;) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(queue, 1u);
        qi = __shfl_sync(0xffffffffu, qi, 0);
        if (qi >= n_work) break;
        const uint32_t pid = order[qi];
        DevLocal res = out[pid];
        __syncwarp();
        if (res.score <= 0 || (res.flags & AFF_FLAG_NO_MATCH)) continue;
        const PairDesc pd = pairs[pid];
        const AffSeq s1{packed + pd.row_off}, s2{packed + pd.col_off};
        aff_epilogue_warp(s1, s2, P, res.score, res.end1, res.end2, work, w, &res);
        if (lane == 0) out[pid] = res;
        __syncwarp();
    }
}

inline cudaError_t affine_configure()
{
    return cudaFuncSetAttribute(affine_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AF_SMEM_BYTES);
}
#endif // __CUDACC__

} // namespace gp
