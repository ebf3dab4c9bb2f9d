// flank_place.cuh -- semi-global placement of a flank inside a contig (sm_100a): BASELINE.json configs[1].
//
// What it stands in for: GAPPadder places a gap's two flanks on every contig with `bwa mem -T <s> -a contigs.fa flanks.fa`
// (/root/reference/pick_contigs.py:83-86) and keeps strand, clip type, aligned length and leftmost contig position of the
// best record per contig and side (:99-147).  BWA is neither vendored nor pinned, so there is nothing in the reference to
// be bit-exact against: PARITY UNPINNED.  This kernel is bit-exact against the builder-written definition in
// oracle/overlap_oracle.c (gpo_semiglobal), the semi-global form SURVEY.md 8c prescribes, with Evaluate's linear scoring
// (match +1, mismatch, indel: ContigsCompactor-v0.2.0/ContigsMerger/ContigsCompactor.cpp:1596,1640,1654):
//
//   rows = flank (m bases, consumed end to end), columns = contig (n bases, free ends)
//   H(0,j) = 0, H(i,0) = i*indel, H(i,j) = max(diag + s, up + indel, left + indel)
//   score = max_j H(m,j); col_end = the smallest such j; col_start = the largest start column over all optimal
//   alignments ending there (start(i,j) = max of start over the predecessors that reach H(i,j)).
//
// Arithmetic.  One 32-bit word per cell: V = (Q << 14) | start with Q = H - indel*(i + j) >= 0 (a potential in rows and
// columns under which the up and left moves add nothing and the diagonal adds s - 2*indel: 5 for a match, 2 for a
// mismatch with GAPPadder's scores).  A cell is then ONE add and ONE 3-input signed max (VIMNMX3.S32):
//   V(i,j) = max3(V(i,j-1), V(i-1,j), V(i-1,j-1) + inc << 14)
// and the max over (Q, start) words IS the definition above: the larger score wins, equal scores keep the larger start.
// The traceback is therefore "compact" in the sense of the north star: the start column rides along in the low bits, no
// table is kept.  Ranges: n <= 16383 (14-bit start), Q <= m + |indel|*(m+n) < 2^17.
//
// Wavefront.  One warp per pair; the table is cut into strips of 512 rows, lane l owns 16 consecutive rows and is two
// columns behind lane l-1, so the shuffle that hands a lane's last row (and the column's symbol) to the next lane is
// issued a whole step before its result is used.  Rows are padded at the TOP (virtual rows whose diagonal increment
// is -infinity simply copy row 0 downwards), so that row m is always the last row of lane 31 in the last strip and the
// best-cell bookkeeping is three instructions on one value per step.  The strip's bottom row goes through a per-warp
// boundary line in global scratch (L2), read 32 columns at a time one chunk ahead.  The diagonal increments come from a
// per-warp shared-memory table (4 column symbols x 16 rows per lane, conflict-free LDS.128) when both sequences are pure
// A/C/G/T; sequences with N or other letters take a compare-and-select per cell (same results, slower).
#pragma once
#include "common.cuh"

namespace gp {

constexpr int FP_R = 16;                         // rows per lane
constexpr int FP_STRIP = 32 * FP_R;              // 512 rows per strip
constexpr int FP_THREADS = 128;                  // 4 warps per CTA
constexpr int FP_CTAS_PER_SM = 3;
constexpr int FP_SKEW = 2;                       // columns between neighbouring lanes
constexpr int FP_START_BITS = 14;
constexpr uint32_t FP_MAX_N = (1u << FP_START_BITS) - 1u;      // 16383
constexpr uint32_t FP_MAX_Q = (1u << 17) - 1u;
constexpr int FP_TAB_WORDS = 4 * FP_R * 32;      // 8 KB per warp
constexpr int FP_RING = 64;                      // lane 0's inputs: {V(above), symbol} of the next columns
constexpr int FP_WARP_WORDS = FP_TAB_WORDS + 2 * FP_RING;
constexpr size_t FP_SMEM_BYTES = (size_t)(FP_THREADS / 32) * FP_WARP_WORDS * sizeof(uint32_t);
constexpr int FP_NEG = -(1 << 29);               // increment of a virtual row: the diagonal never wins

struct DevPlace {                                // same layout as gp_place_result
    int32_t score;
    int32_t col_start;
    int32_t col_end;
    uint32_t flags;
};

inline bool fp_pair_ok(uint32_t m, uint32_t n, int indel)
{
    const uint64_t slope = (uint64_t)(indel < 0 ? -indel : 0);
    return indel <= 0 && n <= FP_MAX_N && (uint64_t)m + slope * ((uint64_t)m + n) <= FP_MAX_Q;
}

template <bool TABLE>
__global__ void __launch_bounds__(FP_THREADS, FP_CTAS_PER_SM)
flank_place_kernel(const uint32_t* __restrict__ packed, const PairDesc* __restrict__ pairs, const uint32_t* __restrict__ order,
                   uint32_t n_work, unsigned int* __restrict__ queue, int mismatch, int indel,
                   uint32_t* __restrict__ scratch, uint32_t scratch_stride, DevPlace* __restrict__ out)
{
    extern __shared__ uint32_t fp_smem[];
    constexpr uint32_t FULL = 0xffffffffu;
    constexpr int R = FP_R, D = FP_SKEW;
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    uint32_t* const bnd = scratch + (size_t)warp_global * scratch_stride;
    uint32_t* const tab = fp_smem + (threadIdx.x >> 5) * FP_WARP_WORDS;
    uint32_t* const ring = tab + FP_TAB_WORDS;                      // [slot] = V above, [FP_RING + slot] = symbol
    const uint32_t tab_addr = (uint32_t)__cvta_generic_to_shared(tab) + (uint32_t)lane * 16u;
    const int slope = -indel;
    const int inc_match = (1 + 2 * slope) << FP_START_BITS, inc_mism = (mismatch + 2 * slope) << FP_START_BITS;

    for (;;) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(queue, 1u);
        qi = __shfl_sync(FULL, qi, 0);
        if (qi >= n_work) break;
        const uint32_t pid = order[qi];
        const PairDesc pd = pairs[pid];
        const int m = (int)pd.m, n = (int)pd.n;
        const int n_strips = (m + FP_STRIP - 1) / FP_STRIP;
        const int pad = n_strips * FP_STRIP - m;                     // virtual rows above row 1

        // boundary line = row 0: H = 0 -> Q = slope*j, start = j
        for (int j = lane; j <= n; j += 32) bnd[j] = ((uint32_t)(slope * j) << FP_START_BITS) | (uint32_t)j;
        __syncwarp();

        int best_hq = 0, best_j = 0, best_start = 0;                 // column 0: H(m,0) = m*indel, i.e. Q - slope*j = 0

        for (int s = 0; s < n_strips; ++s) {
            const int itop = s * FP_STRIP + lane * R - pad;          // my rows are itop+1 .. itop+R (1-based); <= 0: virtual
            const bool last = s == n_strips - 1;
            uint32_t rc[TABLE ? 1 : R];
            if constexpr (TABLE) {
                uint32_t code[R];
#pragma unroll
                for (int r = 0; r < R; ++r) code[r] = itop + r >= 0 ? load_code(packed, pd.row_off, (uint32_t)(itop + r)) : 0xffu;
                __syncwarp();
#pragma unroll 1
                for (uint32_t c = 0; c < 4; ++c) {
#pragma unroll
                    for (int q = 0; q < R / 4; ++q) {
                        uint32_t w[4];
#pragma unroll
                        for (int x = 0; x < 4; ++x) {
                            const uint32_t rcx = code[4 * q + x];
                            w[x] = (uint32_t)(rcx == 0xffu ? FP_NEG : (rcx == c ? inc_match : inc_mism));
                        }
                        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(tab_addr + c * (R * 128u) + q * 512u),
                                     "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) rc[r] = itop + r >= 0 ? load_code(packed, pd.row_off, (uint32_t)(itop + r)) : 0x100u;
            }
            int32_t V[R];
#pragma unroll
            for (int r = 0; r < R; ++r) V[r] = 0;                    // column 0: Q = 0, start = 0 (real and virtual rows alike)
            int32_t up_prev = 0;                                     // V(itop, j-1): the diagonal of my first row
            int32_t bottom = 0;                                      // my last row at the column of my previous step
            uint32_t mysym = 0;                                      // that column's symbol
            int32_t recv_next = 0;
            uint32_t sym_next = 0;
            // lane 0's inputs, one chunk of 32 columns ahead
            uint32_t nextV, nextC;
            {
                const int j = 1 + lane;
                nextV = j <= n ? __ldcg(bnd + j) : 0u;
                nextC = j <= n ? load_code(packed, pd.col_off, (uint32_t)(j - 1)) : 0u;
            }
            const int steps = n + D * 31;
#pragma unroll 1
            for (int t = 1; t <= steps; ++t) {
                if (((t - 1) & 31) == 0) {
                    __syncwarp();
                    const int slot = (t - 1 + lane) & (FP_RING - 1);
                    ring[slot] = nextV;
                    ring[FP_RING + slot] = nextC;
                    const int j = t + 32 + lane;
                    nextV = j <= n ? __ldcg(bnd + j) : 0u;
                    nextC = j <= n ? load_code(packed, pd.col_off, (uint32_t)(j - 1)) : 0u;
                    __syncwarp();
                }
                int32_t recv = recv_next;
                uint32_t csym = sym_next;
                recv_next = __shfl_up_sync(FULL, bottom, 1);           // used one step from now
                sym_next = __shfl_up_sync(FULL, mysym, 1);
                if (lane == 0) {
                    recv = (int32_t)ring[(t - 1) & (FP_RING - 1)];
                    csym = ring[FP_RING + ((t - 1) & (FP_RING - 1))];
                }
                const int j = t - D * lane;
                if ((uint32_t)(j - 1) < (uint32_t)n) {
                    int32_t d[R];
                    if constexpr (TABLE) {
                        uint32_t inc[R];
                        const uint32_t a = tab_addr + (csym & 3u) * (R * 128u);
#pragma unroll
                        for (int q = 0; q < R / 4; ++q)
                            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(inc[4 * q]), "=r"(inc[4 * q + 1]), "=r"(inc[4 * q + 2]), "=r"(inc[4 * q + 3])
                                         : "r"(a + q * 512u));
                        d[0] = up_prev + (int32_t)inc[0];
#pragma unroll
                        for (int r = 1; r < R; ++r) d[r] = V[r - 1] + (int32_t)inc[r];
                    } else {
                        d[0] = up_prev + (rc[0] == csym ? inc_match : rc[0] == 0x100u ? FP_NEG : inc_mism);
#pragma unroll
                        for (int r = 1; r < R; ++r) d[r] = V[r - 1] + (rc[r] == csym ? inc_match : rc[r] == 0x100u ? FP_NEG : inc_mism);
                    }
                    int32_t up = recv;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        const int32_t w = __vimax3_s32(V[r], up, d[r]);   // left and up moves are free under the potential
                        up = w;
                        V[r] = w;
                    }
                    up_prev = recv;
                    mysym = csym;
                    bottom = up;
                    if (lane == 31) {
                        if (!last) bnd[j] = (uint32_t)bottom;               // the next strip's row above
                        else {
                            const int hq = (bottom >> FP_START_BITS) - slope * j;      // H(m,j) + slope*m
                            if (hq > best_hq) { best_hq = hq; best_j = j; best_start = bottom & (int)FP_MAX_N; }
                        }
                    }
                }
            }
            __syncwarp();
        }
        if (lane == 31) {
            DevPlace r;
            r.score = best_hq - slope * m;
            r.col_start = best_start;
            r.col_end = best_j;
            r.flags = TABLE ? 1u : 0u;
            out[pid] = r;
        }
        __syncwarp();
    }
}

inline cudaError_t fp_configure()
{
    cudaError_t e = cudaFuncSetAttribute(flank_place_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FP_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(flank_place_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FP_SMEM_BYTES);
}

} // namespace gp
