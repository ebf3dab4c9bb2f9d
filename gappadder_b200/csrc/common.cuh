// common.cuh -- device-side types shared by the overlap kernels (sm_100a).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gp {

// One Evaluate(s1 = rows, s2 = columns) request on the device: word offsets into the packed
// sequence table (4-bit codes, eight per word) and lengths in bases.
struct PairDesc {
    uint32_t row_off;
    uint32_t m;
    uint32_t col_off;
    uint32_t n;
};

// Device-side result, same layout as gp_result (include/gappadder_b200.h).
struct DevResult {
    int32_t score;
    int32_t row_end;
    int32_t col_end;
    int32_t nclip;
    uint32_t flags;
};

constexpr uint32_t FLAG_ROW0 = 1u, FLAG_COL0 = 2u, FLAG_CONTAINED = 4u, FLAG_KERNEL16 = 8u;

// 4-bit code of base `pos` (0-based) of the sequence starting at word `off`.
__device__ __forceinline__ uint32_t load_code(const uint32_t* __restrict__ packed, uint32_t off, uint32_t pos)
{
    return (__ldg(packed + off + (pos >> 3)) >> ((pos & 7u) * 4u)) & 15u;
}

// The same through L2 (ld.global.cg): for sequences written by other CTAs of the same launch (relax chain arena).
__device__ __forceinline__ uint32_t load_code_cg(const uint32_t* packed, uint32_t off, uint32_t pos)
{
    return (__ldcg(packed + off + (pos >> 3)) >> ((pos & 7u) * 4u)) & 15u;
}

// Best-cell bookkeeping.  The reference scans, for c = 0..C, column n-c top to bottom and then row
// m-c left to right, and keeps the FIRST strict maximum (ContigsCompactor.cpp:1679-1709).  That is
// an order-independent reduction with key (score descending, scan rank ascending) where
//   rank(column loop, c, i) = c*W + i          rank(row loop, c, j) = c*W + (m+1) + j,   W = m+n+2
// and a cell reachable by both loops counts with its smaller rank.  key64 packs
//   [score:32 | (RANK_MAX - rank):30 | origin:2]  so that a plain signed max implements it.
constexpr uint32_t RANK_MAX = 0x3fffffffu;

__host__ __device__ __forceinline__ long long make_key(int score, uint32_t rank, uint32_t origin)
{
    return ((long long)score << 32) | (long long)(((RANK_MAX - rank) << 2) | origin);
}

// Candidate rank of interior cell (i,j), 1<=i<=m, 1<=j<=n; returns RANK_MAX+1 if the cell is in
// neither scan.
__host__ __device__ __forceinline__ uint32_t cell_rank(int i, int j, int m, int n, int C)
{
    uint32_t W = (uint32_t)(m + n + 2);
    uint32_t r = RANK_MAX + 1u;
    int cc = n - j;                       // found by the column loop of c = n-j
    if (cc <= C) r = (uint32_t)cc * W + (uint32_t)i;
    int cr = m - i;                       // found by the row loop of c = m-i
    if (cr <= C) {
        uint32_t r2 = (uint32_t)cr * W + (uint32_t)(m + 1 + j);
        r = r2 < r ? r2 : r;
    }
    return r;
}

__device__ __forceinline__ long long warp_max_key(long long k)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        long long other = __shfl_xor_sync(0xffffffffu, k, o);
        k = other > k ? other : k;
    }
    return k;
}

// Decodes the winning key into the reference's outputs.
__host__ __device__ __forceinline__ void store_result(DevResult* out, long long key, int m, int n, uint32_t extra_flags)
{
    int score = (int)(key >> 32);
    uint32_t lo = (uint32_t)(key & 0xffffffffll);
    uint32_t origin = lo & 3u;
    uint32_t rank = RANK_MAX - (lo >> 2);
    uint32_t W = (uint32_t)(m + n + 2);
    int c = (int)(rank / W);
    uint32_t rem = rank - (uint32_t)c * W;
    int row, col;
    if (rem <= (uint32_t)m) { row = (int)rem; col = n - c; }
    else { row = m - c; col = (int)(rem - (uint32_t)(m + 1)); }
    uint32_t flags = origin | extra_flags;
    // bcontained, ContigsCompactor.cpp:1834-1837
    if ((row + c == m && (origin & FLAG_ROW0)) || (col + c == n && (origin & FLAG_COL0))) flags |= FLAG_CONTAINED;
    out->score = score; out->row_end = row; out->col_end = col; out->nclip = c; out->flags = flags;
}

} // namespace gp
