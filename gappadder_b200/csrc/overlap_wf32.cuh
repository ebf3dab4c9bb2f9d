// overlap_wf32.cuh -- general overlap-DP kernel, one warp per pair, 32-bit lanes (sm_100a).
//
// Computes, for each pair, exactly what ContigsCompactor::Evaluate computes before the
// significance test (ContigsCompactor-v0.2.0/ContigsMerger/ContigsCompactor.cpp:1596-1709 and
// :1736-1837): the overlap DP, the best-cell scan over the last max_clip+1 rows/columns and
// whether the predecessor walk from the best cell ends in row 0 / column 0.
//
// This is the any-length, any-score, any-alphabet (<=16 symbols) kernel; the packed 16-bit kernel
// in overlap_wf16.cuh is the fast path for pairs that fit its ranges.
//
// Layout: the (m x n) table is cut into horizontal strips of 32*R rows.  Lane l owns R
// consecutive rows of the strip and walks the columns left to right, one column per step, one
// step behind lane l-1 (anti-diagonal wavefront across the warp).  Per step a lane needs the cell
// above its first row (the last row of lane l-1, one SHFL) and the column's base (passed down the
// same way); lane 0 takes both from the previous strip's last row, kept in a per-warp global
// scratch line (L2 resident) that is read 32 columns at a time with one coalesced load.
//
// Cell encoding: V = (H << 4) | (prio << 2) | origin with
//   origin bit0: the predecessor walk from the cell ends in row 0, bit1: it ends in column 0;
//   prio: 2 for the diagonal candidate, 1 for "up", 0 for "left", cleared after the max.
// max3 over the three candidates then IS the reference's tie rule (replace only on strict '<',
// :1654,:1660  =>  diagonal beats up beats left on equal scores) and carries the origin of the
// chosen predecessor along, which is all the reference's traceback is used for (:1834-1837).
#pragma once
#include "common.cuh"

namespace gp {

template <int R>
__global__ void __launch_bounds__(128)
overlap_wf32_kernel(const uint32_t* __restrict__ packed, const PairDesc* __restrict__ pairs,
                    const uint32_t* __restrict__ order, uint32_t n_work, unsigned int* __restrict__ queue,
                    int mismatch, int indel, int max_clip,
                    int32_t* __restrict__ scratch, uint32_t scratch_stride, DevResult* __restrict__ out,
                    const unsigned int* __restrict__ n_work_dev)
{
    if (n_work_dev) n_work = *n_work_dev;             // work list filled on the device (overlap_wf16c.cuh's retries)
    const int lane = threadIdx.x & 31;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int32_t* bnd = scratch + (size_t)warp_global * scratch_stride;
    const int sub_match = (1 << 4) + 8;            // diagonal candidate carries prio 2
    const int sub_mism = mismatch * 16 + 8;
    const int g_up = indel * 16 + 4;               // up candidate: prio 1
    const int g_left = indel * 16;                 // left candidate: prio 0

    for (;;) {
        uint32_t qi = 0;
        if (lane == 0) qi = atomicAdd(queue, 1u);
        qi = __shfl_sync(0xffffffffu, qi, 0);
        if (qi >= n_work) break;
        const uint32_t pid = order[qi];
        const PairDesc pd = pairs[pid];
        const int m = (int)pd.m, n = (int)pd.n;
        const int C = max_clip;

        // row 0 of the table: H = 0, walk ends in row 0 (and in column 0 for the corner)
        for (int j = lane; j <= n; j += 32) bnd[j] = (j == 0) ? 3 : 1;
        __syncwarp();

        // (0,n) is the first cell the reference scans (c = 0, column loop, i = 0): rank 0, H = 0.
        long long best = make_key(0, 0u, 1u | (n == 0 ? 2u : 0u));

        for (int i0 = 0; i0 < m; i0 += 32 * R) {
            const int itop = i0 + lane * R;        // rows itop+1 .. itop+R (1-based)
            uint32_t rc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                int i = itop + r;                  // 0-based base index of row i+1
                rc[r] = (i < m) ? load_code(packed, pd.row_off, (uint32_t)i) : 0xffu;
            }
            int32_t V[R];
#pragma unroll
            for (int r = 0; r < R; ++r) V[r] = 2;  // column 0: H = 0, walk ends in column 0
            // V(itop, 0): the cell diagonal-above my first row at column 1
            int32_t up_prev = (itop == 0) ? 3 : 2;
            int32_t bottom = 2;                    // my last row at the column of my previous step
            uint32_t mycode = 0;
            int32_t chunkV = 0; uint32_t chunkC = 0;
            const bool row_tail = (itop + R >= m - C) && (itop + 1 <= m);

            const int steps = n + 31;
            for (int t = 1; t <= steps; ++t) {
                if (((t - 1) & 31) == 0) {         // lane 0's inputs for columns t .. t+31
                    int j = t + lane;
                    if (j <= n) {
                        chunkV = bnd[j];
                        chunkC = load_code(packed, pd.col_off, (uint32_t)(j - 1));
                    }
                }
                int32_t v0 = __shfl_sync(0xffffffffu, chunkV, (t - 1) & 31);
                uint32_t c0 = __shfl_sync(0xffffffffu, chunkC, (t - 1) & 31);
                int32_t upin = __shfl_up_sync(0xffffffffu, bottom, 1);
                uint32_t ccode = __shfl_up_sync(0xffffffffu, mycode, 1);
                if (lane == 0) { upin = v0; ccode = c0; }
                const int j = t - lane;            // my column at this step
                if (j >= 1 && j <= n) {
                    int32_t up = upin, dg = up_prev;
#pragma unroll
                    for (int r = 0; r < R; ++r) {
                        int32_t left = V[r];
                        int32_t d = dg + ((rc[r] == ccode) ? sub_match : sub_mism);
                        int32_t u = up + g_up;
                        int32_t l = left + g_left;
                        int32_t h = __vimax3_s32(d, u, l) & ~12;
                        dg = left;
                        up = h;
                        V[r] = h;
                    }
                    up_prev = upin;
                    mycode = ccode;
                    bottom = V[R - 1];
                    if (lane == 31) bnd[j] = bottom;   // last row of the strip, for the next strip
                    if (row_tail || j >= n - C) {
#pragma unroll
                        for (int r = 0; r < R; ++r) {
                            int i = itop + r + 1;
                            if (i <= m) {
                                uint32_t rk = cell_rank(i, j, m, n, C);
                                if (rk <= RANK_MAX) {
                                    long long k = make_key(V[r] >> 4, rk, (uint32_t)V[r] & 3u);
                                    best = k > best ? k : best;
                                }
                            }
                        }
                    }
                }
            }
            __syncwarp();
        }
        best = warp_max_key(best);
        if (lane == 0) store_result(out + pid, best, m, n, 0u);
        __syncwarp();
    }
}

} // namespace gp
