// relax_chain.cuh -- the relax chains of ContigsMerger on the device, as ONE dependency-driven launch (sm_100a).
//
// Reference: ContigsCompactor::FormMergedSeqFromPath, ContigsCompactor-v0.2.0/ContigsMerger/ContigsCompactor.cpp:1456-1515:
//   merged = node_0;  for k = 1..: Evaluate(merged, node_k, relax) (:1491);  merged = ccAct.GetMerged() (:1512)
// i.e. step k needs the merged contig step k-1 produced: sequential inside a path, independent across paths and gaps.
// Paths of one gap that share a prefix share those steps, so the steps of a batch form a forest; a work item is one
// node of it: Evaluate(merged contig of the parent item, node) followed by SetMergedStringConcat (:108-153).
//
// Round 1 ran the forest level by level from the host: 32 blocking calls (pack, copy, launch, copy back, build strings
// on the host), each as long as its longest pair.  Here the merged contigs never leave the GPU and nothing synchronises
// levels:
//   * a CTA of four warps FOLLOWS a chain: it runs the certificate kernel's CTA-per-pair machinery on (row = merged
//     contig so far, column = next node) -- the pair's 512-row strips pipelined across the four warps --, builds the new
//     merged contig as 4-bit codes in the item's arena slot (a funnel-shift nibble copy by all 128 threads), and goes
//     straight on to the item's first child with the contig still in L2;
//   * where the forest branches, the other children are pushed to a ready ring (release store after the contig and
//     its length are in memory); a CTA without a chain takes the next ring ticket and waits for that slot to be
//     filled.  The ring starts out holding the roots, longest chains first.  Every item is consumed exactly once, a
//     global counter of finished items ends the launch, and no CTA ever waits for work that is not already running
//     or queued: no deadlock whatever the number of resident CTAs.
// The host gets one gp_result per item and rebuilds the strings from the original letters (gp_merged_concat).
// An item whose walk end no certificate system proves (a genuine tie between walks ending on different borders; none
// seen on any synthetic set) is marked unresolved, its subtree is skipped; the caller runs those chains through
// gp_overlap_batch (exact kernels) instead.
#pragma once
#include "overlap_wf16c.cuh"

namespace gp {

struct RelaxItem {                 // 32 bytes
    int32_t parent;                // item that produced the row sequence, or -1: table sequence (row_off, row_len)
    uint32_t row_off, row_len;     // parent < 0 only: the path's first node in the packed table (word offset, bases)
    uint32_t col_off, col_len;     // the node met at this step, in the packed table
    uint32_t arena_off;            // this item's merged contig in the arena (word offset, 128-byte aligned)
    int32_t first_child;           // the child this CTA goes on with (-1: the chain ends here)
    int32_t next_sibling;          // the parent's next child (-1: none): pushed to the ready ring by whoever finishes the parent
};

constexpr uint32_t RELAX_EMPTY = 0xffffffffu, RELAX_EXIT = 0xfffffffeu;
constexpr uint32_t FLAG_UNRESOLVED = 32u;       // gp_result.flags: GP_FLAG_UNRESOLVED
constexpr uint32_t RELAX_RING_SLACK = 2048;     // ring slots beyond the item count: one pending ticket per CTA

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p)
{
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t* p, uint32_t v)
{
    asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}

// Eight 4-bit codes starting at base q of the sequence at word pointer s (through L2)
__device__ __forceinline__ uint32_t relax_fetch8(const uint32_t* s, uint32_t q)
{
    const uint32_t w = q >> 3, sh = (q & 7u) * 4u;
    const uint32_t lo = __ldcg(s + w), hi = sh ? __ldcg(s + w + 1) : 0u;
    return __funnelshift_r(lo, hi, sh);
}

// dst := A[a0, a0 + la) ++ B[b0, b0 + lb) as packed 4-bit codes, unused codes of the last word zero.  Whole CTA.
__device__ __forceinline__ void relax_concat(uint32_t* __restrict__ dst, const uint32_t* A, uint32_t a0, uint32_t la,
                                             const uint32_t* B, uint32_t b0, uint32_t lb)
{
    const uint32_t total = la + lb, nw = (total + 7u) >> 3;
    for (uint32_t w = threadIdx.x; w < nw; w += blockDim.x) {
        const uint32_t p0 = 8u * w;
        uint32_t v = 0u;
        if (p0 < la) {
            v = relax_fetch8(A, a0 + p0);
            if (p0 + 8u > la) v &= (1u << (4u * (la - p0))) - 1u;
        }
        if (p0 + 8u > la && lb) {
            const uint32_t ps = p0 > la ? p0 : la, sh = ps - p0;                 // first base of this word that comes from B
            v |= relax_fetch8(B, b0 + (ps - la)) << (4u * sh);
        }
        if (p0 + 8u > total) v &= (1u << (4u * (total - p0))) - 1u;
        dst[w] = v;
    }
}

// ctrl: [0] ring head (next ticket), [1] ring tail (next free slot), [2] finished items, [8] sub-table passes, [9] unresolved items
template <bool POT2>
__global__ void __launch_bounds__(WF16C_THREADS, WF16C_CTAS_PER_SM)
relax_chain_kernel(const uint32_t* __restrict__ packed, uint32_t* __restrict__ arena, const RelaxItem* __restrict__ items,
                   const uint32_t* __restrict__ subtree, uint32_t n_items, uint32_t ring_total, uint32_t* __restrict__ ring, unsigned int* __restrict__ ctrl,
                   Wf16cParams P, uint32_t* __restrict__ scratch, uint32_t scratch_stride, uint32_t* __restrict__ mlen,
                   DevResult* __restrict__ out, unsigned long long* __restrict__ trace)
{
    constexpr int TEAM = WF16C_THREADS / 32;
    extern __shared__ uint32_t wf16c_smem[];
    __shared__ uint32_t team_prog[TEAM];
    __shared__ long long team_keys[TEAM];
    __shared__ uint32_t next_item;
    const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    Wf16cWarp w;
    w.packed = packed;
    w.team_warp = (int)(threadIdx.x >> 5);
    w.strip_idx = 0;
    w.prog_addr = (uint32_t)__cvta_generic_to_shared(team_prog);
    w.bnd = scratch + (size_t)(warp_global - (uint32_t)w.team_warp) * scratch_stride;
    w.smem = wf16c_smem + (threadIdx.x >> 5) * WF16C_WARP_WORDS;
    const bool leader = threadIdx.x == 0;
    int32_t cur = -1;                                               // the item this CTA goes on with (CTA-uniform)
    for (;;) {
        if (cur < 0) {                                              // no chain to follow: the next ring ticket
            if (leader) {
                const uint32_t t = atomicAdd(ctrl, 1u);
                uint32_t v = RELAX_EXIT;
                if (t < ring_total) {                               // a ticket no push will ever fill: this CTA is done
                    while ((v = ld_acquire_u32(ring + t)) == RELAX_EMPTY) {
                        if (ld_acquire_u32(ctrl + 2) >= n_items) { v = RELAX_EXIT; break; }      // everything is finished
                        __nanosleep(200);
                    }
                }
                next_item = v;
            }
            __syncthreads();
            const uint32_t v = next_item;
            __syncthreads();
            if (v == RELAX_EXIT) break;
            cur = (int32_t)v;
        }
        const uint32_t id = (uint32_t)cur;
        const RelaxItem it = items[id];
        unsigned long long t_begin = 0;
        if (trace && leader) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_begin));   // diagnostic (GP_RELAX_TRACE): per-item timeline
        const uint32_t* row_base = packed;
        uint32_t row_off = it.row_off, m = it.row_len;
        if (it.parent >= 0) {                                       // the parent is finished: we followed it, or its push released us
            row_base = arena;
            row_off = items[it.parent].arena_off;
            m = __ldcg(mlen + it.parent);
        }
        w.packed_row = row_base;
        w.pd = PairDesc{row_off, m, it.col_off, it.col_len};
        const int n = (int)it.col_len;
        DevResult r;
        long long key;
        const uint32_t origin = wf16c_solve_pair<true, TEAM, POT2>(w, P, team_keys, 0u, leader, ctrl + 8, r, key);
        if (origin == 0u) {                                         // handed back to the host with everything below it
            if (leader) { r.flags |= FLAG_UNRESOLVED; out[id] = r; atomicAdd(ctrl + 9, 1u); atomicAdd(ctrl + 2, subtree[id]); }
            cur = -1;
            continue;
        }
        store_result(&r, key | (long long)origin, (int)m, n, FLAG_KERNEL16);
        // SetMergedStringConcat (:108-153) on 4-bit codes
        const bool contained = (r.flags & FLAG_CONTAINED) != 0u;
        const uint32_t len1 = m, len2 = (uint32_t)n, rowe = (uint32_t)r.row_end, cole = (uint32_t)r.col_end, clip = (uint32_t)r.nclip;
        uint32_t* dst = arena + it.arena_off;
        uint32_t total;
        const uint32_t* s1 = row_base + row_off;                      // word pointers of the two sequences
        const uint32_t* s2 = packed + it.col_off;
        if (it.first_child < 0) total = 0;                            // a leaf: nobody reads its contig (the host rebuilds the letters)
        else if (contained && rowe + clip == len1 && len1 < len2) { relax_concat(dst, s2, 0u, len2, s2, 0u, 0u); total = len2; }       // merged = s2 (:116-121)
        else if (contained && cole + clip == len2 && len2 < len1) { relax_concat(dst, s1, 0u, len1, s1, 0u, 0u); total = len1; }       // merged = s1 (:122-127)
        else if (rowe + clip == len1) { relax_concat(dst, s1, 0u, len1 - clip, s2, cole, len2 - cole); total = len1 - clip + len2 - cole; }   // :131-139
        else { relax_concat(dst, s2, 0u, len2 - clip, s1, rowe, len1 - rowe); total = len2 - clip + len1 - rowe; }                     // :141-149
        __threadfence();
        __syncthreads();
        if (leader) {
            out[id] = r;
            if (it.first_child >= 0) {
                mlen[id] = total;
                __threadfence();
                for (int32_t c = items[it.first_child].next_sibling; c >= 0; c = items[c].next_sibling)    // the other children: ready now
                    st_release_u32(ring + atomicAdd(ctrl + 1, 1u), (uint32_t)c);
            }
            atomicAdd(ctrl + 2, 1u);
            if (trace) {
                unsigned long long t_end;
                asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_end));
                trace[3 * (size_t)id] = t_begin; trace[3 * (size_t)id + 1] = t_end; trace[3 * (size_t)id + 2] = ((unsigned long long)m << 32) | (unsigned long long)(uint32_t)n;
            }
        }
        cur = it.first_child;
        __syncthreads();                                             // mlen[id] is written before anyone of this CTA reads it
    }
}

inline cudaError_t relax_configure()
{
    cudaError_t e = cudaFuncSetAttribute(relax_chain_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wf16c_smem_bytes<WF16C_THREADS / 32>());
    if (e != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(relax_chain_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared)) != cudaSuccess) return e;
    e = cudaFuncSetAttribute(relax_chain_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wf16c_smem_bytes<WF16C_THREADS / 32>());
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(relax_chain_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

} // namespace gp
