// quick_check.cuh -- the candidate filter of the pairwise phase on the device (sm_100a).
//
// Reference: MultiThreadQuickChecker::threadQuickCheck and QuickCheckerContigsMatch
// (ContigsCompactor-v0.2.0/ContigsMerger/ContigsCompactor.cpp:992-1100, :1982-2095) with the 2-bit k-mer code of
// KmerUtils.cpp:22-115: pair (i, j), i <= j (j == i included), of a gap's nodes is aligned iff some k-mer of the
// first or the last 30 bases of node j occurs anywhere in node i.  The host form is gp_candidate_pairs
// (gp_host.cpp); this one works on the packed table that is already in HBM for the DP kernels, so that a batch
// of gaps needs no host pass over the bases at all.
//
// Work item = (gap, a range of its nodes to scan); the host cuts every gap into items of roughly equal numbers of
// bases so that a batch of any shape -- 50 000 small gaps, or 20 gaps of 400 long contigs -- fills the chip.
// Persistent CTAs (one per SM: the set takes most of its shared memory) pull items from an atomic queue.  Per item:
//   phase 1  the probe k-mers of the WHOLE gap (the k-mers of the first and last 30 bases of every node) go into a
//            4^k-bit set in shared memory (128 KB for k = 10, GAPPadder's value) and into hash chains (heads in shared
//            memory, (k-mer, owner) records in a per-CTA slab of global scratch that stays in L1/L2);
//   phase 2  the item's nodes stream past the set: a thread takes 32 consecutive bases (one 16-byte load of packed
//            codes plus the word before it), rolls the k-mer along them -- a shift, an or and one shared-memory bit
//            test per base -- and on a hit walks the chain and stores hit(i, owner) = 1 (a plain byte store:
//            idempotent, no atomics; the matrix is zeroed before the launch).
// The packed codes arrive at 0.5 B per base and are read exactly once (plus 4 B of carry-in per 32 bases): an
// HBM-bound byte-stream scan with no arithmetic to speak of.  bench.py reports its achieved GB/s.
//
// Limits: k <= 10 (the set must fit shared memory), QC_MAX_NODES nodes per gap (the hit matrix is n x n bytes);
// gp_quick_check_device returns GP_ERR_RANGE beyond them and the host filter has no such limits.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cuda_runtime.h>

namespace gp {

constexpr int QC_THREADS = 512;
constexpr int QC_MAX_K = 10;                      // 4^10 bits = 128 KB of shared memory
constexpr int QC_MAX_NODES = 4096;                // nodes per gap (contigs and their reverse complements)
constexpr int QC_WINDOW = 30;                     // lenContigLen, ContigsCompactor.cpp:2024
constexpr int QC_HEADS = 8192;                    // hash heads (32-bit indices into the probe slab)
constexpr int QC_CHUNK = 32;                      // bases per thread step: four packed words, one 16-byte load

struct QcItem {                                   // one unit of work: scan nodes [node_lo, node_hi) of gap `gap`
    uint32_t gap, node_lo, node_hi, pad;
};

inline uint32_t qc_probes_per_node(int k) { return 2u * (uint32_t)(QC_WINDOW - k + 1); }
inline size_t qc_smem_bytes(int k)
{
    return std::max<size_t>(16, ((size_t)1 << (2 * k)) / 8) + (size_t)QC_HEADS * 4 + 64;     // 160 KB for k = 10
}

// 2-bit k-mer letters of eight packed 4-bit codes: A C G T -> 0..3, everything else 0 (KmerUtils.cpp:22-58)
__device__ __forceinline__ uint32_t qc_letters(uint32_t w)
{
    const uint32_t other = (w | (w >> 1)) & 0x44444444u;          // bit 2 or bit 3 of a nibble set: not A/C/G/T
    const uint32_t keep = ~((other >> 2) * 0xfu);                  // nibble mask 0x0 for others, 0xf for A/C/G/T
    return w & keep & 0x33333333u;
}

// 2-bit letter of base `pos` of the sequence at word offset `off`
__device__ __forceinline__ uint32_t qc_letter(const uint32_t* __restrict__ packed, uint32_t off, uint32_t pos)
{
    const uint32_t c = (__ldg(packed + off + (pos >> 3)) >> ((pos & 7u) * 4u)) & 15u;
    return c <= 3u ? c : 0u;
}

// k-mer that ENDS at base `end` (end >= k-1), first base in the highest bits like the reference's shift register
__device__ __forceinline__ uint32_t qc_kmer(const uint32_t* __restrict__ packed, uint32_t off, uint32_t end, int k)
{
    uint32_t v = 0;
    for (int t = k - 1; t >= 0; --t) v = (v << 2) | qc_letter(packed, off, end - (uint32_t)t);
    return v;
}

__device__ __forceinline__ uint32_t qc_hash(uint32_t v) { return (v * 0x9E3779B1u) >> (32 - 13); }

// hit: for gap g, n_g * n_g bytes at hit_off[g] (zeroed by the caller), hit[i * n_g + j] = 1 iff (i, j), j >= i, is a candidate.
// slab: per CTA, 3 * slab_probes words of global scratch: probe k-mers, their owner nodes, the chain links.
__global__ void __launch_bounds__(QC_THREADS, 1)
quick_check_kernel(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ seq_off, const uint32_t* __restrict__ seq_len,
                   const uint32_t* __restrict__ gap_first, const uint64_t* __restrict__ hit_off, const QcItem* __restrict__ items,
                   uint32_t n_items, unsigned int* __restrict__ queue, int k, uint32_t* __restrict__ slab, uint32_t slab_probes,
                   uint8_t* __restrict__ hit)
{
    extern __shared__ uint32_t qc_smem[];
    const uint32_t set_words = k >= 3 ? (1u << (2 * k)) / 32u : 1u;       // k = 1, 2: 4 / 16 bits still take a whole word
    uint32_t* kset = qc_smem;
    uint32_t* head = kset + (set_words < 4u ? 4u : set_words);            // QC_HEADS chain heads, 0xffffffff ends a chain
    uint32_t* pk = slab + (size_t)blockIdx.x * 3u * slab_probes;          // probe k-mers
    uint32_t* po = pk + slab_probes;                                      // probe owners
    uint32_t* pn = po + slab_probes;                                      // chain links
    __shared__ uint32_t n_probes, item_idx;
    const uint32_t kmask = k >= 16 ? 0xffffffffu : ((1u << (2 * k)) - 1u);
    uint32_t cur_gap = 0xffffffffu;
    for (;;) {
        if (threadIdx.x == 0) item_idx = atomicAdd(queue, 1u);
        __syncthreads();
        const uint32_t it = item_idx;
        __syncthreads();
        if (it >= n_items) break;
        const QcItem item = items[it];
        const uint32_t first = gap_first[item.gap], n = gap_first[item.gap + 1] - first;
        if (item.gap != cur_gap) {                // consecutive items of one gap reuse the set (the queue hands them out in order)
            cur_gap = item.gap;
            for (uint32_t e = threadIdx.x; e < set_words; e += blockDim.x) kset[e] = 0u;
            for (uint32_t e = threadIdx.x; e < QC_HEADS; e += blockDim.x) head[e] = 0xffffffffu;
            if (threadIdx.x == 0) n_probes = 0;
            __syncthreads();
            // phase 1: the k-mers of the first and last 30 bases of every node (:2026-2029); windows are clipped to the node
            const uint32_t per_side = (uint32_t)(QC_WINDOW - k + 1), per_node = 2u * per_side;
            for (uint32_t e = threadIdx.x; e < n * per_node; e += blockDim.x) {
                const uint32_t j = e / per_node, r = e % per_node, side = r / per_side, a = r % per_side;
                const uint32_t len = seq_len[first + j], wlen = len < (uint32_t)QC_WINDOW ? len : (uint32_t)QC_WINDOW;
                if (a + (uint32_t)k > wlen) continue;
                const uint32_t start = side == 0 ? 0u : len - wlen;
                const uint32_t v = qc_kmer(packed, seq_off[first + j], start + a + (uint32_t)k - 1u, k);
                atomicOr(&kset[v >> 5], 1u << (v & 31u));
                const uint32_t q = atomicAdd(&n_probes, 1u);
                pk[q] = v;
                po[q] = j;
                pn[q] = atomicExch(&head[qc_hash(v)], q);            // push front
            }
            __threadfence_block();
            __syncthreads();
        }
        // phase 2: the item's nodes against the set, 32 bases per thread step
        uint8_t* out = hit + hit_off[item.gap];
        for (uint32_t i = item.node_lo; i < item.node_hi; ++i) {
            const uint32_t len = seq_len[first + i], off = seq_off[first + i];
            if (len < (uint32_t)k) continue;
            const uint32_t n_chunks = (len + QC_CHUNK - 1) / QC_CHUNK;
            for (uint32_t c = threadIdx.x; c < n_chunks; c += blockDim.x) {
                const uint4 w4 = __ldg(reinterpret_cast<const uint4*>(packed + off) + c);       // sequences start 16-byte aligned
                const uint32_t prev = c ? qc_letters(__ldg(packed + off + 4u * c - 1u)) : 0u, prev2 = c ? qc_letters(__ldg(packed + off + 4u * c - 2u)) : 0u;
                // carry-in: the k-1 <= 9 bases before the chunk (the previous two words hold 16)
                uint32_t v = 0;
#pragma unroll
                for (int t = 0; t < 8; ++t) v = (v << 2) | ((prev2 >> (4 * t)) & 3u);
#pragma unroll
                for (int t = 0; t < 8; ++t) v = (v << 2) | ((prev >> (4 * t)) & 3u);
                const uint32_t words[4] = {qc_letters(w4.x), qc_letters(w4.y), qc_letters(w4.z), qc_letters(w4.w)};
                const uint32_t base0 = c * QC_CHUNK;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
#pragma unroll
                    for (int t = 0; t < 8; ++t) {
                        v = ((v << 2) | ((words[q] >> (4 * t)) & 3u)) & kmask;
                        const uint32_t pos = base0 + 8u * q + t;
                        if (pos < len && pos + 1u >= (uint32_t)k && ((kset[v >> 5] >> (v & 31u)) & 1u)) {
                            for (uint32_t p = head[qc_hash(v)]; p != 0xffffffffu; p = __ldcg(pn + p)) {     // slab reads through L2
                                if (__ldcg(pk + p) != v) continue;
                                const uint32_t j = __ldcg(po + p);
                                if (j >= i) out[(size_t)i * n + j] = 1;
                            }
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
}

} // namespace gp
