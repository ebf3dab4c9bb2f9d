// quick_check.cuh -- the candidate filter of the pairwise phase on the device (sm_100a).
//
// Reference: MultiThreadQuickChecker::threadQuickCheck and QuickCheckerContigsMatch
// (ContigsCompactor-v0.2.0/ContigsMerger/ContigsCompactor.cpp:992-1100, :1982-2095) with the 2-bit k-mer code of
// KmerUtils.cpp:22-115: pair (i, j), i <= j (j == i included), of a gap's nodes is aligned iff some k-mer of the
// first or the last 30 bases of node j occurs anywhere in node i.  The host form is gp_candidate_pairs
// (gp_host.cpp); this one works on the packed table that is already in HBM for the DP kernels, so that a batch
// of gaps needs no host pass over the bases at all.
//
// One CTA per gap.  Shared memory: a 4^k-bit set of the gap's probe k-mers (128 KB for k = 10, GAPPadder's value),
// the probes themselves (k-mer, owner) chained by a small hash, and the gap's hit matrix.  Phase 1 puts every window
// k-mer of every node into the set and the chains; phase 2 streams every k-mer of every node past the set (one
// shared-memory bit test per base; packed codes come through L1/L2 at 0.5 B per base) and, on a hit, walks the
// chain and sets hit(i, owner).  It is a byte-stream scan: HBM- and shared-memory-bound, no arithmetic to speak of.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gp {

constexpr int QC_THREADS = 512;
constexpr int QC_MAX_K = 10;                      // 4^10 bits = 128 KB of shared memory
constexpr int QC_MAX_NODES = 256;                 // nodes per gap (contigs and their reverse complements)
constexpr int QC_WINDOW = 30;                     // lenContigLen, ContigsCompactor.cpp:2024
constexpr int QC_HEADS = 8192;

inline uint32_t qc_max_probes(int k) { return (uint32_t)QC_MAX_NODES * 2u * (uint32_t)(QC_WINDOW - k + 1); }   // 10752 for k = 10
inline size_t qc_smem_bytes(int k)
{
    return ((size_t)1 << (2 * k)) / 8 + (size_t)qc_max_probes(k) * 4 + (size_t)qc_max_probes(k) * 2 + (size_t)QC_HEADS * 2 +
           (size_t)QC_MAX_NODES * QC_MAX_NODES / 8 + 64;                  // 217 KB for k = 10
}

// 2-bit k-mer letter of the packed 4-bit code at base `pos`: A C G T -> 0..3, everything else 0 (KmerUtils.cpp:22-58)
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

// hit: for gap g, n_g * n_g bytes at hit_off[g], hit[i * n_g + j] = 1 iff (i, j), j >= i, is a candidate
__global__ void __launch_bounds__(QC_THREADS, 1)
quick_check_kernel(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ seq_off, const uint32_t* __restrict__ seq_len,
                   const uint32_t* __restrict__ gap_first, const uint64_t* __restrict__ hit_off, uint32_t n_gaps, int k,
                   uint32_t max_probes, uint8_t* __restrict__ hit)
{
    extern __shared__ uint32_t qc_smem[];
    const uint32_t set_words = (1u << (2 * k)) / 32u;
    uint32_t* kset = qc_smem;
    uint32_t* probe = kset + set_words;                                   // k-mer | owner << 20
    uint16_t* next = reinterpret_cast<uint16_t*>(probe + max_probes);     // chain, 0xffff ends it
    uint16_t* head = next + max_probes;
    uint32_t* mat = reinterpret_cast<uint32_t*>(head + QC_HEADS);         // n x n bits
    __shared__ uint32_t n_probes;
    for (uint32_t g = blockIdx.x; g < n_gaps; g += gridDim.x) {
        const uint32_t first = gap_first[g], n = gap_first[g + 1] - first;
        for (uint32_t e = threadIdx.x; e < set_words; e += blockDim.x) kset[e] = 0u;
        for (uint32_t e = threadIdx.x; e < QC_HEADS / 2; e += blockDim.x) reinterpret_cast<uint32_t*>(head)[e] = 0xffffffffu;
        for (uint32_t e = threadIdx.x; e < (n * n + 31) / 32; e += blockDim.x) mat[e] = 0u;
        if (threadIdx.x == 0) n_probes = 0;
        __syncthreads();
        // phase 1: the k-mers of the first and last 30 bases of every node (:2026-2029); windows are clipped to the node
        const uint32_t per_node = 2u * (uint32_t)(QC_WINDOW - k + 1);
        for (uint32_t e = threadIdx.x; e < n * per_node; e += blockDim.x) {
            const uint32_t j = e / per_node, r = e % per_node, side = r / (uint32_t)(QC_WINDOW - k + 1), a = r % (uint32_t)(QC_WINDOW - k + 1);
            const uint32_t len = seq_len[first + j], wlen = len < (uint32_t)QC_WINDOW ? len : (uint32_t)QC_WINDOW;
            if (a + (uint32_t)k > wlen) continue;
            const uint32_t start = side == 0 ? 0u : len - wlen;
            const uint32_t v = qc_kmer(packed, seq_off[first + j], start + a + (uint32_t)k - 1u, k);
            atomicOr(&kset[v >> 5], 1u << (v & 31u));
            const uint32_t q = atomicAdd(&n_probes, 1u);
            probe[q] = v | (j << 20);
            const uint32_t h = (v * 0x9E3779B1u) >> (32 - 13);
            // push front (16-bit exchange on the containing word)
            uint32_t* hw = reinterpret_cast<uint32_t*>(head) + (h >> 1);
            uint32_t old = *hw, assumed;
            do {
                assumed = old;
                const uint32_t prev = (h & 1u) ? (assumed >> 16) : (assumed & 0xffffu);
                next[q] = (uint16_t)prev;
                const uint32_t repl = (h & 1u) ? ((assumed & 0xffffu) | (q << 16)) : ((assumed & 0xffff0000u) | q);
                old = atomicCAS(hw, assumed, repl);
            } while (old != assumed);
        }
        __syncthreads();
        // phase 2: every k-mer of every node against the set; the warp walks a node's bases together (coalesced words)
        for (uint32_t i = 0; i < n; ++i) {
            const uint32_t len = seq_len[first + i], off = seq_off[first + i];
            if (len < (uint32_t)k) continue;
            for (uint32_t end = (uint32_t)k - 1u + threadIdx.x; end < len; end += blockDim.x) {
                const uint32_t v = qc_kmer(packed, off, end, k);
                if (!((kset[v >> 5] >> (v & 31u)) & 1u)) continue;
                for (uint32_t q = head[(v * 0x9E3779B1u) >> (32 - 13)]; q != 0xffffu; q = next[q]) {
                    const uint32_t pv = probe[q];
                    if ((pv & 0xfffffu) != v) continue;
                    const uint32_t j = pv >> 20;
                    if (j >= i) { const uint32_t b = i * n + j; atomicOr(&mat[b >> 5], 1u << (b & 31u)); }
                }
            }
        }
        __syncthreads();
        uint8_t* out = hit + hit_off[g];
        for (uint32_t e = threadIdx.x; e < n * n; e += blockDim.x) out[e] = (uint8_t)((mat[e >> 5] >> (e & 31u)) & 1u);
        __syncthreads();
    }
}

} // namespace gp
