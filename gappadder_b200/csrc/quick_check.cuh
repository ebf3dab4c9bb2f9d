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
//            4^k-bit set in shared memory (128 KB for k = 10, GAPPadder's value) and, as (k-mer | owner << 20) words
//            sorted by hash bucket (count, scan, scatter), into shared memory next to it; a gap with more probes
//            than fit there (about 400 nodes at k = 10) keeps them in a per-CTA slab of global scratch instead;
//   phase 2  the item's nodes stream past the set as one flat list of 32-base chunks: a thread takes a chunk (one
//            16-byte load of packed codes plus the two words before it), rolls the k-mer along it -- a shift, an or
//            and one shared-memory bit test per base -- and on a hit compares the bucket's probes and stores
//            hit(i, owner) = 1 (a plain byte store: idempotent, no atomics; the matrix is zeroed before the launch).
// The packed codes arrive at 0.5 B per base and are read exactly once (plus 8 B of carry-in per 32 bases): an
// HBM-bound byte-stream scan with no arithmetic to speak of.  bench.py reports its achieved GB/s.
//
// Limits: k <= 10 (the set must fit shared memory), QC_MAX_NODES nodes per gap (12-bit owner field; the hit matrix is
// n x n bytes); gp_quick_check_device returns GP_ERR_RANGE beyond them and the host filter has no such limits.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cuda_runtime.h>

namespace gp {

constexpr int QC_THREADS = 512;
constexpr int QC_MAX_K = 10;                      // 4^10 bits = 128 KB of shared memory
constexpr int QC_MAX_NODES = 4096;                // nodes per gap (contigs and their reverse complements): 12-bit owner field
constexpr int QC_WINDOW = 30;                     // lenContigLen, ContigsCompactor.cpp:2024
constexpr int QC_BUCKETS = 4096;                  // hash buckets of the probe table
constexpr int QC_CHUNK = 32;                      // bases per thread step: four packed words, one 16-byte load
constexpr size_t QC_SMEM_MAX = 227 * 1024 - 2048; // dynamic shared memory the kernel may ask for

struct QcItem {                                   // one unit of work: scan nodes [node_lo, node_hi) of gap `gap`
    uint32_t gap, node_lo, node_hi;
    uint32_t chunk_begin, n_chunks;               // the nodes' 32-base chunks in the table-wide chunk numbering
    uint32_t pad[3];
};

__host__ __device__ inline uint32_t qc_probes_per_node(int k) { return 2u * (uint32_t)(QC_WINDOW - k + 1); }
inline size_t qc_set_bytes(int k) { return std::max<size_t>(16, ((size_t)1 << (2 * k)) / 8); }
// shared memory: the set, the bucket ends, and room for `smem_probes` probe words
inline size_t qc_smem_bytes(int k, uint32_t smem_probes) { return qc_set_bytes(k) + (size_t)QC_BUCKETS * 4 + (size_t)smem_probes * 4 + 64; }
inline uint32_t qc_smem_probe_capacity(int k) { return (uint32_t)((QC_SMEM_MAX - qc_set_bytes(k) - (size_t)QC_BUCKETS * 4 - 64) / 4); }

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

__device__ __forceinline__ uint32_t qc_hash(uint32_t v) { return (v * 0x9E3779B1u) >> (32 - 12); }

// The window k-mer number e of a gap (e = node * per_node + side * per_side + a): its value and owner; false when the
// window is shorter than a + k bases.
__device__ __forceinline__ bool qc_probe(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ seq_off,
                                         const uint32_t* __restrict__ seq_len, uint32_t first, uint32_t e, int k, uint32_t& v, uint32_t& j)
{
    const uint32_t per_side = (uint32_t)(QC_WINDOW - k + 1), per_node = 2u * per_side;
    j = e / per_node;
    const uint32_t r = e % per_node, side = r / per_side, a = r % per_side;
    const uint32_t len = seq_len[first + j], wlen = len < (uint32_t)QC_WINDOW ? len : (uint32_t)QC_WINDOW;
    if (a + (uint32_t)k > wlen) return false;
    const uint32_t start = side == 0 ? 0u : len - wlen;
    v = qc_kmer(packed, seq_off[first + j], start + a + (uint32_t)k - 1u, k);
    return true;
}

// hit: for gap g, n_g * n_g bytes at hit_off[g] (zeroed by the caller), hit[i * n_g + j] = 1 iff (i, j), j >= i, is a candidate;
// full_matrix != 0: also for j < i (the dedup stage asks "do the ends of j occur in i" for every ordered pair).
// chunk_off: per table sequence, the number of 32-base chunks of all sequences before it (n_seq + 1 entries).
// slab: per CTA, slab_probes words of global scratch for gaps whose probes do not fit shared memory (smem_probes words).
__global__ void __launch_bounds__(QC_THREADS, 1)
quick_check_kernel(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ seq_off, const uint32_t* __restrict__ seq_len,
                   const uint32_t* __restrict__ chunk_off, const uint32_t* __restrict__ gap_first, const uint64_t* __restrict__ hit_off,
                   const QcItem* __restrict__ items, uint32_t n_items, unsigned int* __restrict__ queue, int k,
                   uint32_t smem_probes, uint32_t* __restrict__ slab, uint32_t slab_probes, uint8_t* __restrict__ hit, uint32_t full_matrix)
{
    extern __shared__ uint32_t qc_smem[];
    const uint32_t set_words = k >= 3 ? (1u << (2 * k)) / 32u : 4u;       // k = 1, 2: 4 / 16 bits still take whole words
    uint32_t* kset = qc_smem;
    uint32_t* bend = kset + set_words;                                    // QC_BUCKETS bucket ends (exclusive prefix while filling)
    uint32_t* sprobes = bend + QC_BUCKETS;
    __shared__ uint32_t item_idx, warp_sums[QC_THREADS / 32];
    const uint32_t kmask = (1u << (2 * k)) - 1u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t cur_gap = 0xffffffffu;
    const uint32_t* probes = sprobes;
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
            const uint32_t n_probe_slots = n * qc_probes_per_node(k);
            uint32_t* wprobes = n_probe_slots <= smem_probes ? sprobes : slab + (size_t)blockIdx.x * slab_probes;
            probes = wprobes;
            for (uint32_t e = threadIdx.x; e < set_words; e += blockDim.x) kset[e] = 0u;
            for (uint32_t e = threadIdx.x; e < QC_BUCKETS; e += blockDim.x) bend[e] = 0u;
            __syncthreads();
            // phase 1a: the k-mers of the first and last 30 bases of every node (:2026-2029) into the set; bucket counts
            for (uint32_t e = threadIdx.x; e < n_probe_slots; e += blockDim.x) {
                uint32_t v, j;
                if (!qc_probe(packed, seq_off, seq_len, first, e, k, v, j)) continue;
                atomicOr(&kset[v >> 5], 1u << (v & 31u));
                atomicAdd(&bend[qc_hash(v)], 1u);
            }
            __syncthreads();
            // phase 1b: exclusive prefix sum of the counts (QC_BUCKETS / QC_THREADS = 8 buckets per thread)
            {
                constexpr int PER = QC_BUCKETS / QC_THREADS;
                uint32_t local[PER], sum = 0;
#pragma unroll
                for (int x = 0; x < PER; ++x) { local[x] = bend[threadIdx.x * PER + x]; sum += local[x]; }
                uint32_t incl = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += y; }
                if (lane == 31) warp_sums[warp] = incl;
                __syncthreads();
                uint32_t base = 0;
                for (int w = 0; w < warp; ++w) base += warp_sums[w];
                uint32_t run = base + incl - sum;
#pragma unroll
                for (int x = 0; x < PER; ++x) { bend[threadIdx.x * PER + x] = run; run += local[x]; }
            }
            __syncthreads();
            // phase 1c: scatter (k-mer | owner << 20) into the buckets; afterwards bend[h] is the END of bucket h
            for (uint32_t e = threadIdx.x; e < n_probe_slots; e += blockDim.x) {
                uint32_t v, j;
                if (!qc_probe(packed, seq_off, seq_len, first, e, k, v, j)) continue;
                wprobes[atomicAdd(&bend[qc_hash(v)], 1u)] = v | (j << 20);
            }
            __threadfence_block();
            __syncthreads();
        }
        // phase 2: the item's nodes against the set, as one flat list of 32-base chunks
        uint8_t* out = hit + hit_off[item.gap];
        for (uint32_t f = threadIdx.x; f < item.n_chunks; f += blockDim.x) {
            const uint32_t gchunk = item.chunk_begin + f;
            uint32_t lo = first + item.node_lo, hi = first + item.node_hi;      // the sequence s with chunk_off[s] <= gchunk < chunk_off[s+1]
            while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (__ldg(chunk_off + mid) <= gchunk) lo = mid; else hi = mid; }
            const uint32_t i = lo - first, c = gchunk - __ldg(chunk_off + lo);
            const uint32_t len = seq_len[lo], off = seq_off[lo];
            if (len < (uint32_t)k) continue;
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
                        const uint32_t h = qc_hash(v);
                        for (uint32_t p = h ? bend[h - 1] : 0u, pe = bend[h]; p < pe; ++p) {
                            const uint32_t pv = probes[p];
                            if ((pv & 0xfffffu) == v && ((pv >> 20) >= i || full_matrix)) out[(size_t)i * n + (pv >> 20)] = 1;
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
}

} // namespace gp
