// gp_api.cu -- CUDA half of the C ABI (include/gappadder_b200.h): context, device buffers, pair
// classification and kernel launches for the overlap DP (ContigsCompactor::Evaluate,
// ContigsCompactor-v0.2.0/ContigsMerger/ContigsCompactor.cpp:1572-1873).  sm_100a only.
#include "gappadder_b200.h"
#include "common.cuh"
#include "overlap_wf32.cuh"
#include "overlap_wf16.cuh"
#include "overlap_wf16t.cuh"
#include "overlap_wf16c.cuh"
#include "quick_check.cuh"
#include "flank_place.cuh"
#include "relax_chain.cuh"
#include "affine_local.cuh"

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local std::string g_create_error;

struct DeviceBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = 2 * bytes + 4096;       // doubling: a relax chain's batches grow a little every step, and a
                                              // (pinned) reallocation costs about a millisecond
        cudaError_t e = cudaMalloc(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    cudaError_t reserve_exact(size_t bytes)     // for the few buffers that are large and do not creep
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct HostBuf {   // pinned staging
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaFreeHost(p); p = nullptr; cap = 0; }
        size_t want = 2 * bytes + 4096;       // doubling: a relax chain's batches grow a little every step, and a
                                              // (pinned) reallocation costs about a millisecond
        cudaError_t e = cudaMallocHost(&p, want);
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

} // namespace

struct gp_ctx {
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    std::string error;
    uint64_t launches = 0;

    // sequence table
    DeviceBuf d_packed;
    std::vector<uint32_t> seq_off, seq_len;
    std::vector<uint8_t> seq_acgt;            // 1: the sequence holds codes 0..3 only (table kernel eligible)
    uint32_t n_symbols = 0;
    uint32_t kernel_mask = GP_KERNEL_ALL;

    // pair work lists
    DeviceBuf d_qc_meta, d_qc_hit, d_qc_slab;  // quick check on the device: offsets/lengths/gap bounds/items, hit matrices, probe slabs
    // relax chains on the device (relax_chain.cuh)
    DeviceBuf d_arena, d_rx_items, d_rx_order, d_rx_status, d_rx_results, d_rx_queue;
    HostBuf h_rx_stage, h_rx_out;
    uint64_t rx_second_passes = 0, rx_unresolved = 0, rx_cells_bound = 0;
    double rx_kernel_ms = 0;
    cudaEvent_t rx_ev[2] = {nullptr, nullptr};
    gp_launch_hook relax_hook = nullptr;        // gp_set_relax_launch_hook
    void* relax_hook_user = nullptr;
    // flank placement (semi-global; flank_place.cuh)
    DeviceBuf d_fp_pairs, d_fp_order, d_fp_results, d_fp_queue, d_fp_scratch;
    HostBuf h_fp_stage;
    uint64_t fp_pairs = 0, fp_n_table = 0, fp_n_generic = 0, fp_cells = 0;
    uint32_t fp_max_n = 0;
    std::vector<uint32_t> fp_host_ids;         // pairs with an empty flank or contig: answered on the host
    gp_dp_params fp_params{};
    cudaEvent_t fp_ev[2] = {nullptr, nullptr};
    bool fp_ev_valid = false;
    // affine local aligner (affine_local.cuh)
    DeviceBuf d_af_pairs, d_af_order, d_af_results, d_af_queue, d_af_scratch, d_ae_scratch;
    HostBuf h_af_stage;
    uint64_t af_pairs = 0, af_work = 0, af_cells = 0;
    uint32_t af_max_m = 0, af_max_n = 0;
    std::vector<uint32_t> af_host_ids;         // pairs with an empty sequence: answered on the host
    gp::AffParams af_params{};
    cudaEvent_t af_ev[3] = {nullptr, nullptr, nullptr};
    bool af_ev_valid = false;
    cudaEvent_t qc_ev[2] = {nullptr, nullptr};
    double qc_kernel_ms = 0;
    uint64_t qc_bases = 0;
    uint32_t qc_items = 0;
    DeviceBuf d_pairs, d_order16c, d_order16t, d_order16, d_order32, d_results, d_queue, d_scratch32, d_scratch16, d_scratch16t, d_scratch16c;
    HostBuf h_stage, h_pack, h_results, h_queue;   // h_queue: [0,32) initial d_queue image, [32,64) counters read back
    std::vector<uint32_t> pack_off;
    double timing[GP_TIMING_SLOTS] = {0};       // milliseconds of the last gp_overlap_batch, see gp_last_timing
    uint64_t n_pairs = 0, n16c = 0, n16t = 0, n16 = 0, n32 = 0, cells = 0;
    uint32_t max_n16c = 0, max_n16c_small = 0, max_n16c_orig = 0, max_n16t = 0, max_n16 = 0, max_n32 = 0;
    uint64_t n16c_transposed = 0;
    uint32_t orientation = 0;                   // certificate kernel: 0 the longer sequence becomes the row sequence, 1 never transpose, 2 always (tests)
    uint32_t cert_system = 0;                   // 0: probe decides, 1: start with system U, 2: with L, 3: with C (tests)
    uint32_t team_mode = 0;                     // certificate kernel: 0 auto, 1 one warp per pair, 2 one CTA per pair
    uint64_t max_cells16c = 0;                  // largest m*n routed to the certificate kernel
    bool last_team = false;                     // what the last launch used
    uint32_t cert_layout = 0;                   // certificate kernel: 0 free-moves layout when the launch allows it, 1 never
    bool last_pot2 = false;
    uint64_t second_passes = 0, exact_retries = 0;   // of the last fetched run
    uint64_t cells16c = 0, cells16t = 0, cells16 = 0, cells32 = 0;     // host-routed DP cells per kernel
    std::vector<uint32_t> closed_ids;           // pairs answered in closed form (a sequence against itself), no DP
    uint64_t cells_closed = 0;                  // m*n of those pairs (what the reference computes for them)
    cudaEvent_t kev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};   // kernel boundaries of the last launch
    bool kev_valid = false;
    gp::Wf16cParams p16c{};
    gp_dp_params params{};
    gp::Wf16Params p16{};
    gp::Wf16tParams p16t{};

    int fail(int code, const char* fmt, ...)
    {
        char buf[512];
        va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
        error = buf;
        return code;
    }
};

#define GP_CUDA(ctx, call)                                                                        \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return (ctx)->fail(GP_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

extern "C" {

int gp_create(int device, gp_ctx** out)
{
    if (!out) return GP_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        g_create_error = std::string("no CUDA device: ") + cudaGetErrorString(e);
        return GP_ERR_NO_DEVICE;
    }
    if (device < 0 || device >= count) { g_create_error = "device index out of range"; return GP_ERR_INVALID; }
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return GP_ERR_CUDA; }
    if (prop.major != 10) {
        g_create_error = "device is not sm_100 (this library is built for B200 only, no fallback)";
        return GP_ERR_NO_DEVICE;
    }
    if ((e = cudaSetDevice(device)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); return GP_ERR_CUDA; }
    gp_ctx* c = new (std::nothrow) gp_ctx();
    if (!c) return GP_ERR_NOMEM;
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) {
        g_create_error = cudaGetErrorString(e); delete c; return GP_ERR_CUDA;
    }
    for (auto& ev : c->qc_ev) cudaEventCreate(&ev);
    for (auto& ev : c->fp_ev) cudaEventCreate(&ev);
    for (auto& ev : c->af_ev) cudaEventCreate(&ev);
    for (auto& ev : c->rx_ev) cudaEventCreate(&ev);
    for (auto& ev : c->kev)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) { g_create_error = cudaGetErrorString(e); cudaStreamDestroy(c->stream); delete c; return GP_ERR_CUDA; }
    if ((e = gp::wf16_configure()) != cudaSuccess || (e = gp::wf16t_configure()) != cudaSuccess || (e = gp::wf16c_configure()) != cudaSuccess || (e = gp::fp_configure()) != cudaSuccess || (e = gp::affine_configure()) != cudaSuccess || (e = gp::relax_configure()) != cudaSuccess ||
        // set once, to the maximum: function attributes are per device, and several contexts (workers) may launch concurrently
        (e = cudaFuncSetAttribute(gp::quick_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gp::QC_SMEM_MAX)) != cudaSuccess) {
        g_create_error = std::string("kernel attribute setup failed: ") + cudaGetErrorString(e);
        cudaStreamDestroy(c->stream); delete c; return GP_ERR_CUDA;
    }
    *out = c;
    return GP_OK;
}

void gp_destroy(gp_ctx* c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) { cudaStreamSynchronize(c->stream); cudaStreamDestroy(c->stream); }
    for (auto& ev : c->kev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : c->qc_ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : c->fp_ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : c->af_ev) if (ev) cudaEventDestroy(ev);
    for (auto& ev : c->rx_ev) if (ev) cudaEventDestroy(ev);
    c->d_arena.release(); c->d_rx_items.release(); c->d_rx_order.release(); c->d_rx_status.release(); c->d_rx_results.release(); c->d_rx_queue.release(); c->h_rx_stage.release(); c->h_rx_out.release();
    c->d_fp_pairs.release(); c->d_fp_order.release(); c->d_fp_results.release(); c->d_fp_queue.release(); c->d_fp_scratch.release(); c->h_fp_stage.release();
    c->d_af_pairs.release(); c->d_af_order.release(); c->d_af_results.release(); c->d_af_queue.release(); c->d_af_scratch.release(); c->d_ae_scratch.release(); c->h_af_stage.release();
    c->d_qc_slab.release();
    c->d_packed.release(); c->d_pairs.release(); c->d_order16t.release(); c->d_order16.release(); c->d_order32.release();
    c->d_scratch16t.release(); c->d_scratch16c.release(); c->d_order16c.release(); c->h_queue.release();
    c->d_results.release(); c->d_queue.release(); c->d_scratch32.release(); c->d_scratch16.release();
    c->h_stage.release(); c->h_pack.release(); c->h_results.release(); c->d_qc_meta.release(); c->d_qc_hit.release();
    delete c;
}

const char* gp_last_error(const gp_ctx* c) { return c ? c->error.c_str() : g_create_error.c_str(); }
void* gp_stream(gp_ctx* c) { return c ? (void*)c->stream : nullptr; }
uint64_t gp_kernel_launches(const gp_ctx* c) { return c ? c->launches : 0; }

int gp_pair_stats(const gp_ctx* c, uint64_t* cells, uint64_t* pairs16, uint64_t* pairs32)
{
    if (!c) return GP_ERR_INVALID;
    if (cells) *cells = c->cells;
    if (pairs16) *pairs16 = c->n16c + c->n16t + c->n16;
    if (pairs32) *pairs32 = c->n32;
    return GP_OK;
}

int gp_pair_split(const gp_ctx* c, uint64_t* table16, uint64_t* prmt16, uint64_t* wide32)
{
    if (!c) return GP_ERR_INVALID;
    if (table16) *table16 = c->n16t;      // routed by the host; the certificate kernel's retries come on top
    if (prmt16) *prmt16 = c->n16;
    if (wide32) *wide32 = c->n32;
    return GP_OK;
}

int gp_cert_stats(const gp_ctx* c, uint64_t* cert16, uint64_t* second_passes, uint64_t* exact_retries)
{
    if (!c) return GP_ERR_INVALID;
    if (cert16) *cert16 = c->n16c;
    if (second_passes) *second_passes = c->second_passes;
    if (exact_retries) *exact_retries = c->exact_retries;
    return GP_OK;
}

int gp_closed_form_stats(const gp_ctx* c, uint64_t* pairs, uint64_t* cells)
{
    if (!c) return GP_ERR_INVALID;
    if (pairs) *pairs = c->closed_ids.size();
    if (cells) *cells = c->cells_closed;
    return GP_OK;
}

int gp_kernel_times(gp_ctx* c, double* ms, uint64_t* cells)
{
    if (!c) return GP_ERR_INVALID;
    if (cells) { cells[0] = c->cells16c; cells[1] = c->cells16t; cells[2] = c->cells16; cells[3] = c->cells32; }
    if (ms) {
        for (int k = 0; k < 4; ++k) ms[k] = 0.0;
        if (c->kev_valid) {
            GP_CUDA(c, cudaSetDevice(c->device));
            GP_CUDA(c, cudaEventSynchronize(c->kev[4]));
            for (int k = 0; k < 4; ++k) {
                float t = 0.f;
                GP_CUDA(c, cudaEventElapsedTime(&t, c->kev[k], c->kev[k + 1]));
                ms[k] = t;
            }
        }
    }
    return GP_OK;
}

int gp_set_cert_system(gp_ctx* c, uint32_t system)
{
    if (!c || system > 3) return GP_ERR_INVALID;
    c->cert_system = system;
    return GP_OK;
}

int gp_set_team_mode(gp_ctx* c, uint32_t mode)
{
    if (!c || mode > 3) return GP_ERR_INVALID;
    c->team_mode = mode;
    return GP_OK;
}

int gp_last_team(const gp_ctx* c) { return c && c->last_team ? 1 : 0; }

int gp_set_orientation(gp_ctx* c, uint32_t mode)
{
    if (!c || mode > 2) return GP_ERR_INVALID;
    c->orientation = mode;
    return GP_OK;
}

uint64_t gp_transposed_pairs(const gp_ctx* c) { return c ? c->n16c_transposed : 0; }

int gp_quick_check_device(gp_ctx* c, const uint32_t* gap_first, uint32_t n_gaps, int32_t k, uint8_t* hit, uint64_t hit_bytes)
{
    return gp_quick_check_matrix(c, gap_first, n_gaps, k, hit, hit_bytes, 0);
}

int gp_quick_check_matrix(gp_ctx* c, const uint32_t* gap_first, uint32_t n_gaps, int32_t k, uint8_t* hit, uint64_t hit_bytes, int full_matrix)
{
    if (!c) return GP_ERR_INVALID;
    if ((!gap_first || !hit) && n_gaps) return c->fail(GP_ERR_INVALID, "null gap bounds / hit buffer");
    if (k <= 0 || k > gp::QC_MAX_K) return c->fail(GP_ERR_RANGE, "quick check on the device supports k <= %d", gp::QC_MAX_K);
    if (n_gaps == 0) return GP_OK;
    const uint32_t n_seq = (uint32_t)c->seq_len.size();
    std::vector<uint64_t> hit_off(n_gaps + 1, 0);
    uint64_t total_bases = 0, max_nodes = 0;
    for (uint32_t g = 0; g < n_gaps; ++g) {
        if (gap_first[g + 1] < gap_first[g] || gap_first[g + 1] > n_seq) return c->fail(GP_ERR_INVALID, "gap %u: bad sequence range", g);
        const uint64_t n = gap_first[g + 1] - gap_first[g];
        if (n > (uint64_t)gp::QC_MAX_NODES) return c->fail(GP_ERR_RANGE, "gap %u has %llu nodes (device quick check: <= %d)", g, (unsigned long long)n, gp::QC_MAX_NODES);
        hit_off[g + 1] = hit_off[g] + n * n;
        max_nodes = std::max(max_nodes, n);
        for (uint32_t s = gap_first[g]; s < gap_first[g + 1]; ++s) total_bases += c->seq_len[s];
    }
    const uint64_t total = hit_off[n_gaps];
    if (total > hit_bytes) return c->fail(GP_ERR_INVALID, "hit buffer too small: %llu bytes needed", (unsigned long long)total);
    // Work items: every gap cut into node ranges of about `item_bases` bases.  An item's first cost is the gap's probe
    // table (42 k-mers per node, rebuilt by every CTA that meets the gap), so items are at least 32 kbases; beyond that
    // they are sized for four items per SM so that few large gaps still fill the chip.
    const uint64_t item_bases = std::min<uint64_t>(1u << 20, std::max<uint64_t>(32u << 10, total_bases / (4ull * (uint64_t)c->sm_count) + 1));
    std::vector<uint32_t> chunk_off(n_seq + 1, 0);                       // 32-base chunks before every table sequence
    for (uint32_t s = 0; s < n_seq; ++s) chunk_off[s + 1] = chunk_off[s] + (c->seq_len[s] + gp::QC_CHUNK - 1) / gp::QC_CHUNK;
    std::vector<gp::QcItem> items;
    for (uint32_t g = 0; g < n_gaps; ++g) {
        const uint32_t first = gap_first[g], n = gap_first[g + 1] - first;
        uint32_t lo = 0;
        uint64_t acc = 0;
        for (uint32_t i = 0; i < n; ++i) {
            acc += c->seq_len[first + i];
            if (acc >= item_bases || i + 1 == n) {
                gp::QcItem it{};
                it.gap = g; it.node_lo = lo; it.node_hi = i + 1;
                it.chunk_begin = chunk_off[first + lo]; it.n_chunks = chunk_off[first + i + 1] - chunk_off[first + lo];
                items.push_back(it);
                lo = i + 1; acc = 0;
            }
        }
    }
    c->qc_bases = total_bases;
    c->qc_items = (uint32_t)items.size();
    if (items.empty()) { memset(hit, 0, (size_t)total); return GP_OK; }
    GP_CUDA(c, cudaSetDevice(c->device));
    // meta: [seq_off n_seq][seq_len n_seq][chunk_off n_seq+1][gap_first n_gaps+1][pad][hit_off (n_gaps+1) x u64][items][queue]
    const size_t w32 = (size_t)3 * n_seq + 1 + n_gaps + 1, w32p = (w32 + 3) & ~(size_t)3;
    const size_t items_at = w32p * 4 + (((size_t)(n_gaps + 1) * 8 + 15) & ~(size_t)15);
    const size_t queue_at = items_at + items.size() * sizeof(gp::QcItem);
    const size_t meta_bytes = queue_at + 16;
    GP_CUDA(c, c->d_qc_meta.reserve(meta_bytes));
    GP_CUDA(c, c->d_qc_hit.reserve(total ? total : 16));
    const int blocks = (int)std::min<size_t>(items.size(), (size_t)c->sm_count);
    const uint32_t need_probes = (uint32_t)max_nodes * gp::qc_probes_per_node(k);
    const uint32_t smem_probes = std::min(need_probes, gp::qc_smem_probe_capacity(k));
    const uint32_t slab_probes = need_probes > smem_probes ? need_probes + 32u : 0u;   // only gaps too big for shared memory use it
    GP_CUDA(c, c->d_qc_slab.reserve(slab_probes ? (size_t)blocks * slab_probes * sizeof(uint32_t) : 16));
    char* dmb = (char*)c->d_qc_meta.p;
    uint32_t* dm = (uint32_t*)dmb;
    uint32_t* d_chunk = dm + 2 * (size_t)n_seq;
    uint32_t* d_gapfirst = d_chunk + n_seq + 1;
    GP_CUDA(c, cudaMemcpyAsync(dm, c->seq_off.data(), (size_t)n_seq * 4, cudaMemcpyHostToDevice, c->stream));
    GP_CUDA(c, cudaMemcpyAsync(dm + n_seq, c->seq_len.data(), (size_t)n_seq * 4, cudaMemcpyHostToDevice, c->stream));
    GP_CUDA(c, cudaMemcpyAsync(d_chunk, chunk_off.data(), (size_t)(n_seq + 1) * 4, cudaMemcpyHostToDevice, c->stream));
    GP_CUDA(c, cudaMemcpyAsync(d_gapfirst, gap_first, (size_t)(n_gaps + 1) * 4, cudaMemcpyHostToDevice, c->stream));
    GP_CUDA(c, cudaMemcpyAsync(dm + w32p, hit_off.data(), (size_t)(n_gaps + 1) * 8, cudaMemcpyHostToDevice, c->stream));
    GP_CUDA(c, cudaMemcpyAsync(dmb + items_at, items.data(), items.size() * sizeof(gp::QcItem), cudaMemcpyHostToDevice, c->stream));
    GP_CUDA(c, cudaMemsetAsync(dmb + queue_at, 0, 16, c->stream));
    if (total) GP_CUDA(c, cudaMemsetAsync(c->d_qc_hit.p, 0, total, c->stream));
    const size_t smem = gp::qc_smem_bytes(k, smem_probes);
    GP_CUDA(c, cudaEventRecord(c->qc_ev[0], c->stream));
    gp::quick_check_kernel<<<blocks, gp::QC_THREADS, smem, c->stream>>>(
        (const uint32_t*)c->d_packed.p, dm, dm + n_seq, d_chunk, d_gapfirst, (const uint64_t*)(dm + w32p),
        (const gp::QcItem*)(dmb + items_at), (uint32_t)items.size(), (unsigned int*)(dmb + queue_at), (int)k,
        smem_probes, (uint32_t*)c->d_qc_slab.p, slab_probes, (uint8_t*)c->d_qc_hit.p, full_matrix ? 1u : 0u);
    GP_CUDA(c, cudaGetLastError());
    GP_CUDA(c, cudaEventRecord(c->qc_ev[1], c->stream));
    c->launches += 1;
    if (total) GP_CUDA(c, cudaMemcpyAsync(hit, c->d_qc_hit.p, total, cudaMemcpyDeviceToHost, c->stream));
    GP_CUDA(c, cudaStreamSynchronize(c->stream));       // the host vectors and the caller's buffers are free again
    float ms = 0.f;
    GP_CUDA(c, cudaEventElapsedTime(&ms, c->qc_ev[0], c->qc_ev[1]));
    c->qc_kernel_ms = ms;
    return GP_OK;
}

int gp_quick_check_stats(const gp_ctx* c, double* kernel_ms, uint64_t* bases, uint32_t* items)
{
    if (!c) return GP_ERR_INVALID;
    if (kernel_ms) *kernel_ms = c->qc_kernel_ms;
    if (bases) *bases = c->qc_bases;
    if (items) *items = c->qc_items;
    return GP_OK;
}

int gp_set_cert_layout(gp_ctx* c, uint32_t mode)
{
    if (!c || mode > 1) return GP_ERR_INVALID;
    c->cert_layout = mode;
    return GP_OK;
}

int gp_last_layout(const gp_ctx* c) { return c && c->last_pot2 ? 1 : 0; }

int gp_set_kernel_mask(gp_ctx* c, uint32_t mask)
{
    if (!c) return GP_ERR_INVALID;
    c->kernel_mask = mask & GP_KERNEL_ALL;
    return GP_OK;
}

static int set_sequences_async(gp_ctx* c, const uint32_t* packed, size_t packed_bytes, const uint32_t* seq_word_off,
                               const uint32_t* seq_len, uint32_t n_seq, uint32_t n_symbols)
{
    if ((!packed || !seq_word_off || !seq_len) && n_seq) return c->fail(GP_ERR_INVALID, "null sequence table");
    if (n_symbols > 16) return c->fail(GP_ERR_ALPHABET, "more than 16 symbols");
    GP_CUDA(c, cudaSetDevice(c->device));
    GP_CUDA(c, c->d_packed.reserve(packed_bytes ? packed_bytes : 16));
    if (packed_bytes) GP_CUDA(c, cudaMemcpyAsync(c->d_packed.p, packed, packed_bytes, cudaMemcpyHostToDevice, c->stream));
    c->seq_off.assign(seq_word_off, seq_word_off + n_seq);
    c->seq_len.assign(seq_len, seq_len + n_seq);
    c->n_symbols = n_symbols;
    c->seq_acgt.assign(n_seq, 1);
    if (n_symbols > 4)                            // some sequence holds N or another letter: find which
        for (uint32_t s = 0; s < n_seq; ++s) {
            const uint32_t* w = packed + seq_word_off[s];
            uint32_t any = 0;
            for (uint32_t k = 0, nw = (seq_len[s] + 7) / 8; k < nw; ++k) any |= w[k] & 0xccccccccu;
            c->seq_acgt[s] = any ? 0 : 1;
        }
    c->n_pairs = 0;
    return GP_OK;
}

int gp_set_sequences(gp_ctx* c, const uint32_t* packed, size_t packed_bytes, const uint32_t* seq_word_off,
                     const uint32_t* seq_len, uint32_t n_seq, uint32_t n_symbols)
{
    if (!c) return GP_ERR_INVALID;
    int rc = set_sequences_async(c, packed, packed_bytes, seq_word_off, seq_len, n_seq, n_symbols);
    if (rc != GP_OK) return rc;
    GP_CUDA(c, cudaStreamSynchronize(c->stream));   // caller may reuse `packed` after return
    return GP_OK;
}

int gp_upload_sequences(gp_ctx* c, const char* const* seqs, const uint32_t* seq_len, uint32_t n_seq)
{
    if (!c) return GP_ERR_INVALID;
    if ((!seqs || !seq_len) && n_seq) return c->fail(GP_ERR_INVALID, "null sequences");
    const size_t bytes = gp_packed_size(seq_len, n_seq);
    GP_CUDA(c, cudaSetDevice(c->device));
    GP_CUDA(c, c->h_pack.reserve(bytes ? bytes : 16));
    c->pack_off.resize(n_seq);
    uint32_t nsym = 0;
    int rc = gp_pack_sequences(seqs, seq_len, n_seq, (uint32_t*)c->h_pack.p, c->pack_off.data(), &nsym);
    if (rc != GP_OK) return c->fail(rc, rc == GP_ERR_ALPHABET ? "more than 16 distinct sequence symbols" : "gp_pack_sequences failed");
    rc = set_sequences_async(c, (const uint32_t*)c->h_pack.p, bytes, c->pack_off.data(), seq_len, n_seq, nsym);
    if (rc != GP_OK) return rc;
    GP_CUDA(c, cudaStreamSynchronize(c->stream));
    return GP_OK;
}

// Evaluate(s, s) in closed form.  With match +1, mismatch <= 1 and indel <= 0 no cell (i,j) scores above
// min(i,j), and on the diagonal of a sequence against itself H(i,i) = i.  The scan (ContigsCompactor.cpp:1679-1709)
// starts with column n top to bottom: every cell above (m,n) scores at most i < m, so (m,n) with score m is the
// first strict maximum and nothing later beats it: scoreMax = m, posRowEnd = m, posColEnd = n, nclip = 0.  The
// predecessor of (i,i) is the diagonal one (it reaches i, tried first and only replaced on strict '<', :1651-1665;
// up and left give at most i-1), so the walk ends in the corner: tbCur = (0,0), bcontained (:1834-1837).
static void closed_form_self(gp_result* r, uint32_t m)
{
    r->score = (int32_t)m; r->row_end = (int32_t)m; r->col_end = (int32_t)m; r->nclip = 0;
    r->flags = GP_FLAG_ROW0 | GP_FLAG_COL0 | GP_FLAG_CONTAINED | GP_FLAG_CLOSED;
}

static void patch_closed(const gp_ctx* c, gp_result* out)
{
    const gp::PairDesc* hd = (const gp::PairDesc*)c->h_stage.p;
    for (uint32_t id : c->closed_ids) closed_form_self(out + id, hd[id].m);
}

static void reset_pairs(gp_ctx* c)
{
    c->n_pairs = 0; c->n16c = c->n16t = c->n16 = c->n32 = 0; c->cells = 0;
    c->cells16c = c->cells16t = c->cells16 = c->cells32 = 0; c->kev_valid = false;
    c->closed_ids.clear(); c->cells_closed = 0; c->max_cells16c = 0;
}

static int upload_pairs_impl(gp_ctx* c, const gp_pair* pairs, uint64_t n_pairs, const gp_dp_params* params)
{
    if (!params || (!pairs && n_pairs)) return c->fail(GP_ERR_INVALID, "null pairs/params");
    if (params->max_clip < 0) return c->fail(GP_ERR_INVALID, "max_clip must be >= 0");
    if (n_pairs > 0xfffffff0ull) return c->fail(GP_ERR_RANGE, "too many pairs in one batch");
    GP_CUDA(c, cudaSetDevice(c->device));
    c->params = *params;
    c->n_pairs = n_pairs; c->n16c = c->n16t = c->n16 = c->n32 = 0; c->cells = 0;
    c->max_n16c = c->max_n16c_small = c->max_n16c_orig = c->max_n16t = c->max_n16 = c->max_n32 = 0; c->n16c_transposed = 0;
    c->cells16c = c->cells16t = c->cells16 = c->cells32 = 0; c->kev_valid = false;
    c->closed_ids.clear(); c->cells_closed = 0; c->max_cells16c = 0;
    if (n_pairs == 0) return GP_OK;

    const uint32_t n_seq = (uint32_t)c->seq_len.size();
    const bool params_ok = gp::wf16_params_ok(params->mismatch, params->indel);
    const bool params16 = params_ok && c->n_symbols <= 8 && (c->kernel_mask & GP_KERNEL_PRMT16);
    const bool params16t = params_ok && (c->kernel_mask & GP_KERNEL_TABLE16);
    const bool params16c = params_ok && (c->kernel_mask & GP_KERNEL_CERT16);
    if (params16) c->p16 = gp::wf16_make_params(params->mismatch, params->indel, params->max_clip, c->n_symbols <= 4);
    if (params16t || params16c) c->p16t = gp::wf16t_make_params(params->mismatch, params->indel, params->max_clip);
    if (params16c) c->p16c = gp::wf16c_make_params(params->mismatch, params->indel, params->max_clip);
    const bool closed_ok = (c->kernel_mask & GP_KERNEL_CLOSED) && params->mismatch <= 1 && params->indel <= 0;

    // stage: [PairDesc n][order16c n][order16t n][order16 n][order32 n]
    const size_t desc_bytes = n_pairs * sizeof(gp::PairDesc);
    const size_t ord_bytes = n_pairs * sizeof(uint32_t);
    GP_CUDA(c, c->h_stage.reserve(desc_bytes + 4 * ord_bytes));
    gp::PairDesc* hd = (gp::PairDesc*)c->h_stage.p;
    uint32_t* ho16c = (uint32_t*)((char*)c->h_stage.p + desc_bytes);
    uint32_t* ho16t = ho16c + n_pairs;
    uint32_t* ho16 = ho16t + n_pairs;
    uint32_t* ho32 = ho16 + n_pairs;
    uint64_t max_total = 0;
    for (uint64_t i = 0; i < n_pairs; ++i) {
        const uint32_t a = pairs[i].row_seq, b = pairs[i].col_seq;
        if (a >= n_seq || b >= n_seq) return c->fail(GP_ERR_INVALID, "pair %llu references sequence out of range", (unsigned long long)i);
        const uint32_t m = c->seq_len[a], n = c->seq_len[b];
        hd[i] = gp::PairDesc{c->seq_off[a], m, c->seq_off[b], n};
        max_total = std::max<uint64_t>(max_total, (uint64_t)m + n);
        // A sequence against itself has a closed form (closed_form_self below): no DP cells are computed.
        if (closed_ok && a == b && m >= 1) { c->closed_ids.push_back((uint32_t)i); c->cells_closed += (uint64_t)m * n; continue; }
        c->cells += (uint64_t)m * n;
        // certificate kernel: everything A/C/G/T except a sequence against itself (closed form above when the scores
        // allow it; otherwise an exact kernel -- system C would certify its corner walk, but it is not worth a probe)
        if (params16c && a != b && gp::wf16c_pair_ok(m, n) && c->seq_acgt[a] && c->seq_acgt[b]) {
            // Orientation (Wf16cPass::tr): the kernel computes the transposed table -- rows = the column sequence -- when the
            // column sequence is the longer one: more full 512-row strips, shorter pipeline fill and drain per strip (measured
            // on cfg1: never 7412, always 7220, shorter-as-rows 7151, longer-as-rows 7531 GCUPS; a per-strip cost model chose
            // the same pairs to within noise).  Results come back in the reference's orientation either way.  Never at the
            // price of the launch's free-moves layout (columns <= WF16C_POT2_MAX_N).
            bool tr = false;
            if (c->orientation != 1 && gp::wf16c_pair_ok(n, m) && (m <= gp::WF16C_POT2_MAX_N || n > gp::WF16C_POT2_MAX_N))
                tr = c->orientation == 2 || (c->orientation == 0 && n > m);   // default: the longer sequence as rows
            const uint32_t cn = tr ? m : n;                             // columns of the computed table
            ho16c[c->n16c++] = (uint32_t)i | (tr ? 0x80000000u : 0u); c->max_n16c = std::max(c->max_n16c, cn); c->cells16c += (uint64_t)m * n;
            c->n16c_transposed += tr ? 1 : 0;
            c->max_cells16c = std::max(c->max_cells16c, (uint64_t)m * n);
            c->max_n16c_orig = std::max(c->max_n16c_orig, n);           // retries go to the exact kernels in the reference's orientation
            if (n <= gp::WF16T_MAX_N) c->max_n16c_small = std::max(c->max_n16c_small, n);
        }
        else if (params16t && gp::wf16t_pair_ok(m, n) && c->seq_acgt[a] && c->seq_acgt[b]) { ho16t[c->n16t++] = (uint32_t)i; c->max_n16t = std::max(c->max_n16t, n); c->cells16t += (uint64_t)m * n; }
        else if (params16 && gp::wf16_pair_ok(m, n)) { ho16[c->n16++] = (uint32_t)i; c->max_n16 = std::max(c->max_n16, n); c->cells16 += (uint64_t)m * n; }
        else { ho32[c->n32++] = (uint32_t)i; c->max_n32 = std::max(c->max_n32, n); c->cells32 += (uint64_t)m * n; }
    }
    // 28-bit score field / 30-bit rank field of the kernels
    const uint64_t amax = (uint64_t)std::max(std::max(std::abs(params->mismatch), std::abs(params->indel)), 1);
    if (amax * max_total >= (1ull << 26) || (uint64_t)(params->max_clip + 1) * (max_total + 2) >= (1ull << 30))
        return c->fail(GP_ERR_RANGE, "sequence lengths / penalties / max_clip exceed the kernels' score or rank range");
    // longest first: the tail of the queue is then made of short pairs (load balance).  The order only has
    // to be roughly by size, so it is a counting sort on 2048 size classes of equal width.
    auto by_cells = [&](uint32_t* ord, uint64_t cnt) {
        if (cnt < 2) return;
        uint64_t mx = 0;
        std::vector<uint64_t> cells(cnt);
        for (uint64_t k = 0; k < cnt; ++k) { const gp::PairDesc& d = hd[ord[k] & 0x7fffffffu]; cells[k] = (uint64_t)d.m * d.n; mx = std::max(mx, cells[k]); }
        constexpr uint32_t B = 2048;
        const double inv_width = (double)B / ((double)mx + 1.0);            // size class without a division per pair
        std::vector<uint32_t> start(B + 1, 0), tmp(ord, ord + cnt), cls(cnt);
        for (uint64_t k = 0; k < cnt; ++k) {
            uint32_t b = (uint32_t)((double)cells[k] * inv_width);
            b = b < B ? b : B - 1;
            cls[k] = B - 1 - b;                                             // largest class first
            ++start[cls[k]];
        }
        uint32_t run = 0;
        for (uint32_t b = 0; b <= B; ++b) { const uint32_t v = start[b]; start[b] = run; run += v; }
        for (uint64_t k = 0; k < cnt; ++k) ord[start[cls[k]]++] = tmp[k];
    };
    by_cells(ho16c, c->n16c);
    by_cells(ho16t, c->n16t);
    by_cells(ho16, c->n16);
    by_cells(ho32, c->n32);

    GP_CUDA(c, c->d_pairs.reserve(desc_bytes));
    GP_CUDA(c, c->d_order16c.reserve(ord_bytes));
    GP_CUDA(c, c->d_order16t.reserve(ord_bytes));
    GP_CUDA(c, c->d_order16.reserve(ord_bytes));
    GP_CUDA(c, c->d_order32.reserve(ord_bytes));
    GP_CUDA(c, c->d_results.reserve(n_pairs * sizeof(gp::DevResult)));
    GP_CUDA(c, c->d_queue.reserve(128));
    GP_CUDA(c, c->h_queue.reserve(256));
    {   // initial image of the queue block: work-queue heads 0, fill counts of the exact kernels' lists
        uint32_t* q = (uint32_t*)c->h_queue.p;
        memset(q, 0, 128);
        q[4] = (uint32_t)c->n16t;
        q[12] = (uint32_t)c->n32;
    }
    GP_CUDA(c, cudaMemcpyAsync(c->d_pairs.p, hd, desc_bytes, cudaMemcpyHostToDevice, c->stream));
    if (c->n16c) GP_CUDA(c, cudaMemcpyAsync(c->d_order16c.p, ho16c, c->n16c * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    if (c->n16t) GP_CUDA(c, cudaMemcpyAsync(c->d_order16t.p, ho16t, c->n16t * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    if (c->n16) GP_CUDA(c, cudaMemcpyAsync(c->d_order16.p, ho16, c->n16 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    if (c->n32) GP_CUDA(c, cudaMemcpyAsync(c->d_order32.p, ho32, c->n32 * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    return GP_OK;       // h_stage is the context's own pinned memory: no need to wait for the copies here
}

// A failed upload leaves NO batch behind: a later gp_launch_resident is then a no-op instead of a launch on work lists
// that were never written.
static int upload_pairs_async(gp_ctx* c, const gp_pair* pairs, uint64_t n_pairs, const gp_dp_params* params)
{
    const int rc = upload_pairs_impl(c, pairs, n_pairs, params);
    if (rc != GP_OK) reset_pairs(c);
    return rc;
}

int gp_upload_pairs(gp_ctx* c, const gp_pair* pairs, uint64_t n_pairs, const gp_dp_params* params)
{
    if (!c) return GP_ERR_INVALID;
    int rc = upload_pairs_async(c, pairs, n_pairs, params);
    if (rc != GP_OK) return rc;
    GP_CUDA(c, cudaStreamSynchronize(c->stream));
    return GP_OK;
}

int gp_launch_resident(gp_ctx* c)
{
    if (!c) return GP_ERR_INVALID;
    if (c->n_pairs == 0) return GP_OK;
    GP_CUDA(c, cudaSetDevice(c->device));
    // queue block (32 words): [0] wf16 head, [8] wf32 head, [16] wf16t head, [24] wf16c head,
    // [4] fill count of the wf16t work list, [12] of the wf32 work list (the certificate kernel appends its
    // retries to both), [20] second passes, [21] exact retries
    GP_CUDA(c, cudaMemcpyAsync(c->d_queue.p, c->h_queue.p, 128, cudaMemcpyHostToDevice, c->stream));
    unsigned int* queue = (unsigned int*)c->d_queue.p;
    const bool cert = c->n16c != 0;
    // One warp per pair or one CTA (team of warps) per pair?  With W warps in flight a batch takes about
    // max(largest pair, total / W) warp-seconds of DP in warp mode; a team runs a pair's strips concurrently
    // (about 3.4x faster per pair at 4 warps, 15 % less aggregate throughput).  Few long pairs -- the relax
    // chain's batches -- are bound by the largest pair and take the team kernel.
    // Eight warps per pair (one CTA per SM) go further the same way: about 6x per pair, 57 % of the aggregate.
    int team = c->team_mode == 2 ? gp::WF16C_TEAM : c->team_mode == 3 ? gp::WF16C_TEAM_BIG : 0;
    if (cert && c->team_mode == 0) {
        const double W = (double)c->sm_count * gp::WF16C_CTAS_PER_SM * gp::WF16C_TEAM;
        const double total = (double)c->cells16c, big = (double)c->max_cells16c;
        const double warp_t = std::max(big, total / W);
        const double team_t = std::max(big / 3.4, total / (0.85 * W));
        const double big_t = std::max(big / 6.0, total / (0.57 * W));
        team = warp_t <= team_t && warp_t <= big_t ? 0 : team_t <= big_t ? gp::WF16C_TEAM : gp::WF16C_TEAM_BIG;
    }
    c->last_team = cert && team != 0;
    c->p16c.pot2 = (cert && c->cert_layout == 0 && c->p16c.std_scores && c->max_n16c <= gp::WF16C_POT2_MAX_N) ? 1 : 0;
    c->last_pot2 = c->p16c.pot2 != 0;
    GP_CUDA(c, cudaEventRecord(c->kev[0], c->stream));
    if (cert) {
        int rc = gp::wf16c_launch(c->stream, c->sm_count, (const uint32_t*)c->d_packed.p, (const gp::PairDesc*)c->d_pairs.p,
                                  (const uint32_t*)c->d_order16c.p, (uint32_t)c->n16c, queue + 24, c->p16c, c->max_n16c,
                                  &c->d_scratch16c.p, &c->d_scratch16c.cap,
                                  (uint32_t*)c->d_order16t.p, queue + 4, (uint32_t*)c->d_order32.p, queue + 12,
                                  queue + 20, c->cert_system, team, (gp::DevResult*)c->d_results.p);
        if (rc != 0) return c->fail(GP_ERR_CUDA, "wf16c launch failed: %s", cudaGetErrorString((cudaError_t)rc));
        c->launches += 1;
    }
    GP_CUDA(c, cudaEventRecord(c->kev[1], c->stream));
    const bool big_retries = cert && c->max_n16c_orig > gp::WF16T_MAX_N; // retries the table kernel cannot take
    if (c->n16t || cert) {
        int rc = gp::wf16t_launch(c->stream, c->sm_count, (const uint32_t*)c->d_packed.p, (const gp::PairDesc*)c->d_pairs.p,
                                  (const uint32_t*)c->d_order16t.p, (uint32_t)c->n16t, queue + 16, c->p16t,
                                  std::max(c->max_n16t, c->max_n16c_small),
                                  &c->d_scratch16t.p, &c->d_scratch16t.cap, (gp::DevResult*)c->d_results.p, cert ? queue + 4 : nullptr);
        if (rc != 0) return c->fail(GP_ERR_CUDA, "wf16t launch failed: %s", cudaGetErrorString((cudaError_t)rc));
        c->launches += 1;
    }
    GP_CUDA(c, cudaEventRecord(c->kev[2], c->stream));
    if (c->n16) {
        int rc = gp::wf16_launch(c->stream, c->sm_count, (const uint32_t*)c->d_packed.p, (const gp::PairDesc*)c->d_pairs.p,
                                 (const uint32_t*)c->d_order16.p, (uint32_t)c->n16, queue, c->p16, c->max_n16,
                                 &c->d_scratch16.p, &c->d_scratch16.cap, (gp::DevResult*)c->d_results.p);
        if (rc != 0) return c->fail(GP_ERR_CUDA, "wf16 launch failed: %s", cudaGetErrorString((cudaError_t)rc));
        c->launches += 1;
    }
    GP_CUDA(c, cudaEventRecord(c->kev[3], c->stream));
    if (c->n32 || big_retries) {
        constexpr int R = 8, THREADS = 128;
        const int blocks = c->sm_count * 8;
        const uint32_t warps = (uint32_t)blocks * (THREADS / 32);
        const uint32_t stride = (std::max(c->max_n32, big_retries ? c->max_n16c_orig : 0u) + 1 + 31) & ~31u;
        GP_CUDA(c, c->d_scratch32.reserve((size_t)warps * stride * sizeof(int32_t)));
        gp::overlap_wf32_kernel<R><<<blocks, THREADS, 0, c->stream>>>(
            (const uint32_t*)c->d_packed.p, (const gp::PairDesc*)c->d_pairs.p, (const uint32_t*)c->d_order32.p,
            (uint32_t)c->n32, queue + 8, c->params.mismatch, c->params.indel, c->params.max_clip,
            (int32_t*)c->d_scratch32.p, stride, (gp::DevResult*)c->d_results.p, big_retries ? queue + 12 : nullptr);
        GP_CUDA(c, cudaGetLastError());
        c->launches += 1;
    }
    GP_CUDA(c, cudaEventRecord(c->kev[4], c->stream));
    c->kev_valid = true;
    return GP_OK;
}

int gp_fetch_results(gp_ctx* c, gp_result* out, uint64_t n_pairs)
{
    if (!c) return GP_ERR_INVALID;
    if (n_pairs != c->n_pairs) return c->fail(GP_ERR_INVALID, "n_pairs does not match the uploaded batch");
    if (n_pairs == 0) return GP_OK;
    if (!out) return c->fail(GP_ERR_INVALID, "null output");
    GP_CUDA(c, cudaSetDevice(c->device));
    static_assert(sizeof(gp_result) == sizeof(gp::DevResult), "result layouts must match");
    GP_CUDA(c, cudaMemcpyAsync(out, c->d_results.p, n_pairs * sizeof(gp_result), cudaMemcpyDeviceToHost, c->stream));
    GP_CUDA(c, cudaMemcpyAsync((uint32_t*)c->h_queue.p + 32, c->d_queue.p, 128, cudaMemcpyDeviceToHost, c->stream));
    GP_CUDA(c, cudaStreamSynchronize(c->stream));
    c->second_passes = ((const uint32_t*)c->h_queue.p)[32 + 20];
    c->exact_retries = ((const uint32_t*)c->h_queue.p)[32 + 21];
    patch_closed(c, out);
    return GP_OK;
}

// Launch + copy the results into the context's pinned buffer + one synchronisation + memcpy to `out`.
static int run_and_fetch(gp_ctx* c, gp_result* out, uint64_t n_pairs)
{
    int rc = gp_launch_resident(c);
    if (rc != GP_OK) return rc;
    if (n_pairs == 0) return GP_OK;
    if (!out) return c->fail(GP_ERR_INVALID, "null output");
    const size_t bytes = n_pairs * sizeof(gp_result);
    GP_CUDA(c, c->h_results.reserve(bytes));
    GP_CUDA(c, cudaMemcpyAsync(c->h_results.p, c->d_results.p, bytes, cudaMemcpyDeviceToHost, c->stream));
    GP_CUDA(c, cudaMemcpyAsync((uint32_t*)c->h_queue.p + 32, c->d_queue.p, 128, cudaMemcpyDeviceToHost, c->stream));
    GP_CUDA(c, cudaStreamSynchronize(c->stream));
    c->second_passes = ((const uint32_t*)c->h_queue.p)[32 + 20];
    c->exact_retries = ((const uint32_t*)c->h_queue.p)[32 + 21];
    memcpy(out, c->h_results.p, bytes);
    patch_closed(c, out);
    return GP_OK;
}

int gp_overlap_pairs(gp_ctx* c, const gp_pair* pairs, uint64_t n_pairs, const gp_dp_params* params, gp_result* out)
{
    if (!c) return GP_ERR_INVALID;
    int rc = upload_pairs_async(c, pairs, n_pairs, params);
    if (rc != GP_OK) return rc;
    return run_and_fetch(c, out, n_pairs);
}

int gp_overlap_batch(gp_ctx* c, const char* const* seqs, const uint32_t* seq_len, uint32_t n_seq,
                     const gp_pair* pairs, uint64_t n_pairs, const gp_dp_params* params, gp_result* out)
{
    if (!c) return GP_ERR_INVALID;
    if ((!seqs || !seq_len) && n_seq) return c->fail(GP_ERR_INVALID, "null sequences");
    using clk = std::chrono::steady_clock;
    auto ms = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto t0 = clk::now();
    // pack into the context's pinned staging buffer (kept across calls), then everything is enqueued
    // on the stream without waiting: H2D of the table, pair descriptors and work orders (built on the
    // host while the table is in flight), kernels, D2H of the results; one synchronisation at the end.
    const size_t bytes = gp_packed_size(seq_len, n_seq);
    GP_CUDA(c, cudaSetDevice(c->device));
    GP_CUDA(c, c->h_pack.reserve(bytes ? bytes : 16));
    if (bytes / sizeof(uint32_t) > 0xffffffffull) return c->fail(GP_ERR_RANGE, "packed table exceeds 2^32 words");
    // The bases are packed on helper threads while this thread classifies, orders and enqueues the pairs: that
    // only needs the lengths and the table offsets, which follow from the lengths (the layout of
    // gp_pack_sequences), plus the assumption that everything is A/C/G/T -- checked after the join; input with N
    // or other letters classifies again (its routing differs).
    c->pack_off.resize(n_seq);
    {
        size_t w = 0;
        for (uint32_t s = 0; s < n_seq; ++s) { c->pack_off[s] = (uint32_t)w; w += gp_packed_size(seq_len + s, 1) / sizeof(uint32_t); }
    }
    c->seq_off.assign(c->pack_off.begin(), c->pack_off.end());
    c->seq_len.assign(seq_len, seq_len + n_seq);
    c->n_symbols = 4;
    c->seq_acgt.assign(n_seq, 1);
    uint32_t nsym = 0;
    int rc_pack = GP_OK;
    std::vector<uint32_t> off_check(n_seq);
    std::thread packer([&] { rc_pack = gp_pack_sequences(seqs, seq_len, n_seq, (uint32_t*)c->h_pack.p, off_check.data(), &nsym); });
    int rc = upload_pairs_async(c, pairs, n_pairs, params);
    const auto t_cls = clk::now();
    packer.join();
    const auto t1 = clk::now();
    if (rc_pack != GP_OK || rc != GP_OK || off_check != c->pack_off) {
        cudaStreamSynchronize(c->stream);
        reset_pairs(c);
        // the host metadata describes the NEW table, the device still holds the old one: drop both, so that later
        // calls on this context fail with GP_ERR_INVALID (pair indices out of range) instead of reading stale bases
        c->seq_off.clear(); c->seq_len.clear(); c->seq_acgt.clear(); c->n_symbols = 0;
        if (rc_pack != GP_OK) return c->fail(rc_pack, rc_pack == GP_ERR_ALPHABET ? "more than 16 distinct sequence symbols" : "gp_pack_sequences failed");
        if (rc != GP_OK) return rc;
        return c->fail(GP_ERR_INVALID, "internal: table layout mismatch");
    }
    GP_CUDA(c, c->d_packed.reserve(bytes ? bytes : 16));
    if (bytes) GP_CUDA(c, cudaMemcpyAsync(c->d_packed.p, c->h_pack.p, bytes, cudaMemcpyHostToDevice, c->stream));
    if (nsym > 4) {                               // N or other letters somewhere: exact per-sequence flags, classify again
        GP_CUDA(c, cudaStreamSynchronize(c->stream));            // the staging buffer of the first classification is in flight
        rc = set_sequences_async(c, (const uint32_t*)c->h_pack.p, 0, c->pack_off.data(), seq_len, n_seq, nsym);
        if (rc == GP_OK) rc = upload_pairs_async(c, pairs, n_pairs, params);
        if (rc != GP_OK) { reset_pairs(c); c->seq_off.clear(); c->seq_len.clear(); c->seq_acgt.clear(); c->n_symbols = 0; return rc; }
    }
    (void)t_cls;
    const auto t2 = clk::now();
    rc = run_and_fetch(c, out, n_pairs);
    const auto t3 = clk::now();
    c->timing[0] = ms(t0, t1);      // pack (helper threads) with pair classification, ordering and upload enqueue beside it
    c->timing[1] = ms(t1, t2);      // table upload enqueue (and a second classification for input with N or other letters)
    c->timing[2] = ms(t2, t3);      // copies + kernels + result copy, until the stream is idle
    c->timing[3] = ms(t0, t3);
    return rc;
}

// One-time allocations up front (pinned staging and device buffers grow on demand otherwise, the first batch paying for
// them: about 20 ms for a 200-gap batch).  Sizes are hints; everything still grows when a batch needs more.
int gp_reserve(gp_ctx* c, uint64_t n_bases, uint32_t n_seq, uint64_t n_pairs, uint64_t n_relax_steps)
{
    if (!c) return GP_ERR_INVALID;
    GP_CUDA(c, cudaSetDevice(c->device));
    const size_t packed = (size_t)(n_bases / 2 + (uint64_t)n_seq * 16 + 64);
    GP_CUDA(c, c->h_pack.reserve(packed));
    GP_CUDA(c, c->d_packed.reserve(packed));
    if (n_pairs) {
        GP_CUDA(c, c->h_stage.reserve((size_t)n_pairs * (sizeof(gp::PairDesc) + 4 * sizeof(uint32_t))));
        GP_CUDA(c, c->d_pairs.reserve((size_t)n_pairs * sizeof(gp::PairDesc)));
        GP_CUDA(c, c->d_order16c.reserve((size_t)n_pairs * 4));
        GP_CUDA(c, c->d_order16t.reserve((size_t)n_pairs * 4));
        GP_CUDA(c, c->d_order16.reserve((size_t)n_pairs * 4));
        GP_CUDA(c, c->d_order32.reserve((size_t)n_pairs * 4));
        GP_CUDA(c, c->d_results.reserve((size_t)n_pairs * sizeof(gp::DevResult)));
        GP_CUDA(c, c->h_results.reserve((size_t)n_pairs * sizeof(gp::DevResult)));
        GP_CUDA(c, c->d_queue.reserve(128));
        GP_CUDA(c, c->h_queue.reserve(256));
    }
    if (n_relax_steps) {
        GP_CUDA(c, c->h_rx_stage.reserve((size_t)n_relax_steps * (sizeof(gp::RelaxItem) + sizeof(uint32_t))));
        GP_CUDA(c, c->d_rx_items.reserve((size_t)n_relax_steps * sizeof(gp::RelaxItem)));
        GP_CUDA(c, c->d_rx_order.reserve((size_t)n_relax_steps * 8 + (size_t)gp::RELAX_RING_SLACK * 4));
        GP_CUDA(c, c->d_rx_status.reserve((size_t)n_relax_steps * 4));
        GP_CUDA(c, c->d_rx_results.reserve((size_t)n_relax_steps * sizeof(gp::DevResult)));
        GP_CUDA(c, c->h_rx_out.reserve((size_t)n_relax_steps * (sizeof(gp::DevResult) + 4) + 64));
        GP_CUDA(c, c->d_rx_queue.reserve(64));
        GP_CUDA(c, c->h_queue.reserve(256));
    }
    return GP_OK;
}

// ---- relax chains on the device (relax_chain.cuh) --------------------------------------------------------------------

int gp_relax_chains(gp_ctx* c, const gp_relax_step* steps, uint64_t n_steps, const gp_dp_params* params, gp_result* out, uint32_t* merged_len)
{
    if (!c) return GP_ERR_INVALID;
    if (!params || ((!steps || !out || !merged_len) && n_steps)) return c->fail(GP_ERR_INVALID, "null steps/params/output");
    c->rx_second_passes = c->rx_unresolved = 0; c->rx_cells_bound = 0; c->rx_kernel_ms = 0;
    if (n_steps == 0) return GP_OK;
    if (n_steps > 0x7ffffff0ull) return c->fail(GP_ERR_RANGE, "too many relax steps in one call");
    if (!(params->mismatch == -2 && params->indel == -2) || params->max_clip < 0)
        return c->fail(GP_ERR_RANGE, "gp_relax_chains runs the standard scores (-2, -2) only; use gp_overlap_batch step by step");
    const uint32_t n_seq = (uint32_t)c->seq_len.size();
    const uint32_t n = (uint32_t)n_steps;
    // items, depth, merged-length bounds (len(merged) <= len(row) + len(column), :108-153), arena slots
    GP_CUDA(c, cudaSetDevice(c->device));
    GP_CUDA(c, c->h_rx_stage.reserve((size_t)n * (sizeof(gp::RelaxItem) + sizeof(uint32_t))));
    gp::RelaxItem* items = (gp::RelaxItem*)c->h_rx_stage.p;
    uint32_t* order = (uint32_t*)((char*)c->h_rx_stage.p + (size_t)n * sizeof(gp::RelaxItem));
    std::vector<uint32_t> depth(n), bound(n);
    std::vector<uint64_t> prio(n, 0);
    uint64_t arena_words = 0, max_total = 0;
    uint32_t max_n = 0;
    for (uint32_t k = 0; k < n; ++k) {
        const gp_relax_step& st = steps[k];
        if (st.col_seq >= n_seq || (st.parent < 0 && st.row_seq >= n_seq) || st.parent >= (int32_t)k)
            return c->fail(GP_ERR_INVALID, "relax step %u: bad sequence index or parent (a parent must precede its children)", k);
        const uint32_t cl = c->seq_len[st.col_seq];
        if (!c->seq_acgt[st.col_seq] || cl < 1 || cl > gp::WF16C_MAX_N || (st.parent < 0 && (!c->seq_acgt[st.row_seq] || c->seq_len[st.row_seq] < 1)))
            return c->fail(GP_ERR_RANGE, "relax step %u: sequences outside the certificate kernel's domain (A/C/G/T, 1..16382 column bases)", k);
        gp::RelaxItem it{};
        it.parent = st.parent;
        it.col_off = c->seq_off[st.col_seq]; it.col_len = cl;
        uint32_t rl;
        if (st.parent < 0) { it.row_off = c->seq_off[st.row_seq]; it.row_len = c->seq_len[st.row_seq]; rl = it.row_len; depth[k] = 0; }
        else { rl = bound[st.parent]; depth[k] = depth[st.parent] + 1; }
        if ((uint64_t)rl + cl > 0xffffffull) return c->fail(GP_ERR_RANGE, "relax step %u: merged contig beyond 2^24 bases", k);
        bound[k] = rl + cl;
        prio[k] = (uint64_t)rl * cl;
        c->rx_cells_bound += prio[k];
        it.arena_off = (uint32_t)arena_words;
        arena_words += (((uint64_t)bound[k] + 7) / 8 + 1 + 31) & ~31ull;          // 128-byte slots: no cache line is shared between items
        if (arena_words > 0xfffffff0ull) return c->fail(GP_ERR_RANGE, "relax arena beyond 2^32 words; split the batch");
        max_n = std::max(max_n, cl);
        max_total = std::max<uint64_t>(max_total, (uint64_t)rl + cl);
        items[k] = it;
    }
    if ((uint64_t)(params->max_clip + 1) * (max_total + 2) >= (1ull << 30)) return c->fail(GP_ERR_RANGE, "max_clip x length exceeds the kernels' rank range");
    // priority = the longest chain from the item downwards (cells, upper bound); children come after parents in `steps`
    std::vector<uint32_t> subtree(n, 1);
    {
        std::vector<uint64_t> below(n, 0);
        for (uint32_t k = n; k-- > 0;) {
            prio[k] += below[k];
            const int32_t p = steps[k].parent;
            if (p >= 0) { below[p] = std::max(below[p], prio[k]); subtree[p] += subtree[k]; }
        }
    }
    // children lists, longest chain first: a CTA follows first_child, the siblings go to the ready ring; roots fill the ring
    std::vector<uint32_t> roots;
    {
        std::vector<uint32_t> by_prio(n);
        for (uint32_t k = 0; k < n; ++k) { by_prio[k] = k; items[k].first_child = -1; items[k].next_sibling = -1; }
        std::stable_sort(by_prio.begin(), by_prio.end(), [&](uint32_t a, uint32_t b) { return prio[a] < prio[b]; });    // ascending: push front
        for (uint32_t k : by_prio) {
            const int32_t p = steps[k].parent;
            if (p < 0) continue;
            items[k].next_sibling = items[p].first_child;
            items[p].first_child = (int32_t)k;
        }
        for (uint32_t q = n; q-- > 0;) if (steps[by_prio[q]].parent < 0) roots.push_back(by_prio[q]);                  // descending priority
    }
    const uint32_t ring_slots = n + gp::RELAX_RING_SLACK;
    uint32_t ring_total = (uint32_t)roots.size();                  // entries the ring will ever hold: roots + non-first children
    for (uint32_t k = 0; k < n; ++k) if (steps[k].parent >= 0 && items[steps[k].parent].first_child != (int32_t)k) ++ring_total;
    (void)depth;

    const size_t item_bytes = (size_t)n * sizeof(gp::RelaxItem);
    GP_CUDA(c, c->d_rx_items.reserve(item_bytes));
    GP_CUDA(c, c->d_rx_order.reserve((size_t)n * 4 + (size_t)ring_slots * 4));         // subtree sizes, then the ready ring
    GP_CUDA(c, c->d_rx_status.reserve((size_t)n * 4));                                   // merged lengths
    GP_CUDA(c, c->d_rx_results.reserve((size_t)n * sizeof(gp::DevResult)));
    GP_CUDA(c, c->d_rx_queue.reserve(64));
    GP_CUDA(c, c->d_arena.reserve((size_t)arena_words * 4 + 256));
    GP_CUDA(c, c->h_rx_out.reserve((size_t)n * (sizeof(gp::DevResult) + 4) + 64));
    c->p16c = gp::wf16c_make_params(params->mismatch, params->indel, params->max_clip);
    const bool pot2 = c->cert_layout == 0 && max_n <= gp::WF16C_POT2_MAX_N;
    c->p16c.pot2 = pot2 ? 1 : 0;
    c->last_pot2 = pot2;
    const int blocks = c->sm_count * gp::WF16C_CTAS_PER_SM;
    const uint32_t warps = (uint32_t)blocks * (gp::WF16C_THREADS / 32);
    const uint32_t stride = (max_n + 2 + 31 + 32) & ~31u;
    GP_CUDA(c, c->d_scratch16c.reserve((size_t)warps * stride * sizeof(uint32_t)));
    uint32_t* d_subtree = (uint32_t*)c->d_rx_order.p;
    uint32_t* d_ring = d_subtree + n;
    uint32_t* d_mlen = (uint32_t*)c->d_rx_status.p;
    memcpy(order, subtree.data(), (size_t)n * 4);                                        // staging: subtree sizes
    GP_CUDA(c, cudaMemcpyAsync(c->d_rx_items.p, items, item_bytes, cudaMemcpyHostToDevice, c->stream));
    GP_CUDA(c, cudaMemcpyAsync(d_subtree, order, (size_t)n * 4, cudaMemcpyHostToDevice, c->stream));
    GP_CUDA(c, cudaMemsetAsync(d_ring, 0xff, (size_t)ring_slots * 4, c->stream));       // RELAX_EMPTY
    GP_CUDA(c, cudaMemcpyAsync(d_ring, roots.data(), roots.size() * 4, cudaMemcpyHostToDevice, c->stream));
    GP_CUDA(c, cudaMemsetAsync(d_mlen, 0, (size_t)n * 4, c->stream));
    {
        uint32_t* q = (uint32_t*)c->h_queue.p;                                           // pinned: [0] head, [1] tail, [2] done, [8] passes, [9] unresolved
        GP_CUDA(c, c->h_queue.reserve(256));
        q = (uint32_t*)c->h_queue.p;
        memset(q, 0, 64);
        q[1] = (uint32_t)roots.size();
        GP_CUDA(c, cudaMemcpyAsync(c->d_rx_queue.p, q, 64, cudaMemcpyHostToDevice, c->stream));
    }
    unsigned int* ctrl = (unsigned int*)c->d_rx_queue.p;
    // diagnostic: GP_RELAX_TRACE=<file> dumps every item's begin / end time (ns) and size after the launch
    static const char* trace_path = getenv("GP_RELAX_TRACE");
    unsigned long long* d_trace = nullptr;
    if (trace_path) { GP_CUDA(c, cudaMalloc(&d_trace, (size_t)n * 24)); GP_CUDA(c, cudaMemsetAsync(d_trace, 0, (size_t)n * 24, c->stream)); }
    const size_t smem = gp::wf16c_smem_bytes<gp::WF16C_THREADS / 32>();
    GP_CUDA(c, cudaEventRecord(c->rx_ev[0], c->stream));
    if (pot2)
        gp::relax_chain_kernel<true><<<blocks, gp::WF16C_THREADS, smem, c->stream>>>(
            (const uint32_t*)c->d_packed.p, (uint32_t*)c->d_arena.p, (const gp::RelaxItem*)c->d_rx_items.p, d_subtree, n, ring_total, d_ring, ctrl,
            c->p16c, (uint32_t*)c->d_scratch16c.p, stride, d_mlen, (gp::DevResult*)c->d_rx_results.p, d_trace);
    else
        gp::relax_chain_kernel<false><<<blocks, gp::WF16C_THREADS, smem, c->stream>>>(
            (const uint32_t*)c->d_packed.p, (uint32_t*)c->d_arena.p, (const gp::RelaxItem*)c->d_rx_items.p, d_subtree, n, ring_total, d_ring, ctrl,
            c->p16c, (uint32_t*)c->d_scratch16c.p, stride, d_mlen, (gp::DevResult*)c->d_rx_results.p, d_trace);
    GP_CUDA(c, cudaGetLastError());
    GP_CUDA(c, cudaEventRecord(c->rx_ev[1], c->stream));
    if (c->relax_hook) c->relax_hook(c->relax_hook_user);         // the kernel is enqueued: a caller may queue other work behind it
    c->launches += 1;
    char* ho = (char*)c->h_rx_out.p;
    GP_CUDA(c, cudaMemcpyAsync(ho, c->d_rx_results.p, (size_t)n * sizeof(gp::DevResult), cudaMemcpyDeviceToHost, c->stream));
    GP_CUDA(c, cudaMemcpyAsync(ho + (size_t)n * sizeof(gp::DevResult), d_mlen, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    GP_CUDA(c, cudaMemcpyAsync((uint32_t*)c->h_queue.p + 32, c->d_rx_queue.p, 64, cudaMemcpyDeviceToHost, c->stream));
    GP_CUDA(c, cudaStreamSynchronize(c->stream));
    if (d_trace) {
        std::vector<unsigned long long> tr((size_t)n * 3);
        cudaMemcpy(tr.data(), d_trace, (size_t)n * 24, cudaMemcpyDeviceToHost);
        cudaFree(d_trace);
        if (FILE* f = fopen(trace_path, "w")) {
            for (uint32_t k = 0; k < n; ++k)
                fprintf(f, "%u %d %llu %llu %llu %llu\n", k, steps[k].parent, tr[3 * (size_t)k], tr[3 * (size_t)k + 1], tr[3 * (size_t)k + 2] >> 32, tr[3 * (size_t)k + 2] & 0xffffffffull);
            fclose(f);
        }
    }
    memcpy(out, ho, (size_t)n * sizeof(gp_result));
    const uint32_t* dev_mlen = (const uint32_t*)(ho + (size_t)n * sizeof(gp::DevResult));
    c->rx_second_passes = ((const uint32_t*)c->h_queue.p)[32 + 8];
    c->rx_unresolved = ((const uint32_t*)c->h_queue.p)[32 + 9];
    // merged lengths by the host epilogue's own rule; a step below an unresolved one is unresolved too; the device's contig
    // lengths (inner steps only: nobody reads a leaf's contig) must agree
    for (uint32_t k = 0; k < n; ++k) {
        const int32_t p = steps[k].parent;
        if (p >= 0 && (out[p].flags & GP_FLAG_UNRESOLVED)) { out[k].flags = GP_FLAG_UNRESOLVED; merged_len[k] = 0; continue; }
        if (out[k].flags & GP_FLAG_UNRESOLVED) { merged_len[k] = 0; continue; }
        const int32_t len1 = (int32_t)(p < 0 ? c->seq_len[steps[k].row_seq] : merged_len[p]), len2 = (int32_t)c->seq_len[steps[k].col_seq];
        merged_len[k] = (uint32_t)gp_merged_length(len1, len2, &out[k]);
        if (items[k].first_child >= 0 && dev_mlen[k] != merged_len[k])
            return c->fail(GP_ERR_CUDA, "internal: relax step %u: device contig length %u, host epilogue %u", k, dev_mlen[k], merged_len[k]);
    }
    float ms = 0.f;
    GP_CUDA(c, cudaEventElapsedTime(&ms, c->rx_ev[0], c->rx_ev[1]));
    c->rx_kernel_ms = ms;
    return GP_OK;
}

#if defined(GP_WF16C_TRACE)
// Diagnostic build: writes the stamps of gp_trace (overlap_wf16c.cuh) as "time_ns block warp tag value" lines and clears them.
extern "C" int gp_debug_trace_dump(const char* path)
{
    unsigned int n = 0;
    if (cudaDeviceSynchronize() != cudaSuccess) return GP_ERR_CUDA;
    cudaMemcpyFromSymbol(&n, gp::gp_trace_n, sizeof n);
    if (n > (1u << 16)) n = 1u << 16;
    std::vector<unsigned long long> buf((size_t)3 * n);
    if (n) cudaMemcpyFromSymbol(buf.data(), gp::gp_trace_buf, (size_t)24 * n);
    const unsigned int zero = 0;
    cudaMemcpyToSymbol(gp::gp_trace_n, &zero, sizeof zero);
    FILE* f = fopen(path, "w");
    if (!f) return GP_ERR_INVALID;
    for (unsigned int i = 0; i < n; ++i)
        fprintf(f, "%llu %llu %llu %llu %llu\n", buf[3 * i], buf[3 * i + 1] >> 32, (buf[3 * i + 1] >> 16) & 0xffff, buf[3 * i + 1] & 0xffff, buf[3 * i + 2]);
    fclose(f);
    return (int)n;
}
#endif

int gp_set_relax_launch_hook(gp_ctx* c, gp_launch_hook hook, void* user)
{
    if (!c) return GP_ERR_INVALID;
    c->relax_hook = hook;
    c->relax_hook_user = user;
    return GP_OK;
}

int gp_relax_stats(const gp_ctx* c, double* kernel_ms, uint64_t* second_passes, uint64_t* unresolved)
{
    if (!c) return GP_ERR_INVALID;
    if (kernel_ms) *kernel_ms = c->rx_kernel_ms;
    if (second_passes) *second_passes = c->rx_second_passes;
    if (unresolved) *unresolved = c->rx_unresolved;
    return GP_OK;
}

// ---- flank placement: semi-global alignment of a flank inside a contig (flank_place.cuh) -------------------------------

static void sort_longest_first(const gp::PairDesc* hd, uint32_t* ord, uint64_t cnt)
{
    std::stable_sort(ord, ord + cnt, [&](uint32_t a, uint32_t b) { return (uint64_t)hd[a].m * hd[a].n > (uint64_t)hd[b].m * hd[b].n; });
}

int gp_semiglobal_upload_pairs(gp_ctx* c, const gp_pair* pairs, uint64_t n_pairs, const gp_dp_params* params)
{
    if (!c) return GP_ERR_INVALID;
    c->fp_pairs = 0; c->fp_n_table = c->fp_n_generic = 0; c->fp_cells = 0; c->fp_max_n = 0; c->fp_host_ids.clear(); c->fp_ev_valid = false;
    if (!params || (!pairs && n_pairs)) return c->fail(GP_ERR_INVALID, "null pairs/params");
    if (params->indel > 0) return c->fail(GP_ERR_INVALID, "flank placement needs indel <= 0");
    if (n_pairs > 0xfffffff0ull) return c->fail(GP_ERR_RANGE, "too many pairs in one batch");
    if (n_pairs == 0) return GP_OK;
    GP_CUDA(c, cudaSetDevice(c->device));
    const uint32_t n_seq = (uint32_t)c->seq_len.size();
    const size_t desc_bytes = n_pairs * sizeof(gp::PairDesc);
    GP_CUDA(c, c->h_fp_stage.reserve(desc_bytes + n_pairs * sizeof(uint32_t)));
    gp::PairDesc* hd = (gp::PairDesc*)c->h_fp_stage.p;
    uint32_t* ord = (uint32_t*)((char*)c->h_fp_stage.p + desc_bytes);
    std::vector<uint32_t> generic;
    uint64_t n_table = 0, cells = 0;
    uint32_t max_n = 0;
    for (uint64_t i = 0; i < n_pairs; ++i) {
        const uint32_t a = pairs[i].row_seq, b = pairs[i].col_seq;
        if (a >= n_seq || b >= n_seq) return c->fail(GP_ERR_INVALID, "pair %llu references sequence out of range", (unsigned long long)i);
        const uint32_t m = c->seq_len[a], n = c->seq_len[b];
        hd[i] = gp::PairDesc{c->seq_off[a], m, c->seq_off[b], n};
        if (m == 0 || n == 0) { c->fp_host_ids.push_back((uint32_t)i); continue; }
        if (!gp::fp_pair_ok(m, n, params->indel) || std::abs(params->mismatch) > 1000)
            return c->fail(GP_ERR_RANGE, "pair %llu (flank %u, contig %u bases) exceeds the placement kernel's range (contig <= %u bases, m + |indel|*(m+n) < 2^17)",
                           (unsigned long long)i, m, n, gp::FP_MAX_N);
        cells += (uint64_t)m * n;
        max_n = std::max(max_n, n);
        if (c->seq_acgt[a] && c->seq_acgt[b]) ord[n_table++] = (uint32_t)i; else generic.push_back((uint32_t)i);
    }
    std::copy(generic.begin(), generic.end(), ord + n_table);
    sort_longest_first(hd, ord, n_table);
    sort_longest_first(hd, ord + n_table, generic.size());
    GP_CUDA(c, c->d_fp_pairs.reserve(desc_bytes));
    GP_CUDA(c, c->d_fp_order.reserve(n_pairs * sizeof(uint32_t)));
    GP_CUDA(c, c->d_fp_results.reserve(n_pairs * sizeof(gp::DevPlace)));
    GP_CUDA(c, c->d_fp_queue.reserve(64));
    GP_CUDA(c, cudaMemcpyAsync(c->d_fp_pairs.p, hd, desc_bytes, cudaMemcpyHostToDevice, c->stream));
    GP_CUDA(c, cudaMemcpyAsync(c->d_fp_order.p, ord, n_pairs * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    GP_CUDA(c, cudaStreamSynchronize(c->stream));
    c->fp_pairs = n_pairs; c->fp_n_table = n_table; c->fp_n_generic = generic.size(); c->fp_cells = cells; c->fp_max_n = max_n;
    c->fp_params = *params;
    return GP_OK;
}

int gp_semiglobal_launch(gp_ctx* c)
{
    if (!c) return GP_ERR_INVALID;
    if (c->fp_n_table + c->fp_n_generic == 0) return GP_OK;
    GP_CUDA(c, cudaSetDevice(c->device));
    const int blocks = c->sm_count * gp::FP_CTAS_PER_SM;
    const uint32_t warps = (uint32_t)blocks * (gp::FP_THREADS / 32);
    const uint32_t stride = (c->fp_max_n + 1 + 64 + 31) & ~31u;
    GP_CUDA(c, c->d_fp_scratch.reserve((size_t)warps * stride * sizeof(uint32_t)));
    GP_CUDA(c, cudaMemsetAsync(c->d_fp_queue.p, 0, 64, c->stream));
    unsigned int* queue = (unsigned int*)c->d_fp_queue.p;
    const uint32_t* ord = (const uint32_t*)c->d_fp_order.p;
    GP_CUDA(c, cudaEventRecord(c->fp_ev[0], c->stream));
    if (c->fp_n_table) {
        gp::flank_place_kernel<true><<<blocks, gp::FP_THREADS, gp::FP_SMEM_BYTES, c->stream>>>(
            (const uint32_t*)c->d_packed.p, (const gp::PairDesc*)c->d_fp_pairs.p, ord, (uint32_t)c->fp_n_table, queue,
            c->fp_params.mismatch, c->fp_params.indel, (uint32_t*)c->d_fp_scratch.p, stride, (gp::DevPlace*)c->d_fp_results.p);
        GP_CUDA(c, cudaGetLastError());
        c->launches += 1;
    }
    if (c->fp_n_generic) {
        gp::flank_place_kernel<false><<<blocks, gp::FP_THREADS, gp::FP_SMEM_BYTES, c->stream>>>(
            (const uint32_t*)c->d_packed.p, (const gp::PairDesc*)c->d_fp_pairs.p, ord + c->fp_n_table, (uint32_t)c->fp_n_generic, queue + 8,
            c->fp_params.mismatch, c->fp_params.indel, (uint32_t*)c->d_fp_scratch.p, stride, (gp::DevPlace*)c->d_fp_results.p);
        GP_CUDA(c, cudaGetLastError());
        c->launches += 1;
    }
    GP_CUDA(c, cudaEventRecord(c->fp_ev[1], c->stream));
    c->fp_ev_valid = true;
    return GP_OK;
}

int gp_semiglobal_fetch(gp_ctx* c, gp_place_result* out, uint64_t n_pairs)
{
    if (!c) return GP_ERR_INVALID;
    if (n_pairs != c->fp_pairs) return c->fail(GP_ERR_INVALID, "n_pairs does not match the uploaded batch");
    if (n_pairs == 0) return GP_OK;
    if (!out) return c->fail(GP_ERR_INVALID, "null output");
    static_assert(sizeof(gp_place_result) == sizeof(gp::DevPlace), "result layouts must match");
    GP_CUDA(c, cudaSetDevice(c->device));
    GP_CUDA(c, cudaMemcpyAsync(out, c->d_fp_results.p, n_pairs * sizeof(gp_place_result), cudaMemcpyDeviceToHost, c->stream));
    GP_CUDA(c, cudaStreamSynchronize(c->stream));
    const gp::PairDesc* hd = (const gp::PairDesc*)c->h_fp_stage.p;
    for (uint32_t id : c->fp_host_ids) {            // an empty flank scores 0 at column 0; an empty contig takes m indels
        out[id].score = hd[id].n == 0 ? (int32_t)hd[id].m * c->fp_params.indel : 0;
        out[id].col_start = 0; out[id].col_end = 0; out[id].flags = 2u;
    }
    return GP_OK;
}

int gp_semiglobal_batch(gp_ctx* c, const char* const* seqs, const uint32_t* seq_len, uint32_t n_seq,
                        const gp_pair* pairs, uint64_t n_pairs, const gp_dp_params* params, gp_place_result* out)
{
    if (!c) return GP_ERR_INVALID;
    int rc = gp_upload_sequences(c, seqs, seq_len, n_seq);
    if (rc == GP_OK) rc = gp_semiglobal_upload_pairs(c, pairs, n_pairs, params);
    if (rc == GP_OK) rc = gp_semiglobal_launch(c);
    if (rc == GP_OK) rc = gp_semiglobal_fetch(c, out, n_pairs);
    return rc;
}

int gp_semiglobal_stats(gp_ctx* c, uint64_t* cells, uint64_t* table_pairs, uint64_t* generic_pairs, double* kernel_ms)
{
    if (!c) return GP_ERR_INVALID;
    if (cells) *cells = c->fp_cells;
    if (table_pairs) *table_pairs = c->fp_n_table;
    if (generic_pairs) *generic_pairs = c->fp_n_generic;
    if (kernel_ms) {
        *kernel_ms = 0.0;
        if (c->fp_ev_valid) {
            GP_CUDA(c, cudaSetDevice(c->device));
            GP_CUDA(c, cudaEventSynchronize(c->fp_ev[1]));
            float t = 0.f;
            GP_CUDA(c, cudaEventElapsedTime(&t, c->fp_ev[0], c->fp_ev[1]));
            *kernel_ms = t;
        }
    }
    return GP_OK;
}

// ---- TERefiner's affine local aligner (affine_local.cuh) ---------------------------------------------------------------

void gp_affine_params_terefiner(gp_affine_params* p)
{
    if (!p) return;
    p->match = 1; p->mismatch = -3; p->n_score = -2; p->gap_open = 5; p->gap_ext = 2; p->band_width = 50;      // aln_param_blast, local_alignment.cpp:193-206
}

int gp_local_affine_upload_pairs(gp_ctx* c, const gp_pair* pairs, uint64_t n_pairs, const gp_affine_params* params)
{
    if (!c) return GP_ERR_INVALID;
    c->af_pairs = 0; c->af_work = 0; c->af_cells = 0; c->af_max_m = c->af_max_n = 0; c->af_host_ids.clear(); c->af_ev_valid = false;
    if (!pairs && n_pairs) return c->fail(GP_ERR_INVALID, "null pairs");
    gp_affine_params dflt;
    gp_affine_params_terefiner(&dflt);
    if (!params) params = &dflt;
    const gp::AffParams P{params->match, params->mismatch, params->n_score, params->gap_open, params->gap_ext, params->band_width};
    if (!gp::aff_params_ok(P)) return c->fail(GP_ERR_INVALID, "affine parameters out of range (match 1..64, mismatch / n_score -1024..0, gap_open 0..1024, gap_ext 1..64, band_width >= 1)");
    if (n_pairs > 0xfffffff0ull) return c->fail(GP_ERR_RANGE, "too many pairs in one batch");
    c->af_params = P;
    if (n_pairs == 0) return GP_OK;
    GP_CUDA(c, cudaSetDevice(c->device));
    const uint32_t n_seq = (uint32_t)c->seq_len.size();
    const size_t desc_bytes = n_pairs * sizeof(gp::PairDesc);
    GP_CUDA(c, c->h_af_stage.reserve(desc_bytes + n_pairs * sizeof(uint32_t)));
    gp::PairDesc* hd = (gp::PairDesc*)c->h_af_stage.p;
    uint32_t* ord = (uint32_t*)((char*)c->h_af_stage.p + desc_bytes);
    uint64_t n_work = 0, cells = 0;
    uint32_t max_m = 0, max_n = 0;
    for (uint64_t i = 0; i < n_pairs; ++i) {
        const uint32_t a = pairs[i].row_seq, b = pairs[i].col_seq;
        if (a >= n_seq || b >= n_seq) return c->fail(GP_ERR_INVALID, "pair %llu references sequence out of range", (unsigned long long)i);
        const uint32_t m = c->seq_len[a], n = c->seq_len[b];
        hd[i] = gp::PairDesc{c->seq_off[a], m, c->seq_off[b], n};
        if (m == 0 || n == 0) { c->af_host_ids.push_back((uint32_t)i); continue; }
        if (!gp::aff_pair_ok(m, n, P))
            return c->fail(GP_ERR_RANGE, "pair %llu (%u x %u bases) exceeds the affine aligner's range (min(len1, len2) * match + gap_open + gap_ext <= %d, lengths < 2^20)",
                           (unsigned long long)i, m, n, gp::AFF_OVERFLOW);
        cells += (uint64_t)m * n;
        max_m = std::max(max_m, m);
        max_n = std::max(max_n, n);
        ord[n_work++] = (uint32_t)i;
    }
    sort_longest_first(hd, ord, n_work);
    GP_CUDA(c, c->d_af_pairs.reserve(desc_bytes));
    GP_CUDA(c, c->d_af_order.reserve(n_pairs * sizeof(uint32_t)));
    GP_CUDA(c, c->d_af_results.reserve(n_pairs * sizeof(gp::DevLocal)));
    GP_CUDA(c, c->d_af_queue.reserve(64));
    GP_CUDA(c, cudaMemcpyAsync(c->d_af_pairs.p, hd, desc_bytes, cudaMemcpyHostToDevice, c->stream));
    GP_CUDA(c, cudaMemcpyAsync(c->d_af_order.p, ord, n_pairs * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    GP_CUDA(c, cudaStreamSynchronize(c->stream));
    c->af_pairs = n_pairs; c->af_work = n_work; c->af_cells = cells; c->af_max_m = max_m; c->af_max_n = max_n;
    return GP_OK;
}

int gp_local_affine_launch(gp_ctx* c)
{
    if (!c) return GP_ERR_INVALID;
    if (c->af_work == 0) return GP_OK;
    GP_CUDA(c, cudaSetDevice(c->device));
    const int blocks = c->sm_count * gp::AF_CTAS_PER_SM;
    const uint32_t warps = (uint32_t)blocks * (gp::AF_THREADS / 32);
    const uint32_t stride = (c->af_max_n + 1 + 64 + 31) & ~31u;
    GP_CUDA(c, c->d_af_scratch.reserve((size_t)warps * stride * sizeof(uint32_t)));
    // start recovery: one warp per pair, 8 ints of scratch per row of the longest row sequence and warp, 4 GB at most
    const size_t per_warp = (gp::aff_epilogue_words((int)c->af_max_m) + 31) & ~(size_t)31;
    const size_t budget_words = (size_t)1 << 30;
    constexpr size_t ae_warps_per_block = gp::AE_THREADS / 32;
    size_t ae_blocks = std::min<size_t>((size_t)c->sm_count * 8, (c->af_work + ae_warps_per_block - 1) / ae_warps_per_block);
    ae_blocks = std::max<size_t>(1, std::min(ae_blocks, budget_words / (per_warp * ae_warps_per_block)));
    GP_CUDA(c, c->d_ae_scratch.reserve_exact(ae_blocks * ae_warps_per_block * per_warp * sizeof(int)));
    GP_CUDA(c, cudaMemsetAsync(c->d_af_queue.p, 0, 64, c->stream));
    unsigned int* queue = (unsigned int*)c->d_af_queue.p;
    GP_CUDA(c, cudaEventRecord(c->af_ev[0], c->stream));
    gp::affine_forward_kernel<<<blocks, gp::AF_THREADS, gp::AF_SMEM_BYTES, c->stream>>>(
        (const uint32_t*)c->d_packed.p, (const gp::PairDesc*)c->d_af_pairs.p, (const uint32_t*)c->d_af_order.p, (uint32_t)c->af_work, queue,
        c->af_params, (uint32_t*)c->d_af_scratch.p, stride, (gp::DevLocal*)c->d_af_results.p);
    GP_CUDA(c, cudaGetLastError());
    GP_CUDA(c, cudaEventRecord(c->af_ev[1], c->stream));
    gp::affine_epilogue_kernel<<<(unsigned)ae_blocks, gp::AE_THREADS, 0, c->stream>>>(
        (const uint32_t*)c->d_packed.p, (const gp::PairDesc*)c->d_af_pairs.p, (const uint32_t*)c->d_af_order.p, (uint32_t)c->af_work, queue + 8,
        c->af_params, (int*)c->d_ae_scratch.p, per_warp, (gp::DevLocal*)c->d_af_results.p);
    GP_CUDA(c, cudaGetLastError());
    GP_CUDA(c, cudaEventRecord(c->af_ev[2], c->stream));
    c->launches += 2;
    c->af_ev_valid = true;
    return GP_OK;
}

int gp_local_affine_fetch(gp_ctx* c, gp_local_result* out, uint64_t n_pairs)
{
    if (!c) return GP_ERR_INVALID;
    if (n_pairs != c->af_pairs) return c->fail(GP_ERR_INVALID, "n_pairs does not match the uploaded batch");
    if (n_pairs == 0) return GP_OK;
    if (!out) return c->fail(GP_ERR_INVALID, "null output");
    if (c->af_work && !c->af_ev_valid) return c->fail(GP_ERR_INVALID, "gp_local_affine_fetch before gp_local_affine_launch");
    static_assert(sizeof(gp_local_result) == sizeof(gp::DevLocal), "result layouts must match");
    GP_CUDA(c, cudaSetDevice(c->device));
    GP_CUDA(c, cudaMemcpyAsync(out, c->d_af_results.p, n_pairs * sizeof(gp_local_result), cudaMemcpyDeviceToHost, c->stream));
    GP_CUDA(c, cudaStreamSynchronize(c->stream));
    for (uint32_t id : c->af_host_ids) {            // aln_local_core returns -1 for an empty sequence (local_alignment.cpp:545)
        out[id].score = -1; out[id].start1 = out[id].end1 = out[id].start2 = out[id].end2 = 0; out[id].flags = GP_LOCAL_NO_MATCH;
    }
    return GP_OK;
}

int gp_local_affine_batch(gp_ctx* c, const char* const* seqs, const uint32_t* seq_len, uint32_t n_seq,
                          const gp_pair* pairs, uint64_t n_pairs, const gp_affine_params* params, gp_local_result* out)
{
    if (!c) return GP_ERR_INVALID;
    if ((!seqs || !seq_len) && n_seq) return c->fail(GP_ERR_INVALID, "null sequences");
    // the reference's letter classes (aln_nt4_table, local_alignment.cpp:32-49): a/A c/C g/G t/T, everything else one class
    static const struct Nt4 { char t[256]; Nt4() { memset(t, 'N', sizeof t); for (const char* p = "ACGT"; *p; ++p) { t[(unsigned char)*p] = *p; t[(unsigned char)(*p + 32)] = *p; } } } nt4;
    uint64_t total = 0;
    for (uint32_t s = 0; s < n_seq; ++s) { if (!seqs[s] && seq_len[s]) return c->fail(GP_ERR_INVALID, "null sequence %u", s); total += seq_len[s]; }
    std::vector<char> norm(total ? total : 1);
    std::vector<const char*> ptr(n_seq);
    uint64_t at = 0;
    for (uint32_t s = 0; s < n_seq; ++s) {
        ptr[s] = norm.data() + at;
        const unsigned char* src = (const unsigned char*)seqs[s];
        for (uint32_t i = 0; i < seq_len[s]; ++i) norm[at + i] = nt4.t[src[i]];
        at += seq_len[s];
    }
    int rc = gp_upload_sequences(c, ptr.data(), seq_len, n_seq);
    if (rc == GP_OK) rc = gp_local_affine_upload_pairs(c, pairs, n_pairs, params);
    if (rc == GP_OK) rc = gp_local_affine_launch(c);
    if (rc == GP_OK) rc = gp_local_affine_fetch(c, out, n_pairs);
    return rc;
}

int gp_local_affine_stats(gp_ctx* c, uint64_t* cells, double* forward_ms, double* epilogue_ms)
{
    if (!c) return GP_ERR_INVALID;
    if (cells) *cells = c->af_cells;
    if (forward_ms) *forward_ms = 0.0;
    if (epilogue_ms) *epilogue_ms = 0.0;
    if ((forward_ms || epilogue_ms) && c->af_ev_valid) {
        GP_CUDA(c, cudaSetDevice(c->device));
        GP_CUDA(c, cudaEventSynchronize(c->af_ev[2]));
        float t = 0.f;
        GP_CUDA(c, cudaEventElapsedTime(&t, c->af_ev[0], c->af_ev[1]));
        if (forward_ms) *forward_ms = t;
        GP_CUDA(c, cudaEventElapsedTime(&t, c->af_ev[1], c->af_ev[2]));
        if (epilogue_ms) *epilogue_ms = t;
    }
    return GP_OK;
}

int gp_last_timing(const gp_ctx* c, double* out_ms, int n)
{
    if (!c || !out_ms || n < 0) return GP_ERR_INVALID;
    for (int i = 0; i < n && i < GP_TIMING_SLOTS; ++i) out_ms[i] = c->timing[i];
    return GP_OK;
}

} // extern "C"
