// tests/shim_gp_oracle.cpp -- TEST INFRASTRUCTURE ONLY.
// Serves the DP entry points of include/gappadder_b200.h from the CPU oracle
// (oracle/overlap_oracle.c) so that the HOST logic of the drop-in binary (FASTA I/O, candidate
// filter, "Arrange error" rule, overlap graph, path search, merge chains, output formatting) can be
// checked against the reference's golden outputs on a machine without a GPU.  Linked only into
// build/ContigsMerger_hosttest by tests/test_contigsmerger_host.py; the product binary links
// libgappadder_b200.so and has no such path.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "gappadder_b200.h"

extern "C" {
typedef struct { int32_t score, row_end, col_end, nclip, tb_row, tb_col, bcontained; } gpo_dp_result;
int gpo_evaluate(const char* s1, int m, const char* s2, int n, int mismatch, int indel, int maxclip, gpo_dp_result* r);
}

struct gp_ctx {
    std::string err;
    std::vector<std::string> seqs;        // the table of gp_set_sequences, unpacked again
    std::vector<gp_pair> pairs;           // gp_upload_pairs
    gp_dp_params params{};
    gp_launch_hook hook = nullptr;
    void* hook_user = nullptr;
};

extern "C" {
int gp_create(int, gp_ctx** out) { *out = new gp_ctx(); return GP_OK; }
void gp_destroy(gp_ctx* c) { delete c; }
const char* gp_last_error(const gp_ctx* c) { return c ? c->err.c_str() : ""; }
int gp_overlap_batch(gp_ctx*, const char* const* seqs, const uint32_t* seq_len, uint32_t, const gp_pair* pairs,
                     uint64_t n_pairs, const gp_dp_params* p, gp_result* out)
{
    std::atomic<uint64_t> next(0);
    std::atomic<int> rc(GP_OK);
    auto work = [&] {
        for (uint64_t k = next.fetch_add(1); k < n_pairs; k = next.fetch_add(1)) {
            gpo_dp_result r;
            const uint32_t a = pairs[k].row_seq, b = pairs[k].col_seq;
            if (gpo_evaluate(seqs[a], (int)seq_len[a], seqs[b], (int)seq_len[b], p->mismatch, p->indel, p->max_clip, &r) != 0) { rc = GP_ERR_NOMEM; return; }
            out[k].score = r.score; out[k].row_end = r.row_end; out[k].col_end = r.col_end; out[k].nclip = r.nclip;
            out[k].flags = (r.tb_row == 0 ? GP_FLAG_ROW0 : 0u) | (r.tb_col == 0 ? GP_FLAG_COL0 : 0u) | (r.bcontained ? GP_FLAG_CONTAINED : 0u);
        }
    };
    const unsigned T = n_pairs >= 16 ? std::max(1u, std::min(16u, std::thread::hardware_concurrency())) : 1u;   // realistic-size goldens
    std::vector<std::thread> th;
    for (unsigned t = 1; t < T; ++t) th.emplace_back(work);
    work();
    for (auto& x : th) x.join();
    return rc;
}
// The resident-table path of the drop-in (table up, quick check on the device, pairs up, launch, fetch), served from
// the same oracle: the packed codes are turned back into letters (A C G T N, then B D E F ... for any other byte:
// equal codes <-> equal bytes, and none of the others is a k-mer letter), the device filter is the host filter.
int gp_set_sequences(gp_ctx* c, const uint32_t* packed, size_t, const uint32_t* off, const uint32_t* len, uint32_t n, uint32_t)
{
    static const char letters[] = "ACGTNBDEFHIJKLMO";
    c->seqs.assign(n, std::string());
    for (uint32_t s = 0; s < n; ++s) {
        c->seqs[s].resize(len[s]);
        for (uint32_t i = 0; i < len[s]; ++i) c->seqs[s][i] = letters[(packed[off[s] + (i >> 3)] >> ((i & 7u) * 4u)) & 15u];
    }
    return GP_OK;
}
int gp_upload_sequences(gp_ctx* c, const char* const* seqs, const uint32_t* len, uint32_t n)
{
    c->seqs.assign(n, std::string());
    for (uint32_t s = 0; s < n; ++s) c->seqs[s].assign(seqs[s], len[s]);
    return GP_OK;
}
int gp_quick_check_device(gp_ctx* c, const uint32_t* gap_first, uint32_t n_gaps, int32_t k, uint8_t* hit, uint64_t hit_bytes)
{
    uint64_t pos = 0;
    for (uint32_t g = 0; g < n_gaps; ++g) {
        const uint32_t first = gap_first[g], n = gap_first[g + 1] - first;
        if (pos + (uint64_t)n * n > hit_bytes) return GP_ERR_INVALID;
        std::vector<const char*> nodes; std::vector<uint32_t> lens;
        for (uint32_t i = 0; i < n; ++i) { nodes.push_back(c->seqs[first + i].data()); lens.push_back((uint32_t)c->seqs[first + i].size()); }
        std::vector<gp_pair> cand((size_t)n * (n + 1) / 2 + 1);
        const int64_t np = gp_candidate_pairs(nodes.data(), lens.data(), n, k, cand.data(), cand.size());
        if (np < 0) return (int)np;
        memset(hit + pos, 0, (size_t)n * n);
        for (int64_t q = 0; q < np; ++q) hit[pos + (uint64_t)cand[q].row_seq * n + cand[q].col_seq] = 1;
        pos += (uint64_t)n * n;
    }
    return GP_OK;
}
// every ordered pair: hit(i, j) = "the ends of j occur in i", from the host filter on the two-node list {i, j}
int gp_quick_check_matrix(gp_ctx* c, const uint32_t* gap_first, uint32_t n_gaps, int32_t k, uint8_t* hit, uint64_t hit_bytes, int full_matrix)
{
    if (!full_matrix) return gp_quick_check_device(c, gap_first, n_gaps, k, hit, hit_bytes);
    uint64_t pos = 0;
    for (uint32_t g = 0; g < n_gaps; ++g) {
        const uint32_t first = gap_first[g], n = gap_first[g + 1] - first;
        if (pos + (uint64_t)n * n > hit_bytes) return GP_ERR_INVALID;
        memset(hit + pos, 0, (size_t)n * n);
        for (uint32_t i = 0; i < n; ++i)
            for (uint32_t j = 0; j < n; ++j) {
                const char* nodes[2] = {c->seqs[first + i].data(), c->seqs[first + j].data()};
                const uint32_t lens[2] = {(uint32_t)c->seqs[first + i].size(), (uint32_t)c->seqs[first + j].size()};
                gp_pair cand[4];
                const int64_t np = gp_candidate_pairs(nodes, lens, 2, k, cand, 4);
                if (np < 0) return (int)np;
                for (int64_t q = 0; q < np; ++q) if (cand[q].row_seq == 0 && cand[q].col_seq == 1) hit[pos + (uint64_t)i * n + j] = 1;
                if (i == j) for (int64_t q = 0; q < np; ++q) if (cand[q].row_seq == 0 && cand[q].col_seq == 0) hit[pos + (uint64_t)i * n + j] = 1;
            }
        pos += (uint64_t)n * n;
    }
    return GP_OK;
}
int gp_set_relax_launch_hook(gp_ctx* c, gp_launch_hook hook, void* user) { c->hook = hook; c->hook_user = user; return GP_OK; }
int gp_upload_pairs(gp_ctx* c, const gp_pair* pairs, uint64_t n, const gp_dp_params* p) { c->pairs.assign(pairs, pairs + n); c->params = *p; return GP_OK; }
int gp_launch_resident(gp_ctx*) { return GP_OK; }
int gp_fetch_results(gp_ctx* c, gp_result* out, uint64_t n)
{
    if (n != c->pairs.size()) return GP_ERR_INVALID;
    std::vector<const char*> ptr; std::vector<uint32_t> len;
    for (const std::string& s : c->seqs) { ptr.push_back(s.data()); len.push_back((uint32_t)s.size()); }
    return gp_overlap_batch(c, ptr.data(), len.data(), (uint32_t)ptr.size(), c->pairs.data(), n, &c->params, out);
}
// statistics the merger reads after a batch: nothing to report from the CPU shim
int gp_closed_form_stats(const gp_ctx*, uint64_t* pairs, uint64_t* cells) { if (pairs) *pairs = 0; if (cells) *cells = 0; return GP_OK; }
int gp_cert_stats(const gp_ctx*, uint64_t* a, uint64_t* b, uint64_t* c) { if (a) *a = 0; if (b) *b = 0; if (c) *c = 0; return GP_OK; }
int gp_last_team(const gp_ctx*) { return 0; }
int gp_reserve(gp_ctx*, uint64_t, uint32_t, uint64_t, uint64_t) { return GP_OK; }
// The device relax chain, served from the oracle step by step on the table of the last upload (parents precede children).
// GP_SHIM_NO_RELAX=1 answers GP_ERR_RANGE instead, which sends the merger to its step-by-step loop.
int gp_relax_chains(gp_ctx* c, const gp_relax_step* steps, uint64_t n, const gp_dp_params* p, gp_result* out, uint32_t* merged_len)
{
    if (getenv("GP_SHIM_NO_RELAX")) return GP_ERR_RANGE;
    if (c->hook) c->hook(c->hook_user);                   // "the kernel is enqueued"
    std::vector<std::string> merged(n);
    for (uint64_t k = 0; k < n; ++k) {
        const std::string& row = steps[k].parent < 0 ? c->seqs[steps[k].row_seq] : merged[steps[k].parent];
        const std::string& col = c->seqs[steps[k].col_seq];
        const char* ptr[2] = {row.data(), col.data()};
        const uint32_t len[2] = {(uint32_t)row.size(), (uint32_t)col.size()};
        const gp_pair pr{0, 1};
        if (int rc = gp_overlap_batch(c, ptr, len, 2, &pr, 1, p, out + k)) return rc;
        std::string m(row.size() + col.size() + 1, '\0');
        const int32_t l = gp_merged_concat(row.data(), (int32_t)row.size(), col.data(), (int32_t)col.size(), out + k, &m[0]);
        m.resize((size_t)l);
        merged[k] = std::move(m);
        merged_len[k] = (uint32_t)l;
    }
    return GP_OK;
}
int gp_relax_stats(const gp_ctx*, double* ms, uint64_t* a, uint64_t* b) { if (ms) *ms = 0; if (a) *a = 0; if (b) *b = 0; return GP_OK; }
int gp_quick_check_stats(const gp_ctx*, double* ms, uint64_t* b, uint32_t* i) { if (ms) *ms = 0; if (b) *b = 0; if (i) *i = 0; return GP_OK; }
int gp_last_timing(const gp_ctx*, double* out_ms, int n) { for (int i = 0; i < n; ++i) out_ms[i] = 0.0; return GP_OK; }
}
