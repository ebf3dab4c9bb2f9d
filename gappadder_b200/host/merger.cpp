#include "merger.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <map>
#include <numeric>
#include <thread>

#include "fasta.hpp"
#include "merge_graph.hpp"

namespace gpm {

namespace {

struct Chain {               // one path being merged left to right (FormMergedSeqFromPath, :1456-1515)
    int gap;
    std::vector<int> path;
    size_t next = 1;
    std::string merged;
};

struct GapState {
    std::vector<FastaRecord> contigs;
    std::vector<std::string> node_seq, node_name;   // [c0, c0_R, c1, c1_R, ...]  (:794-799)
    // Letters other than A C G T N.  The DP only tests bytes for equality (:1641) and pairs never cross gaps, so such a
    // gap's nodes go to the packer with its own letters renamed to a fixed placeholder set: the batch table's 16-code
    // alphabet is then a per-gap limit (11 letters besides A C G T N), not a limit on the union over the batch.
    // dp_seq is empty when the gap needs no renaming; outputs are always built from node_seq.
    std::vector<std::string> dp_seq;
    unsigned char letter_map[256];
    const std::string& dp(size_t v) const { return dp_seq.empty() ? node_seq[v] : dp_seq[v]; }
    uint32_t node_base = 0;                          // index of node 0 in the batch sequence table
    uint64_t pair_begin = 0, pair_end = 0;           // slice of the batch pair list
    bool dead = false;                               // fatal input error, nothing more to do
    bool acgt_only = true;                           // every node is pure A/C/G/T and at most 16382 bases (device relax chains)
    bool arranged = false;                           // candidate pairs were generated
    bool want_pairs = false;                         // ... by the quick check on the device (after the parallel phase)
    std::vector<gp_pair> cand;                       // ... with node indices local to the gap
    int cand_rc = 0;
    std::vector<std::vector<int>> paths;             // after RemoveDupRevCompPaths
    std::vector<std::string> merged;                 // one per path with size > 1
};

// MultiThreadQuickChecker::runMultiThreadChecker, ContigsCompactor.cpp:992-1038: how many thread ranges
// the reference manages to form.  Fewer than T => "Arrange error!" and an EMPTY candidate list.
long arranged_ranges(long n_nodes, int T)
{
    long long total = (long long)n_nodes * n_nodes;
    total -= n_nodes;
    total /= 2;
    const long avrg = (long)(total / T);
    long ncnt = 0, ranges = 0;
    for (long i = 0; i < n_nodes; ++i)
        for (long j = i; j < n_nodes; ++j) {
            ++ncnt;
            if (ncnt == avrg) { ++ranges; ncnt = 0; }
        }
    return ranges;
}

// Runs fn(g) for g in [0, G) on up to `threads` host threads (gaps are independent).
template <class F>
void for_each_gap(size_t G, unsigned threads, F fn)
{
    if (threads <= 1 || G < 2) { for (size_t g = 0; g < G; ++g) fn(g); return; }
    std::atomic<size_t> next(0);
    std::vector<std::thread> th;
    const unsigned T = (unsigned)std::min<size_t>(threads, G);
    for (unsigned t = 0; t < T; ++t)
        th.emplace_back([&] { for (size_t g = next.fetch_add(1); g < G; g = next.fetch_add(1)) fn(g); });
    for (auto& x : th) x.join();
}

} // namespace

uint64_t estimate_gap_cells(const std::vector<uint32_t>& contig_len)
{
    return gp_estimate_gap_cells(contig_len.data(), (uint32_t)contig_len.size());
}

std::vector<int> partition_gaps(const std::vector<uint64_t>& cost, int n_parts)
{
    std::vector<int32_t> part(cost.size(), 0);
    if (n_parts > 1) gp_partition_gaps(cost.data(), cost.size(), n_parts, part.data());
    return std::vector<int>(part.begin(), part.end());
}

namespace {
struct PairwisePhase {          // scope of one device phase of the pairwise class (see device_gate.hpp)
    DeviceGate* g;
    bool open;
    explicit PairwisePhase(DeviceGate* gate) : g(gate), open(gate != nullptr) { if (g) g->begin_pairwise(); }
    void end(bool relax_follows = false) { if (open) { g->end_pairwise(relax_follows); open = false; } }
    ~PairwisePhase() { end(false); }
};
struct RelaxHook { DeviceGate* g; bool fired; };
void relax_hook_fn(void* user)
{
    RelaxHook* h = (RelaxHook*)user;
    if (!h->fired) { h->fired = true; h->g->relax_launched(); }
}
} // namespace

int merge_gaps(gp_ctx* ctx, const MergeOptions& opt, std::vector<GapInput>& in, std::vector<GapOutput>& out,
               std::string& error, MergeTimings* timings, DeviceGate* gate)
{
    using clk = std::chrono::steady_clock;
    auto t_prev = clk::now();
    auto lap = [&](double MergeTimings::*slot) {
        const auto now = clk::now();
        if (timings) timings->*slot += std::chrono::duration<double, std::milli>(now - t_prev).count();
        t_prev = now;
    };
    auto t_mark = clk::now();
    auto mark = [&](const char* name) {                   // finer split of the phases above, for --stats (does not touch `lap`)
        const auto now = clk::now();
        if (timings) timings->detail[name] += std::chrono::duration<double, std::milli>(now - t_mark).count();
        t_mark = now;
    };
    const size_t G = in.size();
    out.assign(G, GapOutput());
    std::vector<GapState> st(G);

    // ---- scoring parameters as the reference ends up with them ---------------------------------
    gp_dp_params dp;
    dp.mismatch = (int)opt.score_mismatch;                       // `int matchScoreStep = scoreMismatch` (:1640)
    if (opt.score_indel != std::floor(opt.score_indel)) {
        error = "a fractional -i2 (indel score) is outside this implementation's integer contract";
        return GP_ERR_INVALID;
    }
    dp.indel = (int)opt.score_indel;
    const bool scan_runs = opt.max_overlap_clip_len >= 0;        // `for (c = 0; c <= maxOverlapClipLen; ...)` (:1679)
    // the quick check runs on the device when the library's device form covers the k-mer length (k <= 10)
    const bool device_qc = !opt.host_quick_check && opt.quick_kmer_len >= 1 && opt.quick_kmer_len <= GP_QC_MAX_K;
    dp.max_clip = scan_runs ? (int)std::floor(opt.max_overlap_clip_len) : 0;
    gp_thresholds thr{opt.max_frac_score_loss, opt.min_frac_overlap, opt.min_overlap_len, opt.min_overlap_len_with_scaffold};
    const int max_per_root = opt.max_count_contig_in_path > 0 ? opt.max_count_contig_in_path : 20;   // :33-34,:168

    // ---- read inputs, build nodes, candidate pairs ----------------------------------------------
    std::vector<const char*> seq_ptr;
    std::vector<uint32_t> seq_len;
    std::vector<gp_pair> pairs;
    const unsigned host_threads = opt.host_threads > 0 ? (unsigned)opt.host_threads : std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    for_each_gap(G, host_threads, [&](size_t g) {
        GapState& s = st[g];
        std::string fatal;
        bool ok;
        if (in[g].loaded) { s.contigs = std::move(in[g].records); ok = in[g].read_ok; fatal = in[g].fatal; }
        else ok = read_fasta(in[g].fasta_path, s.contigs, fatal);
        if (!ok) {
            out[g].stdout_text = "FATAL ERROR: " + fatal + "\n";          // THROW, fastareader.cpp:11-15
            out[g].exit_code = 1;
            s.dead = true;
            return;
        }
        const size_t nc = s.contigs.size();
        s.node_seq.reserve(2 * nc); s.node_name.reserve(2 * nc);
        for (FastaRecord& r : s.contigs) {                                // the record's letters MOVE into node 2c (no copy); contigs keep the names
            s.node_seq.push_back(std::move(r.seq));
            const std::string& fwd = s.node_seq.back();
            s.node_name.push_back(r.name);
            std::string rc(fwd.size(), 'N');
            if (!fwd.empty()) gp_revcomp(fwd.data(), (uint32_t)fwd.size(), &rc[0]);
            s.node_seq.push_back(std::move(rc));
            s.node_name.push_back(r.name + "_R");                         // :785-787
        }
        {   // letters besides A C G T N: rename per gap (see GapState::dp_seq).  One pass over the bases: which letters occur
            // (in order of first appearance) and whether everything is A/C/G/T.
            static const char placeholder[] = "BDEFHIJKLMO";
            struct Classes {
                unsigned char other[256], not_acgt[256];
                Classes()
                {
                    for (int b = 0; b < 256; ++b) { not_acgt[b] = !(b == 'A' || b == 'C' || b == 'G' || b == 'T'); other[b] = not_acgt[b] && b != 'N'; }
                }
            };
            static const Classes K;
            bool seen[256] = {false};
            int n_other = 0;
            for (int b = 0; b < 256; ++b) s.letter_map[b] = (unsigned char)b;
            for (size_t c = 0; c < nc; ++c) {
                const std::string& seq = s.node_seq[2 * c];
                if (seq.size() > 16382 || seq.empty()) s.acgt_only = false;
                unsigned any_not_acgt = 0;
                for (unsigned char ch : seq) {
                    any_not_acgt |= K.not_acgt[ch];
                    if (K.other[ch] && !seen[ch]) {
                        seen[ch] = true;
                        if (n_other < 11) s.letter_map[ch] = (unsigned char)placeholder[n_other];
                        ++n_other;
                    }
                }
                if (any_not_acgt) s.acgt_only = false;
            }
            if (n_other > 11) {
                out[g].error = in[g].fasta_path + ": " + std::to_string(n_other + 5) + " distinct sequence letters; this implementation "
                               "handles A C G T N plus 11 others per gap (the reference accepts any letter)";
                out[g].exit_code = 3;
                s.dead = true;
                return;
            }
            if (n_other > 0) {
                s.dp_seq = s.node_seq;
                for (std::string& q : s.dp_seq) for (char& ch : q) ch = (char)s.letter_map[(unsigned char)ch];
            }
        }
        // QuickCheckerContigsMatch::Init on every node (:843-849): a node shorter than k is fatal there
        bool too_short = false;
        for (const std::string& q : s.node_seq) if ((int)q.size() < opt.quick_kmer_len) too_short = true;
        if (too_short && !s.node_seq.empty()) {
            out[g].stdout_text = "FATAL ERROR: k-mer length is too large.\n";   // GetKmersList, :2060-2064
            out[g].exit_code = 1;
            s.dead = true;
            return;
        }
        // candidate pairs of this gap (the quick check), or the reference's "Arrange error!" line
        const long N = (long)s.node_seq.size();
        const int T = opt.num_threads;
        const long ranges = T > 0 ? arranged_ranges(N, T) : 0;
        if (T <= 0 || ranges < T) {
            out[g].stdout_text += "Arrange error! " + std::to_string(ranges) + " " + std::to_string(T) + "\n";   // :1034-1038
            return;
        }
        s.arranged = true;
        if (!scan_runs) return;        // no scan => every Evaluate is rejected (:1674-1722): no edges
        if (device_qc && N <= GP_QC_MAX_NODES) { s.want_pairs = true; return; }   // quick check on the device, below
        std::vector<const char*> nodes;
        std::vector<uint32_t> lens;
        for (const std::string& q : s.node_seq) { nodes.push_back(q.data()); lens.push_back((uint32_t)q.size()); }
        const uint64_t cap = (uint64_t)N * (N + 1) / 2;
        s.cand.resize(cap);
        const int64_t np = gp_candidate_pairs(nodes.data(), lens.data(), (uint32_t)N, opt.quick_kmer_len, s.cand.data(), cap);
        if (np < 0) { s.cand_rc = (int)np; s.cand.clear(); return; }
        s.cand.resize((size_t)np);
    });
    mark("read.nodes");
    // serial: one sequence table and one pair list for the batch (node_seq vectors no longer grow).  Gaps whose
    // candidate filter runs on the device come first: gp_quick_check_device takes contiguous node ranges.  A gap
    // beyond the device filter's limits (GP_QC_MAX_NODES) was filtered on the host above -- that gap only.
    std::vector<size_t> order;
    for (size_t g = 0; g < G; ++g) if (!st[g].dead && st[g].want_pairs) order.push_back(g);
    const size_t n_dev = order.size();
    for (size_t g = 0; g < G; ++g) if (!st[g].dead && !st[g].want_pairs) order.push_back(g);
    std::vector<uint32_t> gap_first;                 // device quick check: node range of every device-filtered gap
    for (size_t q = 0; q < order.size(); ++q) {
        GapState& s = st[order[q]];
        if (s.cand_rc != 0) { error = "gp_candidate_pairs failed"; return s.cand_rc; }
        s.node_base = (uint32_t)seq_ptr.size();
        if (q < n_dev) gap_first.push_back(s.node_base);
        if (q == n_dev) gap_first.push_back(s.node_base);
        for (size_t v = 0; v < s.node_seq.size(); ++v) { seq_ptr.push_back(s.dp(v).data()); seq_len.push_back((uint32_t)s.node_seq[v].size()); }
    }
    if (order.size() == n_dev) gap_first.push_back((uint32_t)seq_ptr.size());

    // Quick check on the device: the nodes go to HBM first (they are needed there for the DP anyway), the filter
    // runs on the packed table (gp_quick_check_device), the pair list is read off the hit matrices in the host
    // filter's order (row by row, j >= i).
    bool table_resident = false;
    if (n_dev > 0) {
        mark("read.table");
        const uint32_t n_seq = (uint32_t)seq_ptr.size();
        // Not gated: the copy runs beside any kernel, and the filter kernel (0.3 ms) simply queues behind whatever persistent
        // kernel another merger has on the SMs -- by the time that one drains, this chunk's pair list is one step away.
        int rc = gp_upload_sequences(ctx, seq_ptr.data(), seq_len.data(), n_seq);   // packs into the context's pinned buffer
        if (rc != GP_OK) { error = gp_last_error(ctx); return rc; }
        table_resident = true;
        mark("read.upload");
        std::vector<uint64_t> hoff(n_dev + 1, 0);
        for (size_t q = 0; q < n_dev; ++q) { const uint64_t n = gap_first[q + 1] - gap_first[q]; hoff[q + 1] = hoff[q] + n * n; }
        std::vector<uint8_t> hit(hoff.back() ? hoff.back() : 1);
        rc = gp_quick_check_device(ctx, gap_first.data(), (uint32_t)n_dev, opt.quick_kmer_len, hit.data(), hit.size());
        if (rc != GP_OK) { error = gp_last_error(ctx); return rc; }
        mark("read.quick_check");
        if (timings) {
            double ms = 0; uint64_t bases = 0; uint32_t items = 0;
            gp_quick_check_stats(ctx, &ms, &bases, &items);
            timings->qc_kernel_ms += ms; timings->qc_bases += bases; timings->qc_items += items;
        }
        // the pair list off the hit matrices, row by row (the host filter's order): counts, offsets, then every gap fills its slice
        std::vector<uint64_t> cnt(n_dev + 1, 0);
        for_each_gap(n_dev, host_threads, [&](size_t q) {
            const uint64_t n = gap_first[q + 1] - gap_first[q];
            const uint8_t* h = hit.data() + hoff[q];
            uint64_t k = 0;
            for (uint64_t i = 0; i < n; ++i) for (uint64_t j = i; j < n; ++j) k += h[i * n + j] != 0;
            cnt[q + 1] = k;
        });
        for (size_t q = 0; q < n_dev; ++q) cnt[q + 1] += cnt[q];
        pairs.resize(cnt[n_dev]);
        for_each_gap(n_dev, host_threads, [&](size_t q) {
            GapState& s = st[order[q]];
            s.pair_begin = cnt[q]; s.pair_end = cnt[q + 1];
            const uint32_t n = gap_first[q + 1] - gap_first[q];
            const uint8_t* h = hit.data() + hoff[q];
            gp_pair* w = pairs.data() + cnt[q];
            for (uint32_t i = 0; i < n; ++i)
                for (uint32_t j = i; j < n; ++j)
                    if (h[(size_t)i * n + j]) *w++ = gp_pair{i + s.node_base, j + s.node_base};
        });
    }
    for (size_t q = n_dev; q < order.size(); ++q) {                       // host-filtered gaps
        GapState& s = st[order[q]];
        s.pair_begin = pairs.size();
        for (const gp_pair& c : s.cand) pairs.push_back(gp_pair{c.row_seq + s.node_base, c.col_seq + s.node_base});
        s.pair_end = pairs.size();
        std::vector<gp_pair>().swap(s.cand);
    }

    mark("read.pairs");
    lap(&MergeTimings::read_ms);
    // ---- pairwise phase: one batch for all gaps (replaces runMultiThreadMergeV2, :696-721) ------
    std::vector<gp_result> res(pairs.size());
    bool relax_announced = false;                      // end_pairwise(true) was called: begin_relax or cancel_relax must follow
    if (!pairs.empty()) {
        int rc;
        PairwisePhase dl(gate);
        mark("pairwise.wait_device");
        if (table_resident) {                                            // the table is in HBM already: pairs up, kernels, results down
            rc = gp_upload_pairs(ctx, pairs.data(), pairs.size(), &dp);
            mark("pairwise.upload_pairs");
            if (rc == GP_OK) rc = gp_launch_resident(ctx);
            if (rc == GP_OK) rc = gp_fetch_results(ctx, res.data(), pairs.size());
        } else {
            rc = gp_overlap_batch(ctx, seq_ptr.data(), seq_len.data(), (uint32_t)seq_ptr.size(), pairs.data(), pairs.size(), &dp, res.data());
        }
        if (rc != GP_OK) { error = gp_last_error(ctx); return rc; }
        if (timings) {
            uint64_t cp = 0, cc = 0;
            gp_closed_form_stats(ctx, &cp, &cc);
            timings->closed_pairs += cp; timings->closed_cells += cc;
        }
        relax_announced = gate != nullptr && !opt.host_relax;
        dl.end(relax_announced);                                         // the relax launch comes next: nobody's pairwise phase before it
    }
    struct CancelRelax {                                                   // whatever path leaves this function
        DeviceGate* g; bool* on;
        ~CancelRelax() { if (g && *on) { g->cancel_relax(); *on = false; } }
    } cancel_relax{gate, &relax_announced};

    mark("pairwise.kernels_fetch");
    lap(&MergeTimings::pairwise_ms);
    // ---- edges, graph, paths (threadMergeContigV2 :652-685, addEdges :724-770, :898-931) --------
    std::vector<Chain> chains;
    std::vector<std::vector<Chain>> gap_chains(G);
    for_each_gap(G, host_threads, [&](size_t g) {
        GapState& s = st[g];
        if (s.dead) return;
        const int N = (int)s.node_seq.size();
        OverlapGraph graph(N);
        for (uint64_t k = s.pair_begin; k < s.pair_end; ++k) {
            const int i = (int)(pairs[k].row_seq - s.node_base), j = (int)(pairs[k].col_seq - s.node_base);
            const int32_t l1 = (int32_t)s.node_seq[i].size(), l2 = (int32_t)s.node_seq[j].size();
            const gp_result& r = res[k];
            out[g].pair_cells += (uint64_t)l1 * l2;
            const int sig = gp_is_score_significant(&thr, r.score, l1, l2, r.row_end, r.col_end, r.nclip);
            if (sig != 2) continue;                                        // OVERLAP_LARGER_MINLEN only (:653,:673)
            if (gp_is_containment(l1, l2, &r)) continue;                   // :673
            const double len = -1.0 * gp_overlap_size(l1, l2, &r);          // :675
            if (r.row_end + r.nclip != l1) graph.add_edge(j, i, len);      // MODE_2_1: edge j -> i (:656-662,:758-763)
            else graph.add_edge(i, j, len);                                // MODE_1_2: edge i -> j
        }
        out[g].n_pairs = (uint32_t)(s.pair_end - s.pair_begin);
        if (opt.build_gml) out[g].gml_text = graph.gml(s.node_name);       // :898-899
        s.paths = remove_revcomp_duplicates(graph.find_paths(max_per_root));   // :907,:926
        for (const std::vector<int>& p : s.paths) {
            if (p.size() > 1) {
                Chain c;
                c.gap = (int)g; c.path = p; c.merged = s.node_seq[p[0]];
                gap_chains[g].push_back(std::move(c));
            }
        }
    });
    for (size_t g = 0; g < G; ++g) for (Chain& c : gap_chains[g]) chains.push_back(std::move(c));

    mark("graph");
    lap(&MergeTimings::graph_ms);
    // ---- relax chains on the device: the whole forest of steps in one launch (gp_relax_chains) ------------------
    // A step = Evaluate(merged contig so far, next node of the path) + SetMergedStringConcat (:1463-1513).  Chains of one
    // gap whose paths share a prefix share those steps (the reference runs them once per path, with the same result), so
    // the steps of a gap form a trie over its paths.  Gaps outside the device entry point's domain (letters other than
    // A C G T, other scores, nodes beyond 16382 bases) and gaps with a step no certificate resolves take the step-by-step
    // loop below (exact kernels through gp_overlap_batch).
    if (!chains.empty() && !opt.host_relax) {
        std::vector<gp_relax_step> steps;
        std::vector<std::vector<int32_t>> chain_steps(chains.size());     // per chain: step index of path position 1, 2, ...
        std::vector<char> gap_on_device(G, 0);
        for (size_t g = 0; g < G; ++g) gap_on_device[g] = !st[g].dead && st[g].acgt_only && dp.mismatch == -2 && dp.indel == -2;
        {
            // the trie of every gap on the host threads (step indices local to the gap), then one list with offsets
            std::vector<size_t> first_chain;                               // chains are grouped by gap
            for (size_t c = 0; c < chains.size(); ++c) if (c == 0 || chains[c].gap != chains[c - 1].gap) first_chain.push_back(c);
            first_chain.push_back(chains.size());
            std::vector<std::vector<gp_relax_step>> gap_steps(first_chain.size() - 1);
            for_each_gap(first_chain.size() - 1, host_threads, [&](size_t q) {
                const size_t c0 = first_chain[q], c1 = first_chain[q + 1];
                const int g = chains[c0].gap;
                if (!gap_on_device[g]) return;
                std::vector<gp_relax_step>& mine = gap_steps[q];
                std::map<std::vector<int>, int32_t> seen;                 // path prefix -> step
                for (size_t c = c0; c < c1; ++c) {
                    const std::vector<int>& p = chains[c].path;
                    int32_t parent = -1;
                    std::vector<int> key(1, p[0]);
                    for (size_t k = 1; k < p.size(); ++k) {
                        key.push_back(p[k]);
                        auto it = seen.find(key);
                        if (it == seen.end()) {
                            it = seen.emplace(key, (int32_t)mine.size()).first;
                            mine.push_back(gp_relax_step{parent, st[g].node_base + (uint32_t)p[0], st[g].node_base + (uint32_t)p[k]});
                        }
                        chain_steps[c].push_back(it->second);
                        parent = it->second;
                    }
                }
            });
            for (size_t q = 0; q + 1 < first_chain.size(); ++q) {
                const int32_t base = (int32_t)steps.size();
                for (gp_relax_step stp : gap_steps[q]) { if (stp.parent >= 0) stp.parent += base; steps.push_back(stp); }
                size_t uses = 0;
                for (size_t c = first_chain[q]; c < first_chain[q + 1]; ++c) { for (int32_t& k : chain_steps[c]) k += base; uses += chain_steps[c].size(); }
                if (timings) timings->relax_shared_pairs += uses - gap_steps[q].size();
            }
        }
        mark("relax.trie");
        if (!steps.empty() && getenv("GP_RELAX_DEBUG")) {                  // per depth: steps, cells (upper bound: rows = sum of the prefix's nodes)
            std::vector<uint32_t> depth(steps.size()), bound(steps.size());
            std::map<uint32_t, std::pair<uint64_t, uint64_t>> per;
            for (size_t k = 0; k < steps.size(); ++k) {
                const uint32_t cl = seq_len[steps[k].col_seq];
                const uint32_t rl = steps[k].parent < 0 ? seq_len[steps[k].row_seq] : bound[steps[k].parent];
                depth[k] = steps[k].parent < 0 ? 0 : depth[steps[k].parent] + 1;
                bound[k] = rl + cl;
                per[depth[k]].first += 1; per[depth[k]].second += (uint64_t)rl * cl;
            }
            for (const auto& kv : per) fprintf(stderr, "relax depth %u: %llu steps, %.3f Gcells (bound)\n", kv.first, (unsigned long long)kv.second.first, kv.second.second / 1e9);
        }
        if (!steps.empty()) {
            std::vector<gp_result> rr(steps.size());
            std::vector<uint32_t> mlen(steps.size());
            int rc;
            if (gate) {
                gate->begin_relax();
                mark("relax.wait_device");
                RelaxHook hook{gate, !relax_announced};                    // fires once, when the kernel is enqueued
                gp_set_relax_launch_hook(ctx, relax_hook_fn, &hook);
                rc = gp_relax_chains(ctx, steps.data(), steps.size(), &dp, rr.data(), mlen.data());
                gp_set_relax_launch_hook(ctx, nullptr, nullptr);
                if (!hook.fired) gate->relax_launched();                   // the call failed before launching
                relax_announced = false;
                gate->end_relax();
            } else {
                rc = gp_relax_chains(ctx, steps.data(), steps.size(), &dp, rr.data(), mlen.data());
            }
            mark("relax.device_call");
            if (rc == GP_ERR_RANGE) {
                for (size_t g = 0; g < G; ++g) gap_on_device[g] = 0;       // outside the entry point's domain after all: step by step
            } else if (rc != GP_OK) {
                error = gp_last_error(ctx);
                return rc;
            } else {
                if (timings) {
                    double ms = 0; uint64_t sp2 = 0, un = 0;
                    gp_relax_stats(ctx, &ms, &sp2, &un);
                    timings->relax_device_ms += ms; timings->relax_pairs += steps.size(); timings->relax_second_passes += sp2;
                    timings->relax_exact_retries += un; timings->relax_steps += 1; timings->relax_team_steps += 1;
                }
                for (size_t c = 0; c < chains.size(); ++c)
                    for (int32_t k : chain_steps[c]) if (rr[k].flags & GP_FLAG_UNRESOLVED) gap_on_device[chains[c].gap] = 0;
                // merged contigs from the original letters, chain by chain; consecutive chains of a gap share long prefixes
                // (the paths come out of a sorted set), so a stack of the previous chain's intermediate contigs is reused
                std::vector<size_t> first_chain_of_gap;
                for (size_t c = 0; c < chains.size(); ++c) if (c == 0 || chains[c].gap != chains[c - 1].gap) first_chain_of_gap.push_back(c);
                first_chain_of_gap.push_back(chains.size());
                std::atomic<int> bad(0);
                for_each_gap(first_chain_of_gap.size() - 1, host_threads, [&](size_t q) {
                    const size_t c0 = first_chain_of_gap[q], c1 = first_chain_of_gap[q + 1];
                    const int g = chains[c0].gap;
                    if (!gap_on_device[g]) return;
                    std::vector<std::string> stack;                        // stack[k] = merged contig after path position k+1
                    std::vector<int32_t> stack_step;
                    for (size_t c = c0; c < c1; ++c) {
                        Chain& ch = chains[c];
                        const std::vector<int32_t>& cs = chain_steps[c];
                        size_t keep = 0;
                        while (keep < cs.size() && keep < stack_step.size() && stack_step[keep] == cs[keep]) ++keep;
                        stack.resize(keep); stack_step.resize(keep);
                        for (size_t k = keep; k < cs.size(); ++k) {
                            const std::string& prev = k == 0 ? st[g].node_seq[ch.path[0]] : stack[k - 1];
                            const std::string& nodeseq = st[g].node_seq[ch.path[k + 1]];
                            std::string m(prev.size() + nodeseq.size() + 1, '\0');
                            const int32_t len = gp_merged_concat(prev.data(), (int32_t)prev.size(), nodeseq.data(), (int32_t)nodeseq.size(), &rr[cs[k]], &m[0]);
                            if ((uint32_t)len != mlen[cs[k]]) bad = 1;     // the device built a different contig: never expected
                            m.resize((size_t)len);
                            stack.push_back(std::move(m));
                            stack_step.push_back(cs[k]);
                        }
                        for (size_t k = 0; k < cs.size(); ++k) {
                            const size_t rows = k == 0 ? st[g].node_seq[ch.path[0]].size() : stack[k - 1].size();
                            out[g].relax_cells += (uint64_t)rows * st[g].node_seq[ch.path[k + 1]].size();
                            out[g].n_relax += 1;
                        }
                        if (getenv("GP_RELAX_DEBUG")) {
                            uint64_t cc = 0;
                            for (size_t k = 0; k < cs.size(); ++k) cc += (uint64_t)(k == 0 ? st[g].node_seq[ch.path[0]].size() : stack[k - 1].size()) * st[g].node_seq[ch.path[k + 1]].size();
                            fprintf(stderr, "relax chain gap %d: %zu steps, %.1f Mcells, final %zu bases\n", g, cs.size(), cc / 1e6, cs.empty() ? (size_t)0 : stack.back().size());
                        }
                        ch.merged = cs.empty() ? ch.merged : stack.back();
                        ch.next = ch.path.size();
                    }
                });
                mark("relax.strings");
                if (bad) { error = "internal: merged contig lengths of the device relax chain and the host epilogue differ"; return GP_ERR_INVALID; }
                // a gap with an unresolved step starts over in the loop below
                for (Chain& ch : chains) if (!gap_on_device[ch.gap]) { ch.next = 1; ch.merged = st[ch.gap].node_seq[ch.path[0]]; }
            }
        }
    }
    // no relax launch after all (no chain, host relax, nothing in the device entry point's domain): the other mergers'
    // pairwise phases need not wait any longer
    if (gate && relax_announced) { gate->cancel_relax(); relax_announced = false; }
    // ---- relax chains, step by step: step k of every remaining chain in one batch (gaps the device path did not take) ------
    for (;;) {
        std::vector<size_t> active;
        for (size_t c = 0; c < chains.size(); ++c) if (chains[c].next < chains[c].path.size()) active.push_back(c);
        if (active.empty()) break;
        if (timings) ++timings->relax_steps;
        std::vector<const char*> sp;
        std::vector<uint32_t> sl;
        std::vector<gp_pair> pp;
        // Chains of one gap whose paths share the prefix path[0..next] hold the same merged string and meet the same
        // node: one Evaluate serves them all (the reference runs it once per path, :1463-1513, with the same result).
        std::deque<std::string> renamed;                                  // keeps the renamed merged strings alive
        std::vector<size_t> rep(active.size());                           // active chain -> index into pp
        std::map<std::pair<size_t, std::vector<int>>, size_t> seen;
        for (size_t a = 0; a < active.size(); ++a) {
            const Chain& ch = chains[active[a]];
            const std::string& nodeseq = st[ch.gap].node_seq[ch.path[ch.next]];
            std::pair<size_t, std::vector<int>> key(ch.gap, std::vector<int>(ch.path.begin(), ch.path.begin() + ch.next + 1));
            auto it = seen.find(key);
            if (it != seen.end()) {
                rep[a] = it->second;
                if (timings) { ++timings->relax_shared_pairs; timings->relax_shared_cells += (uint64_t)ch.merged.size() * nodeseq.size(); }
                continue;
            }
            rep[a] = pp.size();
            seen.emplace(std::move(key), pp.size());
            pp.push_back(gp_pair{(uint32_t)sp.size(), (uint32_t)sp.size() + 1});
            const GapState& gs = st[ch.gap];
            if (gs.dp_seq.empty()) {
                sp.push_back(ch.merged.data());
                sp.push_back(nodeseq.data());
            } else {                                                       // renamed letters (GapState::dp_seq)
                renamed.emplace_back(ch.merged);
                for (char& b : renamed.back()) b = (char)gs.letter_map[(unsigned char)b];
                sp.push_back(renamed.back().data());
                sp.push_back(gs.dp_seq[ch.path[ch.next]].data());
            }
            sl.push_back((uint32_t)ch.merged.size());
            sl.push_back((uint32_t)nodeseq.size());
        }
        std::vector<gp_result> rr(pp.size());
        int rc;
        {
            if (gate && relax_announced) { gate->cancel_relax(); relax_announced = false; }   // no device relax launch for these chains
            PairwisePhase dl(gate);
            rc = gp_overlap_batch(ctx, sp.data(), sl.data(), (uint32_t)sp.size(), pp.data(), pp.size(), &dp, rr.data());
        }
        if (rc != GP_OK) { error = gp_last_error(ctx); return rc; }
        if (timings) {
            uint64_t c16 = 0, sp = 0, er = 0;
            gp_cert_stats(ctx, &c16, &sp, &er);
            timings->relax_pairs += pp.size(); timings->relax_second_passes += sp; timings->relax_exact_retries += er;
            timings->relax_team_steps += gp_last_team(ctx) ? 1 : 0;
            double tm[GP_TIMING_SLOTS] = {0};
            gp_last_timing(ctx, tm, GP_TIMING_SLOTS);
            timings->relax_device_ms += tm[2];
            timings->relax_call_ms += tm[3];
            timings->relax_pack_ms += tm[0] + tm[1];
        }
        for (size_t a = 0; a < active.size(); ++a) {                       // per-gap counters: serial
            const Chain& ch = chains[active[a]];
            out[ch.gap].relax_cells += (uint64_t)ch.merged.size() * st[ch.gap].node_seq[ch.path[ch.next]].size();
            out[ch.gap].n_relax += 1;
        }
        for_each_gap(active.size(), active.size() >= 64 ? host_threads : 1, [&](size_t a) {   // the merged strings: host threads
            Chain& ch = chains[active[a]];
            const std::string& nodeseq = st[ch.gap].node_seq[ch.path[ch.next]];
            std::string m(ch.merged.size() + nodeseq.size() + 1, '\0');
            const int32_t len = gp_merged_concat(ch.merged.data(), (int32_t)ch.merged.size(), nodeseq.data(),
                                                 (int32_t)nodeseq.size(), &rr[rep[a]], &m[0]);   // ccAct.GetMerged() (:1512)
            m.resize((size_t)len);
            ch.merged.swap(m);
            ++ch.next;
        });
    }
    for (Chain& ch : chains) st[ch.gap].merged.push_back(std::move(ch.merged));
    lap(&MergeTimings::relax_ms);
    if (timings) timings->relax_host_ms = timings->relax_ms - timings->relax_device_ms;

    // ---- output (ContigsCompactor.cpp:945-971, CM/main.cpp:281-288) -------------------------------
    for_each_gap(G, host_threads, [&](size_t g) {
        GapState& s = st[g];
        if (s.dead) return;
        std::map<std::string, std::string> name_to_path;                  // mapNewContigNameToPath
        int next_id = 1;                                                   // static contigNumNext, one process per gap
        size_t mi = 0;
        for (const std::vector<int>& p : s.paths) {
            if (p.size() <= 1) continue;
            std::string sub;
            for (int v : p) { sub += " "; sub += s.node_name[v]; }
            const std::string name = "NEW_CONTIG_MERGE_" + std::to_string(next_id++);
            append_fasta(out[g].stdout_text, name, s.merged[mi++], 60);    // printFasta(cout): DEFAULT_LINE_LENGTH
            name_to_path.insert(std::make_pair(name, sub));
        }
        for (size_t c = 0; c < s.contigs.size(); ++c) append_fasta(out[g].stdout_text, s.contigs[c].name, s.node_seq[2 * c], opt.line_length);
        for (const auto& kv : name_to_path) out[g].info_text += kv.first + "  " + kv.second + "\n";   // :1555-1560
        out[g].wrote_info = true;
    });
    lap(&MergeTimings::output_ms);
    return GP_OK;
}

} // namespace gpm
