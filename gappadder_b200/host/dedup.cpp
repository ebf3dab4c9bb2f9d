#include "dedup.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>

namespace gpm {

namespace {

struct RawRecord {               // one FASTA record as rmCotigs sees it (TERefiner/refiner.cpp:418-446)
    std::string header;          // the '>' line without its '\n'
    std::string body;            // the sequence lines, each with its '\n' (a last line without one gains it, as in rmCotigs)
    std::string name;            // header up to the first blank: what samtools faidx / BWA call the contig
    std::string seq;             // letters for the DP: upper case, anything but A C G T becomes N
};

bool read_raw(const std::string& path, std::string& data, std::vector<RawRecord>& recs)
{
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof buf, f)) > 0) data.append(buf, got);
    fclose(f);
    size_t i = 0;
    const size_t n = data.size();
    bool have = false;
    while (i < n) {
        size_t e = data.find('\n', i);
        const bool ended = e != std::string::npos;
        if (!ended) e = n;
        const std::string line = data.substr(i, e - i);
        i = ended ? e + 1 : n;
        if (!line.empty() && line[0] == '>') {
            recs.emplace_back();
            have = true;
            recs.back().header = line;
            size_t b = 1;
            while (b < line.size() && line[b] != ' ' && line[b] != '\t' && line[b] != '\r') ++b;
            recs.back().name = line.substr(1, b - 1);
            continue;
        }
        if (!have) continue;                                     // text before the first header: rmCotigs drops it
        recs.back().body += line;
        recs.back().body += '\n';
        for (char ch : line) {
            if (ch == '\r' || ch == ' ' || ch == '\t') continue;
            const char u = (char)std::toupper((unsigned char)ch);
            recs.back().seq += (u == 'A' || u == 'C' || u == 'G' || u == 'T') ? u : 'N';
        }
    }
    return true;
}

struct SetState {
    std::string data;
    std::vector<RawRecord> recs;
    std::vector<uint8_t> keep;                 // after the unique-name step
    std::vector<uint32_t> kept;                // indices of the kept records: the contigs of the -P step
    std::vector<std::string> node_seq;         // [c0, c0_R, c1, c1_R, ...] of the kept contigs
    uint32_t node_base = 0;
    uint64_t pair_begin = 0, pair_end = 0;
    bool dead = false;
};

template <class F>
void parallel_for(size_t n, F fn)
{
    const unsigned T = (unsigned)std::min<size_t>(std::max(1u, std::min(16u, std::thread::hardware_concurrency())), n);
    if (T <= 1) { for (size_t i = 0; i < n; ++i) fn(i); return; }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; ++t) th.emplace_back([&, t] { for (size_t i = t; i < n; i += T) fn(i); });
    for (auto& x : th) x.join();
}

} // namespace

int dedup_sets(gp_ctx* ctx, const MergeOptions& opt, const std::vector<DedupInput>& in, std::vector<DedupOutput>& out,
               std::string& error, DedupTimings* timings)
{
    using clk = std::chrono::steady_clock;
    auto t0 = clk::now();
    auto lap = [&](double DedupTimings::*slot) {
        const auto now = clk::now();
        if (timings) timings->*slot += std::chrono::duration<double, std::milli>(now - t0).count();
        t0 = now;
    };
    const size_t S = in.size();
    out.assign(S, DedupOutput());
    std::vector<SetState> st(S);
    gp_dp_params dp;
    dp.mismatch = (int)opt.score_mismatch;
    if (opt.score_indel != std::floor(opt.score_indel)) { error = "a fractional -i2 (indel score) is outside this implementation's integer contract"; return GP_ERR_INVALID; }
    dp.indel = (int)opt.score_indel;
    dp.max_clip = opt.max_overlap_clip_len >= 0 ? (int)std::floor(opt.max_overlap_clip_len) : 0;
    if (opt.quick_kmer_len < 1 || opt.quick_kmer_len > GP_QC_MAX_K) { error = "the dedup stage needs 1 <= -k <= 10 (the device quick check)"; return GP_ERR_RANGE; }

    // ---- read, unique names (TERefiner_1 -U), nodes ---------------------------------------------------------------
    parallel_for(S, [&](size_t s) {
        SetState& x = st[s];
        if (!read_raw(in[s].fasta_path, x.data, x.recs)) { out[s].error = "cannot open " + in[s].fasta_path; x.dead = true; return; }
        const uint32_t n = (uint32_t)x.recs.size();
        out[s].n_contigs = n;
        std::vector<const char*> names(n);
        for (uint32_t i = 0; i < n; ++i) names[i] = x.recs[i].name.c_str();
        x.keep.assign(n ? n : 1, 1);
        gp_dedup_unique_names(names.data(), n, x.keep.data());
        for (uint32_t i = 0; i < n; ++i) if (x.keep[i]) x.kept.push_back(i);
        out[s].n_unique = (uint32_t)x.kept.size();
        if (2 * x.kept.size() > GP_QC_MAX_NODES) {
            out[s].error = in[s].fasta_path + ": more than " + std::to_string(GP_QC_MAX_NODES / 2) + " contigs in one set";
            x.dead = true;
            return;
        }
        for (uint32_t c : x.kept) {
            const std::string& q = x.recs[c].seq;
            x.node_seq.push_back(q);
            std::string rc(q.size(), 'N');
            if (!q.empty()) gp_revcomp(q.data(), (uint32_t)q.size(), &rc[0]);
            x.node_seq.push_back(rc);
        }
    });
    std::vector<const char*> seq_ptr;
    std::vector<uint32_t> seq_len, gap_first;
    std::vector<size_t> live;
    for (size_t s = 0; s < S; ++s) {
        if (st[s].dead || st[s].node_seq.empty()) continue;
        live.push_back(s);
        st[s].node_base = (uint32_t)seq_ptr.size();
        gap_first.push_back(st[s].node_base);
        for (const std::string& q : st[s].node_seq) { seq_ptr.push_back(q.data()); seq_len.push_back((uint32_t)q.size()); }
    }
    gap_first.push_back((uint32_t)seq_ptr.size());
    lap(&DedupTimings::read_ms);

    // ---- candidates (every ordered pair whose query ends occur in the reference) and the overlap DP -----------------
    std::vector<gp_pair> pairs;
    std::vector<gp_result> res;
    if (!live.empty()) {
        int rc = gp_upload_sequences(ctx, seq_ptr.data(), seq_len.data(), (uint32_t)seq_ptr.size());
        if (rc != GP_OK) { error = gp_last_error(ctx); return rc; }
        std::vector<uint64_t> hoff(live.size() + 1, 0);
        for (size_t q = 0; q < live.size(); ++q) { const uint64_t n = gap_first[q + 1] - gap_first[q]; hoff[q + 1] = hoff[q] + n * n; }
        std::vector<uint8_t> hit(hoff.back() ? hoff.back() : 1);
        rc = gp_quick_check_matrix(ctx, gap_first.data(), (uint32_t)live.size(), opt.quick_kmer_len, hit.data(), hit.size(), 1);
        if (rc != GP_OK) { error = gp_last_error(ctx); return rc; }
        if (timings) { double ms = 0; uint64_t b = 0; uint32_t it = 0; gp_quick_check_stats(ctx, &ms, &b, &it); timings->qc_kernel_ms += ms; }
        for (size_t q = 0; q < live.size(); ++q) {
            SetState& x = st[live[q]];
            const uint32_t n = gap_first[q + 1] - gap_first[q];
            const uint8_t* h = hit.data() + hoff[q];
            x.pair_begin = pairs.size();
            // query = the forward strand of contig j/2 (rows), reference = node i of another contig, either strand (columns);
            // the reverse-complement query against r is the mirror image of the forward query against r's reverse complement
            for (uint32_t j = 0; j < n; j += 2)
                for (uint32_t i = 0; i < n; ++i)
                    if ((i >> 1) != (j >> 1) && h[(size_t)i * n + j] && seq_len[x.node_base + i] > 0 && seq_len[x.node_base + j] > 0)
                        pairs.push_back(gp_pair{x.node_base + j, x.node_base + i});
            x.pair_end = pairs.size();
        }
        res.resize(pairs.size());
        if (!pairs.empty()) {
            rc = gp_upload_pairs(ctx, pairs.data(), pairs.size(), &dp);
            if (rc == GP_OK) rc = gp_launch_resident(ctx);
            if (rc == GP_OK) rc = gp_fetch_results(ctx, res.data(), pairs.size());
            if (rc != GP_OK) { error = gp_last_error(ctx); return rc; }
        }
    }
    lap(&DedupTimings::device_ms);

    // ---- records, the -P rule, output (rmCotigs) ----------------------------------------------------------------
    parallel_for(S, [&](size_t s) {
        SetState& x = st[s];
        if (x.dead) return;
        const uint32_t nk = (uint32_t)x.kept.size();
        std::vector<gp_dedup_record> recs;
        std::vector<const char*> names(nk);
        std::vector<uint32_t> lens(nk);
        for (uint32_t c = 0; c < nk; ++c) { names[c] = x.recs[x.kept[c]].name.c_str(); lens[c] = (uint32_t)x.recs[x.kept[c]].seq.size(); }
        for (uint32_t c = 0; c < nk; ++c)                            // bwa mem -a reports every contig on itself
            if (lens[c]) recs.push_back(gp_dedup_record{c, c, 1u, lens[c], 0u});
        for (uint64_t k = x.pair_begin; k < x.pair_end; ++k) {
            const uint32_t q = (pairs[k].row_seq - x.node_base) >> 1, r = (pairs[k].col_seq - x.node_base) >> 1;
            gp_dedup_record two[2];
            const int n = gp_dedup_records(q, r, (int32_t)lens[q], (int32_t)lens[r], &res[k], opt.max_frac_score_loss, two);
            for (int t = 0; t < n; ++t) recs.push_back(two[t]);
            out[s].pair_cells += (uint64_t)lens[q] * lens[r];
        }
        out[s].n_pairs = (uint32_t)(x.pair_end - x.pair_begin);
        out[s].n_records = (uint32_t)recs.size();
        std::vector<uint8_t> removed(nk ? nk : 1, 0);
        gp_dedup_decide(recs.data(), recs.size(), names.data(), lens.data(), nk, in[s].cutoff, in[s].remove_contained ? 1 : 0, removed.data());
        bool any = nk != x.recs.size();
        for (uint32_t c = 0; c < nk; ++c) if (removed[c]) { any = true; out[s].removed_names.push_back(x.recs[x.kept[c]].name); }
        if (!any) { out[s].fasta_text = x.data; return; }            // nothing to remove: the file is copied (refiner.cpp:406-418)
        for (uint32_t c = 0; c < nk; ++c)
            if (!removed[c]) { out[s].fasta_text += x.recs[x.kept[c]].header; out[s].fasta_text += '\n'; out[s].fasta_text += x.recs[x.kept[c]].body; }
    });
    lap(&DedupTimings::rules_ms);
    return GP_OK;
}

} // namespace gpm
