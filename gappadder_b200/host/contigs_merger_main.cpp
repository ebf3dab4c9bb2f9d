// contigs_merger_main.cpp -- drop-in for GAPPadder's ContigsMerger binary
// (ContigsCompactor-v0.2.0/ContigsMerger/main.cpp:53-231 flags, :235-289 driver):
//
//   ContigsMerger_b200 -s F -i1 F -i2 F -x F -y F -k I -t I -m I -o INFO IN.fa > OUT.fa
//
// same flags, same defaults, same stdout / INFO / ./tmp.gml bytes; the DP runs on the B200 through
// libgappadder_b200.so.  Because one process per gap cannot amortise CUDA context creation, a batch
// form runs any number of gaps through one context (and through several GPUs):
//
//   ContigsMerger_b200 <flags> --batch LIST [--gpus N] [--streams S] [--no-gml]
//   ContigsMerger_b200 --serve SOCKET [--gpus N] [--window-ms W]     resident service (server.hpp): thin clients -- this same
//       binary with GAPPADDER_B200_SOCKET=SOCKET in its environment -- send one gap each; concurrent requests share a launch
//
// LIST holds one gap per line: IN.fa <TAB> OUT.fa <TAB> INFO.  Each OUT/INFO pair is byte-identical
// to what the single-gap form (and the reference) writes.  tmp.gml is written next to each OUT.fa as
// OUT.fa.gml in batch mode (the reference drops ./tmp.gml in the working directory of each process).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include "fasta.hpp"
#include "gappadder_b200.h"
#include "merger.hpp"
#include "server.hpp"

using namespace gpm;

namespace {

struct Cli {
    MergeOptions opt;
    std::string input;
    bool have_input = false;
    std::string batch;
    int gpus = 1;
    int streams = 1;            // workers (host thread + context + stream) per GPU
    bool write_gml = true;
    bool stats = false;
    std::string serve;          // --serve SOCKET: run as the resident service (server.hpp)
    int window_ms = 3;
    bool shutdown = false;      // --shutdown: ask the server behind GAPPADDER_B200_SOCKET to exit
};

int parse_int(const char* s, int dflt) { int v = dflt; if (s) sscanf(s, "%d", &v); return v; }

// The reference's flags (CheckArguments, CM/main.cpp:53-231: recognised by their second and third character only) are
// parsed by parse_request (server.cpp), shared with the resident server; the flags below are this binary's own and are
// taken out of argv first.
bool parse_args(int argc, char** argv, Cli& c, std::vector<char*>& ref_argv)
{
    ref_argv.assign(1, argv[0]);
    for (int pos = 1; pos < argc; ++pos) {
        const char* a = argv[pos];
        const char* val = pos + 1 < argc ? argv[pos + 1] : nullptr;
        if (!strcmp(a, "--batch")) { if (!val) return false; c.batch = val; ++pos; continue; }
        if (!strcmp(a, "--serve")) { if (!val) return false; c.serve = val; ++pos; continue; }
        if (!strcmp(a, "--window-ms")) { c.window_ms = parse_int(val, 3); ++pos; continue; }
        if (!strcmp(a, "--gpus")) { c.gpus = parse_int(val, 1); ++pos; continue; }
        if (!strcmp(a, "--streams")) { c.streams = parse_int(val, 1); ++pos; continue; }
        if (!strcmp(a, "--stats")) { c.stats = true; continue; }
        if (!strcmp(a, "--shutdown")) { c.shutdown = true; continue; }
        ref_argv.push_back(argv[pos]);
    }
    Request r;
    if (!parse_request((int)ref_argv.size(), ref_argv.data(), r)) return false;
    c.opt = r.opt;
    c.input = r.input;
    c.have_input = !r.input.empty();
    c.write_gml = r.write_gml;
    if (c.opt.verbose) printf("Turn on Verbose\n");
    return true;
}

bool write_file(const std::string& path, const std::string& data)
{
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return false;
    fwrite(data.data(), 1, data.size(), f);
    fclose(f);
    return true;
}

struct BatchLine { std::string in, out, info; };

int run_batch(const Cli& c)
{
    std::vector<BatchLine> lines;
    {
        std::ifstream f(c.batch);
        if (!f) { fprintf(stderr, "ContigsMerger_b200: cannot open batch list %s\n", c.batch.c_str()); return 2; }
        std::string ln;
        while (std::getline(f, ln)) {
            if (ln.empty()) continue;
            BatchLine b;
            size_t t1 = ln.find('\t'), t2 = t1 == std::string::npos ? t1 : ln.find('\t', t1 + 1);
            if (t2 == std::string::npos) { fprintf(stderr, "ContigsMerger_b200: bad batch line: %s\n", ln.c_str()); return 2; }
            b.in = ln.substr(0, t1); b.out = ln.substr(t1 + 1, t2 - t1 - 1); b.info = ln.substr(t2 + 1);
            lines.push_back(b);
        }
    }
    // Balance gaps over workers by estimated pairwise cells (contig lengths only).  A worker is a host thread with
    // its own context and stream; --streams S puts S of them on every GPU, so that one worker's host phases and the
    // thin tail of its relax chains (a few long pairs per launch, 32 launches deep) run beside another worker's
    // kernels.  Gaps are independent: no ordering between workers, results do not depend on the split.
    const int n_dev = c.gpus < 1 ? 1 : c.gpus;
    // Default 1.  Measured on cfg1 (200 gaps, one B200, several runs each): 1 worker 287-306 ms; 3 workers 237-430 ms:
    // sometimes a quarter faster, sometimes slower, because a worker's short relax launch can queue behind another
    // worker's persistent pairwise kernel, which holds every SM until its work queue is empty.
    const int per_dev = c.streams >= 1 ? c.streams : 1;
    int n_gpus = n_dev * per_dev;                                  // number of workers from here on
    if ((size_t)n_gpus > lines.size() && !lines.empty()) n_gpus = (int)lines.size();
    std::vector<uint64_t> cost(lines.size(), 0);
    const auto part0 = std::chrono::steady_clock::now();
    // Every FASTA is read once, here, on the host cores: the contig lengths balance the gaps over the workers
    // (gp_partition_gaps), the records go to the workers as they are.
    std::vector<GapInput> loaded(lines.size());
    {
        std::atomic<size_t> next(0);
        auto reader = [&] {
            for (size_t g = next.fetch_add(1); g < lines.size(); g = next.fetch_add(1)) {
                GapInput& gi = loaded[g];
                gi.fasta_path = lines[g].in;
                gi.read_ok = read_fasta(gi.fasta_path, gi.records, gi.fatal);
                gi.loaded = true;
                std::vector<uint32_t> lens;
                for (const FastaRecord& r : gi.records) lens.push_back((uint32_t)r.seq.size());
                cost[g] = estimate_gap_cells(lens);
            }
        };
        const unsigned T = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
        std::vector<std::thread> rd;
        for (unsigned t = 0; t < T; ++t) rd.emplace_back(reader);
        for (auto& t : rd) t.join();
    }
    const std::vector<int> part = partition_gaps(cost, n_gpus);
    const double partition_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - part0).count();
    std::vector<int> rc(n_gpus, 0);
    std::vector<std::string> err(n_gpus);
    std::vector<uint64_t> cells(n_gpus, 0), pcells(n_gpus, 0);
    std::vector<MergeTimings> tim(n_gpus);
    std::vector<double> wall(n_gpus, 0);
    std::atomic<bool> gap_failed(false);
    auto worker = [&](int dev) {
        std::vector<GapInput> in;
        std::vector<size_t> which;
        for (size_t g = 0; g < lines.size(); ++g) if (part[g] == dev) { in.push_back(std::move(loaded[g])); which.push_back(g); }
        if (in.empty()) return;
        gp_ctx* ctx = nullptr;
        int r = gp_create(dev % n_dev, &ctx);
        if (r != GP_OK) { rc[dev] = r; err[dev] = gp_last_error(nullptr); return; }
        {   // buffers for this worker's share up front (sizes from the FASTA records: every contig and its reverse complement,
            // every node pair of a gap at most, about two relax steps per node)
            uint64_t bases = 0, pairs = 0, nodes = 0;
            for (const GapInput& gi : in) {
                const uint64_t n = 2 * gi.records.size();
                nodes += n; pairs += n * (n + 1) / 2;
                for (const FastaRecord& rec : gi.records) bases += 2 * rec.seq.size();
            }
            gp_reserve(ctx, bases, (uint32_t)nodes, pairs / 2, 2 * nodes);
        }
        std::vector<GapOutput> out;
        const auto w0 = std::chrono::steady_clock::now();
        r = merge_gaps(ctx, c.opt, in, out, err[dev], &tim[dev]);
        wall[dev] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
        gp_destroy(ctx);
        if (r != GP_OK) { rc[dev] = r; return; }
        for (size_t k = 0; k < which.size(); ++k) {
            const BatchLine& b = lines[which[k]];
            if (!out[k].error.empty()) {                       // this gap only: nothing written, the others go on
                fprintf(stderr, "ContigsMerger_b200: %s\n", out[k].error.c_str());
                gap_failed = true;
                continue;
            }
            if (!write_file(b.out, out[k].stdout_text)) { rc[dev] = GP_ERR_INVALID; err[dev] = "cannot write " + b.out; return; }
            if (out[k].wrote_info) write_file(b.info, out[k].info_text);
            if (c.write_gml && out[k].wrote_info) write_file(b.out + ".gml", out[k].gml_text);
            cells[dev] += out[k].pair_cells + out[k].relax_cells;
            pcells[dev] += out[k].pair_cells;
        }
    };
    std::vector<std::thread> th;
    for (int d = 0; d < n_gpus; ++d) th.emplace_back(worker, d);
    for (auto& t : th) t.join();
    for (int d = 0; d < n_gpus; ++d)
        if (rc[d] != 0) { fprintf(stderr, "ContigsMerger_b200: worker %d (GPU %d) failed (%d): %s\n", d, d % n_dev, rc[d], err[d].c_str()); return 3; }
    if (c.stats) {
        // one JSON line per run: cells and the slowest GPU's phase times (merge_gaps only, context creation excluded)
        uint64_t tot = 0, ptot = 0; for (uint64_t x : cells) tot += x; for (uint64_t x : pcells) ptot += x;
        int slow = 0; for (int d = 1; d < n_gpus; ++d) if (wall[d] > wall[slow]) slow = d;
        const MergeTimings& t = tim[slow];
        // dp_gcells / pairwise_gcells: m*n of every Evaluate the reference runs for these gaps; closed_gcells of them
        // (a node against itself) are answered in closed form here, so computed cells = dp_gcells - closed_gcells
        uint64_t closed = 0; for (const MergeTimings& x : tim) closed += x.closed_cells;
        uint64_t shared_pairs = 0, shared_cells = 0;      // relax steps shared between chains with a common path prefix: not computed twice
        for (const MergeTimings& x : tim) { shared_pairs += x.relax_shared_pairs; shared_cells += x.relax_shared_cells; }
        double qc_ms = 0; uint64_t qc_bases = 0; uint32_t qc_items = 0;
        for (const MergeTimings& x : tim) { qc_ms = std::max(qc_ms, x.qc_kernel_ms); qc_bases += x.qc_bases; qc_items += x.qc_items; }
        std::string per_worker = "\"worker_wall_ms\": [";
        for (int d = 0; d < n_gpus; ++d) per_worker += (d ? ", " : "") + std::to_string(wall[d]);
        per_worker += "], \"worker_gcells\": [";
        for (int d = 0; d < n_gpus; ++d) per_worker += (d ? ", " : "") + std::to_string(cells[d] / 1e9);
        per_worker += "], \"worker_gaps\": [";
        for (int d = 0; d < n_gpus; ++d) { size_t cnt = 0; for (size_t g = 0; g < lines.size(); ++g) cnt += part[g] == d; per_worker += (d ? ", " : "") + std::to_string(cnt); }
        per_worker += "], \"detail_ms\": {";
        { bool first = true; for (const auto& kv : t.detail) { char b[96]; snprintf(b, sizeof b, "%s\"%s\": %.3f", first ? "" : ", ", kv.first.c_str(), kv.second); per_worker += b; first = false; } }
        per_worker += "}";
        fprintf(stderr, "{\"gaps\": %zu, \"gpus\": %d, \"workers\": %d, \"dp_gcells\": %.6f, \"pairwise_gcells\": %.6f, \"closed_gcells\": %.6f, \"merge_ms\": %.3f, "
                        "\"read_ms\": %.3f, \"pairwise_ms\": %.3f, \"graph_ms\": %.3f, \"relax_ms\": %.3f, \"relax_steps\": %u, \"output_ms\": %.3f, "
                        "\"relax_device_ms\": %.3f, \"relax_host_ms\": %.3f, \"relax_team_steps\": %u, \"relax_pairs\": %llu, "
                        "\"relax_second_passes\": %llu, \"relax_exact_retries\": %llu, \"relax_shared_pairs\": %llu, \"relax_shared_gcells\": %.6f, \"relax_call_ms\": %.3f, \"relax_pack_ms\": %.3f, "
                        "\"qc_kernel_ms\": %.4f, \"qc_bases\": %llu, \"qc_items\": %u, \"partition_ms\": %.3f, %s}\n",
                lines.size(), n_dev, n_gpus, tot / 1e9, ptot / 1e9, closed / 1e9, wall[slow], t.read_ms, t.pairwise_ms, t.graph_ms, t.relax_ms, t.relax_steps, t.output_ms,
                t.relax_device_ms, t.relax_host_ms, t.relax_team_steps, (unsigned long long)t.relax_pairs,
                (unsigned long long)t.relax_second_passes, (unsigned long long)t.relax_exact_retries,
                (unsigned long long)shared_pairs, shared_cells / 1e9, t.relax_call_ms, t.relax_pack_ms,
                qc_ms, (unsigned long long)qc_bases, qc_items, partition_ms, per_worker.c_str());
    }
    return gap_failed ? 3 : 0;
}

} // namespace

int main(int argc, char** argv)
{
    Cli c;
    std::vector<char*> ref_argv;
    if (!parse_args(argc, argv, c, ref_argv)) { printf("Wrong input.\n"); return 1; }           // CM/main.cpp:216-220
    if (!c.serve.empty()) {
        ServeOptions so;
        so.socket_path = c.serve; so.gpus = c.gpus < 1 ? 1 : c.gpus; so.window_ms = c.window_ms < 0 ? 0 : c.window_ms;
        return serve(so);
    }
    if (!c.batch.empty()) return run_batch(c);
    // single gap, exactly the reference's process: argv[repeatfileArgIndex] defaults to argv[1]
    if (!c.have_input && !c.shutdown) { fprintf(stderr, "usage: ContigsMerger_b200 <flags> contigs.fa\n"); return 1; }
    // A resident server (ContigsMerger_b200 --serve SOCKET) answers when GAPPADDER_B200_SOCKET names it: no CUDA context
    // in this process, and the gaps of concurrent callers share one launch.  Same bytes either way.
    if (const char* sock = getenv("GAPPADDER_B200_SOCKET")) {
        Reply rp;
        std::vector<char*> send_argv = ref_argv;
        char shut[] = "--shutdown";
        if (c.shutdown) send_argv.insert(send_argv.begin() + 1, shut);
        if (request_from_server(sock, (int)send_argv.size(), send_argv.data(), rp)) {
            if (!rp.err.empty()) fwrite(rp.err.data(), 1, rp.err.size(), stderr);
            fwrite(rp.out.data(), 1, rp.out.size(), stdout);
            if (rp.wrote_info) {
                if (c.write_gml) write_file("tmp.gml", rp.gml);
                if (!write_file(c.opt.info_file, rp.info)) { printf("Can not open file: %s\n", c.opt.info_file.c_str()); return 1; }
            }
            return rp.exit_code;
        }
        if (c.shutdown) return 0;
    }
    gp_ctx* ctx = nullptr;
    int rc = gp_create(0, &ctx);
    if (rc != GP_OK) { fprintf(stderr, "ContigsMerger_b200: %s (no CPU fallback)\n", gp_last_error(nullptr)); return 3; }
    std::vector<GapOutput> out;
    std::string err;
    std::vector<GapInput> single(1);
    single[0].fasta_path = c.input;
    rc = merge_gaps(ctx, c.opt, single, out, err);
    gp_destroy(ctx);
    if (rc != GP_OK) { fprintf(stderr, "ContigsMerger_b200: %s\n", err.c_str()); return 3; }   // never partial stdout
    if (!out[0].error.empty()) { fprintf(stderr, "ContigsMerger_b200: %s\n", out[0].error.c_str()); return 3; }
    fwrite(out[0].stdout_text.data(), 1, out[0].stdout_text.size(), stdout);
    if (out[0].wrote_info) {
        if (c.write_gml) write_file("tmp.gml", out[0].gml_text);
        if (!write_file(c.opt.info_file, out[0].info_text)) {
            printf("Can not open file: %s\n", c.opt.info_file.c_str());                 // ContigsCompactor.cpp:1548-1552
            return 1;
        }
    }
    return out[0].exit_code;
}
