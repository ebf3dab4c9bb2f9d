// contigs_merger_main.cpp -- drop-in for GAPPadder's ContigsMerger binary
// (ContigsCompactor-v0.2.0/ContigsMerger/main.cpp:53-231 flags, :235-289 driver):
//
//   ContigsMerger_b200 -s F -i1 F -i2 F -x F -y F -k I -t I -m I -o INFO IN.fa > OUT.fa
//
// same flags, same defaults, same stdout / INFO / ./tmp.gml bytes; the DP runs on the B200 through
// libgappadder_b200.so.  Because one process per gap cannot amortise CUDA context creation, a batch
// form runs any number of gaps through one context (and through several GPUs):
//
//   ContigsMerger_b200 <flags> --batch LIST [--gpus N] [--streams S] [--chunk-gaps C] [--chunk-mb M] [--no-gml]
//   ContigsMerger_b200 <flags> --dedup IN.fa OUT.fa --cutoff C [--contained]      the dedup stage around a merge (dedup.hpp):
//   ContigsMerger_b200 <flags> --dedup-batch LIST [--gpus N]                      MergeContigs.py:15-70 without BWA / samtools;
//       LIST: IN.fa <TAB> OUT.fa <TAB> CUTOFF <TAB> g|p per line (g: `-P -g`, contained contigs; p: `-P`, duplicates)
//   ContigsMerger_b200 --serve SOCKET [--gpus N] [--window-ms W]     resident service (server.hpp): thin clients -- this same
//       binary with GAPPADDER_B200_SOCKET=SOCKET in its environment -- send one gap each; concurrent requests share a launch
//
// LIST holds one gap per line: IN.fa <TAB> OUT.fa <TAB> INFO.  Each OUT/INFO pair is byte-identical
// to what the single-gap form (and the reference) writes.  tmp.gml is written next to each OUT.fa as
// OUT.fa.gml in batch mode (the reference drops ./tmp.gml in the working directory of each process).
#include <malloc.h>
#include <sys/stat.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "dedup.hpp"
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
    int streams = 2;            // mergers (host thread + context + stream) per GPU; they alternate on the device (run_batch)
    int chunk_gaps = 256;       // batch pipeline: gaps per chunk at most ...
    int chunk_mb = 256;         // ... and MB of FASTA per chunk at most
    bool write_gml = true;
    bool stats = false;
    std::string serve;          // --serve SOCKET: run as the resident service (server.hpp)
    int window_ms = 3;
    bool shutdown = false;      // --shutdown: ask the server behind GAPPADDER_B200_SOCKET to exit
    // dedup stage (dedup.hpp): --dedup IN.fa OUT.fa --cutoff C [--contained], or --dedup-batch LIST
    std::string dedup_in, dedup_out, dedup_batch;
    double dedup_cutoff = 0.99;
    bool dedup_contained = false;
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
        if (!strcmp(a, "--streams")) { c.streams = parse_int(val, 2); ++pos; continue; }
        if (!strcmp(a, "--chunk-gaps")) { c.chunk_gaps = parse_int(val, 256); ++pos; continue; }
        if (!strcmp(a, "--chunk-mb")) { c.chunk_mb = parse_int(val, 256); ++pos; continue; }
        if (!strcmp(a, "--stats")) { c.stats = true; continue; }
        if (!strcmp(a, "--dedup")) { if (pos + 2 >= argc) return false; c.dedup_in = argv[pos + 1]; c.dedup_out = argv[pos + 2]; pos += 2; continue; }
        if (!strcmp(a, "--dedup-batch")) { if (!val) return false; c.dedup_batch = val; ++pos; continue; }
        if (!strcmp(a, "--cutoff")) { if (!val) return false; c.dedup_cutoff = atof(val); ++pos; continue; }
        if (!strcmp(a, "--contained")) { c.dedup_contained = true; continue; }
        if (!strcmp(a, "--shutdown")) { c.shutdown = true; continue; }
        ref_argv.push_back(argv[pos]);
    }
    Request r;
    if (!parse_request((int)ref_argv.size(), ref_argv.data(), r)) return false;
    c.opt = r.opt;
    c.input = r.input;
    c.have_input = !r.input.empty();
    c.write_gml = r.write_gml;
    c.opt.build_gml = r.write_gml;
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

// One unit of the batch pipeline: some gaps of one GPU's share, read, merged and written together.
struct Chunk {
    std::vector<size_t> gaps;                       // indices into the batch list
    std::vector<GapInput> in;
    std::vector<GapOutput> out;
    std::atomic<size_t> to_read{0};                 // gaps whose FASTA is not in memory yet
};

// The batch driver.  Gaps are independent, so the batch is a pipeline over CHUNKS of gaps:
//   partition  gaps -> GPUs, longest-processing-time on (FASTA bytes)^2 (a stat per gap; no sequence is read for it), then
//              each GPU's share is cut into chunks of at most --chunk-gaps gaps / --chunk-mb MB of FASTA;
//   readers    a pool of host threads parses the FASTA files chunk by chunk, a bounded number of chunks ahead of the GPU;
//   mergers    --streams S host threads per GPU (default 2), each with its own context, take that GPU's chunks in order
//              and run merge_gaps on them; they share one DeviceGate per GPU (device_gate.hpp) that is held around the
//              device phases only: the host phases of one chunk (nodes, graph, strings, output text) run beside the
//              kernels of another, two pairwise launches never compete for the SMs, and the next chunk's pairwise
//              kernels are queued right behind a chunk's relax kernel, whose long tail leaves most SMs idle;
//   writers    one thread per GPU writes the finished chunks' files and frees them.
// Memory is bounded by the chunks in flight, whatever the batch size (BASELINE configs[3]: 50 000 gaps).  Results do not
// depend on the split: every gap's bytes are what the single-gap form writes.
int run_batch(const Cli& c)
{
    using clk = std::chrono::steady_clock;
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
    // A batch allocates and frees the same big buffers chunk after chunk (hit matrices, pair lists, results, texts): keep them
    // in the heap instead of mapping and unmapping them every time (an mmap, a page fault per 4 KB and an munmap each; with
    // eight GPUs' mergers in one process that was a third of the host's time in the kernel).
    mallopt(M_MMAP_THRESHOLD, 32 << 20);            // the largest value glibc accepts
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    mallopt(M_TOP_PAD, 64 << 20);
    const auto t_start = clk::now();
    auto ms_since = [&](clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); };
    const int n_dev = c.gpus < 1 ? 1 : c.gpus;
    const size_t L = lines.size();

    // ---- partition: cost from the file size alone ---------------------------------------------------------------
    std::vector<uint64_t> fsize(L, 0), cost(L, 0);
    for (size_t g = 0; g < L; ++g) {
        struct stat sb;
        if (stat(lines[g].in.c_str(), &sb) == 0) fsize[g] = (uint64_t)sb.st_size;
        cost[g] = fsize[g] * fsize[g] + 1;                       // pairwise cells grow with the square of the bases
    }
    const std::vector<int> part = partition_gaps(cost, n_dev);
    const uint64_t chunk_bytes = (uint64_t)(c.chunk_mb < 1 ? 1 : c.chunk_mb) << 20;
    const size_t chunk_gaps = (size_t)(c.chunk_gaps < 1 ? 1 : c.chunk_gaps);
    std::vector<std::vector<std::unique_ptr<Chunk>>> chunks(n_dev);
    for (int d = 0; d < n_dev; ++d) {
        // an even split of the share (the last chunk is not a sliver), each chunk within both limits
        std::vector<size_t> mine;
        uint64_t bytes = 0;
        for (size_t g = 0; g < L; ++g) if (part[g] == d) { mine.push_back(g); bytes += fsize[g]; }
        if (mine.empty()) continue;
        size_t n_chunks = std::max<size_t>((mine.size() + chunk_gaps - 1) / chunk_gaps, (size_t)((bytes + chunk_bytes - 1) / chunk_bytes));
        n_chunks = std::max<size_t>(1, std::min(n_chunks, mine.size()));
        for (size_t k = 0; k < n_chunks; ++k) {
            std::unique_ptr<Chunk> ch(new Chunk);
            for (size_t q = k * mine.size() / n_chunks; q < (k + 1) * mine.size() / n_chunks; ++q) ch->gaps.push_back(mine[q]);
            ch->in.resize(ch->gaps.size());
            ch->to_read = ch->gaps.size();
            chunks[d].push_back(std::move(ch));
        }
    }
    const double partition_ms = ms_since(t_start);

    // ---- pipeline state -------------------------------------------------------------------------------------------
    std::mutex mu;                                               // guards the counters below
    std::condition_variable cv;
    std::vector<size_t> merged_chunks(n_dev, 0);                 // chunks a merger has taken (readers stay LOOKAHEAD ahead)
    std::vector<size_t> next_chunk(n_dev, 0);                    // next chunk a merger takes
    std::vector<std::deque<Chunk*>> to_write(n_dev);
    std::vector<size_t> written(n_dev, 0);
    std::atomic<bool> failed(false), gap_failed(false);
    constexpr size_t LOOKAHEAD = 3;
    // read tasks in the order the GPUs need them: round r of every GPU, then round r + 1
    struct ReadTask { int dev; size_t chunk, slot; };
    std::vector<ReadTask> tasks;
    {
        size_t rounds = 0;
        for (int d = 0; d < n_dev; ++d) rounds = std::max(rounds, chunks[d].size());
        for (size_t r = 0; r < rounds; ++r)
            for (int d = 0; d < n_dev; ++d)
                if (r < chunks[d].size())
                    for (size_t q = 0; q < chunks[d][r]->gaps.size(); ++q) tasks.push_back(ReadTask{d, r, q});
    }
    std::atomic<size_t> next_task(0);
    auto reader = [&] {
        for (size_t t = next_task.fetch_add(1); t < tasks.size(); t = next_task.fetch_add(1)) {
            const ReadTask& rt = tasks[t];
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return failed || rt.chunk < merged_chunks[rt.dev] + LOOKAHEAD; });
                if (failed) return;
            }
            Chunk& ch = *chunks[rt.dev][rt.chunk];
            GapInput& gi = ch.in[rt.slot];
            gi.fasta_path = lines[ch.gaps[rt.slot]].in;
            gi.read_ok = read_fasta(gi.fasta_path, gi.records, gi.fatal);
            gi.loaded = true;
            if (ch.to_read.fetch_sub(1) == 1) { std::lock_guard<std::mutex> lk(mu); cv.notify_all(); }
        }
    };

    const int per_dev = c.streams >= 1 ? c.streams : 1;
    const int n_workers = n_dev * per_dev;
    // host threads of a merger's parallel phases: the machine's threads shared out over the mergers, twice over (a merger
    // spends most of its time waiting for the device), between 2 and 16
    MergeOptions mopt = c.opt;
    if (mopt.host_threads <= 0) {
        const int hw = (int)std::max(1u, std::thread::hardware_concurrency());
        mopt.host_threads = std::max(2, std::min(16, 2 * hw / n_workers));
    }
    std::vector<int> rc(n_workers, 0);
    std::vector<std::string> err(n_workers);
    std::vector<MergeTimings> tim(n_workers);
    std::vector<uint64_t> cells(n_dev, 0), pcells(n_dev, 0);
    std::vector<double> wall(n_dev, 0), write_ms(n_dev, 0);
    std::vector<DeviceGate> dev_gate(n_dev);
    auto add_timings = [](MergeTimings& a, const MergeTimings& b) {
        a.read_ms += b.read_ms; a.pairwise_ms += b.pairwise_ms; a.graph_ms += b.graph_ms; a.relax_ms += b.relax_ms; a.output_ms += b.output_ms;
        a.relax_steps += b.relax_steps; a.relax_team_steps += b.relax_team_steps; a.relax_pairs += b.relax_pairs;
        a.relax_second_passes += b.relax_second_passes; a.relax_exact_retries += b.relax_exact_retries;
        a.closed_pairs += b.closed_pairs; a.closed_cells += b.closed_cells; a.relax_call_ms += b.relax_call_ms; a.relax_pack_ms += b.relax_pack_ms;
        a.relax_shared_pairs += b.relax_shared_pairs; a.relax_shared_cells += b.relax_shared_cells;
        a.relax_device_ms += b.relax_device_ms; a.relax_host_ms += b.relax_host_ms;
        a.qc_kernel_ms += b.qc_kernel_ms; a.qc_bases += b.qc_bases; a.qc_items += b.qc_items;
        for (const auto& kv : b.detail) a.detail[kv.first] += kv.second;
    };
    // Contexts first (CUDA context, module load, buffers for the largest chunk: a per-process constant of a second or
    // so, reported as setup_ms), in parallel; the batch clock starts when they exist.
    std::vector<gp_ctx*> ctxs(n_workers, nullptr);
    {
        std::vector<std::thread> mk;
        for (int w = 0; w < n_workers; ++w) mk.emplace_back([&, w] {
            const int dev = w % n_dev;
            if (chunks[dev].size() <= (size_t)(w / n_dev)) return;        // fewer chunks than mergers on this GPU
            const int r = gp_create(dev, &ctxs[w]);
            if (r != GP_OK) { rc[w] = r; err[w] = gp_last_error(nullptr); return; }
            // sizes from the FASTA bytes: every contig and its reverse complement, every node pair of a gap at most,
            // about two relax steps per node
            uint64_t bases = 0, pairs = 0, nodes = 0;
            for (const auto& ch : chunks[dev]) {
                uint64_t b = 0, p = 0, n = 0;
                for (size_t g : ch->gaps) { b += 2 * fsize[g]; const uint64_t k = 2 * (fsize[g] / 1000 + 2); n += k; p += k * (k + 1) / 2; }
                bases = std::max(bases, b); pairs = std::max(pairs, p); nodes = std::max(nodes, n);
            }
            gp_reserve(ctxs[w], bases, (uint32_t)std::min<uint64_t>(nodes, 0x7fffffffu), pairs / 2, 2 * nodes);
        });
        for (auto& t : mk) t.join();
        for (int w = 0; w < n_workers; ++w)
            if (rc[w] != 0) {
                fprintf(stderr, "ContigsMerger_b200: worker %d (GPU %d) failed (%d): %s\n", w, w % n_dev, rc[w], err[w].c_str());
                for (gp_ctx* x : ctxs) if (x) gp_destroy(x);
                return 3;
            }
    }
    const double setup_ms = ms_since(t_start) - partition_ms;
    const auto t_run = clk::now();
    auto merger = [&](int w) {
        const int dev = w % n_dev;
        gp_ctx* ctx = ctxs[w];
        if (!ctx) return;
        int r;
        for (;;) {
            Chunk* ch = nullptr;
            const auto tw = clk::now();
            {
                std::unique_lock<std::mutex> lk(mu);
                if (failed || next_chunk[dev] >= chunks[dev].size()) break;
                ch = chunks[dev][next_chunk[dev]++].get();
                merged_chunks[dev] = next_chunk[dev];
                cv.notify_all();
                cv.wait(lk, [&] { return failed || ch->to_read.load() == 0; });
                if (failed) break;
            }
            MergeTimings t;
            t.detail["pipeline.wait_for_fasta"] = ms_since(tw);
            r = merge_gaps(ctx, mopt, ch->in, ch->out, err[w], &t, &dev_gate[dev]);
            add_timings(tim[w], t);
            if (r != GP_OK) { rc[w] = r; failed = true; std::lock_guard<std::mutex> lk(mu); cv.notify_all(); break; }
            std::vector<GapInput>().swap(ch->in);
            { std::lock_guard<std::mutex> lk(mu); to_write[dev].push_back(ch); cv.notify_all(); }
        }
    };
    auto writer = [&](int dev) {
        for (;;) {
            Chunk* ch = nullptr;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return failed || !to_write[dev].empty() || written[dev] == chunks[dev].size(); });
                if (to_write[dev].empty()) break;                 // everything written, or a failure with nothing queued
                ch = to_write[dev].front(); to_write[dev].pop_front();
            }
            // the chunk's files on a few threads (two or three small files per gap)
            const auto tw0 = clk::now();
            std::atomic<size_t> next_gap(0);
            std::atomic<uint64_t> c_all(0), c_pair(0);
            std::atomic<bool> write_failed(false);
            std::string bad_path;
            std::mutex bad_mu;
            auto write_some = [&] {
                for (size_t k = next_gap.fetch_add(1); k < ch->gaps.size(); k = next_gap.fetch_add(1)) {
                    const BatchLine& b = lines[ch->gaps[k]];
                    GapOutput& o = ch->out[k];
                    if (!o.error.empty()) {                      // this gap only: nothing written, the others go on
                        fprintf(stderr, "ContigsMerger_b200: %s\n", o.error.c_str());
                        gap_failed = true;
                        continue;
                    }
                    if (!write_file(b.out, o.stdout_text)) { write_failed = true; std::lock_guard<std::mutex> lk(bad_mu); bad_path = b.out; continue; }
                    if (o.wrote_info) write_file(b.info, o.info_text);
                    if (c.write_gml && o.wrote_info) write_file(b.out + ".gml", o.gml_text);
                    c_all += o.pair_cells + o.relax_cells;
                    c_pair += o.pair_cells;
                }
            };
            {
                const size_t nt = std::min<size_t>(8, std::max<size_t>(1, ch->gaps.size() / 16));
                std::vector<std::thread> wt;
                for (size_t t = 1; t < nt; ++t) wt.emplace_back(write_some);
                write_some();
                for (auto& t : wt) t.join();
            }
            if (write_failed) {
                std::lock_guard<std::mutex> lk(mu);
                rc[dev] = GP_ERR_INVALID; err[dev] = "cannot write " + bad_path; failed = true; cv.notify_all();
                return;
            }
            cells[dev] += c_all;
            pcells[dev] += c_pair;
            write_ms[dev] += ms_since(tw0);
            std::vector<GapOutput>().swap(ch->out);
            wall[dev] = partition_ms + ms_since(t_run);
            { std::lock_guard<std::mutex> lk(mu); ++written[dev]; cv.notify_all(); }
        }
    };
    std::vector<std::thread> th;
    const unsigned T = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    for (unsigned t = 0; t < T; ++t) th.emplace_back(reader);
    for (int w = 0; w < n_workers; ++w) th.emplace_back(merger, w);
    for (int d = 0; d < n_dev; ++d) th.emplace_back(writer, d);
    for (auto& t : th) t.join();
    const double batch_ms = partition_ms + ms_since(t_run);
    for (gp_ctx* x : ctxs) if (x) gp_destroy(x);
    for (int w = 0; w < n_workers; ++w)
        if (rc[w] != 0) { fprintf(stderr, "ContigsMerger_b200: worker %d (GPU %d) failed (%d): %s\n", w, w % n_dev, rc[w], err[w].c_str()); return 3; }
    if (c.stats) {
        // One JSON line per run.  merge_ms = the whole batch: partition, then from the moment the contexts exist to the last
        // output file written (FASTA reading, every phase, file writing); creating the contexts (setup_ms) is a per-process
        // constant outside it.  The phase times are SUMS over
        // the chunks of the slowest GPU's mergers: phases of different chunks overlap, so they can add up to more than
        // merge_ms.
        uint64_t tot = 0, ptot = 0; for (uint64_t x : cells) tot += x; for (uint64_t x : pcells) ptot += x;
        int slow = 0; for (int d = 1; d < n_dev; ++d) if (wall[d] > wall[slow]) slow = d;
        MergeTimings t;
        for (int w = slow; w < n_workers; w += n_dev) add_timings(t, tim[w]);
        t.detail["pipeline.write_files"] = write_ms[slow];
        // dp_gcells / pairwise_gcells: m*n of every Evaluate the reference runs for these gaps; closed_gcells of them
        // (a node against itself) are answered in closed form here, so computed cells = dp_gcells - closed_gcells
        uint64_t closed = 0; for (const MergeTimings& x : tim) closed += x.closed_cells;
        uint64_t shared_pairs = 0, shared_cells = 0;      // relax steps shared between chains with a common path prefix: not computed twice
        for (const MergeTimings& x : tim) { shared_pairs += x.relax_shared_pairs; shared_cells += x.relax_shared_cells; }
        double qc_ms = 0; uint64_t qc_bases = 0; uint32_t qc_items = 0;
        for (const MergeTimings& x : tim) { qc_ms += x.qc_kernel_ms; qc_bases += x.qc_bases; qc_items += x.qc_items; }
        std::string per_worker = "\"worker_wall_ms\": [";
        for (int d = 0; d < n_dev; ++d) per_worker += (d ? ", " : "") + std::to_string(wall[d]);
        per_worker += "], \"worker_gcells\": [";
        for (int d = 0; d < n_dev; ++d) per_worker += (d ? ", " : "") + std::to_string(cells[d] / 1e9);
        per_worker += "], \"worker_gaps\": [";
        for (int d = 0; d < n_dev; ++d) { size_t cnt = 0; for (size_t g = 0; g < L; ++g) cnt += part[g] == d; per_worker += (d ? ", " : "") + std::to_string(cnt); }
        per_worker += "], \"worker_chunks\": [";
        for (int d = 0; d < n_dev; ++d) per_worker += (d ? ", " : "") + std::to_string(chunks[d].size());
        per_worker += "], \"detail_ms\": {";
        { bool first = true; for (const auto& kv : t.detail) { char b[96]; snprintf(b, sizeof b, "%s\"%s\": %.3f", first ? "" : ", ", kv.first.c_str(), kv.second); per_worker += b; first = false; } }
        per_worker += "}";
        fprintf(stderr, "{\"gaps\": %zu, \"gpus\": %d, \"workers\": %d, \"mergers_per_gpu\": %d, \"dp_gcells\": %.6f, \"pairwise_gcells\": %.6f, \"closed_gcells\": %.6f, \"merge_ms\": %.3f, "
                        "\"read_ms\": %.3f, \"pairwise_ms\": %.3f, \"graph_ms\": %.3f, \"relax_ms\": %.3f, \"relax_steps\": %u, \"output_ms\": %.3f, "
                        "\"relax_device_ms\": %.3f, \"relax_host_ms\": %.3f, \"relax_team_steps\": %u, \"relax_pairs\": %llu, "
                        "\"relax_second_passes\": %llu, \"relax_exact_retries\": %llu, \"relax_shared_pairs\": %llu, \"relax_shared_gcells\": %.6f, \"relax_call_ms\": %.3f, \"relax_pack_ms\": %.3f, "
                        "\"qc_kernel_ms\": %.4f, \"qc_bases\": %llu, \"qc_items\": %u, \"partition_ms\": %.3f, \"setup_ms\": %.3f, %s}\n",
                L, n_dev, n_dev, per_dev, tot / 1e9, ptot / 1e9, closed / 1e9, batch_ms, t.read_ms, t.pairwise_ms, t.graph_ms, t.relax_ms, t.relax_steps, t.output_ms,
                t.relax_device_ms, t.relax_host_ms, t.relax_team_steps, (unsigned long long)t.relax_pairs,
                (unsigned long long)t.relax_second_passes, (unsigned long long)t.relax_exact_retries,
                (unsigned long long)shared_pairs, shared_cells / 1e9, t.relax_call_ms, t.relax_pack_ms,
                qc_ms, (unsigned long long)qc_bases, qc_items, partition_ms, setup_ms, per_worker.c_str());
    }
    return gap_failed ? 3 : 0;
}

// The dedup stage for any number of contig sets: sets are independent, so they are dealt to the GPUs by FASTA size
// (longest-processing-time) and every GPU runs its share as one batch.
int run_dedup(const Cli& c)
{
    struct Line { DedupInput in; std::string out; };
    std::vector<Line> lines;
    if (!c.dedup_batch.empty()) {
        std::ifstream f(c.dedup_batch);
        if (!f) { fprintf(stderr, "ContigsMerger_b200: cannot open dedup list %s\n", c.dedup_batch.c_str()); return 2; }
        std::string ln;
        while (std::getline(f, ln)) {
            if (ln.empty()) continue;
            std::vector<std::string> col;
            size_t b = 0;
            for (;;) { const size_t t = ln.find('\t', b); col.push_back(ln.substr(b, t == std::string::npos ? t : t - b)); if (t == std::string::npos) break; b = t + 1; }
            if (col.size() != 4 || (col[3] != "g" && col[3] != "p")) { fprintf(stderr, "ContigsMerger_b200: bad dedup line: %s\n", ln.c_str()); return 2; }
            Line l;
            l.in.fasta_path = col[0]; l.out = col[1]; l.in.cutoff = atof(col[2].c_str()); l.in.remove_contained = col[3] == "g";
            lines.push_back(l);
        }
    } else {
        Line l;
        l.in.fasta_path = c.dedup_in; l.out = c.dedup_out; l.in.cutoff = c.dedup_cutoff; l.in.remove_contained = c.dedup_contained;
        lines.push_back(l);
    }
    const int n_dev = std::max(1, std::min<int>(c.gpus, (int)std::max<size_t>(1, lines.size())));
    std::vector<uint64_t> cost(lines.size(), 1);
    for (size_t g = 0; g < lines.size(); ++g) { struct stat sb; if (stat(lines[g].in.fasta_path.c_str(), &sb) == 0) cost[g] = (uint64_t)sb.st_size * (uint64_t)sb.st_size + 1; }
    const std::vector<int> part = partition_gaps(cost, n_dev);
    std::vector<int> rc(n_dev, 0);
    std::vector<std::string> err(n_dev);
    std::vector<DedupTimings> tim(n_dev);
    std::vector<double> wall(n_dev, 0);
    std::vector<uint64_t> cells(n_dev, 0), npairs(n_dev, 0), removed(n_dev, 0), contigs(n_dev, 0);
    std::atomic<bool> set_failed(false);
    auto worker = [&](int dev) {
        std::vector<DedupInput> in;
        std::vector<size_t> which;
        for (size_t g = 0; g < lines.size(); ++g) if (part[g] == dev) { in.push_back(lines[g].in); which.push_back(g); }
        if (in.empty()) return;
        gp_ctx* ctx = nullptr;
        int r = gp_create(dev, &ctx);
        if (r != GP_OK) { rc[dev] = r; err[dev] = gp_last_error(nullptr); return; }
        std::vector<DedupOutput> out;
        const auto w0 = std::chrono::steady_clock::now();
        r = dedup_sets(ctx, c.opt, in, out, err[dev], &tim[dev]);
        if (r == GP_OK)
            for (size_t k = 0; k < which.size(); ++k) {
                if (!out[k].error.empty()) { fprintf(stderr, "ContigsMerger_b200: %s\n", out[k].error.c_str()); set_failed = true; continue; }
                if (!write_file(lines[which[k]].out, out[k].fasta_text)) { r = GP_ERR_INVALID; err[dev] = "cannot write " + lines[which[k]].out; break; }
                cells[dev] += out[k].pair_cells; npairs[dev] += out[k].n_pairs; contigs[dev] += out[k].n_contigs;
                removed[dev] += out[k].removed_names.size() + (out[k].n_contigs - out[k].n_unique);
            }
        wall[dev] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - w0).count();
        gp_destroy(ctx);
        rc[dev] = r;
    };
    std::vector<std::thread> th;
    for (int d = 0; d < n_dev; ++d) th.emplace_back(worker, d);
    for (auto& t : th) t.join();
    for (int d = 0; d < n_dev; ++d)
        if (rc[d] != 0) { fprintf(stderr, "ContigsMerger_b200: dedup worker on GPU %d failed (%d): %s\n", d, rc[d], err[d].c_str()); return 3; }
    if (c.stats) {
        int slow = 0; for (int d = 1; d < n_dev; ++d) if (wall[d] > wall[slow]) slow = d;
        uint64_t tc = 0, tp = 0, tr = 0, tn = 0;
        for (int d = 0; d < n_dev; ++d) { tc += cells[d]; tp += npairs[d]; tr += removed[d]; tn += contigs[d]; }
        fprintf(stderr, "{\"sets\": %zu, \"gpus\": %d, \"contigs\": %llu, \"removed\": %llu, \"pairs\": %llu, \"dp_gcells\": %.6f, \"dedup_ms\": %.3f, "
                        "\"read_ms\": %.3f, \"device_ms\": %.3f, \"rules_ms\": %.3f, \"qc_kernel_ms\": %.4f}\n",
                lines.size(), n_dev, (unsigned long long)tn, (unsigned long long)tr, (unsigned long long)tp, tc / 1e9, wall[slow],
                tim[slow].read_ms, tim[slow].device_ms, tim[slow].rules_ms, tim[slow].qc_kernel_ms);
    }
    return set_failed ? 3 : 0;
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
    if (!c.dedup_batch.empty() || !c.dedup_in.empty()) return run_dedup(c);
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
