// merger.hpp -- ContigsMerger's per-gap pipeline (ContigsCompactor::CompactVer3,
// ContigsCompactor.cpp:773-983) with the two Evaluate call sites (:652 pairwise phase, :1491 relax
// chain) replaced by batched calls into the C ABI (include/gappadder_b200.h).  Any number of gaps go
// through one GPU context together: the pairwise phase of ALL gaps is one gp_overlap_pairs call, and
// step k of ALL merge chains is one call.
#pragma once
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "device_gate.hpp"
#include "fasta.hpp"
#include "gappadder_b200.h"

namespace gpm {

// The command line of CM/main.cpp:24-42,53-231 (same flags, same defaults).
struct MergeOptions {
    double max_frac_score_loss = 0.01;       // -s
    double min_frac_overlap = 0.005;         // -c
    double min_overlap_len = 100000;         // -x
    double max_overlap_clip_len = 0;         // -y
    double min_overlap_len_with_scaffold = 6;// -z
    int num_threads = 6;                     // -t   (only decides the "Arrange error" case)
    int min_support_kmer = 5;                // -m   (unused by the reference's live path)
    int line_length = 60;                    // -l
    int quick_kmer_len = 10;                 // -k
    double score_mismatch = -1.0;            // -i1
    double score_indel = -1.0;               // -i2
    std::string info_file = "tmp.info";      // -o
    int max_contig_path_len = -1;            // -p1 (unused by CompactVer3)
    int max_count_contig_in_path = -1;       // -p2 (default MAX_CONTIG_IN_PATH_COUNT = 20)
    bool verbose = false;                    // -V  (accepted, ignored: it pollutes stdout in the reference)
    bool host_quick_check = false;           // --host-quick-check (not a reference flag): candidate filter on the host
    int host_threads = 0;                    // host threads of the parallel host phases; 0: min(16, hardware threads)
    bool build_gml = true;                   // false (--no-gml in batch mode): the tmp.gml text is not even built
    bool host_relax = false;                 // --host-relax (not a reference flag): relax chains step by step from the host (round 1's form)
};

struct GapInput {
    std::string fasta_path;
    // Optional: the file's records when the caller has read it already (the batch driver reads every gap once, in
    // parallel, to balance the gaps over GPUs by contig lengths).  read_ok / fatal are read_fasta's results.
    bool loaded = false;
    bool read_ok = true;
    std::string fatal;
    std::vector<FastaRecord> records;
};

struct GapOutput {
    std::string stdout_text;   // what the reference writes to stdout
    std::string info_text;     // contents of the -o file
    std::string gml_text;      // contents of ./tmp.gml ("" when the reference would not reach it)
    bool wrote_info = false;   // the reference creates the -o file only when it gets that far
    int exit_code = 0;
    // Set when THIS gap is outside the implementation's contract (more than 16 distinct letters: see merger.cpp);
    // nothing is written for it (exit code 3 in the single-gap form), the other gaps of the batch are unaffected.
    std::string error;
    // statistics
    uint64_t pair_cells = 0, relax_cells = 0;
    uint32_t n_pairs = 0, n_relax = 0;
};

// Wall-clock phases of one merge_gaps call, milliseconds (filled when the pointer is given).
struct MergeTimings {
    double read_ms = 0;        // FASTA parsing, reverse complements, candidate pairs (quick check)
    double pairwise_ms = 0;    // gp_overlap_batch of the pairwise phase (pack + H2D + kernels + D2H)
    double graph_ms = 0;       // edges, overlap graph, GML text, path search
    double relax_ms = 0;       // all relax-chain steps (one gp_overlap_batch per step)
    double output_ms = 0;      // output text
    uint32_t relax_steps = 0;
    uint32_t relax_team_steps = 0;     // relax steps the library ran one CTA per pair (gp_last_team)
    uint64_t relax_pairs = 0, relax_second_passes = 0, relax_exact_retries = 0;   // certificate kernel, relax chain
    uint64_t closed_pairs = 0, closed_cells = 0;   // pairwise phase: node-vs-itself pairs answered in closed form
    double relax_call_ms = 0, relax_pack_ms = 0;   // relax chain: inside gp_overlap_batch in total / its packing+classification part
    uint64_t relax_shared_pairs = 0, relax_shared_cells = 0;   // relax steps answered by another chain of the gap with the same path prefix
    double relax_device_ms = 0;        // of relax_ms: inside gp_overlap_batch from first launch to results on the host
    double relax_host_ms = 0;          // of relax_ms: building the step's batch and the merged strings
    double qc_kernel_ms = 0;           // device quick check: kernel time, bases scanned (0.5 B each), work items
    uint64_t qc_bases = 0;
    uint32_t qc_items = 0;
    std::map<std::string, double> detail;  // finer wall-clock split (ms) of the phases above, by name
};

// Runs every gap (preloaded records are moved out of `in`).  Returns GP_OK or the failing gp_status (message via gp_last_error(ctx)); a failure
// here is a GPU/library failure, never an input problem (those are reported per gap like the
// reference does, on stdout with exit code 1).
// gate: when several host threads (each with its own context) drive one GPU, the DeviceGate they share (device_gate.hpp).
// It is held around the device phases only, so one thread's host phases (nodes, graph, strings, output text) run beside
// another thread's kernels, and the next chunk's pairwise kernels are queued right behind this chunk's relax kernel.
int merge_gaps(gp_ctx* ctx, const MergeOptions& opt, std::vector<GapInput>& in, std::vector<GapOutput>& out,
               std::string& error, MergeTimings* timings = nullptr, DeviceGate* gate = nullptr);

// Estimated DP cells of one gap's pairwise phase from contig lengths alone (all node pairs i <= j):
// used to balance gaps over GPUs before any sequence is examined.
uint64_t estimate_gap_cells(const std::vector<uint32_t>& contig_len);

// Longest-processing-time partition of gaps over `n_parts` workers; returns part index per gap.
std::vector<int> partition_gaps(const std::vector<uint64_t>& cost, int n_parts);

} // namespace gpm
