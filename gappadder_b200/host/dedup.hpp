// dedup.hpp -- the dedup stage around every merge: MergeContigs.py:15-70 `remove_duplicate_contained`
// = `TERefiner_1 -U`, a BWA-MEM self-alignment of the contig set, `TERefiner_1 -P -c cutoff [-g]`.
//
// Here: unique names and the removal rules are the reference's, restated exactly in the C ABI (gp_dedup_unique_names,
// gp_dedup_decide; pinned to the prebuilt TERefiner_1 by tests/golden/dedup).  The self-alignment is NOT BWA's (BWA is not
// vendored, not pinned: parity unpinned at that boundary): every ordered contig pair whose ends share a k-mer (the device
// quick check, all ordered pairs) is aligned by the merger's own overlap DP (Evaluate on the B200 kernels, forward and
// reverse-complement strand), and one result stands for two alignment records by the rule of gp_dedup_records.  Many
// contig sets go through one context together, like gaps through merge_gaps.
#pragma once
#include <string>
#include <vector>

#include "gappadder_b200.h"
#include "merger.hpp"

namespace gpm {

struct DedupInput {
    std::string fasta_path;
    double cutoff = 0.99;            // -c
    bool remove_contained = false;   // -g: contained contigs (perfect, full-length cover); otherwise duplicates of similar length
};

struct DedupOutput {
    std::string fasta_text;          // what the stage writes: the kept records, verbatim, in input order
    std::vector<std::string> removed_names;   // contigs the -P rule removed
    uint32_t n_contigs = 0, n_unique = 0, n_pairs = 0, n_records = 0;
    uint64_t pair_cells = 0;
    std::string error;               // this set is outside the implementation's contract; nothing is written for it
};

struct DedupTimings { double read_ms = 0, device_ms = 0, rules_ms = 0, qc_kernel_ms = 0; };

// opt supplies the DP side (-i1, -i2, -y, -k, -s) exactly as for the merger.  Returns GP_OK or the failing gp_status.
int dedup_sets(gp_ctx* ctx, const MergeOptions& opt, const std::vector<DedupInput>& in, std::vector<DedupOutput>& out,
               std::string& error, DedupTimings* timings = nullptr);

} // namespace gpm
