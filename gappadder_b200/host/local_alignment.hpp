// local_alignment.hpp -- host-side mirror of TERefiner's LocalAlignment (TERefiner/algorithms/local_alignment.h:6-20,
// local_alignment.cpp:1036-1160) over the C ABI (gp_local_affine_batch: the affine-gap local aligner on the B200, results
// identical to the reference's own aln_stdaln, see csrc/affine_local.cuh).
//
// The three methods keep the reference's names, argument order and 1-based coordinates.  Each also has a batch form, which
// is the shape the kernels are built for (one warp per pair, thousands of pairs per launch): the rest alignments of a whole
// batch are one more launch, not one per pair.
//
// Where the reference is undefined it reads path[-1] (nothing aligns: local_alignment.cpp:611-614, :817); here such an
// alignment reports 0 0 0 0 and `aligned` false.
#pragma once
#include <string>
#include <vector>

#include "gappadder_b200.h"

namespace gpm {

struct LaPair { std::string ref, sgmt; };                    // sref (seq1), ssgmt (seq2)
struct LaHit {                                               // one aln_stdaln(.., &aln_param_blast, LOCAL, 1)
    int start_ref = 0, end_ref = 0, start_sgmt = 0, end_sgmt = 0;
    int score = 0;
    bool aligned = false;
};
struct LaAlign { LaHit opt, left, right; bool has_left = false, has_right = false; };      // LocalAlignment::align
struct LaRest { LaHit opt, rest; };                                                       // optAlignWithRestSecondOpt

class LocalAlignment {
public:
    explicit LocalAlignment(gp_ctx* ctx) : ctx_(ctx) {}

    // the reference's signatures (local_alignment.h:13-19); false when the library refused the pair (error() says why)
    bool optAlign(const std::string& sref, const std::string& ssgmt, int& optm_start_ref, int& optm_end_ref, int& optm_start_sgmt, int& optm_end_sgmt);
    bool align(const std::string& sref, const std::string& ssgmt, int& optm_start_ref, int& optm_end_ref, int& optm_start_sgmt, int& optm_end_sgmt,
               int& start_ref1, int& end_ref1, int& start_sgmt1, int& end_sgmt1,
               int& start_ref2, int& end_ref2, int& start_sgmt2, int& end_sgmt2);
    bool optAlignWithRestSecondOpt(const std::string& sref, const std::string& ssgmt, int& optm_start_ref, int& optm_end_ref, int& optm_start_sgmt,
                                   int& optm_end_sgmt, int& start_ref1, int& end_ref1, int& start_sgmt1, int& end_sgmt1);

    // many pairs per launch
    bool optAlignBatch(const std::vector<LaPair>& in, std::vector<LaHit>& out);
    bool alignBatch(const std::vector<LaPair>& in, std::vector<LaAlign>& out);
    bool optAlignWithRestSecondOptBatch(const std::vector<LaPair>& in, std::vector<LaRest>& out);

    const std::string& error() const { return error_; }

private:
    gp_ctx* ctx_;
    std::string error_;
};

// RepeatsClassifier::validateRepeats (TERefiner/RepeatsClassifier.cpp:46-114, TERefiner_1 -A): aligned length of the best
// alignment plus that of the second one on the concatenated rest, for seq2 and for its reverse complement
// (StrOperation::getReverseSupplementary, StrOperation.cpp:6-28); the larger of the two sums.
bool validate_repeats_batch(LocalAlignment& la, const std::vector<LaPair>& in, std::vector<int>& out);

} // namespace gpm
