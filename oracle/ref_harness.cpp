// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Thin C-ABI shim around the UNMODIFIED reference ContigsCompactor::Evaluate
// (ContigsCompactor-v0.2.0/ContigsMerger/ContigsCompactor.cpp:1572-1873) so that
// tests and bench.py's cpu_baseline leg can call the reference's own DP on
// arbitrary sequence pairs.  It is compiled by oracle/build_ref.sh against a
// scratch copy of the reference sources (patched only as described there) into
// oracle/_ref/libcm_ref.so.  Nothing in the product path links or loads it.
//
// Evaluate and IsScoreSignificant are private members; the shim opens them with
// the usual test-only preprocessor trick instead of editing the header.
#include <cstdint>
#include <cstring>
#include <string>

#define private public
#include "ContigsCompactor.h"
#undef private

// Filled by the one-line hook build_ref.sh inserts just before
// `int res = 2;` (ContigsCompactor.cpp:1712): the best-cell scan result, which
// the reference otherwise drops for rejected pairs.
static thread_local int g_hook[4];
extern "C" void cmref_hook(int scoreMax, int posRowEnd, int posColEnd, int nclip)
{
    g_hook[0] = scoreMax; g_hook[1] = posRowEnd; g_hook[2] = posColEnd; g_hook[3] = nclip;
}

static thread_local std::string g_merged;

extern "C" {

// Same setters CM/main.cpp:250-262 calls; values are passed as the doubles
// main.cpp ends up with (i.e. widened floats for the sscanf("%f") flags).
void cmref_set_params(double fracLossScore, double fracMinOverlap, double minOverlapLen,
                      double maxOverlapClipLen, double minOverlapLenWithScaffold,
                      double scoreMismatch, double scoreIndel)
{
    ContigsCompactor cc;
    cc.SetVerbose(false);
    cc.SetFractionLossScore(fracLossScore);
    cc.SetMinOverlap(fracMinOverlap);
    cc.SetMinOverlapLen(minOverlapLen);
    cc.SetMaxOverlapLenClip(maxOverlapClipLen);
    cc.SetMinOverlapLenWithScaffold(minOverlapLenWithScaffold);
    cc.SetMismatchScore(scoreMismatch);
    cc.SetIndelScore(scoreIndel);
}

// out[0]=res (0/1/2) out[1]=scoreMax out[2]=posRowEnd out[3]=posColEnd out[4]=nclip
// out[5]=bcontained out[6]=IsContainment() out[7]=GetOverlapSize()
// out[5..7] are -1 when the reference returned before the traceback (res==0, relax==0).
int cmref_evaluate(const char *s1, const char *s2, int relax, int32_t *out)
{
    FastaSequence a, b;
    a.SetName("s1"); a.SetSeq(std::string(s1));
    b.SetName("s2"); b.SetSeq(std::string(s2));
    ContigsCompactor cc;
    ContigsCompactorAction act;
    g_hook[0] = g_hook[1] = g_hook[2] = g_hook[3] = -1;
    int res = cc.Evaluate(&a, &b, act, relax != 0);
    out[0] = res;
    out[1] = g_hook[0]; out[2] = g_hook[1]; out[3] = g_hook[2]; out[4] = g_hook[3];
    if (res == 0 && !relax) {
        out[5] = out[6] = out[7] = -1;
        g_merged.clear();
    } else {
        out[5] = act.bcontained ? 1 : 0;
        out[6] = act.IsContainment() ? 1 : 0;
        out[7] = act.GetOverlapSize();
        g_merged = act.GetMerged();
    }
    return 0;
}

// IsScoreSignificant (ContigsCompactor.cpp:1876-1976) on its own.
int cmref_is_score_significant(int scoreMax, int sz1, int sz2, int rowEnd, int colEnd, int nclip)
{
    ContigsCompactor cc;
    return cc.IsScoreSignificant(scoreMax, sz1, sz2, rowEnd, colEnd, nclip);
}

// Merged string of the last cmref_evaluate on this thread; returns its length.
int64_t cmref_last_merged(char *buf, int64_t cap)
{
    int64_t n = (int64_t)g_merged.size();
    if (buf && cap > 0) {
        int64_t k = n < cap - 1 ? n : cap - 1;
        memcpy(buf, g_merged.data(), (size_t)k);
        buf[k] = 0;
    }
    return n;
}

} // extern "C"

// Candidate filter of the pairwise phase: QuickCheckerContigsMatch
// (ContigsCompactor.cpp:1982-2095) built on node i, asked about node j, exactly as
// threadQuickCheck does (ContigsCompactor.cpp:1089).  Both sequences must be >= 30
// bases (the reference reads out of bounds otherwise).
extern "C" int cmref_quickcheck(const char *si, const char *sj, int kmerLen)
{
    FastaSequence a, b;
    a.SetName("si"); a.SetSeq(std::string(si));
    b.SetName("sj"); b.SetSeq(std::string(sj));
    QuickCheckerContigsMatch q(&a, kmerLen);
    return q.IsMatchFeasible(&b) ? 1 : 0;
}

// FastaSequence::RevsereComplement (fastareader.cpp, GenSeqsUtils.cpp:24-61).
extern "C" void cmref_revcomp(const char *s, char *out)
{
    FastaSequence a;
    a.SetName("s"); a.SetSeq(std::string(s));
    a.RevsereComplement();
    memcpy(out, a.c_str(), (size_t)a.size() + 1);
}
