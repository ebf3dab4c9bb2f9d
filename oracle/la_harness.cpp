// oracle/la_harness.cpp -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
//
// Thin C-ABI shim around the UNMODIFIED reference affine local aligner
// (TERefiner/algorithms/local_alignment.cpp: aln_local_core :512-745, aln_global_core :328-508,
// aln_stdaln_aux :746-825, LocalAlignment::optAlign :1036-1049), compiled by oracle/build_ref.sh
// from the sources where they lie under /root/reference into oracle/_ref/libla_ref.so.
// Tests and bench.py's cpu_baseline leg call it; nothing in the product path links or loads it.
#include <cstdint>
#include <cstring>
#include <string>

#include "local_alignment.h"
#include "stdaln.h"

extern unsigned char aln_nt4_table[256];       // local_alignment.cpp:32 (not declared in stdaln.h)

extern "C" {

// LocalAlignment::optAlign as TERefiner's callers use it (main.cpp:209-212, scaffolding.cpp:103-105):
// out = {start_ref, end_ref, start_sgmt, end_sgmt}, 1-based.
void laref_opt_align(const char *ref, const char *sgmt, int32_t *out)
{
    std::string a(ref), b(sgmt);
    LocalAlignment la;
    int v[4] = {-1, -1, -1, -1};
    la.optAlign(a, b, v[0], v[1], v[2], v[3]);
    for (int k = 0; k < 4; ++k) out[k] = v[k];
}

// The call optAlign makes (aln_stdaln(ref, sgmt, &aln_param_blast, 0, 1)), with the score as well:
// out = {score, start1, end1, start2, end2, path_len}.  The reference reads path[-1] when nothing aligns
// (score < 1, :611-614 then :817-821): the caller must not ask for such pairs; check with laref_forward_score first.
void laref_stdaln_local(const char *ref, const char *sgmt, int32_t *out)
{
    AlnAln *aa = aln_stdaln(ref, sgmt, &aln_param_blast, ALN_TYPE_LOCAL, 1);
    out[0] = aa->score; out[1] = aa->start1; out[2] = aa->end1; out[3] = aa->start2; out[4] = aa->end2; out[5] = aa->path_len;
    aln_free_AlnAln(aa);
}

// LocalAlignment::align (:1053-1090): the best alignment and the two beside it.  out[12], every entry starts at -1 as in
// the reference's callers; entries the method leaves untouched stay -1.  The method reads path[-1] when one of its
// alignments finds nothing: the caller must keep such pairs away (check the three inputs with laref_forward_score).
void laref_align(const char *ref, const char *sgmt, int32_t *out)
{
    std::string a(ref), b(sgmt);
    LocalAlignment la;
    int v[12];
    for (int k = 0; k < 12; ++k) v[k] = -1;
    la.align(a, b, v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11]);
    for (int k = 0; k < 12; ++k) out[k] = v[k];
}

// Forward pass only (aln_local_core with path == 0, :615): the local score.
int32_t laref_forward_score(const char *ref, const char *sgmt)
{
    const int len1 = (int)strlen(ref), len2 = (int)strlen(sgmt);
    if (len1 == 0 || len2 == 0) return -1;
    std::string a(len1, 0), b(len2, 0);
    for (int i = 0; i < len1; ++i) a[i] = (char)aln_nt4_table[(unsigned char)ref[i]];
    for (int j = 0; j < len2; ++j) b[j] = (char)aln_nt4_table[(unsigned char)sgmt[j]];
    int path_len = 0, subo = 0;
    return aln_local_core((unsigned char *)&a[0], len1, (unsigned char *)&b[0], len2, &aln_param_blast, 0, &path_len, 1, &subo);
}

} // extern "C"
